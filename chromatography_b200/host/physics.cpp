// Host-side state containers of the solver boundary (see include/chrom/physics.hpp for the
// reference file:line each type mirrors).  No numerical work of the hot path happens here.
#include "chrom/physics.hpp"

#include <stdexcept>

namespace chrom {

const PhysicalQuantity PhysicalQuantity::Concentration(PhysicalQuantity::kConcentration);
const PhysicalQuantity PhysicalQuantity::Temperature(PhysicalQuantity::kTemperature);
const PhysicalQuantity PhysicalQuantity::Pressure(PhysicalQuantity::kPressure);
const PhysicalQuantity PhysicalQuantity::Velocity(PhysicalQuantity::kVelocity);

std::string PhysicalQuantity::to_string() const {
  switch (kind_) {
    case kConcentration: return "Concentration";
    case kTemperature: return "Temperature";
    case kPressure: return "Pressure";
    case kVelocity: return "Velocity";
    default: return name_;
  }
}

// ---- DMatrix ------------------------------------------------------------------------------------
DMatrix DMatrix::identity(std::size_t r, std::size_t c) {
  DMatrix m(r, c, 0.0);
  for (std::size_t i = 0; i < r && i < c; ++i) m(i, i) = 1.0;
  return m;
}

DMatrix DMatrix::from_row_slice(std::size_t r, std::size_t c, const std::vector<double>& row_major) {
  if (row_major.size() != r * c) throw std::invalid_argument("Matrix init. error: the slice did not contain the right number of elements.");
  DMatrix m(r, c);
  for (std::size_t i = 0; i < r; ++i)
    for (std::size_t j = 0; j < c; ++j) m(i, j) = row_major[i * c + j];
  return m;
}

DMatrix DMatrix::from_column_slice(std::size_t r, std::size_t c, const std::vector<double>& col_major) {
  if (col_major.size() != r * c) throw std::invalid_argument("Matrix init. error: the slice did not contain the right number of elements.");
  DMatrix m(r, c);
  m.data = col_major;
  return m;
}

DVector DMatrix::column(std::size_t j) const {
  return DVector(data.begin() + (std::ptrdiff_t)(j * nrows), data.begin() + (std::ptrdiff_t)((j + 1) * nrows));
}

DVector DMatrix::row(std::size_t i) const {
  DVector r(ncols);
  for (std::size_t j = 0; j < ncols; ++j) r[j] = (*this)(i, j);
  return r;
}

// ---- PhysicalData ---------------------------------------------------------------------------------
PhysicalData PhysicalData::uniform_array(const std::vector<std::size_t>& shape, double value) {
  std::size_t n = 1;
  for (std::size_t s : shape) n *= s;
  return Array(ArrayD{shape, std::vector<double>(n, value)});
}

std::size_t PhysicalData::ndim() const {
  switch (v_.index()) {
    case 0: return 0;
    case 1: return 1;
    case 2: return 2;
    default: return std::get<3>(v_).shape.size();
  }
}

std::vector<std::size_t> PhysicalData::shape() const {
  switch (v_.index()) {
    case 0: return {};
    case 1: return {std::get<1>(v_).size()};
    case 2: return {std::get<2>(v_).nrows, std::get<2>(v_).ncols};
    default: return std::get<3>(v_).shape;
  }
}

std::size_t PhysicalData::len() const {
  switch (v_.index()) {
    case 0: return 1;
    case 1: return std::get<1>(v_).size();
    case 2: return std::get<2>(v_).data.size();
    default: return std::get<3>(v_).data.size();
  }
}

double PhysicalData::as_scalar() const {
  if (const double* s = std::get_if<0>(&v_)) return *s;
  throw std::logic_error("Expected scalar data");
}
const DVector& PhysicalData::as_vector() const {
  if (const DVector* v = std::get_if<1>(&v_)) return *v;
  throw std::logic_error("Expected vector data");
}
const DMatrix& PhysicalData::as_matrix() const {
  if (const DMatrix* m = std::get_if<2>(&v_)) return *m;
  throw std::logic_error("Expected matrix data");
}
const ArrayD& PhysicalData::as_array() const {
  if (const ArrayD* a = std::get_if<3>(&v_)) return *a;
  throw std::logic_error("Expected array data");
}
std::optional<double> PhysicalData::try_as_scalar() const {
  if (const double* s = std::get_if<0>(&v_)) return *s;
  return std::nullopt;
}

std::optional<DVector> PhysicalData::get_column(std::size_t col_index) const {
  const DMatrix* m = std::get_if<2>(&v_);
  if (!m || col_index >= m->ncols) return std::nullopt;
  return m->column(col_index);
}

std::optional<DVector> PhysicalData::get_row(std::size_t row_index) const {
  const DMatrix* m = std::get_if<2>(&v_);
  if (!m || row_index >= m->nrows) return std::nullopt;
  return m->row(row_index);
}

std::vector<double>& PhysicalData::raw() {
  switch (v_.index()) {
    case 1: return std::get<1>(v_);
    case 2: return std::get<2>(v_).data;
    case 3: return std::get<3>(v_).data;
    default: throw std::logic_error("scalar data has no element storage");
  }
}

const double* PhysicalData::values() const {
  switch (v_.index()) {
    case 0: return &std::get<0>(v_);
    case 1: return std::get<1>(v_).data();
    case 2: return std::get<2>(v_).data.data();
    default: return std::get<3>(v_).data.data();
  }
}

namespace {
PhysicalData add_scalar(const PhysicalData& a, double x) {
  PhysicalData r = a;
  r.apply([x](double e) { return e + x; });
  return r;
}
}  // namespace

PhysicalData operator+(const PhysicalData& a, const PhysicalData& b) {
  if (a.is_scalar() && b.is_scalar()) return PhysicalData::Scalar(a.as_scalar() + b.as_scalar());
  if (a.is_scalar()) return add_scalar(b, a.as_scalar());
  if (b.is_scalar()) return add_scalar(a, b.as_scalar());
  if (a.is_vector() && b.is_vector()) {
    const DVector &x = a.as_vector(), &y = b.as_vector();
    if (x.size() != y.size()) throw std::logic_error("Vector lengths must match");
    DVector r(x.size());
    for (std::size_t i = 0; i < x.size(); ++i) r[i] = x[i] + y[i];
    return PhysicalData::Vector(std::move(r));
  }
  if (a.is_matrix() && b.is_matrix()) {
    const DMatrix &x = a.as_matrix(), &y = b.as_matrix();
    if (x.shape() != y.shape()) throw std::logic_error("Matrix dimensions must match");
    DMatrix r(x.nrows, x.ncols);
    for (std::size_t i = 0; i < x.data.size(); ++i) r.data[i] = x.data[i] + y.data[i];
    return PhysicalData::Matrix(std::move(r));
  }
  if (a.is_array() && b.is_array()) {
    const ArrayD &x = a.as_array(), &y = b.as_array();
    if (x.shape != y.shape) throw std::logic_error("Array dimensions must match");
    ArrayD r{x.shape, std::vector<double>(x.data.size())};
    for (std::size_t i = 0; i < x.data.size(); ++i) r.data[i] = x.data[i] + y.data[i];
    return PhysicalData::Array(std::move(r));
  }
  throw std::logic_error("Cannot add different PhysicalData types (except with Scalar)");
}

PhysicalData operator*(const PhysicalData& a, double s) {
  PhysicalData r = a;
  r *= s;
  return r;
}

PhysicalData& PhysicalData::operator*=(double s) {
  apply([s](double e) { return e * s; });
  return *this;
}

// ---- PhysicalState ----------------------------------------------------------------------------------
const PhysicalData* PhysicalState::get(const PhysicalQuantity& q) const {
  auto it = quantities.find(q);
  return it == quantities.end() ? nullptr : &it->second;
}

PhysicalData* PhysicalState::get_mut(const PhysicalQuantity& q) {
  auto it = quantities.find(q);
  return it == quantities.end() ? nullptr : &it->second;
}

std::optional<PhysicalData> PhysicalState::remove(const PhysicalQuantity& q) {
  auto it = quantities.find(q);
  if (it == quantities.end()) return std::nullopt;
  PhysicalData d = std::move(it->second);
  quantities.erase(it);
  return d;
}

std::vector<PhysicalQuantity> PhysicalState::available_quantities() const {
  std::vector<PhysicalQuantity> r;
  for (const auto& kv : quantities) r.push_back(kv.first);
  return r;
}

std::optional<double> PhysicalState::get_metadata(const std::string& key) const {
  auto it = metadata_.find(key);
  if (it == metadata_.end()) return std::nullopt;
  return it->second;
}

std::size_t PhysicalState::memory_bytes() const {
  std::size_t n = 0;
  for (const auto& kv : quantities) n += kv.second.memory_bytes();
  return n;
}

PhysicalState operator+(PhysicalState lhs, const PhysicalState& rhs) {
  for (const auto& kv : rhs.quantities) {
    auto it = lhs.quantities.find(kv.first);
    if (it != lhs.quantities.end()) it->second = it->second + kv.second;  // both have it: add
    else lhs.quantities.emplace(kv.first, kv.second);                     // only rhs has it: include
  }
  return lhs;
}

PhysicalState operator*(PhysicalState lhs, double scalar) {
  for (auto& kv : lhs.quantities) kv.second *= scalar;
  return lhs;
}

// ---- exporter helpers ------------------------------------------------------------------------------------
std::vector<double> outlet_data(const PhysicalQuantity& quantity, const std::vector<PhysicalState>& trajectory,
                                std::size_t idx) {
  if (idx >= trajectory.size()) return {};
  const PhysicalData* d = trajectory[idx].get(quantity);
  if (!d) return {};
  if (d->is_scalar()) return {d->as_scalar()};
  if (d->is_vector()) {
    const DVector& v = d->as_vector();
    if (v.empty()) return {};
    return {v.back()};
  }
  if (d->is_matrix() && d->as_matrix().nrows > 0) {
    const DMatrix& m = d->as_matrix();
    std::vector<double> r(m.ncols);
    for (std::size_t s = 0; s < m.ncols; ++s) r[s] = m(m.nrows - 1, s);
    return r;
  }
  return {};
}

std::vector<std::size_t> sample_indices(std::size_t total, std::optional<std::size_t> n) {
  std::vector<std::size_t> r;
  if (!n || *n == 0 || *n >= total) {
    r.resize(total);
    for (std::size_t i = 0; i < total; ++i) r[i] = i;
    return r;
  }
  if (*n == 1) return {0};
  r.resize(*n);
  for (std::size_t i = 0; i < *n; ++i) r[i] = (i * (total - 1)) / (*n - 1);
  r.back() = total - 1;
  return r;
}

}  // namespace chrom

// ---- Rust float formatting ------------------------------------------------------------------------------
#include <charconv>
#include <cmath>

namespace chrom {

std::string rust_f64_to_string(double x) {
  if (std::isnan(x)) return "NaN";
  if (std::isinf(x)) return x > 0 ? "inf" : "-inf";
  char buf[800];  // fixed notation of a subnormal needs up to ~770 characters
  auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::fixed);  // shortest round-trip, no exponent
  return std::string(buf, r.ptr);
}

std::string rust_f64_debug(double x) {
  std::string s = rust_f64_to_string(x);
  if (std::isfinite(x) && s.find('.') == std::string::npos) s += ".0";
  return s;
}

}  // namespace chrom
