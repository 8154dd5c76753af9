// chrom/physics.hpp — C++ mirror of the state containers that cross chrom-rs's solver boundary.
//
//   PhysicalQuantity   src/physics/traits.rs:45-75
//   PhysicalData       src/physics/data.rs:72-100 (constructors 112-222, accessors 841-1000, operators 1052-1109)
//   PhysicalState      src/physics/traits.rs:177-491
//   ComputeContext     src/physics/context.rs:128-190
//   PhysicalModel      src/physics/traits.rs:555-759
//   outlet_data        src/physics/traits.rs:969-989
//   sample_indices     src/physics/traits.rs:1010-1023
//
// These are host-side containers only.  The right-hand side of the models (compute_physics) is NOT
// evaluated on the host in this backend: it lives in the sm_100a kernels of libchrom_cuda.so, and
// there is no CPU fallback (a model the device code does not know is an Err at solve time).
#pragma once
#include <cstddef>
#include <map>
#include <optional>
#include <string>
#include <unordered_map>
#include <variant>
#include <vector>

#include "chrom/result.hpp"

namespace chrom {

// ---- PhysicalQuantity ---------------------------------------------------------------------------
class PhysicalQuantity {
 public:
  enum Kind { kConcentration, kTemperature, kPressure, kVelocity, kCustom };
  static const PhysicalQuantity Concentration;
  static const PhysicalQuantity Temperature;
  static const PhysicalQuantity Pressure;
  static const PhysicalQuantity Velocity;
  static PhysicalQuantity Custom(std::string name) { return PhysicalQuantity(kCustom, std::move(name)); }

  Kind kind() const { return kind_; }
  std::string to_string() const;  // Display: "Concentration", ..., or the custom name
  bool operator==(const PhysicalQuantity& o) const { return kind_ == o.kind_ && name_ == o.name_; }
  bool operator!=(const PhysicalQuantity& o) const { return !(*this == o); }
  bool operator<(const PhysicalQuantity& o) const { return kind_ != o.kind_ ? kind_ < o.kind_ : name_ < o.name_; }

 private:
  PhysicalQuantity(Kind k, std::string n = {}) : kind_(k), name_(std::move(n)) {}
  Kind kind_;
  std::string name_;
};

// ---- dense containers (nalgebra's DVector / DMatrix as far as the hot path uses them) -------------
using DVector = std::vector<double>;

// Column-major, like nalgebra: element (i, j) at data[j * nrows + i].  For LangmuirMulti the state
// is nz x n_species, so each species' profile is contiguous — this is also the device layout.
struct DMatrix {
  std::size_t nrows = 0, ncols = 0;
  std::vector<double> data;

  DMatrix() = default;
  DMatrix(std::size_t r, std::size_t c, double v = 0.0) : nrows(r), ncols(c), data(r * c, v) {}
  static DMatrix zeros(std::size_t r, std::size_t c) { return DMatrix(r, c, 0.0); }
  static DMatrix from_element(std::size_t r, std::size_t c, double v) { return DMatrix(r, c, v); }
  static DMatrix identity(std::size_t r, std::size_t c);
  static DMatrix from_row_slice(std::size_t r, std::size_t c, const std::vector<double>& row_major);
  static DMatrix from_column_slice(std::size_t r, std::size_t c, const std::vector<double>& col_major);

  double& operator()(std::size_t i, std::size_t j) { return data[j * nrows + i]; }
  double operator()(std::size_t i, std::size_t j) const { return data[j * nrows + i]; }
  std::pair<std::size_t, std::size_t> shape() const { return {nrows, ncols}; }
  DVector column(std::size_t j) const;
  DVector row(std::size_t i) const;
};

// N-dimensional array (ndarray::ArrayD in the reference): shape + row-major data.
struct ArrayD {
  std::vector<std::size_t> shape;
  std::vector<double> data;
};

// ---- PhysicalData -----------------------------------------------------------------------------------
class PhysicalData {
 public:
  PhysicalData() : v_(0.0) {}
  static PhysicalData Scalar(double x) { return PhysicalData(x); }
  static PhysicalData Vector(DVector v) { return PhysicalData(std::move(v)); }
  static PhysicalData Matrix(DMatrix m) { return PhysicalData(std::move(m)); }
  static PhysicalData Array(ArrayD a) { return PhysicalData(std::move(a)); }

  static PhysicalData from_scalar(double x) { return Scalar(x); }
  static PhysicalData from_vec(std::vector<double> v) { return Vector(std::move(v)); }
  static PhysicalData from_vector(DVector v) { return Vector(std::move(v)); }
  static PhysicalData from_matrix(DMatrix m) { return Matrix(std::move(m)); }
  static PhysicalData from_array(ArrayD a) { return Array(std::move(a)); }
  static PhysicalData uniform_vector(std::size_t size, double value) { return Vector(DVector(size, value)); }
  static PhysicalData uniform_matrix(std::size_t rows, std::size_t cols, double value) {
    return Matrix(DMatrix(rows, cols, value));
  }
  static PhysicalData uniform_array(const std::vector<std::size_t>& shape, double value);

  bool is_scalar() const { return v_.index() == 0; }
  bool is_vector() const { return v_.index() == 1; }
  bool is_matrix() const { return v_.index() == 2; }
  bool is_array() const { return v_.index() == 3; }
  std::size_t ndim() const;
  std::vector<std::size_t> shape() const;
  std::size_t len() const;
  bool is_empty() const { return len() == 0; }
  std::size_t memory_bytes() const { return len() * sizeof(double); }

  // as_* panic in the reference on a type mismatch; here they throw std::logic_error.
  double as_scalar() const;
  const DVector& as_vector() const;
  const DMatrix& as_matrix() const;
  const ArrayD& as_array() const;
  std::optional<double> try_as_scalar() const;
  const DVector* try_as_vector() const { return std::get_if<1>(&v_); }
  const DMatrix* try_as_matrix() const { return std::get_if<2>(&v_); }
  const ArrayD* try_as_array() const { return std::get_if<3>(&v_); }

  std::optional<DVector> get_column(std::size_t col_index) const;
  std::optional<DVector> get_row(std::size_t row_index) const;

  template <class F>
  void apply(F f) {
    if (auto* s = std::get_if<0>(&v_)) *s = f(*s);
    else for (double& x : raw()) x = f(x);
  }

  // element-wise `+` and `* f64` with the reference's broadcasting rules (data.rs:1052-1109)
  friend PhysicalData operator+(const PhysicalData& a, const PhysicalData& b);
  friend PhysicalData operator*(const PhysicalData& a, double s);
  friend PhysicalData operator*(double s, const PhysicalData& a) { return a * s; }
  PhysicalData& operator*=(double s);

  // flat view of the stored numbers (Scalar: one element)
  const double* values() const;
  std::size_t n_values() const { return len(); }

 private:
  explicit PhysicalData(double x) : v_(x) {}
  explicit PhysicalData(DVector v) : v_(std::move(v)) {}
  explicit PhysicalData(DMatrix m) : v_(std::move(m)) {}
  explicit PhysicalData(ArrayD a) : v_(std::move(a)) {}
  std::vector<double>& raw();
  std::variant<double, DVector, DMatrix, ArrayD> v_;
};

// ---- PhysicalState ------------------------------------------------------------------------------------
class PhysicalState {
 public:
  std::map<PhysicalQuantity, PhysicalData> quantities;

  PhysicalState() = default;
  PhysicalState(PhysicalQuantity quantity, PhysicalData value) { quantities.emplace(std::move(quantity), std::move(value)); }
  static PhysicalState empty() { return PhysicalState(); }

  const PhysicalData* get(const PhysicalQuantity& q) const;
  PhysicalData* get_mut(const PhysicalQuantity& q);
  void set(PhysicalQuantity q, PhysicalData value) { quantities.insert_or_assign(std::move(q), std::move(value)); }
  std::optional<PhysicalData> remove(const PhysicalQuantity& q);
  std::vector<PhysicalQuantity> available_quantities() const;
  std::optional<double> get_metadata(const std::string& key) const;
  void set_metadata(std::string key, double value) { metadata_[std::move(key)] = value; }
  std::size_t memory_bytes() const;
  std::size_t len() const { return quantities.size(); }
  bool is_empty() const { return quantities.empty(); }

  friend PhysicalState operator+(PhysicalState lhs, const PhysicalState& rhs);
  friend PhysicalState operator*(PhysicalState lhs, double scalar);

 private:
  std::unordered_map<std::string, double> metadata_;
};

// ---- ComputeContext -------------------------------------------------------------------------------------
class ComputeContext {
 public:
  ComputeContext(double time, double time_step) : time_(time), time_step_(time_step) {}
  double time() const { return time_; }
  double time_step() const { return time_step_; }

 private:
  double time_, time_step_;
};

// ---- PhysicalModel ----------------------------------------------------------------------------------------
class TemporalInjection;  // chrom/models.hpp

class PhysicalModel {
 public:
  virtual ~PhysicalModel() = default;
  virtual std::size_t points() const = 0;
  virtual PhysicalState setup_initial_state() const = 0;
  virtual const char* name() const = 0;
  virtual std::optional<std::string> description() const { return std::nullopt; }
  // key: std::nullopt = the default injection, else a species name (physics/traits.rs:738-759)
  virtual Result<Unit> set_injections(const std::map<std::optional<std::string>, TemporalInjection>&) { return Ok(); }
  // compute_physics (physics/traits.rs:601-650) is deliberately absent: the RHS of LangmuirSingle /
  // LangmuirMulti is evaluated by the CUDA kernels only.
};

// `f64::to_string()` / `{}`: shortest round-trip digits, never exponent notation, "600" for 600.0.
std::string rust_f64_to_string(double x);
// `{:?}`: the same digits, but integral values keep a trailing ".0".
std::string rust_f64_debug(double x);

// ---- helpers shared by the exporters ----------------------------------------------------------------------
// Outlet (last cell / last row) of trajectory[idx]: one value per species; empty if idx is out of range.
std::vector<double> outlet_data(const PhysicalQuantity& quantity, const std::vector<PhysicalState>& trajectory,
                                std::size_t idx);
// n evenly spaced indices in [0, total): first and last always included; n = nullopt / 0 / >= total: all.
std::vector<std::size_t> sample_indices(std::size_t total, std::optional<std::size_t> n);

}  // namespace chrom
