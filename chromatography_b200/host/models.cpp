// Host-side model structs of the solver boundary (see include/chrom/models.hpp for the reference
// file:line of each item).  Parameters only: the right-hand sides live in the CUDA kernels.
#include "chrom/models.hpp"

#include <cmath>
#include <set>
#include <stdexcept>

namespace chrom {

// ---- TemporalInjection ------------------------------------------------------------------------------
TemporalInjection TemporalInjection::dirac(double time, double amount) {
  TemporalInjection t;
  t.kind_ = kDirac;
  t.p_[0] = time;
  t.p_[1] = amount;
  return t;
}

TemporalInjection TemporalInjection::gaussian(double center, double width, double peak_concentration) {
  TemporalInjection t;
  t.kind_ = kGaussian;
  t.p_[0] = center;
  t.p_[1] = width;
  t.p_[2] = peak_concentration;
  return t;
}

TemporalInjection TemporalInjection::rectangle(double start, double end, double concentration) {
  if (!(end > start)) throw std::invalid_argument("Rectangle end must be > start");  // injection.rs:376
  TemporalInjection t;
  t.kind_ = kRectangle;
  t.p_[0] = start;
  t.p_[1] = end;
  t.p_[2] = concentration;
  return t;
}

TemporalInjection TemporalInjection::custom(std::function<double(double)> f) {
  TemporalInjection t;
  t.kind_ = kCustom;
  t.f_ = std::make_shared<const std::function<double(double)>>(std::move(f));
  return t;
}

TemporalInjection TemporalInjection::none() { return TemporalInjection(); }

double TemporalInjection::evaluate(double t) const {
  switch (kind_) {
    case kDirac: return t == p_[0] ? p_[1] : 0.0;
    case kGaussian: {
      const double distance = (t - p_[0]) / p_[1];
      return p_[2] * std::exp(-distance * distance / 2.0);
    }
    case kRectangle: return (t >= p_[0] && t < p_[1]) ? p_[2] : 0.0;
    case kCustom: return (*f_)(t);
    default: return 0.0;
  }
}

std::vector<double> TemporalInjection::evaluate_series(const std::vector<double>& times) const {
  std::vector<double> r(times.size());
  for (std::size_t i = 0; i < times.size(); ++i) r[i] = evaluate(times[i]);
  return r;
}

std::string TemporalInjection::debug_string() const {
  switch (kind_) {
    case kDirac: return "Dirac { time: " + rust_f64_debug(p_[0]) + ", amount: " + rust_f64_debug(p_[1]) + " }";
    case kGaussian:
      return "Gaussian { center: " + rust_f64_debug(p_[0]) + ", width: " + rust_f64_debug(p_[1]) +
             ", peak_concentration: " + rust_f64_debug(p_[2]) + " }";
    case kRectangle:
      return "Rectangle { start: " + rust_f64_debug(p_[0]) + ", end: " + rust_f64_debug(p_[1]) +
             ", concentration: " + rust_f64_debug(p_[2]) + " }";
    case kCustom: return "Custom { function: \"<user-defined>\" }";
    default: return "None";
  }
}

bool TemporalInjection::same_profile(const TemporalInjection& o) const {
  if (kind_ != o.kind_) return false;
  if (kind_ == kCustom) return f_ == o.f_;
  // bit patterns, not ==: -0.0 vs 0.0 or NaN parameters must not be merged
  for (int i = 0; i < 3; ++i) {
    std::uint64_t a, b;
    static_assert(sizeof a == sizeof p_[i], "f64");
    __builtin_memcpy(&a, &p_[i], 8);
    __builtin_memcpy(&b, &o.p_[i], 8);
    if (a != b) return false;
  }
  return true;
}

// ---- LangmuirSingle ------------------------------------------------------------------------------------
LangmuirSingle::LangmuirSingle(double lambda, double langmuir_k, double port_number, double porosity, double velocity,
                               double column_length, std::size_t spatial_points, TemporalInjection injection) {
  if (!(porosity > 0.0 && porosity < 1.0))
    throw std::invalid_argument("Porosity must be in ]0,1[, got " + rust_f64_to_string(porosity));
  if (!(column_length > 0.0))
    throw std::invalid_argument("Column length must be positive, got " + rust_f64_to_string(column_length));
  if (!(spatial_points >= 2))
    throw std::invalid_argument("Need at least 2 spatial points, got " + std::to_string(spatial_points));
  lambda_ = lambda;
  langmuir_k_ = langmuir_k;
  port_number_ = port_number;
  column_length_ = column_length;
  n_points_ = spatial_points;
  fe_ = (1.0 - porosity) / porosity;
  ue_ = velocity / porosity;
  dz_ = column_length / (double)spatial_points;
  injection_ = std::move(injection);
}

LangmuirSingle LangmuirSingle::from_fields(double lambda, double langmuir_k, double port_number, double column_length,
                                           std::size_t n_points, double dz, double fe, double ue,
                                           TemporalInjection injection) {
  LangmuirSingle m;
  m.lambda_ = lambda;
  m.langmuir_k_ = langmuir_k;
  m.port_number_ = port_number;
  m.column_length_ = column_length;
  m.n_points_ = n_points;
  m.dz_ = dz;
  m.fe_ = fe;
  m.ue_ = ue;
  m.injection_ = std::move(injection);
  return m;
}

PhysicalState LangmuirSingle::setup_initial_state() const {
  return PhysicalState(PhysicalQuantity::Concentration, PhysicalData::Vector(DVector(n_points_, 0.0)));
}

std::optional<std::string> LangmuirSingle::description() const {
  return std::string("Using Langmuir isotherm with time varying inlet BC. Read from Physical State metadata.");
}

Result<Unit> LangmuirSingle::set_injections(const std::map<std::optional<std::string>, TemporalInjection>& injections) {
  auto it = injections.find(std::nullopt);  // per-species keys are ignored: no named species to target
  if (it != injections.end()) set_injection(it->second);
  return Ok();
}

// ---- SpeciesParams ----------------------------------------------------------------------------------------
Result<Unit> SpeciesParams::validate() const {
  if (lambda < 0.0) return Err("Species '" + name + "' : lambda must be >= 0, got " + rust_f64_to_string(lambda));
  if (langmuir_k <= 0.0)
    return Err("Species '" + name + "' : Langmuir K̃ must be > 0, got " + rust_f64_to_string(langmuir_k));
  if (port_number == 0)
    return Err("Species '" + name + "' : Port number must be > 0 (competitiveness), got " + std::to_string(port_number));
  return Ok();
}

// ---- LangmuirMulti ------------------------------------------------------------------------------------------
Result<LangmuirMulti> LangmuirMulti::create(std::vector<SpeciesParams> species, std::size_t n_points, double porosity,
                                            double velocity, double column_length) {
  if (species.empty()) return Err("Langmuir multi species must have at least one species.");
  for (const auto& sp : species) {
    auto v = sp.validate();
    if (v.is_err()) return Err(v.unwrap_err());
  }
  std::set<std::string> seen;
  for (const auto& sp : species)
    if (!seen.insert(sp.name).second)
      return Err("Duplicate speciees name '" + sp.name + "' in intial species list");  // [sic] langmuir_multi.rs:366-369
  if (n_points < 2)
    return Err("number of points must have at least two points, got " + std::to_string(n_points) + ".");
  if (porosity <= 0.0 || porosity >= 1.0) return Err("porosity must be in ]0, 1[, got " + rust_f64_to_string(porosity) + ".");
  if (velocity <= 0.0) return Err("velocity must be strictly positive, got " + rust_f64_to_string(velocity) + ".");
  if (column_length <= 0.0)
    return Err("column length must be strictly positive, got " + rust_f64_to_string(column_length) + ".");
  LangmuirMulti m;
  m.species_ = std::move(species);
  m.n_points_ = n_points;
  m.porosity_ = porosity;
  m.velocity_ = velocity;
  m.column_length_ = column_length;
  m.dz_ = column_length / (double)n_points;
  m.fe_ = (1.0 - porosity) / porosity;
  m.ue_ = velocity / porosity;
  m.stationary_fraction_ = 1.0 - porosity;
  return m;
}

LangmuirMulti LangmuirMulti::from_fields(std::vector<SpeciesParams> species, std::size_t n_points, double porosity,
                                         double velocity, double column_length, double dz, double fe, double ue,
                                         double stationary_fraction) {
  LangmuirMulti m;
  m.species_ = std::move(species);
  m.n_points_ = n_points;
  m.porosity_ = porosity;
  m.velocity_ = velocity;
  m.column_length_ = column_length;
  m.dz_ = dz;
  m.fe_ = fe;
  m.ue_ = ue;
  m.stationary_fraction_ = stationary_fraction;
  return m;
}

Result<Unit> LangmuirMulti::add_species(SpeciesParams species) {
  auto v = species.validate();
  if (v.is_err()) return v;
  for (const auto& s : species_)
    if (s.name == species.name) return Err("Species '" + species.name + "' already exists in this model");
  species_.push_back(std::move(species));
  return Ok();
}

void LangmuirMulti::set_injection_all(const TemporalInjection& injection) {
  for (auto& sp : species_) sp.injection = injection;
}

Result<Unit> LangmuirMulti::set_injection_for(const std::string& species, TemporalInjection injection) {
  for (auto& sp : species_)
    if (sp.name == species) {
      sp.injection = std::move(injection);
      return Ok();
    }
  return Err("Species '" + species + "' not found in model");
}

std::vector<std::string> LangmuirMulti::species_names() const {
  std::vector<std::string> r;
  for (const auto& s : species_) r.push_back(s.name);
  return r;
}

PhysicalState LangmuirMulti::setup_initial_state() const {
  return PhysicalState(PhysicalQuantity::Concentration, PhysicalData::from_matrix(DMatrix::zeros(n_points_, n_species())));
}

std::optional<std::string> LangmuirMulti::description() const {
  return std::string(
      "Competitive Langmuir adsorption model for n species with upwind spatial discretisation and matrix Jacobian "
      "inversion.");
}

Result<Unit> LangmuirMulti::set_injections(const std::map<std::optional<std::string>, TemporalInjection>& injections) {
  auto def = injections.find(std::nullopt);  // step 1: the default injection to all species
  if (def != injections.end()) set_injection_all(def->second);
  for (const auto& kv : injections)          // step 2: per-species overrides
    if (kv.first) {
      auto r = set_injection_for(*kv.first, kv.second);
      if (r.is_err()) return r;
    }
  return Ok();
}

}  // namespace chrom
