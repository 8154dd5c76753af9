// chrom/models.hpp — C++ mirror of the model structs of chrom-rs that the CUDA backend accepts.
//
//   TemporalInjection   src/models/injection.rs:92-465
//   LangmuirSingle      src/models/langmuir_single.rs:154-370, PhysicalModel impl 418-560
//   SpeciesParams       src/models/langmuir_multi.rs:136-245
//   LangmuirMulti       src/models/langmuir_multi.rs:287-612, PhysicalModel impl 717-940
//
// Constructors, accessors, validation rules and error texts follow the reference.  Where the
// reference `assert!`s (a panic), the C++ mirror throws std::invalid_argument with the same text.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "chrom/physics.hpp"

namespace chrom {

// ---- TemporalInjection --------------------------------------------------------------------------------
class TemporalInjection {
 public:
  enum Kind { kNone = 0, kDirac = 1, kGaussian = 2, kRectangle = 3, kCustom = 4 };  // 0..3 = CHROM_INJ_*

  static TemporalInjection dirac(double time, double amount);
  static TemporalInjection gaussian(double center, double width, double peak_concentration);
  static TemporalInjection rectangle(double start, double end, double concentration);  // throws unless end > start
  static TemporalInjection custom(std::function<double(double)> f);
  static TemporalInjection none();
  TemporalInjection() = default;  // None

  // Dirac: amount iff t == time (bit-exact); Gaussian: peak * exp(-d*d/2.0), d = (t-center)/width;
  // Rectangle: concentration iff start <= t < end; Custom: f(t); None: 0.0   (injection.rs:410-446)
  double evaluate(double t) const;
  std::vector<double> evaluate_series(const std::vector<double>& times) const;

  Kind kind() const { return kind_; }
  double p0() const { return p_[0]; }
  double p1() const { return p_[1]; }
  double p2() const { return p_[2]; }
  std::string debug_string() const;  // the reference's Debug output, e.g. `Dirac { time: 5.0, amount: 0.1 }`
  // Identity used to de-duplicate injection tables of a batch (Custom closures compare by address).
  bool same_profile(const TemporalInjection& o) const;

 private:
  Kind kind_ = kNone;
  double p_[3] = {0.0, 0.0, 0.0};
  std::shared_ptr<const std::function<double(double)>> f_;
};

// ---- LangmuirSingle -------------------------------------------------------------------------------------
class LangmuirSingle : public PhysicalModel {
 public:
  // fe = (1.0 - porosity) / porosity, ue = velocity / porosity, dz = column_length / spatial_points
  // (langmuir_single.rs:261-263).  Throws std::invalid_argument with the reference's assert texts.
  LangmuirSingle(double lambda, double langmuir_k, double port_number, double porosity, double velocity,
                 double column_length, std::size_t spatial_points, TemporalInjection injection);
  // The serde form (src/config/model.rs:111-148): stored fields verbatim, nothing recomputed.
  static LangmuirSingle from_fields(double lambda, double langmuir_k, double port_number, double column_length,
                                    std::size_t n_points, double dz, double fe, double ue,
                                    TemporalInjection injection);

  const TemporalInjection& injection() const { return injection_; }
  void set_injection(TemporalInjection injection) { injection_ = std::move(injection); }
  std::size_t spatial_points() const { return n_points_; }
  double column_length() const { return column_length_; }
  double lambda() const { return lambda_; }
  double langmuir_k() const { return langmuir_k_; }
  double port_number() const { return port_number_; }
  double dz() const { return dz_; }
  double fe() const { return fe_; }
  double ue() const { return ue_; }

  std::size_t points() const override { return n_points_; }
  PhysicalState setup_initial_state() const override;  // Concentration: Vector zeros(n_points)
  const char* name() const override { return "Langmuir single specie with temporal injection"; }
  std::optional<std::string> description() const override;
  Result<Unit> set_injections(const std::map<std::optional<std::string>, TemporalInjection>& injections) override;

 private:
  LangmuirSingle() = default;
  double lambda_ = 0, langmuir_k_ = 0, port_number_ = 0, column_length_ = 0;
  std::size_t n_points_ = 0;
  double dz_ = 0, fe_ = 0, ue_ = 0;
  TemporalInjection injection_;
};

// ---- SpeciesParams ----------------------------------------------------------------------------------------
struct SpeciesParams {
  std::string name;
  double lambda;
  double langmuir_k;
  std::uint32_t port_number;
  TemporalInjection injection;

  SpeciesParams(std::string name, double lambda, double langmuir_k, std::uint32_t port_number,
                TemporalInjection injection)
      : name(std::move(name)), lambda(lambda), langmuir_k(langmuir_k), port_number(port_number),
        injection(std::move(injection)) {}
  Result<Unit> validate() const;  // langmuir_multi.rs:218-245
};

// ---- LangmuirMulti ------------------------------------------------------------------------------------------
class LangmuirMulti : public PhysicalModel {
 public:
  // `LangmuirMulti::new` (langmuir_multi.rs:342-415): validation order and error texts of the reference.
  static Result<LangmuirMulti> create(std::vector<SpeciesParams> species, std::size_t n_points, double porosity,
                                      double velocity, double column_length);
  // The serde form: stored derived fields verbatim.
  static LangmuirMulti from_fields(std::vector<SpeciesParams> species, std::size_t n_points, double porosity,
                                   double velocity, double column_length, double dz, double fe, double ue,
                                   double stationary_fraction);

  Result<Unit> add_species(SpeciesParams species);
  void set_injection_all(const TemporalInjection& injection);
  Result<Unit> set_injection_for(const std::string& species, TemporalInjection injection);
  std::size_t n_species() const { return species_.size(); }
  double porosity() const { return porosity_; }
  double velocity() const { return velocity_; }
  double column_length() const { return column_length_; }
  std::size_t spatial_points() const { return n_points_; }
  const std::vector<SpeciesParams>& species_params() const { return species_; }
  std::vector<std::string> species_names() const;
  double dz() const { return dz_; }
  double fe() const { return fe_; }
  double ue() const { return ue_; }
  double stationary_fraction() const { return stationary_fraction_; }

  std::size_t points() const override { return n_points_; }
  PhysicalState setup_initial_state() const override;  // Concentration: Matrix zeros(n_points, n_species)
  const char* name() const override { return "Langmuir Multi-Species"; }
  std::optional<std::string> description() const override;
  Result<Unit> set_injections(const std::map<std::optional<std::string>, TemporalInjection>& injections) override;

 private:
  LangmuirMulti() = default;
  std::vector<SpeciesParams> species_;
  std::size_t n_points_ = 0;
  double porosity_ = 0, velocity_ = 0, column_length_ = 0, dz_ = 0, fe_ = 0, ue_ = 0, stationary_fraction_ = 0;
};

}  // namespace chrom
