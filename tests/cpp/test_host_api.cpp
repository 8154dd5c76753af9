// CPU-only tests of the C++ host layer (include/chrom.hpp).  They follow the reference's own unit
// tests for the types that cross the solver boundary — no GPU, no compute call:
//   src/models/injection.rs:470-560            TemporalInjection semantics
//   src/models/langmuir_single.rs:736-900      constructor / injection handling
//   src/models/langmuir_multi.rs:1249-1372, 1618-1720   SpeciesParams / LangmuirMulti::new / injections
//   src/solver/traits.rs:656-780               SolverType / SolverConfiguration
//   src/solver/boundary.rs:442-710             boundaries
//   src/solver/scenario.rs:246-330             Scenario
//   src/solver/methods/rk4.rs:740-800          error paths of solve()
//   src/physics/traits.rs (tests)              PhysicalState algebra, sample_indices, outlet_data
//   validation/dissimilarity.rs (tests)        trapezoid / normalize / rsf
#include <cmath>

#include "chrom.hpp"
#include "minitest.hpp"

using namespace chrom;

// ---- TemporalInjection (injection.rs:470-560) -------------------------------------------------------
TEST(test_dirac_injection) {
  auto injection = TemporalInjection::dirac(10.0, 1.0);
  CHECK_EQ(injection.evaluate(10.0), 1.0);
  CHECK_EQ(injection.evaluate(0.0), 0.0);
  CHECK_EQ(injection.evaluate(20.0), 0.0);
  CHECK_EQ(injection.evaluate(10.0 + 1e-9), 0.0);
}

TEST(test_gaussian_injection) {
  auto injection = TemporalInjection::gaussian(10.0, 2.0, 0.1);
  CHECK(std::fabs(injection.evaluate(10.0) - 0.1) < 1e-10);
  const double expected = 0.1 * 0.606;
  CHECK(std::fabs(injection.evaluate(8.0) - expected) < 0.01);
  CHECK(std::fabs(injection.evaluate(12.0) - expected) < 0.01);
  CHECK(injection.evaluate(0.0) < 0.001);
  CHECK(injection.evaluate(20.0) < 0.001);
}

TEST(test_rectangle_injection) {
  auto injection = TemporalInjection::rectangle(5.0, 15.0, 0.05);
  CHECK_EQ(injection.evaluate(4.0), 0.0);
  CHECK_EQ(injection.evaluate(5.0), 0.05);
  CHECK_EQ(injection.evaluate(10.0), 0.05);
  CHECK_EQ(injection.evaluate(14.9), 0.05);
  CHECK_EQ(injection.evaluate(15.0), 0.0);
}

TEST(test_custom_injection) {
  auto injection = TemporalInjection::custom([](double t) { return t < 10.0 ? 0.01 * t : 0.0; });
  CHECK_EQ(injection.evaluate(0.0), 0.0);
  CHECK_EQ(injection.evaluate(5.0), 0.05);
  CHECK_EQ(injection.evaluate(10.0), 0.0);
}

TEST(test_none_injection) {
  auto injection = TemporalInjection::none();
  CHECK_EQ(injection.evaluate(0.0), 0.0);
  CHECK_EQ(injection.evaluate(100.0), 0.0);
}

TEST(test_evaluate_series) {
  auto injection = TemporalInjection::gaussian(10.0, 2.0, 0.1);
  auto values = injection.evaluate_series({0.0, 5.0, 10.0, 15.0, 20.0});
  CHECK_EQ(values.size(), 5u);
  CHECK(std::fabs(values[2] - 0.1) < 1e-10);
}

TEST(test_rectangle_invalid) { CHECK_THROWS(TemporalInjection::rectangle(10.0, 10.0, 0.05)); }

TEST(test_debug_temporal_injection) {
  CHECK_STR_EQ(TemporalInjection::dirac(0.0, 10.0).debug_string(), "Dirac { time: 0.0, amount: 10.0 }");
  CHECK_STR_EQ(TemporalInjection::gaussian(10.0, 3.0, 0.1).debug_string(),
               "Gaussian { center: 10.0, width: 3.0, peak_concentration: 0.1 }");
  CHECK_STR_EQ(TemporalInjection::none().debug_string(), "None");
}

// ---- LangmuirSingle (langmuir_single.rs:736-900) -------------------------------------------------------
static LangmuirSingle tfa(TemporalInjection inj = TemporalInjection::dirac(5.0, 0.1)) {
  return LangmuirSingle(1.2, 0.4, 2.0, 0.4, 0.001, 0.25, 100, std::move(inj));
}

TEST(test_create_langmuir_single) {
  auto model = tfa();
  CHECK_EQ(model.points(), 100u);
  // derived quantities (langmuir_single.rs:261-263): operation order matters for bit parity
  CHECK_EQ(model.fe(), (1.0 - 0.4) / 0.4);
  CHECK_EQ(model.ue(), 0.001 / 0.4);
  CHECK_EQ(model.dz(), 0.25 / 100.0);
}

TEST(test_single_name_description_initial_state) {
  auto model = tfa();
  CHECK_STR_EQ(model.name(), "Langmuir single specie with temporal injection");
  CHECK(model.description().has_value());
  auto st = model.setup_initial_state();
  const PhysicalData* c = st.get(PhysicalQuantity::Concentration);
  CHECK(c && c->is_vector() && c->as_vector().size() == 100);
  for (double v : c->as_vector()) CHECK_EQ(v, 0.0);
}

TEST(test_invalid_porosity) { CHECK_THROWS(LangmuirSingle(1.2, 0.4, 2.0, 1.5, 0.001, 0.25, 100, TemporalInjection::none())); }
TEST(test_invalid_points) { CHECK_THROWS(LangmuirSingle(1.2, 0.4, 2.0, 0.4, 0.001, 0.25, 1, TemporalInjection::none())); }

TEST(test_set_injection_replaces_profile) {
  auto model = tfa(TemporalInjection::dirac(5.0, 0.1));
  CHECK_EQ(model.injection().evaluate(5.0), 0.1);
  model.set_injection(TemporalInjection::gaussian(10.0, 3.0, 0.2));
  CHECK_EQ(model.injection().evaluate(5.0 + 1e-9 - 1e-9) == 0.1, false);
  CHECK(std::fabs(model.injection().evaluate(10.0) - 0.2) < 1e-12);
  std::map<std::optional<std::string>, TemporalInjection> inj;
  inj.emplace(std::nullopt, TemporalInjection::rectangle(0.0, 1.0, 0.3));
  inj.emplace(std::string("ignored species"), TemporalInjection::dirac(0.0, 9.0));
  CHECK(model.set_injections(inj).is_ok());
  CHECK_EQ(model.injection().evaluate(0.5), 0.3);
}

// ---- SpeciesParams / LangmuirMulti (langmuir_multi.rs:1249-1372) ----------------------------------------
static SpeciesParams sp(const char* name, double lambda = 1.2, double k = 0.4, std::uint32_t n = 2) {
  return SpeciesParams(name, lambda, k, n, TemporalInjection::none());
}

TEST(test_species_validation) {
  CHECK(sp("A", 0.0).validate().is_ok());
  CHECK(sp("A", -0.1).validate().is_err());
  CHECK(sp("A", 1.2, 0.0).validate().is_err());
  CHECK(sp("A", 1.2, -1.0).validate().is_err());
  CHECK(sp("A", 1.2, 0.4, 0).validate().is_err());
  CHECK_STR_EQ(sp("A", -0.5).validate().unwrap_err(), "Species 'A' : lambda must be >= 0, got -0.5");
  CHECK_STR_EQ(sp("B", 1.0, 0.4, 0).validate().unwrap_err(), "Species 'B' : Port number must be > 0 (competitiveness), got 0");
}

TEST(test_new_valid_and_rejections) {
  CHECK(LangmuirMulti::create({sp("A")}, 100, 0.4, 0.001, 0.25).is_ok());
  CHECK_STR_EQ(LangmuirMulti::create({}, 100, 0.4, 0.001, 0.25).unwrap_err(),
               "Langmuir multi species must have at least one species.");
  CHECK_STR_EQ(LangmuirMulti::create({sp("A")}, 1, 0.4, 0.001, 0.25).unwrap_err(),
               "number of points must have at least two points, got 1.");
  CHECK_STR_EQ(LangmuirMulti::create({sp("A")}, 100, 0.0, 0.001, 0.25).unwrap_err(), "porosity must be in ]0, 1[, got 0.");
  CHECK(LangmuirMulti::create({sp("A")}, 100, 1.0, 0.001, 0.25).is_err());
  CHECK_STR_EQ(LangmuirMulti::create({sp("A")}, 100, 0.4, 0.0, 0.25).unwrap_err(),
               "velocity must be strictly positive, got 0.");
  CHECK(LangmuirMulti::create({sp("A")}, 100, 0.4, 0.001, 0.0).is_err());
  CHECK(LangmuirMulti::create({sp("A"), sp("A")}, 100, 0.4, 0.001, 0.25).is_err());
}

TEST(test_new_precomputes_derived_quantities) {  // langmuir_multi.rs:1331-1338
  auto m = LangmuirMulti::create({sp("A")}, 100, 0.4, 0.001, 0.25).unwrap();
  CHECK(std::fabs(m.fe() - 1.5) < 1e-10);
  CHECK(std::fabs(m.ue() - 0.0025) < 1e-10);
  CHECK(std::fabs(m.dz() - 0.25 / 100.0) < 1e-15);
  CHECK_EQ(m.stationary_fraction(), 1.0 - 0.4);
}

TEST(test_add_species_and_names) {
  auto m = LangmuirMulti::create({sp("A")}, 100, 0.4, 0.001, 0.25).unwrap();
  CHECK(m.add_species(sp("B")).is_ok());
  CHECK_EQ(m.n_species(), 2u);
  CHECK_STR_EQ(m.add_species(sp("B")).unwrap_err(), "Species 'B' already exists in this model");
  CHECK(m.add_species(sp("C", -1.0)).is_err());
  CHECK_EQ(m.n_species(), 2u);
  auto names = m.species_names();
  CHECK(names.size() == 2 && names[0] == "A" && names[1] == "B");
  CHECK_STR_EQ(m.name(), "Langmuir Multi-Species");
  CHECK(m.description().has_value());
  CHECK_EQ(m.points(), 100u);
}

TEST(test_multi_initial_state_is_empty_column) {  // langmuir_multi.rs:1472-1488
  auto m = LangmuirMulti::create({sp("A"), sp("B")}, 50, 0.4, 0.001, 0.25).unwrap();
  auto st = m.setup_initial_state();
  const DMatrix& c = st.get(PhysicalQuantity::Concentration)->as_matrix();
  CHECK(c.nrows == 50 && c.ncols == 2);
  for (double v : c.data) CHECK_EQ(v, 0.0);
}

TEST(test_set_injection_all_and_for) {  // langmuir_multi.rs:1635-1720
  auto m = LangmuirMulti::create({sp("A"), sp("B")}, 50, 0.4, 0.001, 0.25).unwrap();
  m.set_injection_all(TemporalInjection::dirac(5.0, 0.1));
  for (const auto& s : m.species_params()) CHECK_EQ(s.injection.evaluate(5.0), 0.1);
  CHECK(m.set_injection_for("B", TemporalInjection::dirac(5.0, 0.7)).is_ok());
  CHECK_EQ(m.species_params()[0].injection.evaluate(5.0), 0.1);
  CHECK_EQ(m.species_params()[1].injection.evaluate(5.0), 0.7);
  CHECK_STR_EQ(m.set_injection_for("Z", TemporalInjection::none()).unwrap_err(), "Species 'Z' not found in model");
  CHECK_EQ(m.species_params()[1].lambda, 1.2);
  std::map<std::optional<std::string>, TemporalInjection> inj;
  inj.emplace(std::nullopt, TemporalInjection::dirac(1.0, 0.2));
  inj.emplace(std::string("A"), TemporalInjection::dirac(1.0, 0.9));
  CHECK(m.set_injections(inj).is_ok());
  CHECK_EQ(m.species_params()[0].injection.evaluate(1.0), 0.9);
  CHECK_EQ(m.species_params()[1].injection.evaluate(1.0), 0.2);
  inj.emplace(std::string("nope"), TemporalInjection::none());
  CHECK(m.set_injections(inj).is_err());
}

// ---- SolverType / SolverConfiguration (traits.rs:656-780) -------------------------------------------------
TEST(test_solver_type_name) {
  CHECK_STR_EQ(SolverType::TimeEvolution(10.0, 100).name(), "TimeEvolution");
  CHECK_STR_EQ(SolverType::Other("Iterative").name(), "Iterative");
}

TEST(test_solver_type_validate) {
  CHECK(SolverType::TimeEvolution(10.0, 100).validate().is_ok());
  CHECK_STR_EQ(SolverType::TimeEvolution(-1.0, 100).validate().unwrap_err(), "Total time must be positive");
  CHECK_STR_EQ(SolverType::TimeEvolution(0.0, 100).validate().unwrap_err(), "Total time must be positive");
  CHECK_STR_EQ(SolverType::TimeEvolution(10.0, 0).validate().unwrap_err(), "TimeSteps must be greater than 0");
}

TEST(test_solver_configuration_factories_and_step) {
  auto te = SolverConfiguration::time_evolution(100.0, 1000);
  CHECK(te.solver_type.kind() == SolverType::kTimeEvolution && te.solver_type.total_time == 100.0 && te.solver_type.time_steps == 1000);
  CHECK(!te.step.has_value());
  auto ws = te.with_step(10);
  CHECK(ws.step == std::optional<std::size_t>(10));
  CHECK(ws.solver_type == te.solver_type);
  CHECK(te.with_step(0).step == std::optional<std::size_t>(0));
  CHECK(te.validate().is_ok());
  CHECK(SolverConfiguration::time_evolution(-1.0, 10).validate().is_err());
}

// ---- boundaries (boundary.rs:442-710) ------------------------------------------------------------------------
static PhysicalState scalar_state(double v) { return PhysicalState(PhysicalQuantity::Concentration, PhysicalData::Scalar(v)); }

TEST(test_axis_convention_display) {
  CHECK_STR_EQ(TimeAxisConvention::first().to_string(), "First");
  CHECK_STR_EQ(TimeAxisConvention::last().to_string(), "Last");
  CHECK_STR_EQ(TimeAxisConvention::at(10).to_string(), "Index (10)");
  CHECK_STR_EQ(TimeAxisConvention::none().to_string(), "None");
}

TEST(test_dimension_boundary) {
  DimensionBoundary d("volume", {PhysicalState(PhysicalQuantity::Concentration, PhysicalData::Vector({1., 2., 3., 4.}))});
  CHECK_STR_EQ(d.name, "volume");
  CHECK_EQ(d.size(), 1u);
  CHECK_EQ(d.first()->get(PhysicalQuantity::Concentration)->as_vector()[2], 3.0);
  DimensionBoundary e("x", {});
  CHECK_THROWS(e.first());  // the reference panics "out of bounds"
  CHECK_STR_EQ(e.validate().unwrap_err(), "Dimensions 'x' must have at least one boundary state");
}

TEST(test_temporal_and_general_boundaries) {
  auto t = DomainBoundaries::temporal(scalar_state(1.0));
  CHECK(t.ndim() == 1 && t.sdim() == 0 && t.is_time_dependent());
  CHECK(t.initial_condition() != nullptr && t.initial_condition()->get(PhysicalQuantity::Concentration)->as_scalar() == 1.0);
  CHECK(t.validate().is_ok());
  // a domain without a time axis (what the solvers reject: "No initial condition found in domain boundaries")
  auto s = DomainBoundaries::create({DimensionBoundary("x", {scalar_state(0.), scalar_state(2.)}),
                                     DimensionBoundary("y", {scalar_state(1.), scalar_state(3.)})},
                                    TimeAxisConvention::none());
  CHECK(s.ndim() == 2 && s.sdim() == 2 && !s.is_time_dependent());
  CHECK(s.initial_condition() == nullptr && !s.time_index().has_value());
  CHECK(s.get_boundary("y") != nullptr && s.get_boundary("z") == nullptr);
  auto m = DomainBoundaries::create({DimensionBoundary("x", {scalar_state(0.), scalar_state(1.)}),
                                     DimensionBoundary("t", {scalar_state(7.)})}, std::nullopt);
  CHECK(m.ndim() == 2 && m.sdim() == 1 && m.time_index() == std::optional<std::size_t>(1));
  CHECK_EQ(m.initial_condition()->get(PhysicalQuantity::Concentration)->as_scalar(), 7.0);
  CHECK_EQ(m.spatial_boundaries().size(), 1u);
}

TEST(test_empty_and_duplicate_boundaries) {
  CHECK_STR_EQ(DomainBoundaries().validate().unwrap_err(), "Dimension boundaries cannot be empty.");
  DomainBoundaries dup({DimensionBoundary("x", {scalar_state(0.)}), DimensionBoundary("x", {scalar_state(1.)})});
  CHECK_STR_EQ(dup.validate().unwrap_err(), "It is impossible to store two dimensions with the same name.");
}

// ---- PhysicalState algebra / helpers ---------------------------------------------------------------------------
TEST(test_physical_state_algebra) {
  PhysicalState a(PhysicalQuantity::Concentration, PhysicalData::Vector({1., 2., 3.}));
  PhysicalState b(PhysicalQuantity::Concentration, PhysicalData::Vector({10., 20., 30.}));
  auto c = a + b * 0.5;
  const DVector& v = c.get(PhysicalQuantity::Concentration)->as_vector();
  CHECK(v[0] == 6.0 && v[1] == 12.0 && v[2] == 18.0);
  PhysicalState t(PhysicalQuantity::Temperature, PhysicalData::Scalar(300.0));
  auto u = a + t;
  CHECK_EQ(u.len(), 2u);
  CHECK_THROWS(PhysicalData::Vector({1., 2.}) + PhysicalData::Vector({1.}));
  CHECK_THROWS(PhysicalData::Vector({1., 2.}) + PhysicalData::Matrix(DMatrix(2, 1)));
  auto sv = PhysicalData::Scalar(1.0) + PhysicalData::Vector({1., 2.});
  CHECK(sv.is_vector() && sv.as_vector()[1] == 3.0);
  CHECK_STR_EQ(PhysicalQuantity::Custom("Viscosity").to_string(), "Viscosity");
  CHECK(PhysicalState::empty().is_empty() && PhysicalState::empty().memory_bytes() == 0);
}

TEST(test_sample_indices) {  // physics/traits.rs:1010-1023
  CHECK(sample_indices(5, std::nullopt) == (std::vector<std::size_t>{0, 1, 2, 3, 4}));
  CHECK(sample_indices(5, 0) == (std::vector<std::size_t>{0, 1, 2, 3, 4}));
  CHECK(sample_indices(5, 9) == (std::vector<std::size_t>{0, 1, 2, 3, 4}));
  CHECK(sample_indices(5, 1) == (std::vector<std::size_t>{0}));
  CHECK(sample_indices(10, 4) == (std::vector<std::size_t>{0, 3, 6, 9}));
  CHECK(sample_indices(10001, 3) == (std::vector<std::size_t>{0, 5000, 10000}));
}

TEST(test_outlet_data) {  // physics/traits.rs:969-989
  std::vector<PhysicalState> traj;
  traj.emplace_back(PhysicalQuantity::Concentration, PhysicalData::Vector({1., 2., 3.}));
  traj.emplace_back(PhysicalQuantity::Concentration, PhysicalData::Matrix(DMatrix::from_row_slice(2, 2, {1., 2., 3., 4.})));
  traj.emplace_back(PhysicalQuantity::Concentration, PhysicalData::Scalar(5.0));
  CHECK(outlet_data(PhysicalQuantity::Concentration, traj, 0) == (std::vector<double>{3.}));
  CHECK(outlet_data(PhysicalQuantity::Concentration, traj, 1) == (std::vector<double>{3., 4.}));
  CHECK(outlet_data(PhysicalQuantity::Concentration, traj, 2) == (std::vector<double>{5.}));
  CHECK(outlet_data(PhysicalQuantity::Concentration, traj, 3).empty());
  CHECK(outlet_data(PhysicalQuantity::Temperature, traj, 0).empty());
}

TEST(test_rust_float_formatting) {  // metadata values are `f64::to_string()` (rk4.rs:345-346)
  CHECK_STR_EQ(rust_f64_to_string(0.06), "0.06");
  CHECK_STR_EQ(rust_f64_to_string(600.0), "600");
  CHECK_STR_EQ(rust_f64_to_string(1e-7), "0.0000001");
  CHECK_STR_EQ(rust_f64_to_string(1e21), "1000000000000000000000");
  CHECK_STR_EQ(rust_f64_to_string(600.0 / 10000.0), "0.06");
  CHECK_STR_EQ(rust_f64_to_string(0.1 + 0.2), "0.30000000000000004");
  CHECK_STR_EQ(rust_f64_debug(5.0), "5.0");
}

// ---- solve(): the error paths that never reach the device (rk4.rs:213-245) ----------------------------------------
struct ExponentialDecay : PhysicalModel {  // the reference's mock model (rk4.rs:376-400): unknown to the device code
  std::size_t points() const override { return 1; }
  PhysicalState setup_initial_state() const override { return PhysicalState(PhysicalQuantity::Concentration, PhysicalData::Vector({1.0})); }
  const char* name() const override { return "Exponential Decay"; }
};

static Scenario tfa_scenario() {
  auto model = std::make_unique<LangmuirSingle>(tfa());
  auto initial = model->setup_initial_state();
  return Scenario(std::move(model), DomainBoundaries::temporal(std::move(initial)));
}

TEST(test_solver_names) {
  CHECK_STR_EQ(RK4Solver().name(), "Runge Kutta (RK4)");
  CHECK_STR_EQ(EulerSolver().name(), "Forward Euler");
}

TEST(test_solve_rejects_invalid_configuration) {
  auto sc = tfa_scenario();
  CHECK_STR_EQ(RK4Solver().solve(sc, SolverConfiguration::time_evolution(-1.0, 10)).unwrap_err(), "Total time must be positive");
  CHECK_STR_EQ(EulerSolver().solve(sc, SolverConfiguration::time_evolution(1.0, 0)).unwrap_err(), "TimeSteps must be greater than 0");
  // the RK4 solver reports the wrong-configuration error with the Euler solver's text (rk4.rs:227-230)
  CHECK_STR_EQ(RK4Solver().solve(sc, SolverConfiguration(SolverType::Other("Iterative"))).unwrap_err(),
               "EulerSolver only supports TimeEvolution configuration, got Iterative");
  CHECK_STR_EQ(EulerSolver().solve(sc, SolverConfiguration(SolverType::Other("Analytical"))).unwrap_err(),
               "EulerSolver only supports TimeEvolution configuration, got Analytical");
}

TEST(test_solve_rejects_invalid_scenarios) {
  auto cfg = SolverConfiguration::time_evolution(1.0, 10);
  Scenario empty(std::make_unique<LangmuirSingle>(tfa()), DomainBoundaries());
  CHECK_STR_EQ(RK4Solver().solve(empty, cfg).unwrap_err(), "Dimension boundaries cannot be empty.");
  Scenario no_ic(std::make_unique<LangmuirSingle>(tfa()),
                 DomainBoundaries::create({DimensionBoundary("x", {scalar_state(0.), scalar_state(1.)})},
                                          TimeAxisConvention::none()));
  CHECK_STR_EQ(RK4Solver().solve(no_ic, cfg).unwrap_err(), "No initial condition found in domain boundaries");
  Scenario mock(std::make_unique<ExponentialDecay>(), DomainBoundaries::temporal(scalar_state(1.0)));
  auto e = RK4Solver().solve(mock, cfg);
  CHECK(e.is_err() && e.unwrap_err().find("not supported by the CUDA backend") != std::string::npos);
  Scenario wrong_len(std::make_unique<LangmuirSingle>(tfa()),
                     DomainBoundaries::temporal(PhysicalState(PhysicalQuantity::Concentration, PhysicalData::Vector(DVector(7, 0.0)))));
  CHECK_STR_EQ(RK4Solver().solve(wrong_len, cfg).unwrap_err(), "Concentration profile size 7 vs points discretization 100");
  CHECK_STR_EQ(tfa_scenario().get_model_name(), "Langmuir single specie with temporal injection");
  CHECK(tfa_scenario().debug_string().find("Langmuir single specie") != std::string::npos);
}

TEST(test_validate_state_messages) {  // solver/mod.rs:527-548
  CHECK_STR_EQ(validate_state_message(42),
               "NaN detected in Concentration at step 42. This indicates numerical instability. Try reducing time step "
               "(increase time_steps parameter).");
  CHECK_STR_EQ(validate_state_message(-7),
               "Infinity detected in Concentration at step 7. This indicates numerical overflow. Try reducing time step or "
               "check physics model for division by zero.");
  CHECK_STR_EQ(validate_state_message(0), "");
}

// ---- analytics (validation/dissimilarity.rs tests) ------------------------------------------------------------------
static std::vector<double> linspace(double a, double b, std::size_t n) {
  std::vector<double> r(n);
  for (std::size_t i = 0; i < n; ++i) r[i] = a + (b - a) * (double)i / (double)(n - 1);
  return r;
}
static std::vector<double> gaussian_signal(const std::vector<double>& t, double center, double sigma) {
  std::vector<double> r(t.size());
  for (std::size_t i = 0; i < t.size(); ++i) {
    const double x = (t[i] - center) / sigma;
    r[i] = std::exp(-0.5 * x * x) / (sigma * std::sqrt(2.0 * M_PI));
  }
  return r;
}

TEST(test_trapezoid_normalize_rsf) {
  auto t = linspace(0.0, 100.0, 2000);
  auto c = gaussian_signal(t, 50.0, 5.0);
  CHECK_NEAR(trapezoid(t, c), 1.0, 1e-6);
  auto y = normalize(t, c);
  CHECK_NEAR(trapezoid(t, y), 1.0, 1e-12);
  CHECK_NEAR(rsf(t, c, t, c), 0.0, 1e-12);                       // identical signals
  auto far = gaussian_signal(t, 20.0, 2.0);
  auto c2 = gaussian_signal(t, 80.0, 2.0);
  CHECK_NEAR(rsf(t, far, t, c2), 2.0, 1e-3);                     // disjoint signals
  std::vector<double> scaled(c);
  for (double& v : scaled) v *= 7.5;
  CHECK_NEAR(rsf(t, c, t, scaled), 0.0, 1e-12);                  // scale invariance
  auto t2 = linspace(0.0, 100.0, 777);                           // different grids
  CHECK(rsf(t, c, t2, gaussian_signal(t2, 50.0, 5.0)) < 1e-3);
  CHECK_NEAR(first_moment(t, c), 50.0, 1e-6);
  CHECK_NEAR(peak_sigma(t, c), 5.0, 1e-4);
  CHECK_THROWS(normalize(t, std::vector<double>(t.size(), 0.0)));
}

int main(int argc, char** argv) { return minitest::run(argc, argv); }
