// GPU tests of the C++ host layer through the reference's own call shape
// (`RK4Solver::new().solve(&scenario, &config)`), following the reference's validation suite and
// solver tests:
//   validation/main.rs:292-528, validation/reference.rs:34-356   TFA linear/non-linear, Euler-vs-RK4 Rsf,
//                                                                ascorbic/erythorbic, glucose/fructose
//   src/solver/methods/rk4.rs:588-717, euler.rs:474-600          result structure, metadata, time points
//   src/solver/mod.rs:514-553                                    validate_state error texts
#include <cmath>

#include "chrom.hpp"
#include "minitest.hpp"

using namespace chrom;

// ---- validation/reference.rs constants ----------------------------------------------------------------
static const double COLUMN_LENGTH = 0.3, POROSITY = 0.4, VELOCITY = 1.0e-3, LAMBDA = 1.0, KI = 0.5, PORT_NUMBER = 6.0;
static const std::size_t N_POINTS = 100, N_STEPS = 4500;
static const double T_TOTAL = 900.0;
static const double T_INJ = 2.0 * T_TOTAL / (double)N_STEPS;

struct Signal {
  std::vector<double> t, c;
};

static std::vector<double> outlet_profile(const SimulationResult& result, std::size_t n_points) {
  std::vector<double> out;
  for (const auto& state : result.state_trajectory) {
    const PhysicalData* d = state.get(PhysicalQuantity::Concentration);
    if (!d) continue;
    if (d->is_vector()) out.push_back(d->as_vector()[n_points - 1]);
    else if (d->is_matrix()) out.push_back(d->as_matrix()(n_points - 1, 0));
    else out.push_back(0.0);
  }
  return out;
}

template <class S>
static Signal run_single(double c0) {
  auto injection = TemporalInjection::rectangle(0.0, T_INJ, c0);
  auto model = std::make_unique<LangmuirSingle>(LAMBDA, KI, PORT_NUMBER, POROSITY, VELOCITY, COLUMN_LENGTH, N_POINTS, injection);
  const std::size_t n_pts = model->points();
  auto boundaries = DomainBoundaries::temporal(model->setup_initial_state());
  Scenario scenario(std::move(model), std::move(boundaries));
  auto config = SolverConfiguration::time_evolution(T_TOTAL, N_STEPS);
  auto result = S().solve(scenario, config).expect("solver failed");
  return {result.time_points, outlet_profile(result, n_pts)};
}
static Signal run_rk4(double c0) { return run_single<RK4Solver>(c0); }
static Signal run_euler(double c0) { return run_single<EulerSolver>(c0); }

static std::size_t argmax(const std::vector<double>& v) {
  std::size_t k = 0;
  for (std::size_t i = 1; i < v.size(); ++i)
    if (v[i] >= v[k]) k = i;  // Iterator::max_by returns the LAST maximum
  return k;
}

// ---- Case A: linear regime --------------------------------------------------------------------------------
TEST(linear_tfa_retention_time) {
  auto s = run_rk4(1e-3);
  const double t_r = first_moment(s.t, s.c);
  CHECK_MSG(std::fabs(t_r - 624.0) / 624.0 * 100.0 < 1.0, "tR = " + std::to_string(t_r));
}

TEST(linear_tfa_peak_width) {
  auto s = run_rk4(1e-3);
  const double sigma = peak_sigma(s.t, s.c);
  CHECK_MSG(std::fabs(sigma - 61.39) / 61.39 * 100.0 < 10.0, "sigma = " + std::to_string(sigma));
}

TEST(linear_tfa_mass_conservation) {
  auto s = run_rk4(1e-3);
  const double recovery = trapezoid(s.t, s.c) / (1e-3 * T_INJ);
  CHECK_MSG(recovery >= 0.90, "recovery = " + std::to_string(recovery));
}

// ---- Case B: non-linear regime ------------------------------------------------------------------------------
TEST(nonlinear_tfa_peak_compression) {
  auto nl = run_rk4(1.0), lin = run_rk4(1e-3);
  CHECK(nl.t[argmax(nl.c)] < lin.t[argmax(lin.c)]);
}

TEST(nonlinear_tfa_mass_conservation) {
  auto s = run_rk4(1.0);
  const double recovery = trapezoid(s.t, s.c) / (1.0 * T_INJ);
  CHECK_MSG(recovery >= 0.85, "recovery = " + std::to_string(recovery));
}

// ---- Case C: Euler vs RK4 -------------------------------------------------------------------------------------
TEST(euler_vs_rk4_rsf) {
  auto e = run_euler(1e-3), r = run_rk4(1e-3);
  const double criterion = rsf(e.t, e.c, r.t, r.c);
  CHECK_MSG(criterion < 0.05, "Rsf = " + std::to_string(criterion));
  CHECK_MSG(std::fabs(criterion - 0.016) < 0.004, "published Rsf = 0.016 (CHANGELOG 0.4.0), got " + std::to_string(criterion));
}

// ---- Cases D / E: multi-species ---------------------------------------------------------------------------------
struct SpeciesReference {
  const char* name;
  double lambda, langmuir_k;
  std::uint32_t port_number;
  double t_retention;
};
struct MultiSpeciesCase {
  double column_length, porosity, velocity;
  std::size_t n_points;
  double c0, t_total;
  std::size_t n_steps;
  std::vector<SpeciesReference> species;
  double t_inj() const { return 2.0 * t_total / (double)n_steps; }
};
static MultiSpeciesCase ascorbic_erythorbic() {
  return {0.25, 0.4, 1.0e-3, 100, 1.0e-3, 800.0, 4000, {{"Ascorbic", 1.0, 1.1, 2, 448.0}, {"Erythorbic", 1.0, 1.7, 2, 556.0}}};
}
static MultiSpeciesCase glucose_fructose_linear() {
  return {COLUMN_LENGTH, POROSITY, VELOCITY, N_POINTS, 1.0e-3, 400.0, 2000,
          {{"Glucose", 0.0, 0.27 / (1.0 - POROSITY), 1, 168.6}, {"Fructose", 0.0, 0.46 / (1.0 - POROSITY), 1, 202.8}}};
}

static std::pair<std::vector<double>, std::vector<std::vector<double>>> run_multi_rk4(const MultiSpeciesCase& c) {
  std::vector<SpeciesParams> species;
  for (const auto& sp : c.species)
    species.emplace_back(sp.name, sp.lambda, sp.langmuir_k, sp.port_number, TemporalInjection::rectangle(0.0, c.t_inj(), c.c0));
  const std::size_t n_species = species.size();
  auto model = std::make_unique<LangmuirMulti>(
      LangmuirMulti::create(std::move(species), c.n_points, c.porosity, c.velocity, c.column_length)
          .expect("invalid multi-species reference case parameters"));
  auto boundaries = DomainBoundaries::temporal(model->setup_initial_state());
  Scenario scenario(std::move(model), std::move(boundaries));
  auto result = RK4Solver().solve(scenario, SolverConfiguration::time_evolution(c.t_total, c.n_steps)).expect("RK4 solver failed");
  std::vector<std::vector<double>> per_species(n_species);
  for (const auto& state : result.state_trajectory) {
    const DMatrix& m = state.get(PhysicalQuantity::Concentration)->as_matrix();
    for (std::size_t s = 0; s < n_species; ++s) per_species[s].push_back(m(c.n_points - 1, s));
  }
  return {result.time_points, per_species};
}

TEST(ascorbic_erythorbic_retention_times) {
  auto c = ascorbic_erythorbic();
  auto r = run_multi_rk4(c);
  std::vector<double> t_r;
  for (std::size_t i = 0; i < c.species.size(); ++i) {
    t_r.push_back(first_moment(r.first, r.second[i]));
    CHECK_MSG(std::fabs(t_r[i] - c.species[i].t_retention) / c.species[i].t_retention * 100.0 < 1.0,
              std::string(c.species[i].name) + " tR = " + std::to_string(t_r[i]));
  }
  const double gap_ana = c.species[1].t_retention - c.species[0].t_retention;  // ascorbic_erythorbic_retention_gap
  CHECK(std::fabs((t_r[1] - t_r[0]) - gap_ana) / gap_ana * 100.0 < 1.0);
}

TEST(glucose_fructose_linear_retention_times) {
  auto c = glucose_fructose_linear();
  auto r = run_multi_rk4(c);
  for (std::size_t i = 0; i < c.species.size(); ++i) {
    const double t_r = first_moment(r.first, r.second[i]);
    CHECK_MSG(std::fabs(t_r - c.species[i].t_retention) / c.species[i].t_retention * 100.0 < 1.0,
              std::string(c.species[i].name) + " tR = " + std::to_string(t_r));
  }
}

// ---- result structure (rk4.rs:588-717, euler.rs:474-600) ----------------------------------------------------------
static Scenario tfa_scenario(TemporalInjection inj, double k = 0.4) {
  auto model = std::make_unique<LangmuirSingle>(1.2, k, 2.0, 0.4, 0.001, 0.25, 100, std::move(inj));
  auto initial = model->setup_initial_state();
  return Scenario(std::move(model), DomainBoundaries::temporal(std::move(initial)));
}

TEST(test_rk4_result_structure_and_metadata) {
  auto sc = tfa_scenario(TemporalInjection::gaussian(10.0, 3.0, 0.1));
  auto r = RK4Solver().solve(sc, SolverConfiguration::time_evolution(60.0, 1000)).unwrap();
  CHECK_EQ(r.len(), 1001u);
  CHECK_EQ(r.state_trajectory.size(), 1001u);
  CHECK_EQ(r.time_points[0], 0.0);
  CHECK_EQ(r.time_points[1000], 60.0);
  const double dt = 60.0 / 1000.0;
  for (std::size_t s = 0; s < 1000; ++s) CHECK_EQ(r.time_points[s + 1], ((double)s + 1.0) * dt);
  CHECK_STR_EQ(r.metadata.at("solver"), "Runge-Kutta 4");
  CHECK_STR_EQ(r.metadata.at("time steps"), "1000");
  CHECK_STR_EQ(r.metadata.at("dt"), "0.06");
  CHECK_STR_EQ(r.metadata.at("total time"), "60");
  CHECK_STR_EQ(r.metadata.at("function evaluations"), "4000");
  // the last trajectory entry is the final state, the first is the initial condition
  CHECK(r.state_trajectory.back().get(PhysicalQuantity::Concentration)->as_vector() ==
        r.final_state.get(PhysicalQuantity::Concentration)->as_vector());
  for (double v : r.state_trajectory.front().get(PhysicalQuantity::Concentration)->as_vector()) CHECK_EQ(v, 0.0);
  // the outlet series the device samples equals the last cell of every trajectory state
  CHECK(r.outlet.size() == 1 && r.outlet[0] == outlet_profile(r, 100));
  auto e = EulerSolver().solve(sc, SolverConfiguration::time_evolution(60.0, 1000)).unwrap();
  CHECK_STR_EQ(e.metadata.at("solver"), "Forward Euler");
  CHECK(e.metadata.count("function evaluations") == 0 && e.metadata.size() == 4);
}

TEST(test_custom_injection_equals_builtin) {
  // A closure is tabulated on the host at the solver's stage times: the same profile written as a
  // closure must give the same bits as the built-in Gaussian.
  auto g = TemporalInjection::gaussian(10.0, 3.0, 0.1);
  auto custom = TemporalInjection::custom([](double t) {
    const double distance = (t - 10.0) / 3.0;
    return 0.1 * std::exp(-distance * distance / 2.0);
  });
  RK4Solver solver;
  solver.set_arithmetic(Arithmetic::Exact);
  auto cfg = SolverConfiguration::time_evolution(60.0, 500);
  auto a = solver.solve(tfa_scenario(g), cfg).unwrap();
  auto b = solver.solve(tfa_scenario(custom), cfg).unwrap();
  CHECK(a.final_state.get(PhysicalQuantity::Concentration)->as_vector() ==
        b.final_state.get(PhysicalQuantity::Concentration)->as_vector());
  CHECK(a.outlet[0] == b.outlet[0]);
}

TEST(test_unstable_step_reports_validate_state_error) {
  // CFL >> 1: the reference's validate_state aborts with its NaN / Infinity text (solver/mod.rs:527-548)
  auto sc = tfa_scenario(TemporalInjection::rectangle(0.0, 1e9, 1.0));
  auto r = RK4Solver().solve(sc, SolverConfiguration::time_evolution(60000.0, 600));
  CHECK(r.is_err());
  const std::string& e = r.unwrap_err();
  CHECK_MSG(e.find("detected in Concentration at step ") != std::string::npos &&
                (e.rfind("NaN", 0) == 0 || e.rfind("Infinity", 0) == 0), e);
}

TEST(test_solve_batch_equals_individual_solves) {
  // a sweep in one device batch == the reference's sequential loop over solve()
  std::vector<Scenario> scs;
  for (double k : {0.1, 0.4, 0.8}) scs.push_back(tfa_scenario(TemporalInjection::gaussian(10.0, 3.0, 0.1), k));
  scs.push_back(Scenario(std::make_unique<LangmuirSingle>(1.2, 0.4, 2.0, 0.4, 0.001, 0.25, 100, TemporalInjection::none()),
                         DomainBoundaries()));  // invalid: its own Err, the others still run
  std::vector<const Scenario*> ptrs;
  for (auto& s : scs) ptrs.push_back(&s);
  RK4Solver solver;
  auto cfg = SolverConfiguration::time_evolution(300.0, 2000);
  BatchOptions opt;
  opt.snapshot_stride = 0;
  auto batch = solver.solve_batch(ptrs, cfg, opt).unwrap();
  CHECK_EQ(batch.size(), 4u);
  CHECK_STR_EQ(batch[3].unwrap_err(), "Dimension boundaries cannot be empty.");
  for (std::size_t i = 0; i < 3; ++i) {
    auto one = solver.solve(scs[i], cfg).unwrap();
    const auto& b = batch[i].unwrap();
    CHECK(b.state_trajectory.empty() && b.len() == 2001);
    CHECK(b.final_state.get(PhysicalQuantity::Concentration)->as_vector() ==
          one.final_state.get(PhysicalQuantity::Concentration)->as_vector());
    CHECK(b.outlet[0] == one.outlet[0]);
    // on-device summary vs the host analytics on the returned outlet series
    const double area = trapezoid(one.time_points, one.outlet[0]);
    CHECK(std::fabs(b.summary[0][0] - area) <= 1e-7 * std::fabs(area));
    const double t_r = first_moment(one.time_points, one.outlet[0]);  // summary[1] is the normalised first moment
    CHECK(std::fabs(b.summary[0][1] - t_r) <= 1e-7 * t_r);
  }
}

TEST(test_strided_trajectory) {
  auto sc = tfa_scenario(TemporalInjection::gaussian(10.0, 3.0, 0.1));
  RK4Solver full, strided;
  BatchOptions o;
  o.snapshot_stride = 100;
  o.outlet_stride = 10;
  strided.set_default_options(o);
  auto cfg = SolverConfiguration::time_evolution(60.0, 1000);
  auto a = full.solve(sc, cfg).unwrap();
  auto b = strided.solve(sc, cfg).unwrap();
  CHECK_EQ(b.state_trajectory.size(), 11u);
  CHECK_EQ(b.outlet[0].size(), 101u);
  for (std::size_t s = 0; s < 11; ++s) {
    CHECK_EQ(b.trajectory_steps[s], s * 100);
    CHECK(b.state_trajectory[s].get(PhysicalQuantity::Concentration)->as_vector() ==
          a.state_trajectory[s * 100].get(PhysicalQuantity::Concentration)->as_vector());
  }
  for (std::size_t s = 0; s < 101; ++s) {
    CHECK_EQ(b.outlet[0][s], a.outlet[0][s * 10]);
    CHECK_EQ(b.outlet_time_points[s], a.time_points[s * 10]);
  }
}

TEST(test_solve_scan_equals_individual_solves) {
  // examples/cfl_stability_scan.rs:378-394 runs one solve per CFL target (a different step count each) in a
  // sequential loop; spatial_temporal_convergence.rs:261-277 also varies the grid.  solve_scan runs the groups
  // concurrently; every entry must equal its own solve(), bit for bit, whatever group / context it ran on.
  std::vector<Scenario> scs;
  std::vector<SolverConfiguration> cfgs;
  for (std::size_t steps : {400u, 800u, 1200u, 2000u, 800u, 400u}) {
    scs.push_back(tfa_scenario(TemporalInjection::gaussian(10.0, 3.0, 0.1), 0.2 + 1e-4 * (double)steps));
    cfgs.push_back(SolverConfiguration::time_evolution(120.0, steps));
  }
  {  // another grid (nz = 60) and a two-species model in the same scan
    auto fine = std::make_unique<LangmuirSingle>(1.2, 0.4, 2.0, 0.4, 0.001, 0.25, 60, TemporalInjection::gaussian(10.0, 3.0, 0.1));
    auto init = fine->setup_initial_state();
    scs.push_back(Scenario(std::move(fine), DomainBoundaries::temporal(std::move(init))));
    cfgs.push_back(SolverConfiguration::time_evolution(120.0, 800));
    std::vector<SpeciesParams> sp;
    sp.emplace_back("Ascorbic", 1.0, 1.1, 2, TemporalInjection::gaussian(10.0, 3.0, 0.1));
    sp.emplace_back("Erythorbic", 1.0, 1.7, 2, TemporalInjection::gaussian(10.0, 3.0, 0.1));
    auto multi = std::make_unique<LangmuirMulti>(LangmuirMulti::create(std::move(sp), 100, 0.4, 0.001, 0.25).unwrap());
    auto minit = multi->setup_initial_state();
    scs.push_back(Scenario(std::move(multi), DomainBoundaries::temporal(std::move(minit))));
    cfgs.push_back(SolverConfiguration::time_evolution(120.0, 800));
  }
  scs.push_back(Scenario(std::make_unique<LangmuirSingle>(1.2, 0.4, 2.0, 0.4, 0.001, 0.25, 100, TemporalInjection::none()),
                         DomainBoundaries()));  // invalid: its own Err
  cfgs.push_back(SolverConfiguration::time_evolution(120.0, 400));
  std::vector<const Scenario*> ptrs;
  for (auto& s : scs) ptrs.push_back(&s);
  RK4Solver solver;
  BatchOptions opt;
  opt.snapshot_stride = 0;
  auto scan = solver.solve_scan(ptrs, cfgs, opt, 3).unwrap();
  CHECK_EQ(scan.size(), scs.size());
  CHECK_STR_EQ(scan.back().unwrap_err(), "Dimension boundaries cannot be empty.");
  for (std::size_t i = 0; i + 1 < scs.size(); ++i) {
    auto one = solver.solve(scs[i], cfgs[i]).unwrap();
    const auto& b = scan[i].unwrap();
    CHECK_EQ(b.len(), one.len());
    CHECK(b.time_points == one.time_points);
    CHECK(b.outlet == one.outlet);
    const PhysicalData* fa = b.final_state.get(PhysicalQuantity::Concentration);
    const PhysicalData* fb = one.final_state.get(PhysicalQuantity::Concentration);
    if (fa->is_vector()) CHECK(fa->as_vector() == fb->as_vector());
    else CHECK(fa->as_matrix().data == fb->as_matrix().data);
  }
  // mismatched argument lengths are an error of the whole call
  cfgs.pop_back();
  CHECK(solver.solve_scan(ptrs, cfgs, opt).is_err());
}

int main(int argc, char** argv) { return minitest::run(argc, argv); }
