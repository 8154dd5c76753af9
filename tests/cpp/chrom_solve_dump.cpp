// Parity probe for the C++ host layer: runs one solve through `Solver::solve` and dumps
// time_points, the outlet series, every 50th trajectory state and the final state as raw f64, so
// tests/test_gpu_cpp_host.py can compare them bit for bit with the CPU oracle.
//   usage: chrom_solve_dump single|multi rk4|euler exact|fast <total_time> <steps> <ramp_ic 0|1> <out.bin>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "chrom.hpp"

using namespace chrom;

static void put(std::FILE* f, const std::vector<double>& v) { std::fwrite(v.data(), sizeof(double), v.size(), f); }

int main(int argc, char** argv) {
  if (argc != 8) {
    std::fprintf(stderr, "usage: %s single|multi rk4|euler exact|fast <total_time> <steps> <ramp_ic> <out.bin>\n", argv[0]);
    return 2;
  }
  const bool multi = !std::strcmp(argv[1], "multi");
  const bool rk4 = !std::strcmp(argv[2], "rk4");
  const Arithmetic arith = !std::strcmp(argv[3], "exact") ? Arithmetic::Exact : Arithmetic::Fast;
  const double total_time = std::atof(argv[4]);
  const std::size_t steps = (std::size_t)std::atoll(argv[5]);
  const bool ramp = std::atoi(argv[6]) != 0;
  const std::size_t nz = 100;

  std::unique_ptr<PhysicalModel> model;
  PhysicalState initial;
  if (!multi) {  // BASELINE config 1: examples/tfa.rs:44-64
    auto m = std::make_unique<LangmuirSingle>(1.2, 0.4, 2.0, 0.4, 0.001, 0.25, nz, TemporalInjection::gaussian(10.0, 3.0, 0.1));
    initial = m->setup_initial_state();
    if (ramp) {
      DVector y(nz);
      for (std::size_t i = 0; i < nz; ++i) y[i] = 0.01 * (double)(i % 7) / 7.0;
      initial = PhysicalState(PhysicalQuantity::Concentration, PhysicalData::Vector(y));
    }
    model = std::move(m);
  } else {  // BASELINE config 2: examples/acids_multi.rs:130-174
    std::vector<SpeciesParams> sp;
    sp.emplace_back("Ascorbic", 1.0, 1.1, 2, TemporalInjection::gaussian(10.0, 3.0, 0.1));
    sp.emplace_back("Erythorbic", 1.0, 1.7, 2, TemporalInjection::gaussian(10.0, 3.0, 0.1));
    auto m = std::make_unique<LangmuirMulti>(LangmuirMulti::create(std::move(sp), nz, 0.4, 0.001, 0.25).unwrap());
    initial = m->setup_initial_state();
    if (ramp) {
      DMatrix y(nz, 2);
      for (std::size_t i = 0; i < nz; ++i) {
        y(i, 0) = 0.01 * (double)(i % 7) / 7.0;
        y(i, 1) = 0.02 * (double)(i % 5) / 5.0;
      }
      initial = PhysicalState(PhysicalQuantity::Concentration, PhysicalData::Matrix(y));
    }
    model = std::move(m);
  }
  Scenario scenario(std::move(model), DomainBoundaries::temporal(std::move(initial)));
  auto config = SolverConfiguration::time_evolution(total_time, steps);
  Result<SimulationResult> res = Err("unset");
  if (rk4) {
    RK4Solver s;
    s.set_arithmetic(arith);
    res = s.solve(scenario, config);
  } else {
    EulerSolver s;
    s.set_arithmetic(arith);
    res = s.solve(scenario, config);
  }
  if (res.is_err()) {
    std::fprintf(stderr, "solve failed: %s\n", res.unwrap_err().c_str());
    return 1;
  }
  const SimulationResult& r = res.unwrap();
  std::FILE* f = std::fopen(argv[7], "wb");
  if (!f) return 3;
  put(f, r.time_points);                               // steps + 1
  for (const auto& o : r.outlet) put(f, o);            // n_species x (steps + 1)
  for (std::size_t s = 0; s < r.state_trajectory.size(); s += 50) {  // nz x n (column-major) each
    const PhysicalData* d = r.state_trajectory[s].get(PhysicalQuantity::Concentration);
    std::fwrite(d->values(), sizeof(double), d->n_values(), f);
  }
  const PhysicalData* d = r.final_state.get(PhysicalQuantity::Concentration);
  std::fwrite(d->values(), sizeof(double), d->n_values(), f);
  std::fclose(f);
  std::printf("solver=%s steps=%s dt=%s\n", r.metadata.at("solver").c_str(), r.metadata.at("time steps").c_str(),
              r.metadata.at("dt").c_str());
  return 0;
}
