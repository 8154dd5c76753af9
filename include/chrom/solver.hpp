// chrom/solver.hpp — C++ mirror of chrom-rs's solver boundary, backed by libchrom_cuda.so.
//
//   SolverType / SolverConfiguration     src/solver/traits.rs:62-121, 258-430
//   SimulationResult                     src/solver/traits.rs:449-545
//   trait Solver                         src/solver/traits.rs:606-642          <- the drop-in boundary
//   DimensionBoundary / DomainBoundaries src/solver/boundary.rs:43-400
//   TimeAxisConvention                   src/solver/boundary.rs:404-428
//   Scenario                             src/solver/scenario.rs:50-110
//   RK4Solver::solve / EulerSolver::solve  src/solver/methods/rk4.rs:205-350, euler.rs:166-288
//   validate_state (error texts)         src/solver/mod.rs:514-553
//
// CudaRK4Solver / CudaEulerSolver implement `Solver` with the reference's validation order, error
// texts, time points and metadata; the time integration itself runs in the sm_100a kernels behind
// include/chrom_cuda.h.  There is no CPU integrator here and no fallback: without the CUDA library
// or a B200 the constructor of the context fails, and models other than LangmuirSingle /
// LangmuirMulti are an `Err`.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <vector>

#include "chrom/models.hpp"
#include "chrom/physics.hpp"
#include "chrom/result.hpp"

struct chrom_cuda_ctx;  // include/chrom_cuda.h

namespace chrom {

// ---- SolverType ----------------------------------------------------------------------------------------
class SolverType {
 public:
  // The hot path is the time integrators: TimeEvolution is the one configuration RK4Solver / EulerSolver accept
  // (rk4.rs:222-231, euler.rs:183-192).  The reference's other variants (Iterative, Analytical,
  // SpatialDiscretization, Custom: traits.rs:172-215) belong to solvers outside this path; `Other(name)` stands for
  // any of them so that the rejection ("... only supports TimeEvolution configuration, got <name>") stays testable.
  enum Kind { kTimeEvolution, kOther };
  static SolverType TimeEvolution(double total_time, std::size_t time_steps);
  static SolverType Other(std::string name);

  Kind kind() const { return kind_; }
  const std::string& name() const { return name_; }
  Result<Unit> validate() const;  // traits.rs:172-215

  double total_time = 0.0;                 // TimeEvolution
  std::size_t time_steps = 0;              // TimeEvolution
  bool operator==(const SolverType& o) const;

 private:
  Kind kind_ = kTimeEvolution;
  std::string name_ = "TimeEvolution";
};

// ---- SolverConfiguration ----------------------------------------------------------------------------------
struct SolverConfiguration {
  SolverType solver_type;
  std::optional<std::size_t> step;  // sampling hint for exporters (traits.rs:263-279); None on every factory

  explicit SolverConfiguration(SolverType t) : solver_type(std::move(t)) {}
  static SolverConfiguration create(SolverType t) { return SolverConfiguration(std::move(t)); }  // `new`
  static SolverConfiguration time_evolution(double total_time, std::size_t time_steps);
  SolverConfiguration with_step(std::size_t n) const;
  Result<Unit> validate() const { return solver_type.validate(); }
};

// ---- boundaries ----------------------------------------------------------------------------------------------
struct TimeAxisConvention {
  enum Kind { None, First, Last, Index } kind = Last;
  std::size_t index = 0;
  static TimeAxisConvention none() { return {None, 0}; }
  static TimeAxisConvention first() { return {First, 0}; }
  static TimeAxisConvention last() { return {Last, 0}; }
  static TimeAxisConvention at(std::size_t i) { return {Index, i}; }
  std::string to_string() const;  // "None" | "First" | "Last" | "Index (i)"
  bool operator==(const TimeAxisConvention& o) const { return kind == o.kind && (kind != Index || index == o.index); }
  bool operator!=(const TimeAxisConvention& o) const { return !(*this == o); }
};

struct DimensionBoundary {
  std::string name;
  std::vector<PhysicalState> states;
  DimensionBoundary(std::string name, std::vector<PhysicalState> states)
      : name(std::move(name)), states(std::move(states)) {}
  const PhysicalState* first() const { return &states.at(0); }  // out of range on empty, like the reference's panic
  const PhysicalState* last() const { return &states.at(states.size() - 1); }
  std::size_t size() const { return states.size(); }
  bool is_empty() const { return states.empty(); }
  Result<Unit> validate() const;
};

struct DomainBoundaries {
  std::vector<DimensionBoundary> dimensions;
  TimeAxisConvention convention = TimeAxisConvention::none();  // Default::default(): empty, None

  DomainBoundaries() = default;
  explicit DomainBoundaries(std::vector<DimensionBoundary> dims)  // `new`: convention Last
      : dimensions(std::move(dims)), convention(TimeAxisConvention::last()) {}
  static DomainBoundaries create(std::vector<DimensionBoundary> dims, std::optional<TimeAxisConvention> convention);
  static DomainBoundaries temporal(PhysicalState initial);

  std::size_t ndim() const { return dimensions.size(); }
  std::size_t sdim() const;
  bool is_time_dependent() const { return convention.kind != TimeAxisConvention::None; }
  std::optional<std::size_t> time_index() const;
  const DimensionBoundary* time_boundary() const;
  const PhysicalState* initial_condition() const;
  std::vector<const DimensionBoundary*> spatial_boundaries() const;
  const DimensionBoundary* get_boundary(const std::string& name) const;
  Result<Unit> validate() const;
};

// ---- Scenario ----------------------------------------------------------------------------------------------------
struct Scenario {
  std::unique_ptr<PhysicalModel> model;
  DomainBoundaries conditions;

  Scenario(std::unique_ptr<PhysicalModel> model, DomainBoundaries conditions)
      : model(std::move(model)), conditions(std::move(conditions)) {}
  Result<Unit> validate() const { return conditions.validate(); }
  const char* get_model_name() const { return model->name(); }
  std::size_t ndim() const { return conditions.ndim(); }
  std::size_t sdim() const { return conditions.sdim(); }
  bool is_time_dependent() const { return conditions.is_time_dependent(); }
  std::string debug_string() const;
};

// ---- SimulationResult ----------------------------------------------------------------------------------------------
struct SimulationResult {
  std::vector<double> time_points;              // steps + 1 values, time_points[s + 1] = (s + 1.0) * dt
  std::vector<PhysicalState> state_trajectory;  // the reference keeps every step (rk4.rs:315); see TrajectoryPolicy
  PhysicalState final_state;
  std::map<std::string, std::string> metadata;

  // Extras the device produces anyway (not in the reference's struct; empty unless requested):
  std::vector<std::vector<double>> outlet;   // [species][n_out]: last cell at steps 0, k, 2k, ...
  std::vector<double> outlet_time_points;    // the time of each outlet sample
  std::vector<std::vector<double>> summary;  // [species][CHROM_SUMMARY_LEN]: trapezoid area, first moment, max, time of max, sigma, min
  std::vector<std::size_t> trajectory_steps; // the step index of each state in state_trajectory

  SimulationResult() = default;
  SimulationResult(std::vector<double> time_points, std::vector<PhysicalState> state_trajectory,
                   PhysicalState final_state)
      : time_points(std::move(time_points)), state_trajectory(std::move(state_trajectory)),
        final_state(std::move(final_state)) {}
  void add_metadata(std::string key, std::string value) { metadata[std::move(key)] = std::move(value); }
  std::size_t len() const { return time_points.size(); }
  bool is_empty() const { return time_points.empty(); }
};

// ---- Solver ----------------------------------------------------------------------------------------------------------
class Solver {
 public:
  virtual ~Solver() = default;
  virtual Result<SimulationResult> solve(const Scenario& scenario, const SolverConfiguration& config) const = 0;
  virtual const char* name() const = 0;
};

// The reference's validate_state error strings (solver/mod.rs:527-548) from a device status word:
// +s = NaN first present after step s, -s = Inf (no NaN) after step s, 0 = none.
std::string validate_state_message(std::int64_t status);

// ---- CUDA backend -----------------------------------------------------------------------------------------------------
enum class Arithmetic : int {
  Fast = 0,            // CHROM_ARITH_FAST: FMA, one reciprocal per cell-stage; <= 1e-9 relative to the reference
  Exact = 1,           // CHROM_ARITH_EXACT: the reference's IEEE operations in the reference's order
  FastStructured = 3,  // CHROM_ARITH_FAST_STRUCTURED (LangmuirMulti: the explicit name of Fast)
  FastDense = 4,       // CHROM_ARITH_FAST_DENSE (LangmuirMulti n <= 8: dense register LU with partial pivoting)
};

// One libchrom_cuda context (devices + streams).  Shared between solvers; calls are serialised.
class CudaContext {
 public:
  // device_ids empty: every visible device.  Err when the library finds no usable GPU (no CPU fallback).
  static Result<std::shared_ptr<CudaContext>> create(const std::vector<int>& device_ids = {});
  static Result<std::shared_ptr<CudaContext>> global();  // process-wide default, created on first use
  ~CudaContext();
  CudaContext(const CudaContext&) = delete;
  CudaContext& operator=(const CudaContext&) = delete;
  chrom_cuda_ctx* handle() const { return h_; }
  struct Impl;
  Impl* impl() const { return impl_.get(); }

 private:
  CudaContext() = default;
  chrom_cuda_ctx* h_ = nullptr;
  std::unique_ptr<Impl> impl_;
};

// What a batch call returns per scenario besides the reference's fields.
struct BatchOptions {
  std::size_t snapshot_stride = 1;  // 1 = the reference's full state_trajectory; k = every k-th step; 0 = none
  std::size_t outlet_stride = 1;    // outlet series sampling; 0 = none
  bool summary = true;              // on-device area / first moment / max / time of max / sigma / min of the outlet
};

class CudaSolverBase : public Solver {
 public:
  Result<SimulationResult> solve(const Scenario& scenario, const SolverConfiguration& config) const override;
  // `solve` over independent scenarios of one model type and grid in ONE device batch (the reference
  // runs sweeps as sequential loops: examples/cfl_stability_scan.rs:378-394).  Entry i is what
  // solve(scenarios[i]) would have returned; the outer Err is a failure of the whole batch.
  Result<std::vector<Result<SimulationResult>>> solve_batch(const std::vector<const Scenario*>& scenarios,
                                                            const SolverConfiguration& config,
                                                            const BatchOptions& options = BatchOptions()) const;
  // Heterogeneous sweeps — a different time grid (CFL / stiffness scans: examples/cfl_stability_scan.rs:378-394,
  // stiffness_convergence.rs:360), spatial grid (spatial_temporal_convergence.rs:261-277) or model type per
  // scenario.  Entries are grouped by (model type, n_points, species count, configuration); every group is ONE
  // device batch, and up to `max_concurrent` groups run at the same time, each on its own libchrom_cuda context
  // and host thread (distinct contexts are thread-safe), so that small groups share the GPU instead of
  // queueing.  configs.size() must equal scenarios.size().  Entry i is what solve(scenarios[i], configs[i])
  // would have returned.
  Result<std::vector<Result<SimulationResult>>> solve_scan(const std::vector<const Scenario*>& scenarios,
                                                           const std::vector<SolverConfiguration>& configs,
                                                           const BatchOptions& options = BatchOptions(),
                                                           std::size_t max_concurrent = 4) const;
  void set_arithmetic(Arithmetic a) { arith_ = a; }
  Arithmetic arithmetic() const { return arith_; }
  void set_context(std::shared_ptr<CudaContext> ctx) { ctx_ = std::move(ctx); }
  // Default options used by solve(): the reference's behaviour (every state kept).
  void set_default_options(const BatchOptions& o) { defaults_ = o; }

 protected:
  CudaSolverBase(int method, const char* meta_solver) : method_(method), meta_solver_(meta_solver) {}

 private:
  // solve_batch on an explicit context (solve_scan's workers); nullptr = this solver's / the global context
  Result<std::vector<Result<SimulationResult>>> solve_batch_on(std::shared_ptr<CudaContext> on,
                                                               const std::vector<const Scenario*>& scenarios,
                                                               const SolverConfiguration& config,
                                                               const BatchOptions& options) const;
  int method_;
  const char* meta_solver_;
  Arithmetic arith_ = Arithmetic::Fast;
  BatchOptions defaults_;
  mutable std::shared_ptr<CudaContext> ctx_;
};

class CudaRK4Solver : public CudaSolverBase {
 public:
  CudaRK4Solver() : CudaSolverBase(/*CHROM_RK4=*/1, "Runge-Kutta 4") {}            // rk4.rs:343
  const char* name() const override { return "Runge Kutta (RK4)"; }               // rk4.rs:352-354
};

class CudaEulerSolver : public CudaSolverBase {
 public:
  CudaEulerSolver() : CudaSolverBase(/*CHROM_EULER=*/0, "Forward Euler") {}       // euler.rs:282
  const char* name() const override { return "Forward Euler"; }                   // euler.rs:290-292
};

// In a host that links this backend these ARE the solvers: same names as the reference's types.
using RK4Solver = CudaRK4Solver;
using EulerSolver = CudaEulerSolver;

// ---- chromatogram analytics (validation/dissimilarity.rs, validation/main.rs:190-206) ------------------------------------
double trapezoid(const std::vector<double>& times, const std::vector<double>& values);
std::vector<double> normalize(const std::vector<double>& times, const std::vector<double>& concentrations);
double rsf(const std::vector<double>& t1, const std::vector<double>& c1, const std::vector<double>& t2,
           const std::vector<double>& c2);
double first_moment(const std::vector<double>& times, const std::vector<double>& concs);
double peak_sigma(const std::vector<double>& times, const std::vector<double>& concs);

}  // namespace chrom
