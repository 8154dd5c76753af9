/* chrom_cuda.h — C ABI of libchrom_cuda.so, the B200 (sm_100a) backend for the method-of-lines
 * time integrator of chrom-rs.  This is the drop-in boundary: everything a Rust `impl Solver`
 * needs, as `extern "C"` functions over plain pointers and sizes.  No torch / C++ types.
 *
 * What each entry point replaces in the reference (paths relative to the chrom-rs tree):
 *   chrom_cuda_solve_single_batch   RK4Solver::solve / EulerSolver::solve on a LangmuirSingle scenario
 *                                   src/solver/methods/rk4.rs:205-350, euler.rs:166-288,
 *                                   src/models/langmuir_single.rs:445-516; one call = n_cols solves
 *   chrom_cuda_solve_multi_batch    the same solvers on a LangmuirMulti scenario
 *                                   src/models/langmuir_multi.rs:717-899 (jacobian 627-668, inverse 674-686)
 *   chrom_cuda_tabulate_injection   TemporalInjection::evaluate at the stage times the solvers use
 *                                   src/models/injection.rs:410-446; rk4.rs:267,283,289,295; euler.rs:229
 *   status[]                        validate_state  src/solver/mod.rs:514-553
 *   chrom_cuda_status_message       the Err(String) texts of src/solver/mod.rs:527-548
 *   chrom_cuda_validate_cfg         SolverType::validate  src/solver/traits.rs:172-185
 *   chrom_cuda_time_points          time_points of SimulationResult  rk4.rs:254,327
 *   chrom_cuda_sample_indices       sample_indices  src/physics/traits.rs:1010-1023 (the export / plot sampling rule)
 *   chrom_cuda_node_index           Detector::node_index  src/domain/detector.rs:285-290
 *   summary[]                       trapezoid / first_moment / peak_sigma of the validation suite
 *                                   validation/dissimilarity.rs:81-88, validation/main.rs:190-206
 *   chrom_cuda_plan_rsf             rsf  validation/dissimilarity.rs:173-203 (two series on one time grid)
 *   plan_* functions                no reference counterpart: keep a batch resident in HBM between calls
 *
 * Conventions: every function returns CHROM_OK (0) or a negative CHROM_ERR_*; the text of the last
 * error on a context is returned by chrom_cuda_last_error.  Buffers are caller-allocated; the
 * library never keeps or frees caller memory.  A context is not thread-safe; distinct contexts are.
 * There is no CPU fallback: without a CUDA device chrom_cuda_create fails.
 */
#ifndef CHROM_CUDA_H
#define CHROM_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHROM_OK 0
#define CHROM_ERR_INVALID (-1)     /* bad argument / configuration (message says which) */
#define CHROM_ERR_CUDA (-2)        /* CUDA runtime error */
#define CHROM_ERR_UNSUPPORTED (-3) /* shape outside what the kernels cover */
#define CHROM_ERR_NO_DEVICE (-4)   /* no usable CUDA device: there is no CPU fallback */

#define CHROM_EULER 0
#define CHROM_RK4 1

/* Arithmetic mode of the kernels.
 * FAST  : algebraically reduced RHS (one reciprocal per cell-stage), FMA contraction.  Within
 *         1e-9 relative of the reference on concentration profiles (typically ~1e-13).
 * EXACT : every IEEE operation of the reference in the reference's order, no FMA.  Bit-identical
 *         to the reference arithmetic (n<=4 closed-form inverses and n>=5 nalgebra LU included). */
#define CHROM_ARITH_FAST 0
#define CHROM_ARITH_EXACT 1
/* LangmuirMulti.  I + Fe*M = diag(alpha) - p q^T (langmuir_multi.rs:655-663), so its LU with partial pivoting
 * needs only the O(n) scalars that define it: FAST evaluates that elimination in closed form (same pivots, same
 * partial-pivoting test; a cell that needs a row exchange is redone with the dense row-exchanging LU).
 * CHROM_ARITH_FAST_STRUCTURED is the explicit name of that mode (= FAST for LangmuirMulti).
 * CHROM_ARITH_FAST_DENSE forms the n x n matrix per cell and factorises it with a dense register-resident
 * per-thread LU with partial pivoting (n <= 8; larger n run FAST): the same solution to rounding, more
 * arithmetic, kept selectable because it is the literal algorithm and the pivoting fallback of FAST. */
#define CHROM_ARITH_FAST_STRUCTURED 3
#define CHROM_ARITH_FAST_DENSE 4

#define CHROM_INJ_NONE 0
#define CHROM_INJ_DIRAC 1     /* p0 = time, p1 = amount */
#define CHROM_INJ_GAUSSIAN 2  /* p0 = center, p1 = width, p2 = peak_concentration */
#define CHROM_INJ_RECTANGLE 3 /* p0 = start, p1 = end, p2 = concentration */

typedef struct chrom_cuda_ctx chrom_cuda_ctx;
typedef struct chrom_cuda_plan chrom_cuda_plan;

/* One LangmuirSingle column.  fe, ue, dz are passed verbatim (the reference deserialises them from
 * the model file and never recomputes them: src/config/model.rs:111-148). */
typedef struct {
  double lambda, langmuir_k, port_number, fe, ue, dz;
  uint32_t inj_table;     /* index into inj_tables */
  uint32_t detector_cell; /* where "outlet" and "summary" are sampled: 0 = the column outlet (last cell, the reference's
                             default Detector::outlet()); k > 0 = cell k - 1 (Detector::node_index, detector.rs:285) */
} chrom_single_col;

/* One LangmuirMulti column; its species are species[species_off .. species_off + n_species). */
typedef struct {
  double fe, ue, dz, stationary_fraction;
  uint32_t species_off;
  uint32_t detector_cell; /* as in chrom_single_col */
} chrom_multi_col;

typedef struct {
  double lambda, langmuir_k;
  uint32_t port_number;
  uint32_t inj_table;
} chrom_species;

typedef struct {
  int32_t method;           /* CHROM_EULER | CHROM_RK4 */
  int32_t arith;            /* CHROM_ARITH_FAST | CHROM_ARITH_EXACT (| _FAST_STRUCTURED | _FAST_DENSE: LangmuirMulti) */
  double total_time;        /* > 0 */
  uint64_t time_steps;      /* > 0; dt = total_time / (double)time_steps */
  uint32_t n_points;        /* nz >= 2 */
  uint32_t n_species;       /* 1 for the single-species entry points */
  uint64_t outlet_stride;   /* outlet sampled at steps 0, k, 2k, ...; 1 = every step (reference); 0 = none */
  uint64_t snapshot_stride; /* full profile at steps 0, k, 2k, ...; 1 = the whole trajectory; 0 = none */
  /* Optional explicit sample lists (host pointers, read during the call only): strictly increasing step indices in
   * [0, time_steps] — what sample_indices(time_steps + 1, n) returns (physics/traits.rs:1010-1023).  A non-NULL list
   * replaces the corresponding stride: the outlet / snapshot buffers then hold exactly n_*_steps samples per column. */
  const uint64_t* outlet_steps;
  uint64_t n_outlet_steps;
  const uint64_t* snapshot_steps;
  uint64_t n_snapshot_steps;
} chrom_run_cfg;

/* Values per (column, species) of the on-device chromatogram summary, over ALL time steps at the detector cell:
 *   [0] trapezoid area  sum 0.5 (c_k + c_k+1)(t_k+1 - t_k)           validation/dissimilarity.rs:81-88
 *   [1] first moment    trapezoid(t c) / area   (retention time)     validation/main.rs:190-194
 *   [2] maximum         [3] time of the maximum (first occurrence)
 *   [4] sigma           sqrt(trapezoid((t - [1])^2 c) / area)        validation/main.rs:197-206
 *   [5] minimum         (undershoot / oscillation indicator of the stability scans) */
#define CHROM_SUMMARY_LEN 6

typedef struct {
  double h2d_ms;    /* host->device copies (CUDA events, max over devices) */
  double kernel_ms; /* solver kernels only, CUDA events on the launch stream, max over devices */
  double d2h_ms;    /* device->host copies */
  double total_ms;  /* host wall clock of the whole call */
  uint64_t kernel_launches;
  uint64_t h2d_bytes, d2h_bytes;
} chrom_cuda_timing;

/* ---- context ------------------------------------------------------------------------------- */
int32_t chrom_cuda_device_count(void);
/* device_ids == NULL: use devices 0..n_devices-1 (n_devices <= 0: all visible devices). */
int32_t chrom_cuda_create(const int32_t* device_ids, int32_t n_devices, chrom_cuda_ctx** out);
void chrom_cuda_destroy(chrom_cuda_ctx* ctx);
/* Valid until the next call on ctx.  ctx == NULL: the last error of chrom_cuda_create on this thread. */
const char* chrom_cuda_last_error(const chrom_cuda_ctx* ctx);
const char* chrom_cuda_version(void);
int32_t chrom_cuda_last_timing(const chrom_cuda_ctx* ctx, chrom_cuda_timing* out);

/* Page-locked host memory for buffers handed to solve / download (pageable memory works, slower). */
int32_t chrom_cuda_host_alloc(size_t bytes, void** out);
void chrom_cuda_host_free(void* p);

/* Multi-GPU hosts: pins the CALLING thread to the CPUs of the NUMA node device `device_index` of the context
 * hangs off (sysfs: the device's PCI numa_node and that node's cpulist, intersected with the CPUs the process may
 * use) and prefers that node for the thread's allocations, so that page-locked buffers allocated afterwards and
 * the copy-back of outlet series do not cross the socket interconnect.  Call it before chrom_cuda_host_alloc.
 * *numa_node = the node (-1: unknown, nothing changed), *n_cpus = CPUs in the new affinity mask.  Either may be NULL.
 * No reference counterpart (the reference is a single-process CPU library). */
int32_t chrom_cuda_bind_host_thread(chrom_cuda_ctx* ctx, int32_t device_index, int32_t* numa_node, int32_t* n_cpus);

/* ---- host-side helpers (pure CPU, no device needed) ------------------------------------------ */
/* Error text of SolverType::validate; NULL when the configuration is valid. */
const char* chrom_cuda_validate_cfg(const chrom_run_cfg* cfg);
/* Number of samples a stride produces: 0 if stride == 0 else time_steps / stride + 1. */
uint64_t chrom_cuda_n_samples(uint64_t time_steps, uint64_t stride);
/* Samples per column of the outlet / snapshot buffers for this configuration (list length, else the stride rule). */
uint64_t chrom_cuda_cfg_n_out(const chrom_run_cfg* cfg);
uint64_t chrom_cuda_cfg_n_snap(const chrom_run_cfg* cfg);
/* sample_indices(total, Some(n)) of the reference (n == 0 or n >= total: every index; n == 1: [0]; else
 * (i * (total - 1)) / (n - 1) with the last forced to total - 1).  Writes at most `cap` entries to out (may be NULL)
 * and returns the number of indices the rule produces. */
uint64_t chrom_cuda_sample_indices(uint64_t total, uint64_t n, uint64_t* out, uint64_t cap);
/* Detector::node_index: round(position / (column_length / n_points)) clamped to n_points - 1 (position in metres). */
uint32_t chrom_cuda_node_index(double position, double column_length, uint32_t n_points);
/* out[0] = 0.0, out[s + 1] = ((double)s + 1.0) * dt  — (time_steps + 1) values. */
void chrom_cuda_time_points(double total_time, uint64_t time_steps, double* out);
/* Stages per step of a table: 3 for RK4 (t, t + dt/2.0, t + dt), 1 for Euler (t); t = (double)step * dt. */
uint32_t chrom_cuda_table_stages(int32_t method);
/* Fills out[time_steps][stages] with TemporalInjection::evaluate at the solver's stage times. */
int32_t chrom_cuda_tabulate_injection(int32_t kind, double p0, double p1, double p2, int32_t method,
                                      double total_time, uint64_t time_steps, double* out);
/* The contiguous column range device `device_index` of `n_devices` owns in a batch of n_cols
 * (ceil(n_cols / n_devices) columns per device, trailing devices may get fewer or none). */
void chrom_cuda_shard_range(uint64_t n_cols, uint32_t n_devices, uint32_t device_index, uint64_t* col0,
                            uint64_t* count);
/* The reference's validate_state error string for a nonzero status (0 bytes written when status == 0).
 * Returns the length the full message needs. */
size_t chrom_cuda_status_message(int64_t status, char* buf, size_t buf_len);

/* ---- one-shot solves (host buffers in, host buffers out) ------------------------------------- */
/* inj_tables : [n_tables][time_steps][stages] host-evaluated inlet concentrations.
 * y0         : [n_cols][nz] initial state, or NULL = zeros (setup_initial_state).
 * outlet     : [n_cols][n_out]        last cell at the sampled steps, or NULL
 * snapshots  : [n_cols][n_snap][nz]   full profiles at the sampled steps, or NULL
 * final_state: [n_cols][nz]           or NULL
 * summary    : [n_cols][CHROM_SUMMARY_LEN]  area, first moment, max, time of max, sigma, min of the detector
 *                                     signal over all steps (see CHROM_SUMMARY_LEN), or NULL
 * status     : [n_cols]               0 ok; +s = NaN first present after step s; -s = Inf (no NaN)
 *                                     after step s (validate_state runs after every step), or NULL */
int32_t chrom_cuda_solve_single_batch(chrom_cuda_ctx* ctx, const chrom_run_cfg* cfg, const chrom_single_col* cols,
                                      uint64_t n_cols, const double* inj_tables, uint32_t n_tables,
                                      const double* y0, double* outlet, double* snapshots, double* final_state,
                                      double* summary, int64_t* status);

/* State layout [n_cols][n_species][nz] (= nalgebra's column-major nz x n matrix per column).
 * outlet [n_cols][n_species][n_out]; snapshots [n_cols][n_snap][n_species][nz];
 * final_state [n_cols][n_species][nz]; summary [n_cols][n_species][CHROM_SUMMARY_LEN]; status [n_cols]. */
int32_t chrom_cuda_solve_multi_batch(chrom_cuda_ctx* ctx, const chrom_run_cfg* cfg, const chrom_multi_col* cols,
                                     uint64_t n_cols, const chrom_species* species, uint64_t n_species_total,
                                     const double* inj_tables, uint32_t n_tables, const double* y0,
                                     double* outlet, double* snapshots, double* final_state, double* summary,
                                     int64_t* status);

/* ---- plans: a batch kept resident in HBM (inputs uploaded once, outputs left on the device) -- */
#define CHROM_WANT_OUTLET 1u
#define CHROM_WANT_SNAPSHOTS 2u
#define CHROM_WANT_FINAL 4u
#define CHROM_WANT_SUMMARY 8u
#define CHROM_WANT_STATUS 16u

int32_t chrom_cuda_plan_single(chrom_cuda_ctx* ctx, const chrom_run_cfg* cfg, const chrom_single_col* cols,
                               uint64_t n_cols, const double* inj_tables, uint32_t n_tables, const double* y0,
                               uint32_t want, chrom_cuda_plan** out);
int32_t chrom_cuda_plan_multi(chrom_cuda_ctx* ctx, const chrom_run_cfg* cfg, const chrom_multi_col* cols,
                              uint64_t n_cols, const chrom_species* species, uint64_t n_species_total,
                              const double* inj_tables, uint32_t n_tables, const double* y0, uint32_t want,
                              chrom_cuda_plan** out);
/* Runs the solve from the uploaded initial state; returns after the kernels have finished. */
int32_t chrom_cuda_plan_execute(chrom_cuda_plan* plan);
/* Copies the requested outputs to host buffers (any may be NULL). */
int32_t chrom_cuda_plan_download(chrom_cuda_plan* plan, double* outlet, double* snapshots, double* final_state,
                                 double* summary, int64_t* status);
/* Surface resolution Rsf = integral |Y_a - Y_b| dt (Y = c / area) between the outlet series two executed plans hold
 * on the device, per (column, species): the reference's Euler-vs-RK4 criterion (validation/main.rs:430-440) without
 * moving either series to the host.  Both plans must have the same columns, species and outlet sampling (one time
 * grid: rsf() then resamples nothing).  rsf: host [n_cols][n_species].  A series of zero area gives NaN. */
int32_t chrom_cuda_plan_rsf(chrom_cuda_plan* a, chrom_cuda_plan* b, double* rsf);
/* Human-readable description of the kernel variant and launch geometry chosen. */
int32_t chrom_cuda_plan_describe(const chrom_cuda_plan* plan, char* buf, size_t buf_len);
void chrom_cuda_plan_destroy(chrom_cuda_plan* plan);

/* ---- measurement ----------------------------------------------------------------------------- */
/* Dependent-free DFMA micro-benchmark on device `device_index` of the context: measured FP64
 * peak in TFLOP/s (2 flops per DFMA) and the SM clock-independent duration in ms. */
int32_t chrom_cuda_measure_fp64_peak(chrom_cuda_ctx* ctx, int32_t device_index, double* tflops, double* ms);
/* Writes a buffer larger than the L2 (256 MiB) on every device of the context: cold-cache timing. */
int32_t chrom_cuda_flush_l2(chrom_cuda_ctx* ctx);
/* Device copy bandwidth (read + write bytes / time) in GB/s over `bytes` bytes. */
int32_t chrom_cuda_measure_copy_bw(chrom_cuda_ctx* ctx, int32_t device_index, size_t bytes, double* gbs);
/* Max relative error vs IEEE 1.0/x over n samples of +-[lo, hi] of: max_rel_err[0] the MUFU.RCP64H
 * seed, [1] seed + cubic step, [2] seed + cubic + Newton step (what the FAST kernels use). */
int32_t chrom_cuda_probe_rcp(chrom_cuda_ctx* ctx, double lo, double hi, uint32_t n, double* max_rel_err);

#ifdef __cplusplus
}
#endif
#endif /* CHROM_CUDA_H */
