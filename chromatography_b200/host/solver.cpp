// The reference's `Solver` boundary on top of the C ABI of libchrom_cuda.so
// (include/chrom/solver.hpp lists the reference file:line of every item mirrored here).
//
// CudaSolverBase::solve_batch is `RK4Solver::solve` / `EulerSolver::solve` (src/solver/methods/rk4.rs:205-350,
// euler.rs:166-288) with the time loop replaced by ONE call of chrom_cuda_solve_{single,multi}_batch:
// the validation order, the error strings, the time points and the metadata are the reference's.
#include "chrom/solver.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <map>
#include <thread>
#include <cstring>
#include <limits>
#include <mutex>
#include <set>
#include <stdexcept>

#include "chrom_cuda.h"

namespace chrom {

// ---- SolverType ------------------------------------------------------------------------------------------
SolverType SolverType::TimeEvolution(double total_time, std::size_t time_steps) {
  SolverType t;
  t.kind_ = kTimeEvolution;
  t.name_ = "TimeEvolution";
  t.total_time = total_time;
  t.time_steps = time_steps;
  return t;
}
SolverType SolverType::Other(std::string name) {
  SolverType t;
  t.kind_ = kOther;
  t.name_ = std::move(name);
  return t;
}

Result<Unit> SolverType::validate() const {
  switch (kind_) {
    case kTimeEvolution:
      if (total_time <= 0.0) return Err("Total time must be positive");
      if (time_steps == 0) return Err("TimeSteps must be greater than 0");
      return Ok();
    default: return Ok();
  }
}

bool SolverType::operator==(const SolverType& o) const {
  return kind_ == o.kind_ && name_ == o.name_ && total_time == o.total_time && time_steps == o.time_steps;
}

// ---- SolverConfiguration -------------------------------------------------------------------------------------
SolverConfiguration SolverConfiguration::time_evolution(double total_time, std::size_t time_steps) {
  return SolverConfiguration(SolverType::TimeEvolution(total_time, time_steps));
}
SolverConfiguration SolverConfiguration::with_step(std::size_t n) const {
  SolverConfiguration c = *this;
  c.step = n;
  return c;
}

// ---- boundaries -----------------------------------------------------------------------------------------------
std::string TimeAxisConvention::to_string() const {
  switch (kind) {
    case None: return "None";
    case First: return "First";
    case Last: return "Last";
    default: return "Index (" + std::to_string(index) + ")";
  }
}

Result<Unit> DimensionBoundary::validate() const {
  if (is_empty()) return Err("Dimensions '" + name + "' must have at least one boundary state");
  return Ok();
}

DomainBoundaries DomainBoundaries::create(std::vector<DimensionBoundary> dims, std::optional<TimeAxisConvention> convention) {
  DomainBoundaries b;
  b.dimensions = std::move(dims);
  b.convention = convention.value_or(TimeAxisConvention::last());
  return b;
}

DomainBoundaries DomainBoundaries::temporal(PhysicalState initial) {
  std::vector<PhysicalState> states;
  states.push_back(std::move(initial));
  std::vector<DimensionBoundary> dims;
  dims.emplace_back("t", std::move(states));
  return DomainBoundaries(std::move(dims));
}

std::size_t DomainBoundaries::sdim() const {
  return convention.kind == TimeAxisConvention::None ? ndim() : ndim() - 1;
}

std::optional<std::size_t> DomainBoundaries::time_index() const {
  switch (convention.kind) {
    case TimeAxisConvention::Last: return ndim() - 1;  // wraps on an empty domain, like the reference's usize
    case TimeAxisConvention::First: return std::size_t{0};
    case TimeAxisConvention::None: return std::nullopt;
    default: return convention.index;
  }
}

const DimensionBoundary* DomainBoundaries::time_boundary() const {
  auto i = time_index();
  if (!i || *i >= dimensions.size()) return nullptr;
  return &dimensions[*i];
}

const PhysicalState* DomainBoundaries::initial_condition() const {
  const DimensionBoundary* b = time_boundary();
  if (!b || b->states.empty()) return nullptr;
  return &b->states.front();
}

std::vector<const DimensionBoundary*> DomainBoundaries::spatial_boundaries() const {
  auto excl = time_index();
  std::vector<const DimensionBoundary*> r;
  for (std::size_t i = 0; i < dimensions.size(); ++i)
    if (!excl || *excl != i) r.push_back(&dimensions[i]);
  return r;
}

const DimensionBoundary* DomainBoundaries::get_boundary(const std::string& name) const {
  for (const auto& d : dimensions)
    if (d.name == name) return &d;
  return nullptr;
}

Result<Unit> DomainBoundaries::validate() const {
  if (dimensions.empty()) return Err("Dimension boundaries cannot be empty.");
  for (const auto& d : dimensions) {
    auto r = d.validate();
    if (r.is_err()) return r;
  }
  std::set<std::string> names;
  for (const auto& d : dimensions) names.insert(d.name);
  if (names.size() != dimensions.size()) return Err("It is impossible to store two dimensions with the same name.");
  return Ok();
}

std::string Scenario::debug_string() const {
  std::string s = "Scenario { name: \"" + std::string(get_model_name()) + "\", dimension: " + std::to_string(ndim()) +
                  ", spatial dim: " + std::to_string(sdim()) + ", is time dependent: " +
                  (is_time_dependent() ? "true" : "false") + ", Boundaries / conditions: DomainBoundaries { dimensions: [";
  for (std::size_t i = 0; i < conditions.dimensions.size(); ++i) {
    const auto& d = conditions.dimensions[i];
    s += (i ? ", " : "") + std::string("DimensionBoundary { name: \"") + d.name + "\", states: " +
         std::to_string(d.states.size()) + " }";
  }
  s += "], convention: " + conditions.convention.to_string() + " } }";
  return s;
}

std::string validate_state_message(std::int64_t status) {
  if (status == 0) return {};
  char buf[512];
  chrom_cuda_status_message(status, buf, sizeof buf);
  return buf;
}

// ---- CudaContext --------------------------------------------------------------------------------------------------
struct CudaContext::Impl {
  std::mutex mu;  // a chrom_cuda_ctx is not thread-safe; Solver::solve(&self) must be re-entrant
};

Result<std::shared_ptr<CudaContext>> CudaContext::create(const std::vector<int>& device_ids) {
  chrom_cuda_ctx* h = nullptr;
  std::vector<std::int32_t> ids(device_ids.begin(), device_ids.end());
  const std::int32_t rc = ids.empty() ? chrom_cuda_create(nullptr, 0, &h) : chrom_cuda_create(ids.data(), (std::int32_t)ids.size(), &h);
  if (rc != CHROM_OK) return Err(std::string("CUDA backend unavailable (no CPU fallback): ") + chrom_cuda_last_error(nullptr));
  std::shared_ptr<CudaContext> c(new CudaContext());
  c->h_ = h;
  c->impl_ = std::make_unique<Impl>();
  return c;
}

Result<std::shared_ptr<CudaContext>> CudaContext::global() {
  static std::mutex mu;
  static std::shared_ptr<CudaContext> g;
  std::lock_guard<std::mutex> lk(mu);
  if (!g) {
    auto r = create();
    if (r.is_err()) return r;
    g = r.unwrap();
  }
  return g;
}

CudaContext::~CudaContext() {
  if (h_) chrom_cuda_destroy(h_);
}

// ---- solve ---------------------------------------------------------------------------------------------------------
namespace {

// Page-locked host buffer from the library (outputs of a sweep are GB-sized: pageable memory halves D2H speed).
template <class T>
class PinnedBuf {
 public:
  PinnedBuf() = default;
  ~PinnedBuf() {
    if (p_) chrom_cuda_host_free(p_);
  }
  PinnedBuf(const PinnedBuf&) = delete;
  PinnedBuf& operator=(const PinnedBuf&) = delete;
  bool alloc(std::size_t n) {
    n_ = n;
    if (n == 0) return true;
    void* p = nullptr;
    if (chrom_cuda_host_alloc(n * sizeof(T), &p) != CHROM_OK) return false;
    p_ = static_cast<T*>(p);
    return true;
  }
  T* get() const { return p_; }
  std::size_t size() const { return n_; }

 private:
  T* p_ = nullptr;
  std::size_t n_ = 0;
};

// De-duplicated injection tables of a batch: [n_tables][steps][stages].
struct TableSet {
  int method;
  double total_time;
  std::uint64_t steps;
  std::uint32_t stages;
  std::vector<TemporalInjection> profiles;
  std::vector<double> data;

  std::uint32_t add(const TemporalInjection& inj) {
    for (std::size_t i = 0; i < profiles.size(); ++i)
      if (profiles[i].same_profile(inj)) return (std::uint32_t)i;
    const std::size_t off = data.size();
    data.resize(off + steps * stages);
    double* out = data.data() + off;
    if (inj.kind() == TemporalInjection::kCustom) {
      // A closure cannot cross the C ABI: evaluate it here at exactly the solver's stage times
      // (rk4.rs:267,283,289,295: t = step as f64 * dt, t + dt / 2.0, t + dt; euler.rs:229).
      const double dt = total_time / (double)steps;
      for (std::uint64_t s = 0; s < steps; ++s) {
        const double t = (double)s * dt;
        if (method == CHROM_RK4) {
          out[3 * s + 0] = inj.evaluate(t);
          out[3 * s + 1] = inj.evaluate(t + dt / 2.0);
          out[3 * s + 2] = inj.evaluate(t + dt);
        } else {
          out[s] = inj.evaluate(t);
        }
      }
    } else {
      chrom_cuda_tabulate_injection((std::int32_t)inj.kind(), inj.p0(), inj.p1(), inj.p2(), method, total_time, steps, out);
    }
    profiles.push_back(inj);
    return (std::uint32_t)(profiles.size() - 1);
  }
};

bool any_nonzero_bits(const double* p, std::size_t n) {
  for (std::size_t i = 0; i < n; ++i) {
    std::uint64_t b;
    std::memcpy(&b, p + i, 8);
    if (b != 0) return true;  // -0.0 counts: y0 = NULL means +0.0 everywhere
  }
  return false;
}

}  // namespace

Result<SimulationResult> CudaSolverBase::solve(const Scenario& scenario, const SolverConfiguration& config) const {
  auto r = solve_batch({&scenario}, config, defaults_);
  if (r.is_err()) return Err(r.unwrap_err());
  return std::move(r.unwrap()[0]);
}

Result<std::vector<Result<SimulationResult>>> CudaSolverBase::solve_scan(const std::vector<const Scenario*>& scenarios,
                                                                         const std::vector<SolverConfiguration>& configs,
                                                                         const BatchOptions& opt,
                                                                         std::size_t max_concurrent) const {
  if (configs.size() != scenarios.size()) return Err("solve_scan needs one SolverConfiguration per scenario");
  // group key: model type / grid / species count / time grid.  Anything solve_batch would reject per
  // scenario (unsupported model, bad initial condition) still reaches it and gets its own error entry.
  struct Group {
    std::vector<std::size_t> idx;
  };
  std::vector<Group> groups;
  std::map<std::string, std::size_t> by_key;
  for (std::size_t i = 0; i < scenarios.size(); ++i) {
    const Scenario& sc = *scenarios[i];
    const SolverType& st = configs[i].solver_type;
    const auto* m = dynamic_cast<const LangmuirMulti*>(sc.model.get());
    char key[160];
    std::uint64_t tbits;
    std::memcpy(&tbits, &st.total_time, 8);
    snprintf(key, sizeof key, "%s|%zu|%zu|%d|%llx|%llu", sc.model->name(), (std::size_t)sc.model->points(),
             m ? (std::size_t)m->n_species() : (std::size_t)1, (int)st.kind(), (unsigned long long)tbits,
             (unsigned long long)st.time_steps);
    auto it = by_key.find(key);
    if (it == by_key.end()) {
      it = by_key.emplace(key, groups.size()).first;
      groups.emplace_back();
    }
    groups[it->second].idx.push_back(i);
  }
  std::vector<Result<SimulationResult>> results;
  results.reserve(scenarios.size());
  for (std::size_t i = 0; i < scenarios.size(); ++i) results.emplace_back(Err(std::string("not solved")));
  const std::size_t workers = std::max<std::size_t>(1, std::min(max_concurrent, groups.size()));
  std::atomic<std::size_t> next{0};
  std::mutex err_mu;
  std::string fatal;
  auto work = [&](std::size_t w) {
    // worker 0 keeps this solver's context; the others get their own (one context = one set of streams)
    std::shared_ptr<CudaContext> ctx = (w == 0) ? ctx_ : nullptr;
    for (;;) {
      const std::size_t g = next.fetch_add(1);
      if (g >= groups.size()) return;
      if (!ctx) {
        auto c = (w == 0) ? CudaContext::global() : CudaContext::create();
        if (c.is_err()) {
          std::lock_guard<std::mutex> lk(err_mu);
          fatal = c.unwrap_err();
          return;
        }
        ctx = c.unwrap();
      }
      std::vector<const Scenario*> sub;
      for (std::size_t i : groups[g].idx) sub.push_back(scenarios[i]);
      auto r = solve_batch_on(ctx, sub, configs[groups[g].idx[0]], opt);
      if (r.is_err()) {  // a failure of the whole group (invalid configuration, ...): every member reports it
        for (std::size_t i : groups[g].idx) results[i] = Err(r.unwrap_err());
        continue;
      }
      auto v = std::move(r.unwrap());
      for (std::size_t k = 0; k < v.size(); ++k) results[groups[g].idx[k]] = std::move(v[k]);
    }
  };
  std::vector<std::thread> threads;
  for (std::size_t w = 1; w < workers; ++w) threads.emplace_back(work, w);
  work(0);
  for (auto& t : threads) t.join();
  if (!fatal.empty()) return Err(fatal);
  return results;
}

Result<std::vector<Result<SimulationResult>>> CudaSolverBase::solve_batch(const std::vector<const Scenario*>& scenarios,
                                                                          const SolverConfiguration& config,
                                                                          const BatchOptions& opt) const {
  return solve_batch_on(nullptr, scenarios, config, opt);
}

Result<std::vector<Result<SimulationResult>>> CudaSolverBase::solve_batch_on(std::shared_ptr<CudaContext> on,
                                                                             const std::vector<const Scenario*>& scenarios,
                                                                             const SolverConfiguration& config,
                                                                             const BatchOptions& opt) const {
  // ====== Step 1: validation (rk4.rs:213-231) ======
  {
    auto v = config.validate();
    if (v.is_err()) return Err(v.unwrap_err());
  }
  std::vector<Result<SimulationResult>> results;
  results.reserve(scenarios.size());
  std::vector<std::size_t> live;  // indices of scenarios that reach the device
  std::vector<std::string> early(scenarios.size());
  for (std::size_t i = 0; i < scenarios.size(); ++i) {
    auto v = scenarios[i]->validate();
    if (v.is_err()) early[i] = v.unwrap_err();
  }
  const SolverType& st = config.solver_type;
  if (st.kind() != SolverType::kTimeEvolution)  // the RK4 text names the Euler solver too (rk4.rs:227-230)
    return Err("EulerSolver only supports TimeEvolution configuration, got " + st.name());
  const double total_time = st.total_time;
  const std::uint64_t steps = st.time_steps;
  const double dt = total_time / (double)steps;

  // ====== Step 2: setup — initial conditions and model parameters ======
  const LangmuirSingle* single0 = nullptr;
  const LangmuirMulti* multi0 = nullptr;
  for (std::size_t i = 0; i < scenarios.size(); ++i) {
    if (!early[i].empty()) continue;
    const Scenario& sc = *scenarios[i];
    if (sc.conditions.initial_condition() == nullptr) {
      early[i] = "No initial condition found in domain boundaries";  // rk4.rs:242-245
      continue;
    }
    const auto* s = dynamic_cast<const LangmuirSingle*>(sc.model.get());
    const auto* m = dynamic_cast<const LangmuirMulti*>(sc.model.get());
    if (!s && !m) {
      early[i] = std::string("model '") + sc.model->name() +
                 "' is not supported by the CUDA backend (LangmuirSingle and LangmuirMulti only; no CPU fallback)";
      continue;
    }
    const PhysicalData* c = sc.conditions.initial_condition()->get(PhysicalQuantity::Concentration);
    if (!c) {
      early[i] = "Concentration is required";  // langmuir_single.rs:470, langmuir_multi.rs:738
      continue;
    }
    if (s) {
      if (!c->is_vector()) {
        early[i] = "Expected vector data";
        continue;
      }
      if (c->as_vector().size() != s->points()) {
        early[i] = "Concentration profile size " + std::to_string(c->as_vector().size()) + " vs points discretization " +
                   std::to_string(s->points());  // langmuir_single.rs:474-480
        continue;
      }
    } else {
      if (!c->is_matrix()) {
        early[i] = "Expected matrix data";
        continue;
      }
      const DMatrix& mm = c->as_matrix();
      if (mm.nrows != m->points() || mm.ncols != m->n_species()) {
        early[i] = "State matrix shape [" + std::to_string(m->points()) + ", " + std::to_string(m->n_species()) +
                   "] expected, got [" + std::to_string(mm.nrows) + ", " + std::to_string(mm.ncols) + "]";
        continue;
      }
    }
    if (!single0 && !multi0) {
      single0 = s;
      multi0 = m;
    } else if ((single0 != nullptr) != (s != nullptr) || sc.model->points() != (single0 ? single0->points() : multi0->points()) ||
               (m && m->n_species() != multi0->n_species())) {
      return Err("solve_batch needs scenarios of one model type, grid size and species count");
    }
    live.push_back(i);
  }

  auto finish_early = [&]() {
    for (std::size_t i = 0; i < scenarios.size(); ++i) results.emplace_back(Err(early[i]));
    return results;
  };
  if (live.empty()) return finish_early();

  const bool is_multi = multi0 != nullptr;
  const std::size_t nz = is_multi ? multi0->points() : single0->points();
  const std::size_t n = is_multi ? multi0->n_species() : 1;
  const std::size_t n_cols = live.size();

  chrom_run_cfg cfg;
  std::memset(&cfg, 0, sizeof cfg);
  cfg.method = method_;
  cfg.arith = (std::int32_t)arith_;
  cfg.total_time = total_time;
  cfg.time_steps = steps;
  cfg.n_points = (std::uint32_t)nz;
  cfg.n_species = (std::uint32_t)n;
  cfg.outlet_stride = opt.outlet_stride;
  cfg.snapshot_stride = opt.snapshot_stride;
  if (nz > std::numeric_limits<std::uint32_t>::max()) return Err("n_points exceeds the backend's 32-bit grid index");

  TableSet tables{method_, total_time, steps, chrom_cuda_table_stages(method_), {}, {}};
  std::vector<chrom_single_col> scols;
  std::vector<chrom_multi_col> mcols;
  std::vector<chrom_species> species;
  std::vector<double> y0(n_cols * n * nz);
  for (std::size_t k = 0; k < n_cols; ++k) {
    const Scenario& sc = *scenarios[live[k]];
    const PhysicalData* c = sc.conditions.initial_condition()->get(PhysicalQuantity::Concentration);
    if (!is_multi) {
      const auto* s = static_cast<const LangmuirSingle*>(sc.model.get());
      chrom_single_col col;
      std::memset(&col, 0, sizeof col);
      col.lambda = s->lambda();
      col.langmuir_k = s->langmuir_k();
      col.port_number = s->port_number();
      col.fe = s->fe();  // stored fields verbatim: the reference never recomputes them
      col.ue = s->ue();
      col.dz = s->dz();
      col.inj_table = tables.add(s->injection());
      scols.push_back(col);
      std::memcpy(&y0[k * nz], c->as_vector().data(), nz * sizeof(double));
    } else {
      const auto* m = static_cast<const LangmuirMulti*>(sc.model.get());
      chrom_multi_col col;
      std::memset(&col, 0, sizeof col);
      col.fe = m->fe();
      col.ue = m->ue();
      col.dz = m->dz();
      col.stationary_fraction = m->stationary_fraction();
      col.species_off = (std::uint32_t)species.size();
      mcols.push_back(col);
      for (const auto& sp : m->species_params()) {
        chrom_species cs;
        std::memset(&cs, 0, sizeof cs);
        cs.lambda = sp.lambda;
        cs.langmuir_k = sp.langmuir_k;
        cs.port_number = sp.port_number;
        cs.inj_table = tables.add(sp.injection);
        species.push_back(cs);
      }
      // nalgebra's column-major nz x n matrix IS the device layout [species][nz]
      std::memcpy(&y0[k * n * nz], c->as_matrix().data.data(), n * nz * sizeof(double));
    }
  }
  const bool zero_ic = !any_nonzero_bits(y0.data(), y0.size());  // setup_initial_state: skip the upload

  // the context first: without a usable GPU this is where the call fails (there is no CPU path)
  std::shared_ptr<CudaContext> ctx = on ? on : ctx_;
  if (!ctx) {
    auto g = CudaContext::global();
    if (g.is_err()) return Err(g.unwrap_err());
    ctx = ctx_ = g.unwrap();
  }

  const std::uint64_t n_out = chrom_cuda_n_samples(steps, opt.outlet_stride);
  const std::uint64_t n_snap = chrom_cuda_n_samples(steps, opt.snapshot_stride);
  PinnedBuf<double> outlet, snaps, final_state, summary;
  PinnedBuf<std::int64_t> status;
  if (!outlet.alloc(n_cols * n * n_out) || !snaps.alloc(n_cols * n_snap * n * nz) || !final_state.alloc(n_cols * n * nz) ||
      !summary.alloc(opt.summary ? n_cols * n * CHROM_SUMMARY_LEN : 0) || !status.alloc(n_cols))
    return Err("chrom_cuda_host_alloc failed: not enough page-locked host memory for the requested outputs "
               "(lower BatchOptions::snapshot_stride / outlet_stride)");

  // ====== Step 3: time integration — on the device ======
  {
    std::lock_guard<std::mutex> lk(ctx->impl()->mu);
    std::int32_t rc;
    if (!is_multi)
      rc = chrom_cuda_solve_single_batch(ctx->handle(), &cfg, scols.data(), n_cols, tables.data.data(),
                                         (std::uint32_t)tables.profiles.size(), zero_ic ? nullptr : y0.data(), outlet.get(),
                                         snaps.get(), final_state.get(), summary.get(), status.get());
    else
      rc = chrom_cuda_solve_multi_batch(ctx->handle(), &cfg, mcols.data(), n_cols, species.data(), species.size(),
                                        tables.data.data(), (std::uint32_t)tables.profiles.size(),
                                        zero_ic ? nullptr : y0.data(), outlet.get(), snaps.get(), final_state.get(),
                                        summary.get(), status.get());
    if (rc != CHROM_OK) return Err(std::string(chrom_cuda_last_error(ctx->handle())));
  }

  // ====== Step 4: build results (rk4.rs:254, 315-347) ======
  std::vector<double> time_points(steps + 1);
  chrom_cuda_time_points(total_time, steps, time_points.data());
  auto make_state = [&](const double* p) {
    if (!is_multi) return PhysicalState(PhysicalQuantity::Concentration, PhysicalData::Vector(DVector(p, p + nz)));
    DMatrix m(nz, n);
    std::memcpy(m.data.data(), p, n * nz * sizeof(double));
    return PhysicalState(PhysicalQuantity::Concentration, PhysicalData::Matrix(std::move(m)));
  };
  std::size_t k = 0;
  for (std::size_t i = 0; i < scenarios.size(); ++i) {
    if (k >= live.size() || live[k] != i) {
      results.emplace_back(Err(early[i]));
      continue;
    }
    if (status.get()[k] != 0) {  // validate_state(&state, step + 1)? (rk4.rs:331)
      results.emplace_back(Err(validate_state_message(status.get()[k])));
      ++k;
      continue;
    }
    SimulationResult r;
    r.time_points = time_points;
    r.state_trajectory.reserve(n_snap);
    for (std::uint64_t s = 0; s < n_snap; ++s) {
      r.state_trajectory.push_back(make_state(snaps.get() + (k * n_snap + s) * n * nz));
      r.trajectory_steps.push_back(s * opt.snapshot_stride);
    }
    r.final_state = make_state(final_state.get() + k * n * nz);
    r.add_metadata("solver", meta_solver_);
    r.add_metadata("time steps", std::to_string(steps));
    r.add_metadata("dt", rust_f64_to_string(dt));
    r.add_metadata("total time", rust_f64_to_string(total_time));
    if (method_ == CHROM_RK4) r.add_metadata("function evaluations", std::to_string(4 * steps));
    if (n_out) {
      for (std::size_t sp = 0; sp < n; ++sp) {
        const double* o = outlet.get() + (k * n + sp) * n_out;
        r.outlet.emplace_back(o, o + n_out);
      }
      r.outlet_time_points.resize(n_out);
      for (std::uint64_t s = 0; s < n_out; ++s) r.outlet_time_points[s] = time_points[s * opt.outlet_stride];
    }
    if (opt.summary)
      for (std::size_t sp = 0; sp < n; ++sp) {
        const double* q = summary.get() + (k * n + sp) * CHROM_SUMMARY_LEN;
        r.summary.emplace_back(q, q + CHROM_SUMMARY_LEN);
      }
    results.emplace_back(std::move(r));
    ++k;
  }
  return results;
}

// ---- chromatogram analytics (validation/dissimilarity.rs) ----------------------------------------------------------------
double trapezoid(const std::vector<double>& times, const std::vector<double>& values) {
  if (times.size() != values.size()) throw std::invalid_argument("assertion `left == right` failed");
  double sum = 0.0;
  for (std::size_t i = 0; i + 1 < times.size(); ++i) sum += 0.5 * (values[i] + values[i + 1]) * (times[i + 1] - times[i]);
  return sum;
}

std::vector<double> normalize(const std::vector<double>& times, const std::vector<double>& concentrations) {
  if (times.size() != concentrations.size())
    throw std::invalid_argument("times and concentrations must have the same length");
  const double area = trapezoid(times, concentrations);
  if (!(std::fabs(area) > std::numeric_limits<double>::epsilon()))
    throw std::invalid_argument("cannot normalise a zero signal (area = " + rust_f64_to_string(area) + ")");
  std::vector<double> r(concentrations.size());
  for (std::size_t i = 0; i < r.size(); ++i) r[i] = concentrations[i] / area;
  return r;
}

namespace {
double interpolate_at(const std::vector<double>& times, const std::vector<double>& values, double t) {
  if (t <= times.front()) return values.front();
  if (t >= times.back()) return values.back();
  std::size_t idx = (std::size_t)(std::upper_bound(times.begin(), times.end(), t) - times.begin());  // partition_point(x <= t)
  idx = idx > 0 ? idx - 1 : 0;
  idx = std::min(idx, times.size() - 2);
  const double dt = times[idx + 1] - times[idx];
  if (dt < std::numeric_limits<double>::epsilon()) return values[idx];
  const double alpha = (t - times[idx]) / dt;
  return values[idx] * (1.0 - alpha) + values[idx + 1] * alpha;
}
}  // namespace

double rsf(const std::vector<double>& t1, const std::vector<double>& c1, const std::vector<double>& t2,
           const std::vector<double>& c2) {
  const std::vector<double> y1 = normalize(t1, c1), y2 = normalize(t2, c2);
  double min_dt = std::numeric_limits<double>::infinity();
  for (std::size_t i = 0; i + 1 < t1.size(); ++i) min_dt = std::fmin(min_dt, std::fabs(t1[i + 1] - t1[i]));
  for (std::size_t i = 0; i + 1 < t2.size(); ++i) min_dt = std::fmin(min_dt, std::fabs(t2[i + 1] - t2[i]));
  const double tol = min_dt * 0.01;
  std::vector<double> grid(t1);
  grid.insert(grid.end(), t2.begin(), t2.end());
  std::stable_sort(grid.begin(), grid.end());
  // Vec::dedup_by(|a, b| (a - b).abs() < tol): drop an element that is within tol of the last one KEPT
  std::vector<double> merged;
  for (double g : grid)
    if (merged.empty() || !(std::fabs(g - merged.back()) < tol)) merged.push_back(g);
  std::vector<double> diff(merged.size());
  for (std::size_t i = 0; i < merged.size(); ++i)
    diff[i] = std::fabs(interpolate_at(t1, y1, merged[i]) - interpolate_at(t2, y2, merged[i]));
  return trapezoid(merged, diff);
}

double first_moment(const std::vector<double>& times, const std::vector<double>& concs) {
  const double area = trapezoid(times, concs);
  std::vector<double> num(times.size());
  for (std::size_t i = 0; i < times.size(); ++i) num[i] = times[i] * concs[i];
  return trapezoid(times, num) / area;
}

double peak_sigma(const std::vector<double>& times, const std::vector<double>& concs) {
  const double t_mean = first_moment(times, concs);
  const double area = trapezoid(times, concs);
  std::vector<double> num(times.size());
  for (std::size_t i = 0; i < times.size(); ++i) {
    const double d = times[i] - t_mean;
    num[i] = (d * d) * concs[i];  // powi(2)
  }
  return std::sqrt(trapezoid(times, num) / area);
}

}  // namespace chrom
