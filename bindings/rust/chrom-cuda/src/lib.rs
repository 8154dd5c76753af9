//! `impl Solver for CudaRk4Solver / CudaEulerSolver` — chrom-rs's own `Solver` trait
//! (src/solver/traits.rs:606-642) backed by libchrom_cuda.so.
//!
//! NOT compiled in this repository's environment (no Rust toolchain). The logic below is the Rust
//! twin of chromatography_b200/solver.py, which IS exercised by the GPU test-suite.
//!
//! Parameters are obtained through serde because `PhysicalModel` has no downcast hook and the model
//! fields are private, while the trait is `#[typetag::serde]` (src/physics/traits.rs:555):
//! `serde_json::to_value(&scenario.model)` gives `{"LangmuirSingle": {...}}` / `{"LangmuirMulti": {...}}`
//! (src/config/model.rs:111-148). Injection tables are evaluated on the Rust side with the model's own
//! `TemporalInjection::evaluate` at the solver's stage times, so Dirac equality and `exp` are the
//! reference's own (and `Custom` closures would work too if the model exposed them).
pub mod ffi;

use chrom_rs::models::TemporalInjection;
use chrom_rs::physics::{PhysicalData, PhysicalQuantity, PhysicalState};
use chrom_rs::solver::{Scenario, SimulationResult, Solver, SolverConfiguration, SolverType};
use nalgebra::{DMatrix, DVector};
use serde_json::Value;
use std::ffi::CStr;

pub struct CudaContext(*mut ffi::chrom_cuda_ctx);
unsafe impl Send for CudaContext {}

impl CudaContext {
    /// `devices = &[]` uses every visible GPU (the batch is cut into contiguous column ranges).
    pub fn new(devices: &[i32]) -> Result<Self, String> {
        let mut p = std::ptr::null_mut();
        let rc = unsafe {
            if devices.is_empty() { ffi::chrom_cuda_create(std::ptr::null(), 0, &mut p) }
            else { ffi::chrom_cuda_create(devices.as_ptr(), devices.len() as i32, &mut p) }
        };
        if rc != ffi::CHROM_OK {
            return Err(unsafe { CStr::from_ptr(ffi::chrom_cuda_last_error(std::ptr::null())) }.to_string_lossy().into());
        }
        Ok(Self(p))
    }
    fn last_error(&self) -> String {
        unsafe { CStr::from_ptr(ffi::chrom_cuda_last_error(self.0)) }.to_string_lossy().into()
    }
}
impl Drop for CudaContext {
    fn drop(&mut self) { unsafe { ffi::chrom_cuda_destroy(self.0) } }
}

fn f(v: &Value, k: &str) -> Result<f64, String> {
    v.get(k).and_then(Value::as_f64).ok_or_else(|| format!("CUDA backend: model field '{k}' missing"))
}

/// rk4.rs:267-295 / euler.rs:229: t = step as f64 * dt; t + dt / 2.0; t + dt.
fn tabulate(inj: &TemporalInjection, rk4: bool, dt: f64, steps: usize) -> Vec<f64> {
    let mut out = Vec::with_capacity(steps * if rk4 { 3 } else { 1 });
    for step in 0..steps {
        if rk4 {
            let t = (step as f64) * dt;
            out.push(inj.evaluate(t));
            out.push(inj.evaluate(t + dt / 2.0));
            out.push(inj.evaluate(t + dt));
        } else {
            out.push(inj.evaluate(dt * step as f64));
        }
    }
    out
}

fn solve(ctx: &CudaContext, rk4: bool, scenario: &Scenario, config: &SolverConfiguration)
    -> Result<SimulationResult, String>
{
    config.validate()?;                                              // rk4.rs:213
    scenario.validate()?;                                            // rk4.rs:217
    let (total_time, time_steps) = match &config.solver_type {       // rk4.rs:221-232 (text says Euler in both)
        SolverType::TimeEvolution { total_time, time_steps } => (*total_time, *time_steps),
        other => return Err(format!("EulerSolver only supports TimeEvolution configuration, got {}", other.name())),
    };
    let dt = total_time / (time_steps as f64);
    let initial = scenario.conditions.initial_condition()
        .ok_or_else(|| "No initial condition found in domain boundaries".to_string())?;
    let model = serde_json::to_value(&scenario.model)
        .map_err(|e| format!("CUDA backend cannot read the model parameters: {e}"))?; // Custom injections fail here

    let mut cfg = ffi::chrom_run_cfg {
        method: if rk4 { ffi::CHROM_RK4 } else { ffi::CHROM_EULER }, arith: ffi::CHROM_ARITH_FAST,
        total_time, time_steps: time_steps as u64, n_points: 0, n_species: 1,
        outlet_stride: 0, snapshot_stride: 1,                        // the reference keeps every state (rk4.rs:315)
        outlet_steps: std::ptr::null(), n_outlet_steps: 0, snapshot_steps: std::ptr::null(), n_snapshot_steps: 0,
    };
    let n_snap = time_steps + 1;
    let mut status = [0i64; 1];

    let (nz, n, snapshots, final_state) = if let Some(b) = model.get("LangmuirSingle") {
        let nz = b["n_points"].as_u64().ok_or("n_points")? as usize;
        cfg.n_points = nz as u32;
        let inj: TemporalInjection = serde_json::from_value(b["injection"].clone()).map_err(|e| e.to_string())?;
        let table = tabulate(&inj, rk4, dt, time_steps);
        let col = ffi::chrom_single_col {
            lambda: f(b, "lambda")?, langmuir_k: f(b, "langmuir_k")?, port_number: f(b, "port_number")?,
            fe: f(b, "fe")?, ue: f(b, "ue")?, dz: f(b, "dz")?, inj_table: 0, detector_cell: 0,  // 0 = the column outlet
        };
        let y0: Vec<f64> = initial.get(PhysicalQuantity::Concentration).ok_or("Concentration is required")?
            .as_vector().iter().copied().collect();
        let mut snaps = vec![0.0; n_snap * nz];
        let mut fin = vec![0.0; nz];
        let rc = unsafe { ffi::chrom_cuda_solve_single_batch(ctx.0, &cfg, &col, 1, table.as_ptr(), 1, y0.as_ptr(),
            std::ptr::null_mut(), snaps.as_mut_ptr(), fin.as_mut_ptr(), std::ptr::null_mut(), status.as_mut_ptr()) };
        if rc != ffi::CHROM_OK { return Err(ctx.last_error()); }
        (nz, 1usize, snaps, fin)
    } else if let Some(b) = model.get("LangmuirMulti") {
        let nz = b["n_points"].as_u64().ok_or("n_points")? as usize;
        let sp = b["species"].as_array().ok_or("species")?;
        let n = sp.len();
        cfg.n_points = nz as u32;
        cfg.n_species = n as u32;
        let mut tables = Vec::new();
        let mut species = Vec::new();
        for (k, s) in sp.iter().enumerate() {
            let inj: TemporalInjection = serde_json::from_value(s["injection"].clone()).map_err(|e| e.to_string())?;
            tables.extend(tabulate(&inj, rk4, dt, time_steps));
            species.push(ffi::chrom_species { lambda: f(s, "lambda")?, langmuir_k: f(s, "langmuir_k")?,
                port_number: s["port_number"].as_u64().ok_or("port_number")? as u32, inj_table: k as u32 });
        }
        let col = ffi::chrom_multi_col { fe: f(b, "fe")?, ue: f(b, "ue")?, dz: f(b, "dz")?,
            stationary_fraction: f(b, "stationary_fraction")?, species_off: 0, detector_cell: 0 };
        // nalgebra DMatrix is column-major nz x n == the device layout [species][nz]
        let y0: Vec<f64> = initial.get(PhysicalQuantity::Concentration).ok_or("Concentration is required")?
            .as_matrix().as_slice().to_vec();
        let mut snaps = vec![0.0; n_snap * nz * n];
        let mut fin = vec![0.0; nz * n];
        let rc = unsafe { ffi::chrom_cuda_solve_multi_batch(ctx.0, &cfg, &col, 1, species.as_ptr(), n as u64,
            tables.as_ptr(), n as u32, y0.as_ptr(), std::ptr::null_mut(), snaps.as_mut_ptr(), fin.as_mut_ptr(),
            std::ptr::null_mut(), status.as_mut_ptr()) };
        if rc != ffi::CHROM_OK { return Err(ctx.last_error()); }
        (nz, n, snaps, fin)
    } else {
        return Err("model is not supported by the CUDA backend (LangmuirSingle and LangmuirMulti only; no CPU fallback)".into());
    };

    if status[0] != 0 {                                              // validate_state: rk4.rs:331
        let mut buf = vec![0u8; 512];
        unsafe { ffi::chrom_cuda_status_message(status[0], buf.as_mut_ptr() as *mut _, buf.len()) };
        return Err(CStr::from_bytes_until_nul(&buf).unwrap().to_string_lossy().into());
    }

    let to_state = |s: &[f64]| -> PhysicalState {
        if n == 1 && model.get("LangmuirSingle").is_some() {
            PhysicalState::new(PhysicalQuantity::Concentration, PhysicalData::Vector(DVector::from_column_slice(s)))
        } else {
            PhysicalState::new(PhysicalQuantity::Concentration,
                               PhysicalData::from_matrix(DMatrix::from_column_slice(nz, n, s)))
        }
    };
    let mut time_points = vec![0.0; time_steps + 1];
    unsafe { ffi::chrom_cuda_time_points(total_time, time_steps as u64, time_points.as_mut_ptr()) };
    let trajectory: Vec<PhysicalState> = snapshots.chunks(nz * n).map(to_state).collect();
    let mut result = SimulationResult::new(time_points, trajectory, to_state(&final_state));
    if rk4 {                                                          // rk4.rs:343-347
        result.add_metadata("solver", "Runge-Kutta 4");
        result.add_metadata("time steps", time_steps.to_string());
        result.add_metadata("dt", dt.to_string());
        result.add_metadata("total time", total_time.to_string());
        result.add_metadata("function evaluations", (4 * time_steps).to_string());
    } else {                                                          // euler.rs:282-285
        result.add_metadata("solver", "Forward Euler");
        result.add_metadata("time steps", time_steps.to_string());
        result.add_metadata("dt", dt.to_string());
        result.add_metadata("total time", total_time.to_string());
    }
    Ok(result)
}

pub struct CudaRk4Solver { ctx: std::sync::Mutex<CudaContext> }
pub struct CudaEulerSolver { ctx: std::sync::Mutex<CudaContext> }

impl CudaRk4Solver {
    pub fn new() -> Result<Self, String> { Ok(Self { ctx: std::sync::Mutex::new(CudaContext::new(&[0])?) }) }
}
impl CudaEulerSolver {
    pub fn new() -> Result<Self, String> { Ok(Self { ctx: std::sync::Mutex::new(CudaContext::new(&[0])?) }) }
}

impl Solver for CudaRk4Solver {
    fn solve(&self, scenario: &Scenario, config: &SolverConfiguration) -> Result<SimulationResult, String> {
        solve(&self.ctx.lock().unwrap(), true, scenario, config)
    }
    fn name(&self) -> &str { "Runge Kutta (RK4)" }
}

impl Solver for CudaEulerSolver {
    fn solve(&self, scenario: &Scenario, config: &SolverConfiguration) -> Result<SimulationResult, String> {
        solve(&self.ctx.lock().unwrap(), false, scenario, config)
    }
    fn name(&self) -> &str { "Forward Euler" }
}
