//! Raw declarations of include/chrom_cuda.h (same names, same order, `#[repr(C)]` structs).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

pub const CHROM_OK: i32 = 0;
pub const CHROM_EULER: i32 = 0;
pub const CHROM_RK4: i32 = 1;
pub const CHROM_ARITH_FAST: i32 = 0;
pub const CHROM_ARITH_EXACT: i32 = 1;
pub const CHROM_ARITH_FAST_STRUCTURED: i32 = 3; // LangmuirMulti: the explicit name of FAST
pub const CHROM_ARITH_FAST_DENSE: i32 = 4; // LangmuirMulti n <= 8: dense register LU on every cell

#[repr(C)]
pub struct chrom_cuda_ctx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct chrom_single_col {
    pub lambda: f64,
    pub langmuir_k: f64,
    pub port_number: f64,
    pub fe: f64,
    pub ue: f64,
    pub dz: f64,
    pub inj_table: u32,
    pub _reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct chrom_multi_col {
    pub fe: f64,
    pub ue: f64,
    pub dz: f64,
    pub stationary_fraction: f64,
    pub species_off: u32,
    pub _reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct chrom_species {
    pub lambda: f64,
    pub langmuir_k: f64,
    pub port_number: u32,
    pub inj_table: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct chrom_run_cfg {
    pub method: i32,
    pub arith: i32,
    pub total_time: f64,
    pub time_steps: u64,
    pub n_points: u32,
    pub n_species: u32,
    pub outlet_stride: u64,
    pub snapshot_stride: u64,
}

unsafe extern "C" {
    pub fn chrom_cuda_create(device_ids: *const i32, n_devices: i32, out: *mut *mut chrom_cuda_ctx) -> i32;
    pub fn chrom_cuda_destroy(ctx: *mut chrom_cuda_ctx);
    pub fn chrom_cuda_last_error(ctx: *const chrom_cuda_ctx) -> *const c_char;
    pub fn chrom_cuda_n_samples(time_steps: u64, stride: u64) -> u64;
    pub fn chrom_cuda_time_points(total_time: f64, time_steps: u64, out: *mut f64);
    pub fn chrom_cuda_table_stages(method: i32) -> u32;
    pub fn chrom_cuda_status_message(status: i64, buf: *mut c_char, buf_len: usize) -> usize;
    pub fn chrom_cuda_host_alloc(bytes: usize, out: *mut *mut c_void) -> i32;
    pub fn chrom_cuda_host_free(p: *mut c_void);
    pub fn chrom_cuda_bind_host_thread(
        ctx: *mut chrom_cuda_ctx, device_index: i32, numa_node: *mut i32, n_cpus: *mut i32,
    ) -> i32;
    pub fn chrom_cuda_solve_single_batch(
        ctx: *mut chrom_cuda_ctx, cfg: *const chrom_run_cfg, cols: *const chrom_single_col, n_cols: u64,
        inj_tables: *const f64, n_tables: u32, y0: *const f64, outlet: *mut f64, snapshots: *mut f64,
        final_state: *mut f64, summary: *mut f64, status: *mut i64,
    ) -> i32;
    pub fn chrom_cuda_solve_multi_batch(
        ctx: *mut chrom_cuda_ctx, cfg: *const chrom_run_cfg, cols: *const chrom_multi_col, n_cols: u64,
        species: *const chrom_species, n_species_total: u64, inj_tables: *const f64, n_tables: u32,
        y0: *const f64, outlet: *mut f64, snapshots: *mut f64, final_state: *mut f64, summary: *mut f64,
        status: *mut i64,
    ) -> i32;
}
