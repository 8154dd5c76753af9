// kernels.h — launch descriptors shared by the C-ABI layer (api.cu) and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/chrom_cuda.h"
#include "sampling.cuh"

namespace chrom {

// Kernel-internal arithmetic selector carried in BatchDev::arith: the two public modes plus the
// polynomial FAST variant the host may substitute for CHROM_ARITH_FAST on single-species batches.
#define CHROM_KARITH_POLY 2
#define CHROM_KARITH_STRUCT 3  /* multi-species: structured (diag + rank-1) LU, see multi_resident.cu */

// Everything one device shard of a batch needs; all pointers are device pointers.
struct BatchDev {
  // inputs
  const chrom_single_col* scols;  // single-species columns (or null)
  const chrom_multi_col* mcols;   // multi-species columns (or null)
  const chrom_species* species;   // species table, indices already rebased to this shard
  const double* tables;           // [n_tables][steps][stages]
  const double* y0;               // initial state or null (= zeros)
  // outputs (any may be null)
  double* outlet;
  double* snapshots;
  double* final_state;
  double* summary;
  long long* status;
  // streaming path scratch: two state buffers [n_cols][nz]
  double* ping;
  double* pong;
  double* scratch;  // multi-species warp-per-cell kernels: a dense-matrix block per warp (multi_tiled_scratch_doubles)
  // shape
  uint32_t n_cols;
  uint32_t nz;
  uint32_t n_species;
  uint32_t steps;
  uint32_t step_begin, step_end;  // the launch advances steps [step_begin, step_end) (register-resident single-species
                                  // kernels only: a one-shot solve time-slices its last chunk; y0 = the carried state)
  uint32_t stages;  // table entries per step
  uint32_t outlet_stride, snapshot_stride;
  const uint32_t* out_steps;   // explicit outlet sample list (device, kNoSample-terminated) or null = the stride
  const uint32_t* snap_steps;  // explicit snapshot list or null
  const uint32_t* h_out_steps;   // host copies of the two lists (the streaming path schedules samples on the host)
  const uint32_t* h_snap_steps;
  double* series;  // streaming path with a summary: full-rate detector series [n_cols][steps + 1] (device scratch)
  uint32_t n_out, n_snap;
  int32_t method, arith;
  double dt, half_dt, sixth_dt;  // dt, dt/2.0, dt/6.0 evaluated on the host in IEEE double
};

// Per-launch arguments of the tiled LangmuirMulti kernels (multi_tiled.cuh).
struct TiledArgs {
  const double* src;          // [n_cols][n][nz] state at step0
  double* dst;                // [n_cols][n][nz] state at step0 + nsub
  unsigned long long* code;   // [n_cols] ~0 = clean, else (step << 1) | (NaN ? 0 : 1), merged with atomicMin
  double* gsumm;              // [n_cols][n][5] running summary (area, moment, max, tmax, previous sample)
  uint32_t step0, nsub;       // this launch advances steps step0 .. step0 + nsub - 1
  uint32_t tiles_per_col, tw, halo, pitch;  // tile width (cells), upstream halo, shared-memory row pitch (odd)
};

// How a kernel family maps a batch onto the machine; filled by the *_choose functions.
struct LaunchGeom {
  int kind = 0;           // 1 single_resident, 2 single_resident_mw, 3 single_stream, 4 multi_resident, 5 multi_tiled,
                          // 6 multi_reg (register-resident multi-species)
  int M = 0;              // cells per lane
  int L = 0;              // lanes per column (resident) / threads per column
  int cpw = 0;            // columns per warp (resident, single warp)
  int warps_per_col = 0;  // multi-warp resident
  dim3 grid, block;
  size_t smem = 0;
  uint64_t launches_per_execute = 1;
  uint32_t cols_per_wave = 0;  // columns one full wave of CTAs covers (0: the path is not chunkable)
  // single_resident, batch that cannot fill the machine: the columns beyond `cols_first` run as a SECOND launch with
  // its own cells per lane on a second stream, so that every SM sub-partition carries one warp of each (M2 = 0: off)
  int M2 = 0, L2 = 0, cpw2 = 0;
  uint32_t cols_first = 0;
  // multi_tiled: tile width, halo, shared-memory pitch, steps per launch, species per lane (0 = thread per cell)
  uint32_t tw = 0, halo = 0, pitch = 0, nsub = 0;
  int wj = 0;
  std::string name;
};

// single_resident.cu
bool single_resident_choose(const BatchDev& b, int sm_count, LaunchGeom* g, std::string* why);
cudaError_t single_resident_launch(const BatchDev& b, const LaunchGeom& g, cudaStream_t s);

// single_stream.cu — nz beyond the resident capacity: state streamed through HBM/L2, one launch per step
bool single_stream_choose(const BatchDev& b, int sm_count, LaunchGeom* g, std::string* why);
cudaError_t single_stream_launch(const BatchDev& b, const LaunchGeom& g, cudaStream_t s);
size_t single_stream_extra_doubles();  // per column, behind the pong state buffer: status code + coefficients

// multi_resident.cu — LangmuirMulti, per-cell n x n solve
bool multi_resident_choose(const BatchDev& b, int sm_count, LaunchGeom* g, std::string* why);
cudaError_t multi_resident_launch(const BatchDev& b, const LaunchGeom& g, cudaStream_t s);
int multi_kernel_arith(int n_species, int arith);

// multi_reg.cu — LangmuirMulti n <= 16 with the state in registers (FAST arithmetic)
bool multi_reg_choose(const BatchDev& b, int sm_count, LaunchGeom* g, std::string* why);
cudaError_t multi_reg_launch(const BatchDev& b, const LaunchGeom& g, cudaStream_t s);
cudaError_t multi_reg2_launch(const BatchDev& b, const LaunchGeom& g, cudaStream_t s);  // kind 7 (multi_reg2.cuh)

// multi_tiled.cu — LangmuirMulti beyond the resident capacity (long columns, n > 64): tiles through HBM/L2
bool multi_tiled_choose(const BatchDev& b, int sm_count, LaunchGeom* g, std::string* why);
cudaError_t multi_tiled_launch(const BatchDev& b, const LaunchGeom& g, cudaStream_t s);
size_t multi_tiled_extra_doubles(uint32_t n_species);  // per column, behind the pong buffer: status code + summary
size_t multi_tiled_scratch_doubles(const BatchDev& b, const LaunchGeom& g);

// analytics.cu — surface resolution Rsf between two device-resident outlet series on one time grid
cudaError_t rsf_launch(const double* a, const double* b, size_t n_series, uint32_t n_out, const uint32_t* steps_list,
                       uint32_t stride, double dt, double* out, cudaStream_t s);

// probe.cu
cudaError_t measure_fp64_peak(int sm_count, cudaStream_t s, double* tflops, double* ms);
cudaError_t measure_copy_bw(size_t bytes, cudaStream_t s, double* gbs);
cudaError_t probe_rcp(double lo, double hi, uint32_t n, cudaStream_t s, double* max_rel_err);

}  // namespace chrom
