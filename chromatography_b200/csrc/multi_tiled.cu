// multi_tiled.cu — host side of the tiled LangmuirMulti integrator (device code: multi_tiled.cuh):
// tile geometry, kernel lookup, the per-launch loop.  Covers what the shared-memory-resident path
// (multi_resident.cu) cannot: 4 * n * nz doubles beyond one CTA's shared memory, and n > 64.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "kernels.h"

namespace chrom {

typedef void (*tkern_t)(const BatchDev, double*, const TiledArgs);
tkern_t multi_tiled_pick_0(int method, int arith, int wj);
tkern_t multi_tiled_pick_1(int method, int arith, int wj);
tkern_t multi_tiled_pick_2(int method, int arith, int wj);
tkern_t multi_tiled_pick_3(int method, int arith, int wj);
tkern_t multi_tiled_pick_4(int method, int arith, int wj);
tkern_t multi_tiled_pick_5(int method, int arith, int wj);
tkern_t multi_tiled_pick_6(int method, int arith, int wj);
tkern_t multi_tiled_pick_7(int method, int arith, int wj);
tkern_t multi_tiled_pick_8(int method, int arith, int wj);
tkern_t multi_tiled_pick_9(int method, int arith, int wj);
tkern_t multi_tiled_pick_10(int method, int arith, int wj);
tkern_t multi_tiled_pick_11(int method, int arith, int wj);
tkern_t multi_tiled_pick_12(int method, int arith, int wj);
tkern_t multi_tiled_pick_13(int method, int arith, int wj);
tkern_t multi_tiled_pick_14(int method, int arith, int wj);
tkern_t multi_tiled_pick_15(int method, int arith, int wj);
tkern_t multi_tiled_pick_16(int method, int arith, int wj);
void multi_tiled_init_launch(const BatchDev& b, double* ping, unsigned long long* code, cudaStream_t s);
void multi_tiled_finalize_launch(const BatchDev& b, const unsigned long long* code, const double* gsumm, cudaStream_t s);

namespace {

constexpr int kWarpCellNMax = 256;   // warp per cell: 8 species per lane
constexpr size_t kSmemMax = 227u * 1024u;

tkern_t lookup_tiled(int n, int method, int arith, int wj) {
  if (wj > 0) return multi_tiled_pick_0(method, arith, wj);
  switch (n) {
    case 1: return multi_tiled_pick_1(method, arith, 0);
    case 2: return multi_tiled_pick_2(method, arith, 0);
    case 3: return multi_tiled_pick_3(method, arith, 0);
    case 4: return multi_tiled_pick_4(method, arith, 0);
    case 5: return multi_tiled_pick_5(method, arith, 0);
    case 6: return multi_tiled_pick_6(method, arith, 0);
    case 7: return multi_tiled_pick_7(method, arith, 0);
    case 8: return multi_tiled_pick_8(method, arith, 0);
    case 9: return multi_tiled_pick_9(method, arith, 0);
    case 10: return multi_tiled_pick_10(method, arith, 0);
    case 11: return multi_tiled_pick_11(method, arith, 0);
    case 12: return multi_tiled_pick_12(method, arith, 0);
    case 13: return multi_tiled_pick_13(method, arith, 0);
    case 14: return multi_tiled_pick_14(method, arith, 0);
    case 15: return multi_tiled_pick_15(method, arith, 0);
    case 16: return multi_tiled_pick_16(method, arith, 0);
  }
  return nullptr;
}

// shared memory of one tile: 4 state arrays [n][pitch], inlet [3][n], summary [n][kSummSlots], SpeciesConst [n], table pointers [n]
size_t tiled_smem_bytes(int n, uint32_t pitch) {
  return ((size_t)4 * n * pitch + (3 + kSummSlots) * (size_t)n) * sizeof(double) + (size_t)n * 10 * sizeof(double) + (size_t)n * sizeof(void*);
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

}  // namespace

size_t multi_tiled_extra_doubles(uint32_t n_species) { return 1 + (size_t)kSummSlots * n_species; }

size_t multi_tiled_scratch_doubles(const BatchDev& b, const LaunchGeom& g) {
  const size_t n = b.n_species;
  if (g.wj > 0) {  // a block per warp: FAST n*n + n (dense redo), EXACT 2*n*n + n
    const size_t per_warp = (b.arith == CHROM_ARITH_EXACT ? 2 : 1) * n * n + n;
    return per_warp * g.grid.x * (g.block.x / 32);
  }
  return 0;  // thread per cell: registers only
}

// Environment knobs (experiments and tests): CHROM_TILED_WARP=1 warp per cell also for n = 5..8,
// CHROM_TILED_TW=<cells> caps the tile width, CHROM_TILED_FUSE=<steps> sets the steps per launch.
bool multi_tiled_choose(const BatchDev& b, int sm_count, LaunchGeom* g, std::string* why) {
  const int n = (int)b.n_species;
  if (n < 1 || n > kWarpCellNMax) {
    if (why) *why = "n_species outside 1..256";
    return false;
  }
  const bool rk4 = b.method == CHROM_RK4;
  int wj = 0;
  // Thread per cell: n <= 8 in every mode (n <= 4 always: nalgebra inverts those with closed forms, which are
  // per-thread code), n = 9..16 for the closed-form FAST elimination (4 n doubles per thread; measured 4x the
  // warp-per-cell rate).  Everything else one warp per cell.
  const bool thread_cell = n <= 8 ? !(n >= 5 && env_int("CHROM_TILED_WARP", 0) != 0)
                                  : (n <= 16 && b.arith == CHROM_KARITH_STRUCT && env_int("CHROM_TILED_WARP", 0) == 0);
  if (!thread_cell) wj = n <= 32 ? 1 : (n <= 64 ? 2 : (n <= 128 ? 4 : 8));
  const int karith = b.arith;
  int T;
  if (wj > 0) T = (karith != CHROM_ARITH_EXACT && wj <= 2) ? 512 : 256;
  else T = (n <= 8) ? ((karith == CHROM_KARITH_STRUCT || n <= 2) ? 512 : 256) : 256;  // multi_max_threads
  // widest tile that fits shared memory
  uint32_t tw_max = 0;
  const uint32_t tw_min = std::min<uint32_t>(8u, b.nz);  // columns shorter than 8 cells are one (narrow) tile
  for (uint32_t tw = tw_min; tw <= b.nz + 1 && tw <= 32u * (uint32_t)T; ++tw) {
    if (tiled_smem_bytes(n, tw | 1u) > kSmemMax) break;
    tw_max = tw;
  }
  const int cap = env_int("CHROM_TILED_TW", 0);
  if (cap >= 8) tw_max = std::min<uint32_t>(tw_max, (uint32_t)cap);
  if (tw_max < tw_min) {
    if (why) *why = "n_species too large for a shared-memory tile";
    return false;
  }
  uint32_t tw, tiles, nsub, halo;
  if (b.nz <= tw_max) {  // the whole column is one tile: every time step in a single launch
    tw = b.nz;
    tiles = 1;
    nsub = b.steps;
    halo = 0;
  } else {
    const uint32_t per_step = rk4 ? 4u : 1u;
    int fuse = env_int("CHROM_TILED_FUSE", 0);
    if (fuse <= 0) {
      // Steps per launch.  A batch that fills the machine amortises the tile's trip through HBM/L2 and the launch
      // over as many steps as a halo of ~20 % of the tile allows (measured on n = 100, tile 52..56: 1 step 7.4e8,
      // 2 steps 9.0e8, 3 steps 9.6e8 cell-stage/s); a batch that cannot fill it wants narrow tiles instead
      // (more CTAs per column), i.e. a short halo.
      const bool fills = (uint64_t)b.n_cols * ((b.nz + tw_max - 1) / tw_max) >= (uint64_t)sm_count;
      fuse = (int)std::max<uint32_t>(1u, std::min<uint32_t>(8u, tw_max / (5u * per_step)));
      if (!fills) fuse = std::min(fuse, (int)std::max<uint32_t>(1u, 4u / per_step));
    }
    nsub = std::min<uint32_t>((uint32_t)fuse, b.steps);
    while (nsub > 1 && nsub * per_step * 2 > tw_max) --nsub;  // a tile must own at least half of its cells
    halo = nsub * per_step;
    if (halo >= tw_max) {
      if (why) *why = "tile narrower than its halo";
      return false;
    }
    // fewest tiles at the widest tile, then the narrowest tile that still covers the column with them
    const uint32_t V = tw_max - halo;
    tiles = 1 + (b.nz - tw_max + V - 1) / V;
    // A batch that cannot fill the machine is latency-bound: narrower tiles put more CTAs (warp per cell: fewer
    // cells per warp and stage) on a column; a tile still owns at least max(8, 2 * halo) cells.
    if (wj > 0 && (uint64_t)b.n_cols * tiles < (uint64_t)sm_count) {
      const uint32_t own_min = std::max<uint32_t>(8u, 2u * halo);
      const uint32_t by_machine = (uint32_t)sm_count / std::max<uint32_t>(1u, b.n_cols);
      const uint32_t by_column = std::max<uint32_t>(1u, (b.nz > halo ? b.nz - halo : 1u) / own_min);
      tiles = std::max(tiles, std::min(by_machine, by_column));
    }
    tw = (b.nz + (tiles - 1) * halo + tiles - 1) / tiles;
    tw = std::min(std::max(tw, halo + 1), tw_max);
    // the tile count that exactly covers the column at this width (its last tile owns at least one cell)
    tiles = b.nz > tw ? 1 + (b.nz - tw + (tw - halo) - 1) / (tw - halo) : 1;
  }
  if (wj == 0) {  // threads: a multiple of 32 that keeps the cells per thread
    const uint32_t cpt = (tw + T - 1) / T;
    T = (int)(((tw + cpt - 1) / cpt + 31u) / 32u * 32u);
  } else {
    const uint32_t tmax = (karith != CHROM_ARITH_EXACT && wj <= 2) ? 512u : 256u;  // tiled_warp_threads (multi_tiled.cuh)
    const uint32_t cells_per_pass = (karith != CHROM_ARITH_EXACT) ? (wj == 1 ? 4u : (wj == 2 ? 2u : 1u)) : 1u;
    T = (int)std::min<uint32_t>(tmax, (tw + cells_per_pass - 1) / cells_per_pass * 32u);
    T = std::max(32, T / 32 * 32);
  }
  const uint32_t pitch = tw | 1u;
  const size_t smem = tiled_smem_bytes(n, pitch);
  const uint32_t items = b.n_cols * tiles;
  int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(kSmemMax / (smem + 1024), (size_t)(65536 / (T * 255))));
  ctas_per_sm = std::max(1, ctas_per_sm);
  g->kind = 5;
  g->M = wj > 0 ? (int)((tw + T / 32 - 1) / (T / 32)) : (int)((tw + T - 1) / T);
  g->L = T;
  g->cpw = (int)nsub;
  g->warps_per_col = (int)tiles;
  g->block = dim3(T);
  g->grid = dim3(std::min<uint32_t>(items, (uint32_t)(sm_count * ctas_per_sm)));
  g->smem = smem;
  g->cols_per_wave = 0;
  g->tw = tw;
  g->halo = halo;
  g->pitch = pitch;
  g->nsub = nsub;
  g->wj = wj;
  g->launches_per_execute = ((uint64_t)b.steps + nsub - 1) / nsub + 2;
  char buf[280];
  snprintf(buf, sizeof buf,
           "multi_tiled<n=%d,%s,%s,%s> tiles/col=%u tile=%u halo=%u threads=%d smem=%zuB grid=%u, %u step(s) per launch%s", n,
           rk4 ? "RK4" : "Euler",
           karith == CHROM_ARITH_EXACT ? "exact" : (karith == CHROM_KARITH_STRUCT ? "fast-structured-LU" : "fast-dense-LU"),
           wj > 0 ? "warp/cell" : "thread/cell", tiles, tw, halo, T, smem, g->grid.x, nsub,
           tiles > 1 ? ", CUDA graph" : "");
  g->name = buf;
  return lookup_tiled(n, b.method, karith, wj) != nullptr;
}

// b.ping: [n_cols][n][nz]; b.pong: the same followed by n_cols status codes and n_cols*n*5 summary doubles.
cudaError_t multi_tiled_launch(const BatchDev& b, const LaunchGeom& g, cudaStream_t s) {
  tkern_t k = lookup_tiled((int)b.n_species, b.method, b.arith, g.wj);
  if (!k) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return e;
  const size_t plane = (size_t)b.n_cols * b.n_species * b.nz;
  unsigned long long* code = reinterpret_cast<unsigned long long*>(b.pong + plane);
  double* gsumm = b.pong + plane + b.n_cols;
  multi_tiled_init_launch(b, b.ping, code, s);
  TiledArgs a{};
  a.src = b.ping;
  a.dst = b.pong;
  a.code = code;
  a.gsumm = gsumm;
  a.tiles_per_col = (uint32_t)g.warps_per_col;
  a.tw = g.tw;
  a.halo = g.halo;
  a.pitch = g.pitch;
  for (uint32_t step = 0; step < b.steps; step += g.nsub) {
    a.step0 = step;
    a.nsub = std::min<uint32_t>(g.nsub, b.steps - step);
    k<<<g.grid, g.block, g.smem, s>>>(b, b.scratch, a);
    const double* t = a.src;
    a.src = a.dst;
    a.dst = const_cast<double*>(t);
  }
  if (b.final_state) {
    e = cudaMemcpyAsync(b.final_state, a.src, plane * sizeof(double), cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return e;
  }
  multi_tiled_finalize_launch(b, code, gsumm, s);
  return cudaGetLastError();
}

}  // namespace chrom
