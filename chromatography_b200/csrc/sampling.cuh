// sampling.cuh — what every kernel family shares about OUTPUTS: which steps are sampled (a stride or an explicit
// sorted list: the reference's sample_indices, physics/traits.rs:1010-1023), at which cell the detector sits
// (Detector::node_index, domain/detector.rs:285), and the running chromatogram summary (trapezoid area, first
// moment, maximum, second moment, minimum: validation/dissimilarity.rs:81-88, validation/main.rs:190-206).
#pragma once
#include <stdint.h>

#include "../../include/chrom_cuda.h"

namespace chrom {

constexpr uint32_t kNoSample = 0xffffffffu;
constexpr int kSummSlots = 7;  // running state per (column, species): see Summ

// Position in a sample schedule: sample number k (= its slot in the output buffer) is due after `next` completed
// steps.  list == nullptr: slot k is step k * stride.  Lists are device arrays with a kNoSample sentinel behind the
// last entry.
struct SampleCursor {
  uint32_t k, next;
};

__host__ __device__ inline bool sample_has_initial(const uint32_t* list, uint32_t stride) {
  return list ? list[0] == 0u : stride != 0u;
}
// The first sample strictly after step s0 (s0 == 0: after the initial state; s0 > 0: a launch that continues a
// solve — a time slice, a tile launch).  n = entries of the list.
__host__ __device__ inline SampleCursor sample_begin(const uint32_t* list, uint32_t n, uint32_t stride, uint32_t s0) {
  SampleCursor c;
  if (list) {
    uint32_t lo = 0u, hi = n;  // first k with list[k] > s0
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (list[mid] > s0) hi = mid;
      else lo = mid + 1u;
    }
    c.k = lo;
    c.next = list[lo];  // the sentinel when lo == n
  } else if (stride) {
    c.k = s0 / stride + 1u;
    c.next = c.k * stride;
  } else {
    c.k = 0u;
    c.next = kNoSample;
  }
  return c;
}
__host__ __device__ inline void sample_advance(SampleCursor& c, const uint32_t* list, uint32_t stride) {
  ++c.k;
  c.next = list ? list[c.k] : c.next + stride;
}

// The detector cell of a column: 0 = the outlet (last cell), k > 0 = cell k - 1 (validated < nz by the host).
__host__ __device__ inline uint32_t detector_of(uint32_t detector_cell, uint32_t nz) {
  return detector_cell ? detector_cell - 1u : nz - 1u;
}

// Running summary of one detector signal.  Raw sums while the solve runs (so that a time-sliced solve can park them
// in the summary buffer between slices), finished by summ_store.
struct Summ {
  double area, m1, vmax, tmax, prev, m2, vmin;
};
__host__ __device__ inline void summ_init(Summ& s, double v0) {
  s.area = s.m1 = s.m2 = 0.0;
  s.vmax = s.vmin = s.prev = v0;
  s.tmax = 0.0;
}
// sample v after step `step` (time (step + 1) dt; the previous sample sits at step dt): one trapezoid of each integral
__host__ __device__ inline void summ_step(Summ& s, double v, uint32_t step, double dt) {
  const double t0 = (double)step * dt, t1 = ((double)step + 1.0) * dt;
  const double w = 0.5 * (t1 - t0);
  s.area = fma(w, s.prev + v, s.area);
  s.m1 = fma(w, fma(t0, s.prev, t1 * v), s.m1);
  s.m2 = fma(w, fma(t0 * t0, s.prev, (t1 * t1) * v), s.m2);
  if (v > s.vmax) {
    s.vmax = v;
    s.tmax = t1;
  }
  if (v < s.vmin) s.vmin = v;
  s.prev = v;
}
// o[CHROM_SUMMARY_LEN]: area, first moment, max, time of max, sigma, min.  The second CENTRAL moment of the
// reference (trapezoid((t - mean)^2 c) / area) is linear in the raw trapezoid sums: m2/area - mean^2.
__host__ __device__ inline void summ_store(double* o, const Summ& s) {
  const double mean = (s.area != 0.0) ? s.m1 / s.area : 0.0;
  const double var = (s.area != 0.0) ? s.m2 / s.area - mean * mean : 0.0;
  o[0] = s.area;
  o[1] = mean;
  o[2] = s.vmax;
  o[3] = s.tmax;
  o[4] = var > 0.0 ? sqrt(var) : 0.0;
  o[5] = s.vmin;
}
// park / resume the raw sums in a summary slot (time-sliced solves); prev is recomputed from the carried state
__host__ __device__ inline void summ_park(double* o, const Summ& s) {
  o[0] = s.area;
  o[1] = s.m1;
  o[2] = s.vmax;
  o[3] = s.tmax;
  o[4] = s.m2;
  o[5] = s.vmin;
}
__host__ __device__ inline void summ_resume(Summ& s, const double* o, double prev) {
  s.area = o[0];
  s.m1 = o[1];
  s.vmax = o[2];
  s.tmax = o[3];
  s.m2 = o[4];
  s.vmin = o[5];
  s.prev = prev;
}

}  // namespace chrom
