// world.cuh — device-side point tests against the two collision worlds.
// The worlds stand in for the user's ompl::base::StateValidityChecker::isValid
// (called at /root/reference/src/BatchingPOMP.cpp:518, 654, 777, 784).
#pragma once
#include "common.cuh"

#ifdef __CUDACC__

// floor(v) as an int for |v| < 2^31 without a conversion instruction: adding 2^52 + 2^51 with round-down leaves
// floor(v) in the low word of the sum (FP64 -> int conversions are slow; this is one FP64 add).
__device__ __forceinline__ int bp_floor_int(double v) { return __double2loint(__dadd_rd(v, 6755399441055744.0)); }

// 2-D occupancy image (README.md:56 convention; SURVEY.md §8c definition).
__device__ __forceinline__ bool bp_valid_image(double x0, double x1, const ImageWorldView& im) {
  if (!(x0 >= 0.0 && x0 <= 1.0 && x1 >= 0.0 && x1 <= 1.0)) return false;
  int col = bp_floor_int(__dmul_rn(x0, (double)im.W));  // x in [0,1]: the product is far below 2^31
  int row = bp_floor_int(__dmul_rn(x1, (double)im.H));
  col = min(col, im.W - 1);
  row = min(row, im.H - 1);
  unsigned long long tile = __ldg(im.bits + (size_t)(row >> 3) * im.tiles_x + (col >> 3));
  return (tile >> (((row & 7) << 3) | (col & 7))) & 1ull;
}

// Closed-box containment of x in box b; lohi row = [stride] double2 (lo, hi).
template <int D>
__device__ __forceinline__ bool bp_inside_box(const double* x, const double2* lohi_row) {
#pragma unroll
  for (int j = 0; j < D; ++j) {
    double2 lh = lohi_row[j];
    if (!(lh.x <= x[j] && x[j] <= lh.y)) return false;
  }
  return true;
}

// Same test with the loads of a group of dimensions issued together (one shared-memory round trip per
// group instead of one per dimension): dims [0,D1) first, the rest only if those all pass.  Measured on
// the headline config: D1 = 2 / 3 / 4 / 7 -> 1.048 / 1.042 / 1.041 / 1.016 ms, so all loads go first.
template <int D>
__device__ __forceinline__ bool bp_inside_box_grouped(const double* x, const double2* lohi_row) {
#ifndef BP_D1
#define BP_D1 7
#endif
  constexpr int D1 = D < BP_D1 ? D : BP_D1;
  double2 lh[D];
#pragma unroll
  for (int j = 0; j < D1; ++j) lh[j] = lohi_row[j];
  bool in = true;
#pragma unroll
  for (int j = 0; j < D1; ++j) in = in && (lh[j].x <= x[j]) && (x[j] <= lh[j].y);
  if (!in) return false;
  if (D > D1) {
#pragma unroll
    for (int j = D1; j < D; ++j) lh[j] = lohi_row[j];
#pragma unroll
    for (int j = D1; j < D; ++j) in = in && (lh[j].x <= x[j]) && (x[j] <= lh[j].y);
  }
  return in;
}

// Point query through the range masks (cell range [c,c]).  rmask/lohi may point to shared or global.
template <int D>
__device__ __forceinline__ bool bp_valid_boxes(const double* x, const BoxWorldView& w, const uint4* rmask,
                                               const double2* lohi) {
  if (w.nboxes == 0) return true;
  int cell[D];
#pragma unroll
  for (int j = 0; j < D; ++j) cell[j] = bp_cell(x[j], w.g0[j], w.gs[j]);
  for (int g = 0; g < w.w4; ++g) {
    uint4 m = make_uint4(~0u, ~0u, ~0u, ~0u);
#pragma unroll
    for (int j = 0; j < D; ++j) {
      uint4 v = rmask[bp_midx(bp_rrow(j, cell[j], 0), g, w.rows)];
      m.x &= v.x; m.y &= v.y; m.z &= v.z; m.w &= v.w;
    }
    unsigned words[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      unsigned word = words[k];
      while (word) {
        int b = (g * 4 + k) * 32 + (__ffs(word) - 1);
        word &= word - 1;
        if (bp_inside_box<D>(x, lohi + (size_t)b * w.lohi_stride)) return false;
      }
    }
  }
  return true;
}

#endif
