// fastpath.cuh — arithmetic of the fast box-world edge kernel (edge_check_fast.cu), shared between the
// device code, the host-side table builder (world.cu) and the CPU check of its error analysis
// (tests/cpp/fastpath_check.cpp compiles this header with g++).
//
// What the reference computes (src/BatchingPOMP.cpp:435-461, 496-544): for an edge with n states, sample
// k (k = perm_n[slot]) is x_j = from_j + (to_j - from_j) * t_k, t_k = (1+k)/(n+1), in unfused FP64; slots
// 1..n-2 are tested in order against the closed boxes; the first invalid slot blocks the edge.
//
// The fast kernel decides most (edge, box) pairs without FP64: along the edge, "inside box b" holds for a
// contiguous range of k.  A conservative OUTER range (no truly-inside sample lies outside it) and a
// conservative INNER range (every sample in it is truly inside) are computed in FP32 from 16-bit quantised
// box bounds.  The first slot whose k lies in the outer range is the answer if it also lies in the inner
// range; otherwise (a sample within a few 1e-5 of a box face) that one sample is tested exactly in FP64
// with the reference's arithmetic.  Results are therefore bit-identical to the reference's.
//
// Quantised coordinates ("q-units"), per dimension j:   q_j(x) = fma(x, qs_j, qo_j)
// with qs_j = FP_QSCALE / extent_j, qo_j = FP_QOFF - g0_j*qs_j, so the box region maps to [4, 65004].
//   table:  lo_q = floor(q(lo)) - FP_QSLACK,  hi_q = ceil(q(hi)) + FP_QSLACK      (u16, no clamping occurs)
//   hence   lo_q + 3 <= q(lo) < lo_q + 4   and   hi_q - 4 < q(hi) <= hi_q - 3.
// Per edge and dimension (F = q(from_j), T = q(to_j), Dq = T - F, all FP64):
//   (the kernel works with t = s/(n+1): inv = 1/Dq, and scales the final range by n+1, an exact float)
//   moving dimension (|Dq| >= 1):  inv = (n+1)/Dq (FP32, relative error <= 2^-21),
//                                  c = -(2^23 + F)*inv in FP32 (error <= 0.5 ulp = 0.5*|inv|, plus 2^-8*|inv|
//                                  from F entering as a float)
//   s(bound) = fma(2^23 + bound_q, inv, c)   — 2^23 + bound_q is the exact float built from the u16 by a
//   byte permute — is the sample coordinate s = k+1 at which the model line F + Dq*s/(n+1) crosses the
//   bound, with an error equivalent to <= 0.65 q-units of position:
//     0.5 (rounding of c) + 2^18 * 2^-21 (relative error of inv on |bound - F| <= 2^18) + FP64 noise.
//   OUTER: a truly-inside sample has position >= q(lo) >= lo_q + 3 > lo_q + 0.65, so s_k >= s(lo_q) (and
//          symmetrically for hi): the range [max_j s_enter, min_j s_exit] contains every inside sample.
//   INNER: s_k >= s(lo_q) + FP_PAD_IN*|inv| puts the model position >= lo_q + 6, the true one >= lo_q + 5.35
//          > q(lo) (and symmetrically for hi), per dimension.
//   static dimension (|Dq| < 1 q-unit, e.g. an axis-aligned edge): the coordinate stays within F +- 1;
//          inv = FP_BIG, c = -(2^23 + F)*FP_BIG: s(bound) = (bound_q - F +- 0.5)*FP_BIG is hugely negative
//          whenever F >= lo_q + 2 (every possibly-inside case, since q(lo) - 1 >= lo_q + 2) and hugely
//          positive for hi symmetrically, so the dimension does not constrain the OUTER range; an edge with a
//          static dimension never uses the INNER shortcut (its hits are verified exactly).
// The 16-bit bounds are only meaningful inside the mapped region, so both endpoints must lie in the safe zone
// q in [FP_QOFF-1, FP_QOFF+FP_QSCALE+1] (bounds beyond it are clamped to 0 / 65535, which stays conservative
// for samples inside the zone).  Edges outside the safe zone, with more than FP_NMAX states, with a cell span beyond the mask
// classes, or with too many candidates are not handled here at all: they are marked BPOMP_EDGE_SLOW and the
// general kernel (edge_check.cu) evaluates them.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FP_HD __host__ __device__ __forceinline__
#else
#define FP_HD inline
#endif

#define FP_QSCALE 65000.0
#define FP_QOFF 4.0
#define FP_QSLACK 3
#define FP_PAD_IN 6.0f
#define FP_BIG 1099511627776.0f /* 2^40 */
#define FP_SAFE 131072.0        /* 2^17: |q| of both endpoints must not exceed this */
#define FP_NMAX 255u            /* largest nStates handled by the fast kernel */
#define FP_CELLS 32             /* broad-phase cells per dimension = q >> 11 */
#define FP_CELL_SHIFT 11
#ifndef FP_CLASSES
#define FP_CLASSES 8            /* span classes: class k < FP_CLASSES-1 -> span k, the last -> spans up to FP_SPAN_MAX */
#endif
#define FP_SPAN_MAX 13
#define FP_ROWS_PER_DIM (FP_CLASSES * FP_CELLS) /* row = class*32 + first cell */
#define FP_MAXBOXES 512         /* 4 groups of 128 boxes, one per lane of a team */
#define BPOMP_EDGE_SLOW 3u      /* internal status: left to the general kernel */

// q_j(x): one fused multiply-add, monotone in x; the same instruction on the host (fma) and the device.
FP_HD double fp_q(double x, double qs, double qo) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(x, qs, qo);
#else
  return fma(x, qs, qo);
#endif
}

// Monotone 16-bit saturation of q (NaN -> 0), then the cell is its top 5 bits.
FP_HD uint32_t fp_q16(double q) {
#if defined(__CUDA_ARCH__)
  unsigned short r;
  asm("cvt.rzi.u16.f64 %0, %1;" : "=h"(r) : "d"(q));  // float-to-integer conversions saturate
  return r;
#else
  if (!(q > 0.0)) return 0u;
  if (q >= 65535.0) return 65535u;
  return (uint32_t)q;
#endif
}
FP_HD int fp_cell(double q) { return (int)(fp_q16(q) >> FP_CELL_SHIFT); }

// Span class of a cell range [c0, c0+span]; FP_CLASSES means "not covered" (slow edge).
FP_HD int fp_class(int span) { return span < FP_CLASSES - 1 ? span : (span <= FP_SPAN_MAX ? FP_CLASSES - 1 : FP_CLASSES); }
FP_HD int fp_class_span(int cls) { return cls < FP_CLASSES - 1 ? cls : FP_SPAN_MAX; }

// Per-edge, per-dimension FP32 data of the slab test.
struct FpDim {
  float inv;  // 1/Dq, or FP_BIG for a static dimension (t units: t = s/(n+1))
  float c;    // -(2^23 + F) * inv
  bool stat;
};

FP_HD FpDim fp_edge_dim(double F, double T) {
  FpDim r;
  const double Dq = T - F;
  const float Dqf = (float)Dq;
  r.stat = fabsf(Dqf) < 1.0f;
  float rc;
#if defined(__CUDA_ARCH__)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(Dqf));  // 1 ulp
#else
  rc = 1.0f / Dqf;
#endif
  r.inv = r.stat ? FP_BIG : rc;
  // c = -(2^23 + F)*inv: the 2^23 part is an exact scaling, F enters as a float (its rounding, <= 2^-8 q-units,
  // is negligible); one rounding at the end, 0.5 ulp(c) = 0.5*|inv|
  r.c = fmaf(-r.inv, (float)F, -8388608.0f * r.inv);
  return r;
}

// nStates of initializeEdgePoints (BP.cpp:440-445) as a function of the SQUARED distance, without the FP64 square
// root and division: thr[k] (3 <= k <= FP_NTHR) is the smallest double s with floor(sqrt(s)/L) >= k in the
// reference's arithmetic (host-built by bisection with IEEE sqrt and division; thr[0..2] = 0), so
// n = max(2, #{k : thr[k] <= s}) exactly.  An FP32 estimate (relative error < 2^-21, i.e. far less than one unit
// for n < FP_NTHR) picks the neighbourhood; two exact comparisons fix it.  Returns 0 when the estimate is out of
// the table's range (the caller then uses the FP64 formula).
#define FP_NTHR 64u
FP_HD uint32_t fp_nstates_thr(double dist2, float inv_L, const double* thr) {
  float est;
#if defined(__CUDA_ARCH__)
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(est) : "f"((float)dist2));
#else
  est = sqrtf((float)dist2);
#endif
  est *= inv_L;
  if (!(est < (float)(FP_NTHR - 2u))) return 0u;  // also NaN
  uint32_t ne = (uint32_t)est;
  if (ne < 2u) ne = 2u;
  const double s0 = thr[ne], s1 = thr[ne + 1u];
  return ne + (dist2 >= s1 ? 1u : 0u) - (dist2 < s0 ? 1u : 0u);
}

// The exact float 2^23 + v for a 16-bit v held in the low (HI = 0) or high (HI = 1) half of w.
template <int HI>
FP_HD float fp_magic(uint32_t w) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(__byte_perm(w, 0x4B000000u, HI ? 0x7432 : 0x7410));
#else
  union { uint32_t u; float f; } x;
  x.u = 0x4B000000u | (HI ? (w >> 16) : (w & 0xFFFFu));
  return x.f;
#endif
}

// From t units to sample coordinates s = t*(n+1), clamped to the samples of the edge (s = 1..n).
FP_HD void fp_range_to_s(uint32_t n, float& smin, float& smax, float& imin, float& imax) {
  const float np1 = (float)(n + 1u);
  smin = fmaxf(smin * np1, 0.5f);
  smax = fminf(smax * np1, (float)n + 0.5f);
  imin = fmaxf(imin * np1, 0.5f);
  imax = fminf(imax * np1, (float)n + 0.5f);
}

// Narrows the OUTER range [smin, smax] and the INNER one [imin, imax] by one dimension's slab.
FP_HD void fp_slab(float lo_m, float hi_m, float inv, float c, float& smin, float& smax, float& imin, float& imax) {
  const float a = fmaf(lo_m, inv, c), b = fmaf(hi_m, inv, c);
  const float enter = fminf(a, b), leave = fmaxf(a, b);
  const float p = FP_PAD_IN * fabsf(inv);
  smin = fmaxf(smin, enter);
  smax = fminf(smax, leave);
  imin = fmaxf(imin, enter + p);
  imax = fminf(imax, leave - p);
}

// Byte offset (group 0) of mask row `row` (= class*32 + first cell) of dimension j in the shared-memory table.
// Dimensions are stored in pairs: one 128-byte line holds row r of dimension 2i (bytes 0..63: four groups of
// 128 boxes, 16 B each) and row r of dimension 2i+1 (bytes 64..127).  Lanes 0-3 of a quarter-warp walk the
// pairs as (2i, 2i+1), lanes 4-7 as (2i+1, 2i), so their eight 16-byte loads cover eight different bank
// groups.  The unpaired last dimension of an odd D is stored compactly (64-byte rows).
FP_HD uint32_t fp_row_offset(int D, int j, uint32_t row) {
  if ((D & 1) && j == D - 1) return (uint32_t)(j >> 1) * (FP_ROWS_PER_DIM * 128u) + row * 64u;
  return (uint32_t)(j >> 1) * (FP_ROWS_PER_DIM * 128u) + row * 128u + (uint32_t)(j & 1) * 64u;
}
FP_HD size_t fp_tmask_bytes(int D) {
  return (size_t)(D / 2) * (FP_ROWS_PER_DIM * 128u) + (size_t)(D & 1) * (FP_ROWS_PER_DIM * 64u);
}

// Shared-memory copy of the sample coordinates of short edges, four slots per 16-byte load: row n
// (3 <= n <= FP_PERM4_NMAX) has FP_PERM4_ROW floats, entry i = float(1 + perm_n[1 + i]) for i < n-2 (the checked
// slots 1..n-2, BP.cpp:510) and -1 beyond.
#ifndef FP_PERM4_NMAX
#define FP_PERM4_NMAX 16u
#endif
#define FP_PERM4_ROW (((FP_PERM4_NMAX - 2u) + 3u) / 4u * 4u) /* floats per row */
#define FP_PERM4_BYTES ((FP_PERM4_NMAX - 2u) * FP_PERM4_ROW * 4u)
#define FP_THR_BYTES ((FP_NTHR + 2u) * 8u)
#define FP_PERM4_SMEM_BYTES (FP_PERM4_BYTES + FP_THR_BYTES) /* one block: the perm rows, then thr[] */
FP_HD uint32_t fp_perm4_off(uint32_t n) { return (n - 3u) * FP_PERM4_ROW; }

// Offset of the row of n in the table of float sample coordinates perm_s (rows n = 2..FP_NMAX, row n has n
// entries: perm_s[off(n) + slot] = float(1 + perm_n[slot])).
FP_HD uint32_t fp_perm_off(uint32_t n) { return n * (n - 1u) / 2u - 1u; }
