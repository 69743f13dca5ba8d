// belief_grid.cuh — exact spatial index of the belief model for low-dimensional spaces (D <= 3).
//
// The reference keeps its belief points in a GNAT (cspacebelief/KNNModel.hpp:65, :89-93) and asks it for the k
// nearest (:108-111).  Here the points [0, Mg) of the append-only model are binned into a uniform grid over
// their bounding box (cell-ordered copies of the coordinates and insertion indices, CSR cell starts); points
// appended since the last build, [Mg, M), are the "tail" and are scanned directly.  The index is rebuilt when
// the tail outgrows a fraction of the model (belief.cu: bel_grid_prepare), so the amortised build cost per
// appended point is O(1).
//
// A warp answers one query: the tail first, then the cells at Chebyshev distance 0, 1, 2, ... from the
// query's cell.  Every lane keeps the k best (distance, insertion index) pairs of the points it saw (shared
// memory, as in the brute-force one-launch kernel); after each ring the warp stops as soon as k of the points
// seen so far are closer than the nearest face of the searched block of cells — every point outside the block
// is farther than that face — and the lists are merged in (distance, index) order.  Distances are computed by
// the reference's arithmetic and ties are resolved by insertion index, so the k nearest and their order are
// exactly those of the brute-force scan and of the CPU oracle: the estimate is bit-equal.
#pragma once

#include <cfloat>

struct BelGridView {
  const unsigned long long* start;  // [cells + 1]
  const double* pts;                // [Mg][D] in cell order
  const uint32_t* idx;              // [Mg] insertion index of each cell-ordered point
  const double* geom;               // device: lo[3], inv[3], h[3] (written by bg_bbox_kernel)
  uint32_t G;                       // cells per dimension
  uint64_t Mg;                      // indexed prefix of the model (0: no index)
  uint32_t ring_cap;                // rings searched before a query falls back to the plain scan
};

// Monotone cell index (x <= y  =>  cell(x) <= cell(y)): the subtraction, the multiplication by a non-negative
// factor and the truncation are all monotone.  NaN and negative values go to cell 0.
__device__ __forceinline__ int bg_cell(double x, double lo, double inv, uint32_t G) {
  const double v = __dmul_rn(__dsub_rn(x, lo), inv);
  int c = __double2int_rz(v);  // saturating, NaN -> 0
  c = c < 0 ? 0 : c;
  return c > (int)G - 1 ? (int)G - 1 : c;
}

// Bounding box of the points and the cell geometry: one block.
template <int D>
__global__ void __launch_bounds__(1024) bg_bbox_kernel(const double* __restrict__ pts, uint64_t M, uint32_t G,
                                                       double* __restrict__ geom) {
  __shared__ double s_lo[32][D], s_hi[32][D];
  double lo[D], hi[D];
#pragma unroll
  for (int j = 0; j < D; ++j) { lo[j] = DBL_MAX; hi[j] = -DBL_MAX; }
  for (uint64_t p = threadIdx.x; p < M; p += blockDim.x) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const double v = pts[p * D + j];
      lo[j] = fmin(lo[j], v);  // fmin / fmax ignore a NaN operand
      hi[j] = fmax(hi[j], v);
    }
  }
#pragma unroll
  for (int j = 0; j < D; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[j] = fmin(lo[j], __shfl_xor_sync(0xffffffffu, lo[j], o));
      hi[j] = fmax(hi[j], __shfl_xor_sync(0xffffffffu, hi[j], o));
    }
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5][j] = lo[j]; s_hi[threadIdx.x >> 5][j] = hi[j]; }
  }
  __syncthreads();
  if (threadIdx.x < D) {
    const int j = threadIdx.x;
    double a = DBL_MAX, b = -DBL_MAX;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a = fmin(a, s_lo[w][j]); b = fmax(b, s_hi[w][j]); }
    double ext = b - a;
    if (!(ext > 0.0) || !(ext < DBL_MAX)) ext = 0.0;  // one point, equal coordinates, or nothing finite
    geom[j] = ext > 0.0 ? a : 0.0;
    geom[3 + j] = ext > 0.0 ? (double)G / ext : 0.0;  // inv = 0: everything in cell 0 of this dimension
    geom[6 + j] = ext > 0.0 ? ext / (double)G : DBL_MAX;
  }
}

template <int D>
__device__ __forceinline__ uint32_t bg_cell_of(const double* x, const double* __restrict__ geom, uint32_t G) {
  uint32_t c = 0;
#pragma unroll
  for (int j = D - 1; j >= 0; --j) c = c * G + (uint32_t)bg_cell(x[j], geom[j], geom[3 + j], G);
  return c;
}

template <int D>
__global__ void __launch_bounds__(256) bg_hist_kernel(const double* __restrict__ pts, uint64_t M, uint32_t G,
                                                      const double* __restrict__ geom, uint32_t* __restrict__ cell_of,
                                                      uint32_t* __restrict__ cnt) {
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < M; p += (uint64_t)gridDim.x * blockDim.x) {
    double x[D];
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = pts[p * D + j];
    const uint32_t c = bg_cell_of<D>(x, geom, G);
    cell_of[p] = c;
    atomicAdd(cnt + c, 1u);
  }
}

// cnt holds the cell populations and is counted down to zero: the slot of a point inside its cell is arbitrary
// (queries order candidates by (distance, insertion index) themselves).
template <int D>
__global__ void __launch_bounds__(256) bg_scatter_kernel(const double* __restrict__ pts, uint64_t M,
                                                         const uint32_t* __restrict__ cell_of, uint32_t* __restrict__ cnt,
                                                         const unsigned long long* __restrict__ start,
                                                         double* __restrict__ out_pts, uint32_t* __restrict__ out_idx) {
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < M; p += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t c = cell_of[p];
    const unsigned long long o = start[c] + (atomicSub(cnt + c, 1u) - 1u);
#pragma unroll
    for (int j = 0; j < D; ++j) out_pts[o * D + j] = pts[p * D + j];
    out_idx[o] = (uint32_t)p;
  }
}

// One candidate in the reference's arithmetic, inserted into the lane's list in (distance, index) order; the
// candidates of a lane arrive in no particular index order.
template <int D>
__device__ __forceinline__ void bg_consider(const double* x, const double* __restrict__ pt, uint32_t idx, int K,
                                            double* kd, uint32_t* ki, int& cnt, double& worst, uint32_t& worst_i,
                                            double& thr) {
  double d2 = 0.0;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    const double diff = __dsub_rn(x[j], __ldg(pt + j));
    d2 = __dadd_rn(d2, __dmul_rn(diff, diff));
  }
  if (cnt == K && d2 > thr) return;
  const double dist = __dsqrt_rn(d2);
  if (cnt == K && (dist > worst || (dist == worst && idx > worst_i))) return;
  if (!(dist == dist)) return;  // a NaN distance is never among the nearest
  int pos = cnt < K ? cnt : K - 1;
  while (pos > 0) {
    const double pd = kd[(pos - 1) * 32];
    const uint32_t pi = ki[(pos - 1) * 32];
    if (!(pd > dist || (pd == dist && pi > idx))) break;
    kd[pos * 32] = pd;
    ki[pos * 32] = pi;
    --pos;
  }
  kd[pos * 32] = dist;
  ki[pos * 32] = idx;
  if (cnt < K) ++cnt;
  if (cnt == K) {
    worst = kd[(K - 1) * 32];
    worst_i = ki[(K - 1) * 32];
    thr = __dmul_rn(__dmul_rn(worst, worst), 1.0 + 9.094947017729282e-13);  // conservative (2^-40)
  }
}

// Merge of the 32 per-lane lists and KNNModel::estimate (KNNModel.hpp:101-134); M > 0.
__device__ __forceinline__ double bg_merge_estimate(const double* kd, const uint32_t* ki, int cnt, uint64_t M, int K,
                                                    const double* __restrict__ vals, double prior,
                                                    double prior_weight) {
  const int knn = (uint64_t)K < M ? K : (int)M;  // :108
  int head = 0;
  double dot = 0.0, wsum = 0.0;
  for (int k = 0; k < knn; ++k) {
    double cd = head < cnt ? kd[head * 32] : DBL_MAX;
    uint32_t ci = head < cnt ? ki[head * 32] : 0xFFFFFFFFu;
    const uint32_t mine = ci;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, cd, o);
      const uint32_t oi = __shfl_xor_sync(0xffffffffu, ci, o);
      if (od < cd || (od == cd && oi < ci)) {
        cd = od;
        ci = oi;
      }
    }
    if (ci == 0xFFFFFFFFu) break;                 // fewer finite-distance points than knn
    if (mine == ci) ++head;                       // the winner's list advances (indices are unique)
    const double distance = fmax(0.0001, cd);     // :121
    const double w = __ddiv_rn(1.0, distance);    // :122
    const double v = __ldg(vals + ci);            // :123
    dot = __dadd_rn(dot, __dmul_rn(w, v));        // weights.dot(values), sequential
    wsum = __dadd_rn(wsum, w);                    // weights.sum()
  }
  double result = __ddiv_rn(dot, wsum);                                                                   // :127
  result = __ddiv_rn(__dadd_rn(result, __dmul_rn(prior_weight, prior)), __dadd_rn(1.0, prior_weight));  // :130
  return result;
}

// KNNModel::estimate of one query by a warp through the grid.  kd / ki point at this lane's column of the
// warp's [K][32] lists.
template <int D>
__device__ __forceinline__ double bg_estimate_warp(const double* x, const BelGridView& g,
                                                   const double* __restrict__ pts, const double* __restrict__ vals,
                                                   uint64_t M, int K, double prior, double prior_weight, double* kd,
                                                   uint32_t* ki, int lane) {
  if (M == 0) return prior;  // KNNModel.hpp:104-106
  int cnt = 0;
  double worst = DBL_MAX, thr = DBL_MAX;
  uint32_t worst_i = 0xFFFFFFFFu;
  // the tail: points appended since the index was built
  for (uint64_t p = g.Mg + (uint64_t)lane; p < M; p += 32)
    bg_consider<D>(x, pts + p * D, (uint32_t)p, K, kd, ki, cnt, worst, worst_i, thr);
  bool fallback = false;
  if (g.Mg > 0) {
    int c[D];
    double lo[D], h[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      lo[j] = __ldg(g.geom + j);
      h[j] = __ldg(g.geom + 6 + j);
      c[j] = bg_cell(x[j], lo[j], __ldg(g.geom + 3 + j), g.G);
    }
    const int G = (int)g.G;
    for (int rho = 0;; ++rho) {
      if ((uint32_t)rho > g.ring_cap) {
        fallback = true;
        break;
      }
      // cells at Chebyshev distance exactly rho from c
      if (D == 1) {
        if (lane < 2 && (rho > 0 || lane == 0)) {
          const int cx = lane == 0 ? c[0] - rho : c[0] + rho;
          if (cx >= 0 && cx < G) {
            const unsigned long long b = __ldg(g.start + cx), e = __ldg(g.start + cx + 1);
            for (unsigned long long p = b; p < e; ++p)
              bg_consider<D>(x, g.pts + p * D, __ldg(g.idx + p), K, kd, ki, cnt, worst, worst_i, thr);
          }
        }
      } else if (D == 2) {
        const int ring = rho == 0 ? 1 : 8 * rho;
        for (int t = lane; t < ring; t += 32) {
          int dx = 0, dy = 0;
          if (rho > 0) {
            const int side = t / (2 * rho), o = t - side * 2 * rho;
            dx = side == 0 ? -rho + o : side == 1 ? rho : side == 2 ? rho - o : -rho;
            dy = side == 0 ? -rho : side == 1 ? -rho + o : side == 2 ? rho : rho - o;
          }
          const int cx = c[0] + dx, cy = c[1] + dy;
          if (cx < 0 || cx >= G || cy < 0 || cy >= G) continue;
          const int cell = cy * G + cx;
          const unsigned long long b = __ldg(g.start + cell), e = __ldg(g.start + cell + 1);
          for (unsigned long long p = b; p < e; ++p)
            bg_consider<D>(x, g.pts + p * D, __ldg(g.idx + p), K, kd, ki, cnt, worst, worst_i, thr);
        }
      } else {
        const int side = 2 * rho + 1, block = side * side * side;
        for (int t = lane; t < block; t += 32) {
          const int dz = t / (side * side) - rho, rem = t % (side * side);
          const int dy = rem / side - rho, dx = rem % side - rho;
          const int ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy, az = dz < 0 ? -dz : dz;
          if (max(ax, max(ay, az)) != rho) continue;  // interior: searched by an earlier ring
          const int cx = c[0] + dx, cy = c[1] + dy, cz = c[D - 1] + dz;
          if (cx < 0 || cx >= G || cy < 0 || cy >= G || cz < 0 || cz >= G) continue;
          const int cell = (cz * G + cy) * G + cx;
          const unsigned long long b = __ldg(g.start + cell), e = __ldg(g.start + cell + 1);
          for (unsigned long long p = b; p < e; ++p)
            bg_consider<D>(x, g.pts + p * D, __ldg(g.idx + p), K, kd, ki, cnt, worst, worst_i, thr);
        }
      }
      // nearest face of the searched block that still has cells behind it.  A point behind the low face of
      // dimension j has cell <= c-rho-1, so (p-lo)*inv < c-rho up to 2 roundings, and x_j - p_j is at least
      // (x_j - lo_j) - (c-rho) h_j minus a slack that dwarfs those roundings (and the same on the high side).
      double gap = DBL_MAX;
      bool open = false;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        const double slack = 9.094947017729282e-13 * (fabs(x[j]) + fabs(lo[j]) + (double)(c[j] + rho + 1) * h[j]);
        if (c[j] - rho > 0) {
          open = true;
          gap = fmin(gap, ((x[j] - lo[j]) - (double)(c[j] - rho) * h[j]) - slack);
        }
        if (c[j] + rho < G - 1) {
          open = true;
          gap = fmin(gap, ((lo[j] + (double)(c[j] + rho + 1) * h[j]) - x[j]) - slack);
        }
      }
      if (!open) break;  // the whole grid has been searched
      const double lim = gap * (1.0 - 9.094947017729282e-13);
      int closer = 0;
      for (int k = 0; k < cnt; ++k) closer += kd[k * 32] < lim ? 1 : 0;
      closer = __reduce_add_sync(0xffffffffu, closer);
      if (closer >= K) break;
    }
  }
  if (fallback) {  // a query far from every point: plain scan of the whole model
    cnt = 0;
    worst = DBL_MAX;
    thr = DBL_MAX;
    worst_i = 0xFFFFFFFFu;
    for (uint64_t p = (uint64_t)lane; p < M; p += 32)
      bg_consider<D>(x, pts + p * D, (uint32_t)p, K, kd, ki, cnt, worst, worst_i, thr);
  }
  return bg_merge_estimate(kd, ki, cnt, M, K, vals, prior, prior_weight);
}
