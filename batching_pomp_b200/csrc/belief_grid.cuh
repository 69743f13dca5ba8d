// belief_grid.cuh — exact spatial index of the belief model for low-dimensional spaces (D <= 3).
//
// The reference keeps its belief points in a GNAT (cspacebelief/KNNModel.hpp:65, :89-93) and asks it for the k
// nearest (:108-111).  Here the points [0, Mg) of the append-only model are binned into a uniform grid of
// 2^b cells per dimension over their bounding box, the cells numbered in Morton (Z) order: cell-ordered copies
// of the coordinates and insertion indices plus one CSR array of cell starts.  A block of 2^(D l) consecutive
// cells aligned to its size is a square / cube of side 2^l cells, so the same array is an implicit quadtree /
// octree: node (level l, code) owns the points [start[code], start[code + 2^(D l)]).  Points appended since the
// last build, [Mg, M), are the "tail" and are scanned directly; the index is rebuilt when the tail outgrows a
// fraction of the model (belief.cu: bel_grid_prepare), so the amortised build cost per appended point is O(1).
//
// A warp answers one query and keeps the k best (distance, insertion index) pairs seen so far sorted across
// its lanes, in registers (lane i holds the i-th best; k <= 32):
//   1. the tail is scanned;
//   2. seed: from the query's cell upwards to the smallest enclosing node with >= k points, then downwards
//      into the nearest child that still has >= k points until the node is small; its points are scanned, so
//      the list now holds k real points from the populated region nearest to the query;
//   3. one depth-first pass over the implicit tree in Morton order (stackless: after a node, the next node is
//      the largest aligned block that starts where it ended), skipping empty nodes, the seed node, and every
//      node whose box is farther from the query than the current k-th best distance.
// Distances are computed by the reference's arithmetic, candidates are ordered by (distance, insertion index)
// and a node is only skipped when every point in it is strictly farther than the k-th best, so the k nearest
// and their order are exactly those of the brute-force scan and of the CPU oracle: the estimate is bit-equal.
#pragma once

#include <cfloat>

#define BG_LEAF_MAX 64    // nodes with at most this many points are scanned instead of descended into
#define BG_SEED_MAX 256   // >= 2^D * k for every supported D and k: a larger node has a child with >= k points

struct BelGridView {
  const unsigned long long* start;  // [2^(D bits) + 1] in Morton order
  const double* pts;                // [Mg][D] in cell order
  const uint32_t* idx;              // [Mg] insertion index of each cell-ordered point
  const double* geom;               // device: lo[3], inv[3], h[3] (written by bg_bbox_kernel)
  uint32_t bits;                    // b: 2^b cells per dimension
  uint64_t Mg;                      // indexed prefix of the model (0: no index)
};

// Morton code of D cell coordinates of `bits` bits each (dimension 0 in the lowest bit of every digit).
template <int D>
__device__ __forceinline__ uint32_t bg_morton(const int* c, uint32_t bits) {
  uint32_t code = 0;
  for (uint32_t k = 0; k < bits; ++k)
#pragma unroll
    for (int j = 0; j < D; ++j) code |= (((uint32_t)c[j] >> k) & 1u) << (k * D + j);
  return code;
}
template <int D>
__device__ __forceinline__ void bg_demorton(uint32_t code, uint32_t bits, int* c) {
#pragma unroll
  for (int j = 0; j < D; ++j) c[j] = 0;
  for (uint32_t k = 0; k < bits; ++k)
#pragma unroll
    for (int j = 0; j < D; ++j) c[j] |= (int)((code >> (k * D + j)) & 1u) << k;
}

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
__device__ __forceinline__ uint32_t bg_cell_of(const double* x, const double* __restrict__ geom, uint32_t bits) {
  int c[D];
#pragma unroll
  for (int j = 0; j < D; ++j) c[j] = bg_cell(x[j], geom[j], geom[3 + j], 1u << bits);
  return bg_morton<D>(c, bits);
}

template <int D>
__global__ void __launch_bounds__(256) bg_hist_kernel(const double* __restrict__ pts, uint64_t M, uint32_t bits,
                                                      const double* __restrict__ geom, uint32_t* __restrict__ cell_of,
                                                      uint32_t* __restrict__ cnt) {
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < M; p += (uint64_t)gridDim.x * blockDim.x) {
    double x[D];
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = pts[p * D + j];
    const uint32_t c = bg_cell_of<D>(x, geom, bits);
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

// The warp's list: lane i holds the i-th best (distance, insertion index) pair; empty entries are
// (DBL_MAX, 0xFFFFFFFF).  worst / worst_i mirror entry K-1 in every lane.
struct BgList {
  double d;
  uint32_t i;
  double worst;
  uint32_t worst_i;
};

// Insert a warp-uniform candidate that is known to precede entry K-1.
__device__ __forceinline__ void bg_insert(BgList& L, double cd, uint32_t ci, int K, int lane) {
  const bool before = L.d < cd || (L.d == cd && L.i < ci);  // insertion indices are unique
  const int pos = __popc(__ballot_sync(0xffffffffu, before && lane < K));
  const double pd = __shfl_up_sync(0xffffffffu, L.d, 1);
  const uint32_t pi = __shfl_up_sync(0xffffffffu, L.i, 1);
  if (lane == pos) {
    L.d = cd;
    L.i = ci;
  } else if (lane > pos) {
    L.d = pd;
    L.i = pi;
  }
  if (lane >= K) {
    L.d = DBL_MAX;
    L.i = 0xFFFFFFFFu;
  }
  L.worst = __shfl_sync(0xffffffffu, L.d, K - 1);
  L.worst_i = __shfl_sync(0xffffffffu, L.i, K - 1);
}

// Offer the points [beg, end) of `pts` (insertion index = idx[p], or p itself when idx is null) to the list.
template <int D>
__device__ __forceinline__ void bg_scan(BgList& L, const double* x, const double* __restrict__ pts,
                                        const uint32_t* __restrict__ idx, unsigned long long beg,
                                        unsigned long long end, int K, int lane) {
  for (unsigned long long base = beg; base < end; base += 32) {
    const unsigned long long p = base + (unsigned long long)lane;
    double dist = DBL_MAX;
    uint32_t id = 0xFFFFFFFFu;
    bool pass = false;
    if (p < end) {
      double d2 = 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        const double diff = __dsub_rn(x[j], __ldg(pts + p * D + j));
        d2 = __dadd_rn(d2, __dmul_rn(diff, diff));
      }
      // worst^2 (1 + 2^-40) bounds the squared distance of anything that can still enter (inf when not full)
      if (!(d2 > __dmul_rn(__dmul_rn(L.worst, L.worst), 1.0 + 9.094947017729282e-13))) {
        dist = __dsqrt_rn(d2);
        id = idx ? __ldg(idx + p) : (uint32_t)p;
        pass = dist < L.worst || (dist == L.worst && id < L.worst_i);
      }
    }
    unsigned m = __ballot_sync(0xffffffffu, pass);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const double cd = __shfl_sync(0xffffffffu, dist, src);
      const uint32_t ci = __shfl_sync(0xffffffffu, id, src);
      if (cd < L.worst || (cd == L.worst && ci < L.worst_i)) bg_insert(L, cd, ci, K, lane);
    }
  }
}

// Squared lower bound on the distance from x to any point of node (level l, code).  A point of the node has a
// fine cell in [c0, c1] per dimension; by the monotone cell function its coordinate is at least
// lo + c0 h and at most lo + (c1 + 1) h up to a few roundings, which the slack dwarfs.
template <int D>
__device__ __forceinline__ double bg_node_lb2(const double* x, const double* lo, const double* h, uint32_t code,
                                              int l, uint32_t bits) {
  int c[D];
  bg_demorton<D>(code >> (D * l), bits - (uint32_t)l, c);
  double lb2 = 0.0;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    const double c0 = (double)((unsigned)c[j] << l), c1p = (double)(((unsigned)c[j] + 1u) << l);
    const double slack = 9.094947017729282e-13 * (fabs(x[j]) + fabs(lo[j]) + c1p * h[j]);
    const double below = ((lo[j] + c0 * h[j]) - x[j]) - slack;   // x left of the box
    const double above = (x[j] - (lo[j] + c1p * h[j])) - slack;  // x right of the box
    const double a = fmax(0.0, fmax(below, above));
    if (a > 0.0 && a < DBL_MAX) lb2 += a * a;
  }
  return lb2;
}

// KNNModel::estimate (KNNModel.hpp:101-134) of one query by a warp through the index.
template <int D>
__device__ __forceinline__ double bg_estimate_warp(const double* x, const BelGridView& g,
                                                   const double* __restrict__ pts, const double* __restrict__ vals,
                                                   uint64_t M, int K, double prior, double prior_weight, int lane) {
  if (M == 0) return prior;  // KNNModel.hpp:104-106
  BgList L;
  L.d = DBL_MAX;
  L.i = 0xFFFFFFFFu;
  L.worst = DBL_MAX;
  L.worst_i = 0xFFFFFFFFu;
  // 1. the tail: points appended since the index was built
  bg_scan<D>(L, x, pts, nullptr, g.Mg, M, K, lane);
  if (g.Mg > 0) {
    const uint32_t bits = g.bits;
    const uint32_t ncells = 1u << (D * bits);
    double lo[D], h[D];
    int c[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      lo[j] = __ldg(g.geom + j);
      h[j] = __ldg(g.geom + 6 + j);
      c[j] = bg_cell(x[j], lo[j], __ldg(g.geom + 3 + j), 1u << bits);
    }
    // 2. seed node: up to the smallest enclosing node with >= K points ...
    uint32_t scode = bg_morton<D>(c, bits);
    int sl = 0;
    unsigned long long sb, se;
    for (;; ++sl) {
      scode = (scode >> (D * sl)) << (D * sl);
      sb = __ldg(g.start + scode);
      se = __ldg(g.start + scode + (1u << (D * sl)));
      if (se - sb >= (unsigned long long)K || sl == (int)bits) break;
    }
    // ... then down into the nearest child that still has >= K points, while the node is large
    while (sl > 0 && se - sb > BG_SEED_MAX) {
      int best = -1;
      double best_lb = DBL_MAX;
      unsigned long long bb = 0, be = 0;
      const uint32_t cs = 1u << (D * (sl - 1));
      for (int ch = 0; ch < (1 << D); ++ch) {
        const uint32_t cc = scode + (uint32_t)ch * cs;
        const unsigned long long b0 = __ldg(g.start + cc), e0 = __ldg(g.start + cc + cs);
        if (e0 - b0 < (unsigned long long)K) continue;
        const double lb = bg_node_lb2<D>(x, lo, h, cc, sl - 1, bits);
        if (best < 0 || lb < best_lb) {
          best = ch;
          best_lb = lb;
          bb = b0;
          be = e0;
        }
      }
      if (best < 0) break;
      scode += (uint32_t)best * cs;
      --sl;
      sb = bb;
      se = be;
    }
    bg_scan<D>(L, x, g.pts, g.idx, sb, se, K, lane);
    // 3. depth-first pass in Morton order with pruning
    uint32_t code = 0;
    int l = (int)bits;
    while (code < ncells) {
      const uint32_t span = 1u << (D * l);
      const unsigned long long b0 = __ldg(g.start + code), e0 = __ldg(g.start + code + span);
      bool skip = e0 == b0 || (l == sl && code == scode);
      if (!skip) {
        const double lb2 = bg_node_lb2<D>(x, lo, h, code, l, bits);
        skip = lb2 * (1.0 - 9.313225746154785e-10) > L.worst * L.worst;  // every point of the node is farther
      }
      if (!skip && l > 0 && (e0 - b0 > BG_LEAF_MAX || (l > sl && ((code ^ scode) >> (D * l)) == 0))) {
        --l;  // first child starts at the same code (ancestors of the seed node are always opened)
        continue;
      }
      if (!skip) bg_scan<D>(L, x, g.pts, g.idx, b0, e0, K, lane);
      code += span;
      if (code >= ncells) break;
      l = min((int)bits, (__ffs(code) - 1) / D);
    }
  }
  // KNNModel.hpp:108-130 over the sorted list
  const int knn = (uint64_t)K < M ? K : (int)M;
  const bool have = lane < knn && L.i != 0xFFFFFFFFu;
  const double distance = fmax(0.0001, L.d);                    // :121
  const double w = have ? __ddiv_rn(1.0, distance) : 0.0;       // :122
  const double wv = have ? __dmul_rn(w, __ldg(vals + L.i)) : 0.0;  // :123
  double dot = 0.0, wsum = 0.0;
  for (int k = 0; k < knn; ++k) {
    if (__shfl_sync(0xffffffffu, L.i, k) == 0xFFFFFFFFu) break;
    dot = __dadd_rn(dot, __shfl_sync(0xffffffffu, wv, k));   // weights.dot(values), sequential
    wsum = __dadd_rn(wsum, __shfl_sync(0xffffffffu, w, k));  // weights.sum()
  }
  double result = __ddiv_rn(dot, wsum);                                                                   // :127
  result = __ddiv_rn(__dadd_rn(result, __dmul_rn(prior_weight, prior)), __dadd_rn(1.0, prior_weight));  // :130
  return result;
}
