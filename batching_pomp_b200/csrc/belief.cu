// belief.cu — the C-space belief model: cspacebelief::KNNModel
// (/root/reference/include/batching_pomp/cspacebelief/KNNModel.hpp:56-134, BeliefPoint.hpp:43-71,
// beliefDistanceFunction src/BatchingPOMP.cpp:213-219) and computeAndSetEdgeFreeProbability
// (src/BatchingPOMP.cpp:464-493) as batched device operators.
//
// Exact kNN over the append-only point set, three ways:
//   * dim <= 3: a Morton-ordered grid used as an implicit quadtree / octree, a warp per query with the k best in
//     registers (belief_grid.cuh); rebuilt incrementally as the model grows;
//   * dim >= 4, large batches: brute force tiled through shared memory with a packed-FP32 prefilter and deferred
//     insertion (bk_knn_partial_f32); one thread per query slot, a (distance, insertion index)-ordered top-k list
//     per query in shared memory, the point set split into chunks across blocks and merged afterwards;
//   * small batches (the planner's per-expansion calls): one launch, a block per edge, a warp per query
//     (bk_pfree_fused), through the index where there is one.
// Ordering and summation are sequential in ascending (distance, insertion index), as the CPU oracle defines
// them, so results are bit-equal to the oracle and within 1e-12 relative of the
// reference's Eigen-packetised sums (checked against the reference's own KNNModel, tests/test_oracle_vs_reference.py).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "belief_grid.cuh"

#ifndef BK_QB
#define BK_QB 128    // threads per block
#endif
#ifndef BK_QT
#define BK_QT 2      // queries per thread
#endif
#ifndef BK_TILE
#define BK_TILE 256  // belief points staged per tile
#endif
#ifndef BK_G
#define BK_G 8       // points per prefilter group (independent accumulation chains per query)
#endif
#define BK_KMAX 32   // largest supported k

extern "C" int bpomp_belief_config(bpomp_ctx* ctx, int k, double prior, double prior_weight) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (k < 1 || k > BK_KMAX) return bp_fail(ctx, BPOMP_ERR_INVALID, "k must be in 1..%d", BK_KMAX);
  ctx->knn = k;
  ctx->prior = prior;
  ctx->prior_weight = prior_weight;
  return BPOMP_OK;
}

extern "C" int bpomp_belief_size(bpomp_ctx* ctx, uint64_t* count) {
  if (!ctx || !count) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  *count = ctx->bel_count;
  return BPOMP_OK;
}

extern "C" int bpomp_belief_clear(bpomp_ctx* ctx) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  ctx->bel_count = 0;
  ctx->bg_count = 0;
  return BPOMP_OK;
}

static int belief_grow(bpomp_ctx* ctx, uint64_t extra) {
  uint64_t need = ctx->bel_count + extra;
  // insertion indices are 32-bit in the kNN lists (0xFFFFFFFF marks an empty entry)
  if (need >= 0xFFFFFFFFull) return bp_fail(ctx, BPOMP_ERR_RANGE, "belief model is limited to 2^32-2 points");
  BP_TRY(bp_reserve(ctx, ctx->d_bel_pts, (size_t)need * ctx->dim * sizeof(double), true));
  BP_TRY(bp_reserve(ctx, ctx->d_bel_vals, (size_t)need * sizeof(double), true));
  return BPOMP_OK;
}

extern "C" int bpomp_belief_append(bpomp_ctx* ctx, const double* states, const double* values, int64_t count) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (count < 0 || (count > 0 && (!states || !values))) return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  if (count == 0) return BPOMP_OK;
  BP_TRY(belief_grow(ctx, (uint64_t)count));
  size_t row = (size_t)ctx->dim * sizeof(double);
  BP_CUDA(ctx, cudaMemcpyAsync((char*)ctx->d_bel_pts.p + ctx->bel_count * row, states, (size_t)count * row,
                               cudaMemcpyHostToDevice, ctx->stream));
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->d_bel_vals.as<double>() + ctx->bel_count, values, (size_t)count * sizeof(double),
                               cudaMemcpyHostToDevice, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->bel_count += (uint64_t)count;
  return BPOMP_OK;
}

// ---------------------------------------------------------------------------
// Edge geometry helpers shared by append_checked and edge_pfree
// ---------------------------------------------------------------------------
struct EdgeSrc {
  const double* from;        // explicit [count][D] or null
  const double* to;
  const uint32_t* from_ids;  // indexed or null
  const uint32_t* to_ids;
  const double* verts;
};

template <int D>
__device__ __forceinline__ void bk_load(const EdgeSrc& s, int64_t e, double* f, double* t) {
  const double* pf = s.from_ids ? s.verts + (size_t)s.from_ids[e] * D : s.from + (size_t)e * D;
  const double* pt = s.to_ids ? s.verts + (size_t)s.to_ids[e] * D : s.to + (size_t)e * D;
#pragma unroll
  for (int j = 0; j < D; ++j) { f[j] = pf[j]; t[j] = pt[j]; }
}

// mode 0: number of belief queries of computeAndSetEdgeFreeProbability (BP.cpp:475)
// mode 1: number of states checkAndSetEdgeBlocked appends to the model (BP.cpp:510-537)
template <int D>
__global__ void __launch_bounds__(256) bk_count_kernel(EdgeSrc src, int64_t count, double L, PermView pv, int mode,
                                                       const int32_t* __restrict__ first_invalid,
                                                       uint32_t* __restrict__ nst, uint32_t* __restrict__ cnt) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x) {
    double f[D], t[D];
    bk_load<D>(src, e, f, t);
    uint32_t n = bp_nstates(__dsqrt_rn(bp_dist2<D>(f, t)), L);
    uint32_t c = 0;
    if (n == 0u) {
      atomicAdd(pv.flags + 1, 1);
      n = 2u;
    }
    if (n > 2u) {
      if (__ldg(pv.off + n) == BP_ABSENT) {
        pv.need[n] = 1;
        atomicAdd(pv.flags, 1);
      }
      if (mode == 0) {
        uint32_t step = __ldg(pv.step + n);  // (unsigned)(n / log(n)), BP.cpp:472
        c = (n - 2u + step - 1u) / step;      // i = 1, 1+step, ... < n-1
      } else {
        int32_t fi = first_invalid[e];
        c = fi >= 1 ? (uint32_t)fi : n - 2u;
      }
    }
    nst[e] = n;
    cnt[e] = c;
  }
}

// Writes the states of the selected slots of every edge: mode 0 -> query states (slots 1, 1+step, ..),
// mode 1 -> checked states (slots 1..cnt) plus their belief values.
template <int D>
__global__ void __launch_bounds__(256) bk_expand_kernel(EdgeSrc src, int64_t count, PermView pv, int mode,
                                                        const int32_t* __restrict__ first_invalid,
                                                        const uint32_t* __restrict__ nst,
                                                        const uint64_t* __restrict__ offs, double* __restrict__ out,
                                                        double* __restrict__ out_vals) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x) {
    uint64_t o = offs[e];
    uint32_t c = (uint32_t)(offs[e + 1] - o);
    if (c == 0) continue;
    double f[D], t[D], dtf[D];
    bk_load<D>(src, e, f, t);
#pragma unroll
    for (int j = 0; j < D; ++j) dtf[j] = __dsub_rn(t[j], f[j]);
    uint32_t n = nst[e];
    const double* tf = pv.tfrac + __ldg(pv.off + n);
    uint32_t step = mode == 0 ? __ldg(pv.step + n) : 1u;
    bool blocked = mode == 1 && first_invalid[e] >= 1;
    for (uint32_t k = 0; k < c; ++k) {
      double tt = __ldg(tf + 1u + k * step);
#pragma unroll
      for (int j = 0; j < D; ++j) out[(o + k) * D + j] = __dadd_rn(f[j], __dmul_rn(dtf[j], tt));
      if (mode == 1) out_vals[o + k] = (blocked && k + 1 == c) ? 1.0 : 0.0;  // BP.cpp:525 / :535
    }
  }
}

// ---------------------------------------------------------------------------
// kNN: partial top-k per (query, point chunk)
// ---------------------------------------------------------------------------
// Tile layout: bk_ds(D) doubles per point = the D coordinates, |p|^2 / 2 at index D, padded to an even
// count so a point is read with 16-byte broadcast loads.
__host__ __device__ constexpr int bk_ds(int D) { return (D + 2) & ~1; }

// One (query, point) candidate in the reference's arithmetic (unfused sum of squares, sqrt) and the ordered
// insertion into the query's list kd/ki (entries BK_QT*BK_QB apart).  nq = -query (negation is exact).
template <int D>
__device__ __forceinline__ void bk_consider(const double* nq, const double* pt, uint32_t idx, int K, double* kd,
                                            uint32_t* ki, int& cnt, double& worst, double& thr) {
  constexpr int LS = BK_QT * BK_QB;
  double d2 = 0.0;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    double diff = __dsub_rn(-nq[j], pt[j]);
    d2 = __dadd_rn(d2, __dmul_rn(diff, diff));
  }
  if (cnt == K && d2 > thr) return;
  double dist = __dsqrt_rn(d2);
  if (cnt == K && !(dist < worst)) return;  // an equidistant later point never displaces (GNAT)
  // insert after every entry with distance <= dist (indices ascend within the scan)
  int pos = cnt < K ? cnt : K - 1;
  while (pos > 0 && kd[(pos - 1) * LS] > dist) {
    kd[pos * LS] = kd[(pos - 1) * LS];
    ki[pos * LS] = ki[(pos - 1) * LS];
    --pos;
  }
  kd[pos * LS] = dist;
  ki[pos * LS] = idx;
  if (cnt < K) ++cnt;
  if (cnt == K) {
    worst = kd[(K - 1) * LS];
    thr = __dmul_rn(__dmul_rn(worst, worst), 1.0 + 9.094947017729282e-13);  // conservative (2^-40)
  }
}

// Each thread owns BK_QT queries (tid, tid + BK_QB, ...) of the block's BK_QT*BK_QB, so a point read from the
// tile (a shared-memory broadcast, the scarce resource of this kernel) serves BK_QT distance evaluations.
template <int D>
__global__ void __launch_bounds__(BK_QB) bk_knn_partial(const double* __restrict__ queries, int64_t nq_total,
                                                        const double* __restrict__ pts, uint64_t M, uint64_t chunk,
                                                        int K, double* __restrict__ part_d,
                                                        uint32_t* __restrict__ part_i) {
  constexpr int DS = bk_ds(D);
  constexpr int LS = BK_QT * BK_QB;
  extern __shared__ __align__(16) unsigned char smem[];
  double* s_pts = reinterpret_cast<double*>(smem);                                  // [BK_TILE][DS]
  double* s_kd = s_pts + BK_TILE * DS;                                              // [K][BK_QT][BK_QB]
  uint32_t* s_ki = reinterpret_cast<uint32_t*>(s_kd + (size_t)K * LS);              // [K][BK_QT][BK_QB]
  __shared__ unsigned long long s_pnmax;  // bits of the tile's largest |p|^2 (non-negative doubles order as integers)
  const int tid = threadIdx.x;
  const uint64_t p_begin = (uint64_t)blockIdx.y * chunk;
  const uint64_t p_end = p_begin + chunk < M ? p_begin + chunk : M;
  double nq[BK_QT][D];   // -query
  double qn[BK_QT];      // |q|^2 (approximate: prefilter only)
  double worst[BK_QT], thr[BK_QT], thrp[BK_QT];
  int cnt[BK_QT];
#pragma unroll
  for (int s = 0; s < BK_QT; ++s) {
    const int64_t qi = (int64_t)blockIdx.x * LS + s * BK_QB + tid;
    qn[s] = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const double v = qi < nq_total ? queries[qi * D + j] : 0.0;  // out-of-range slots scan as query 0, unwritten
      nq[s][j] = -v;
      qn[s] = fma(v, v, qn[s]);
    }
    cnt[s] = 0;
    worst[s] = DBL_MAX;  // worst distance of a full list
    thr[s] = DBL_MAX;    // its square, padded: bound on the unfused d2
    thrp[s] = DBL_MAX;   // bound on the prefilter value
  }

  // Prefilter: |q-p|^2 / 2 = |p|^2/2 - q.p + |q|^2/2 as D fused multiply-adds on top of the stored |p|^2/2 (a
  // third of the FP64 work of the unfused difference form).  Its absolute error is below 2^-49 (|q|^2 + |p|^2);
  // the bound is padded by 2^-46 (|q|^2 + max |p|^2 of the tile), so a point it rejects is beyond thr in the
  // reference's arithmetic too.  Survivors go through bk_consider().  A NaN makes the comparison false.
  auto approx = [&](const double2* pp, int s) {
    double acc = (D & 1) ? pp[D / 2].y : pp[D / 2].x;
#pragma unroll
    for (int h = 0; h < D / 2; ++h) {
      const double2 v = pp[h];
      acc = fma(nq[s][2 * h], v.x, acc);
      acc = fma(nq[s][2 * h + 1], v.y, acc);
    }
    if (D & 1) acc = fma(nq[s][D - 1], pp[D / 2].x, acc);
    return acc;
  };

  for (uint64_t base = p_begin; base < p_end; base += BK_TILE) {
    uint32_t tn = (uint32_t)((p_end - base) < BK_TILE ? (p_end - base) : BK_TILE);
    __syncthreads();
    if (tid == 0) s_pnmax = 0ull;
    for (uint32_t i = tid; i < tn * D; i += BK_QB) s_pts[(i / D) * DS + (i % D)] = pts[base * D + i];
    __syncthreads();
    for (uint32_t p = tid; p < tn; p += BK_QB) {
      double pn = 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) pn = fma(s_pts[p * DS + j], s_pts[p * DS + j], pn);
      s_pts[p * DS + D] = 0.5 * pn;
      atomicMax(&s_pnmax, (unsigned long long)__double_as_longlong(pn));  // NaN bits exceed +inf: stays NaN
    }
    __syncthreads();
    const double pnmax = __longlong_as_double((long long)s_pnmax);
    double pad[BK_QT];
#pragma unroll
    for (int s = 0; s < BK_QT; ++s) {
      pad[s] = 2.842170943040401e-14 * (qn[s] + pnmax);  // 2^-45, halved with the bound below
      thrp[s] = 0.5 * ((thr[s] - qn[s]) + pad[s]);
    }
    uint32_t p = 0;
    while (p < tn) {  // cnt[] is the same in every thread: min(points scanned, K)
      if (cnt[0] == K && !(p & (BK_G - 1u)) && p + BK_G <= tn) {
        double acc[BK_G][BK_QT];
        bool far = true;
#pragma unroll
        for (int u = 0; u < BK_G; ++u) {
          const double2* pp = reinterpret_cast<const double2*>(s_pts + (p + u) * DS);
#pragma unroll
          for (int s = 0; s < BK_QT; ++s) {
            acc[u][s] = approx(pp, s);
            far = far && acc[u][s] > thrp[s];
          }
        }
        if (!far) {
#pragma unroll
          for (int u = 0; u < BK_G; ++u) {
#pragma unroll
            for (int s = 0; s < BK_QT; ++s) {
              if (!(acc[u][s] > thrp[s])) {  // a stale (looser) bound after an insertion stays conservative
                bk_consider<D>(nq[s], s_pts + (p + u) * DS, (uint32_t)(base + p + u), K, s_kd + s * BK_QB + tid,
                               s_ki + s * BK_QB + tid, cnt[s], worst[s], thr[s]);
                thrp[s] = 0.5 * ((thr[s] - qn[s]) + pad[s]);
              }
            }
          }
        }
        p += BK_G;
      } else {
        const double2* pp = reinterpret_cast<const double2*>(s_pts + p * DS);
#pragma unroll
        for (int s = 0; s < BK_QT; ++s) {
          if (cnt[s] < K || !(approx(pp, s) > thrp[s])) {
            bk_consider<D>(nq[s], s_pts + p * DS, (uint32_t)(base + p), K, s_kd + s * BK_QB + tid,
                           s_ki + s * BK_QB + tid, cnt[s], worst[s], thr[s]);
            thrp[s] = 0.5 * ((thr[s] - qn[s]) + pad[s]);
          }
        }
        ++p;
      }
    }
  }
#pragma unroll
  for (int s = 0; s < BK_QT; ++s) {
    const int64_t qi = (int64_t)blockIdx.x * LS + s * BK_QB + tid;
    if (qi >= nq_total) continue;
    double* od = part_d + ((size_t)blockIdx.y * nq_total + qi) * K;
    uint32_t* oi = part_i + ((size_t)blockIdx.y * nq_total + qi) * K;
    for (int k = 0; k < K; ++k) {
      od[k] = k < cnt[s] ? s_kd[k * LS + s * BK_QB + tid] : DBL_MAX;
      oi[k] = k < cnt[s] ? s_ki[k * LS + s * BK_QB + tid] : 0xFFFFFFFFu;
    }
  }
}

// ---------------------------------------------------------------------------
// The same partial top-k for dimensions >= 4 (no spatial index there), restructured around the two things the
// FP64 version above was bound by (profiles/r2_belief_summary.md: FP64 pipe 45 % active at 16 % warps active, and
// list insertions running with one or two active lanes):
//  * the prefilter runs on the FP32 pipe with the packed two-lane instructions of sm_100 (FFMA2).  The tile holds
//    the points rounded to float, two points interleaved per record ((x_j of point 2i, x_j of point 2i+1) pairs,
//    then the pair of |p|^2/2 values, computed in FP64 and rounded), so one 16-byte broadcast load feeds two
//    packed fused multiply-adds; a thread owns BKS_QT queries (each coordinate duplicated into a float2).
//        a = |p|^2/2 - q.p - t   (FP32),   t = (thr - |q|^2)/2 + E rounded up,   rejects the point iff a > 0.
//    E = 2^-20 (|q|^2 + max|p|^2) bounds the error of a — inputs rounded to float (2^-24 relative each), D + 1 <= 9
//    roundings of the chain, the rounding of |p|^2/2: below 12 * 2^-24 (|q|^2 + |p|^2) — so a rejected point is
//    beyond thr in exact arithmetic (and thr already pads the reference's own rounding).  The rejection test of a
//    group of 8 points is one OR over the sign bits of its accumulators;
//  * survivors are not inserted on the spot: a thread pushes them (tile position, query slot) to its own small
//    queue in shared memory, and the queues are drained together — at the end of a tile, or when one is about to
//    overflow — so that the exact evaluation (the reference's unfused FP64 arithmetic on the original coordinates,
//    kept beside the float tile) and the ordered insertion run with most lanes of a warp busy.  A queue is
//    drained in push order, i.e. ascending insertion index per list, which the tie rule relies on.
// While a list is still filling (the first k points of a chunk) every point is a survivor.
// ---------------------------------------------------------------------------
#ifndef BKS_QT
#define BKS_QT 2
#endif
#ifndef BKS_TILE
#define BKS_TILE 256  // points per tile: 13.8 ms against 14.9 ms with 128 (half the barriers and queue drains)
#endif
#define BKS_QCAP 8
#ifndef BKS_G
#define BKS_G 16  // points per group (8 pair records): measured 14.8 ms against 16.6 ms with 8 (1e5 x 2e5, 7-D)
#endif
__host__ __device__ constexpr int bks_pp(int D) { return (D + 2) & ~1; }  // float2 entries per pair record
__host__ __device__ constexpr size_t bks_smem(int D, int K) {
  return (size_t)(BKS_TILE / 2) * bks_pp(D) * sizeof(float2) + (size_t)BKS_TILE * D * sizeof(double) +
         (size_t)K * BKS_QT * BK_QB * (sizeof(double) + sizeof(uint32_t)) + (size_t)BKS_QCAP * BK_QB * sizeof(uint16_t);
}

template <int D>
__device__ __forceinline__ void bks_consider(const double* __restrict__ q, const double* pt, uint32_t idx, int K,
                                             double* kd, uint32_t* ki, int& cnt, double& worst, double& thr) {
  constexpr int LS = BKS_QT * BK_QB;
  double d2 = 0.0;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    double diff = __dsub_rn(__ldg(q + j), pt[j]);
    d2 = __dadd_rn(d2, __dmul_rn(diff, diff));
  }
  if (cnt == K && d2 > thr) return;
  double dist = __dsqrt_rn(d2);
  if (cnt == K && !(dist < worst)) return;  // an equidistant later point never displaces (GNAT)
  int pos = cnt < K ? cnt : K - 1;
  while (pos > 0 && kd[(pos - 1) * LS] > dist) {
    kd[pos * LS] = kd[(pos - 1) * LS];
    ki[pos * LS] = ki[(pos - 1) * LS];
    --pos;
  }
  kd[pos * LS] = dist;
  ki[pos * LS] = idx;
  if (cnt < K) ++cnt;
  if (cnt == K) {
    worst = kd[(K - 1) * LS];
    thr = __dmul_rn(__dmul_rn(worst, worst), 1.0 + 9.094947017729282e-13);  // conservative (2^-40)
  }
}

template <int D>
__global__ void __launch_bounds__(BK_QB) bk_knn_partial_f32(const double* __restrict__ queries, int64_t nq_total,
                                                            const double* __restrict__ pts, uint64_t M, uint64_t chunk,
                                                            int K, double* __restrict__ part_d,
                                                            uint32_t* __restrict__ part_i) {
  constexpr int PP = bks_pp(D);
  constexpr int LS = BKS_QT * BK_QB;
  static_assert(BKS_G * BKS_QT <= 32, "hit mask is 32 bits");
  extern __shared__ __align__(16) unsigned char smem[];
  float* s_pts = reinterpret_cast<float*>(smem);                                                       // [TILE/2][PP] float2
  double* s_ptd = reinterpret_cast<double*>(smem + (size_t)(BKS_TILE / 2) * PP * sizeof(float2));      // [TILE][D]
  double* s_kd = s_ptd + (size_t)BKS_TILE * D;                                                         // [K][QT][QB]
  uint32_t* s_ki = reinterpret_cast<uint32_t*>(s_kd + (size_t)K * LS);
  uint16_t* s_q = reinterpret_cast<uint16_t*>(s_ki + (size_t)K * LS) + threadIdx.x;                    // [QCAP][QB]
  __shared__ unsigned int s_pnhi;  // high word of the tile's largest |p|^2 (non-negative doubles order as integers)
  const int tid = threadIdx.x;
  const uint64_t p_begin = (uint64_t)blockIdx.y * chunk;
  const uint64_t p_end = p_begin + chunk < M ? p_begin + chunk : M;
  float2 nq2[BKS_QT][D];  // (-q_j, -q_j)
  float2 nthr2[BKS_QT];   // (-t, -t)
  double qn[BKS_QT];      // |q|^2
  double worst[BKS_QT], thr[BKS_QT], pad[BKS_QT];
  int cnt[BKS_QT];        // the same in every thread and slot: min(points scanned, K)
  const double* qptr[BKS_QT];
#pragma unroll
  for (int s = 0; s < BKS_QT; ++s) {
    int64_t qi = (int64_t)blockIdx.x * LS + s * BK_QB + tid;
    if (qi >= nq_total) qi = 0;  // out-of-range slots scan as query 0 and are not written
    qptr[s] = queries + qi * D;
    qn[s] = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const double v = qptr[s][j];
      const float f = -(float)v;
      nq2[s][j] = make_float2(f, f);
      qn[s] = fma(v, v, qn[s]);
    }
    cnt[s] = 0;
    worst[s] = DBL_MAX;
    thr[s] = DBL_MAX;
    pad[s] = 0.0;
    nthr2[s] = make_float2(0.f, 0.f);
  }
  int qlen = 0;
  uint64_t base = p_begin;  // first point of the current tile (the drain needs it)

  auto refresh = [&]() {
#pragma unroll
    for (int s = 0; s < BKS_QT; ++s) {
      const float t = -__double2float_ru(0.5 * (thr[s] - qn[s]) + pad[s]);
      nthr2[s] = make_float2(t, t);
    }
  };
  auto consider_slot = [&](int s, uint32_t p) {
#pragma unroll
    for (int t = 0; t < BKS_QT; ++t)
      if (t == s)
        bks_consider<D>(qptr[t], s_ptd + p * D, (uint32_t)(base + p), K, s_kd + t * BK_QB + tid, s_ki + t * BK_QB + tid,
                        cnt[t], worst[t], thr[t]);
  };
  auto drain = [&]() {  // every lane empties its queue, oldest entry first
    int i = 0;
    while (__any_sync(0xffffffffu, i < qlen)) {
      if (i < qlen) {
        const uint32_t code = s_q[i * BK_QB];
        consider_slot((int)(code >> 12), code & 0xfffu);
      }
      ++i;
    }
    qlen = 0;
    refresh();
  };

  for (; base < p_end; base += BKS_TILE) {
    const uint32_t tn = (uint32_t)((p_end - base) < BKS_TILE ? (p_end - base) : BKS_TILE);
    const uint32_t tn_pad = (tn + BKS_G - 1) & ~(uint32_t)(BKS_G - 1);
    __syncthreads();  // every thread has drained the previous tile
    if (tid == 0) s_pnhi = 0u;
    __syncthreads();
    unsigned int pnhi = 0u;
    for (uint32_t p = tid; p < tn_pad; p += BK_QB) {
      float* rec = s_pts + (size_t)(p >> 1) * 2 * PP + (p & 1u);
      if (p < tn) {
        double pn = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) {
          const double v = pts[(base + p) * D + j];
          s_ptd[p * D + j] = v;
          rec[2 * j] = (float)v;
          pn = fma(v, v, pn);
        }
        rec[2 * D] = (float)(0.5 * pn);
        pnhi = max(pnhi, (unsigned int)__double2hiint(pn));  // NaN / inf high words exceed every finite one
      } else {  // padding up to a whole group: a point that is never a survivor
#pragma unroll
        for (int j = 0; j < D; ++j) rec[2 * j] = 0.f;
        rec[2 * D] = __int_as_float(0x7f800000);
      }
    }
    pnhi = __reduce_max_sync(0xffffffffu, pnhi);
    if ((tid & 31) == 0) atomicMax(&s_pnhi, pnhi);
    __syncthreads();
    // upper bound of max |p|^2: the next high word (non-finite stays non-finite: every point is then evaluated exactly)
    const double pnmax = __hiloint2double((int)(s_pnhi < 0x7ff00000u ? s_pnhi + 1u : 0x7ff80000u), 0);
#pragma unroll
    for (int s = 0; s < BKS_QT; ++s) pad[s] = 9.5367431640625e-07 * (qn[s] + pnmax);  // 2^-20
    refresh();
    for (uint32_t p = 0; p < tn; p += BKS_G) {
      uint32_t hits = 0;
      if (cnt[0] < K) {  // lists still filling (block-uniform): every real point of the group is a survivor
#pragma unroll
        for (int u = 0; u < BKS_G; ++u)
          if (p + u < tn) hits |= ((1u << BKS_QT) - 1u) << (u * BKS_QT);
      } else {
        const float4* g4 = reinterpret_cast<const float4*>(s_pts) + (size_t)(p >> 1) * (PP / 2);
        float2 acc[BKS_G / 2][BKS_QT];
        uint32_t sign = 0;
#pragma unroll
        for (int u = 0; u < BKS_G / 2; ++u) {
          float2 x2[PP];
#pragma unroll
          for (int h = 0; h < PP / 2; ++h) {
            const float4 t = g4[u * (PP / 2) + h];
            x2[2 * h] = make_float2(t.x, t.y);
            x2[2 * h + 1] = make_float2(t.z, t.w);
          }
#pragma unroll
          for (int s = 0; s < BKS_QT; ++s) {
            float2 a = __fadd2_rn(x2[D], nthr2[s]);
#pragma unroll
            for (int j = 0; j < D; ++j) a = __ffma2_rn(nq2[s][j], x2[j], a);
            acc[u][s] = a;
            sign |= __float_as_uint(a.x) | __float_as_uint(a.y);
          }
        }
        if (sign >> 31) {  // some accumulator is not positive: collect the survivors of this group
#pragma unroll
          for (int u = 0; u < BKS_G / 2; ++u)
#pragma unroll
            for (int s = 0; s < BKS_QT; ++s) {
              if (!(acc[u][s].x > 0.f)) hits |= 1u << ((2 * u) * BKS_QT + s);
              if (!(acc[u][s].y > 0.f)) hits |= 1u << ((2 * u + 1) * BKS_QT + s);
            }
        }
      }
      if (__any_sync(0xffffffffu, hits != 0u)) {
        const int nh = __popc(hits);
        if (__any_sync(0xffffffffu, qlen + nh > BKS_QCAP)) drain();
        if (nh > BKS_QCAP) {  // more than a queue holds (lists filling): evaluate this group in place
          while (hits) {
            const int b = __ffs(hits) - 1;
            hits &= hits - 1;
            consider_slot(b % BKS_QT, p + (uint32_t)(b / BKS_QT));
          }
          refresh();
        } else {
          while (hits) {  // ascending bit = ascending point, so a list's candidates stay in index order
            const int b = __ffs(hits) - 1;
            hits &= hits - 1;
            s_q[qlen * BK_QB] = (uint16_t)(((uint32_t)(b % BKS_QT) << 12) | (p + (uint32_t)(b / BKS_QT)));
            ++qlen;
          }
        }
      }
    }
    drain();
  }
#pragma unroll
  for (int s = 0; s < BKS_QT; ++s) {
    const int64_t qi = (int64_t)blockIdx.x * LS + s * BK_QB + tid;
    if (qi >= nq_total) continue;
    double* od = part_d + ((size_t)blockIdx.y * nq_total + qi) * K;
    uint32_t* oi = part_i + ((size_t)blockIdx.y * nq_total + qi) * K;
    for (int k = 0; k < K; ++k) {
      od[k] = k < cnt[s] ? s_kd[k * LS + s * BK_QB + tid] : DBL_MAX;
      oi[k] = k < cnt[s] ? s_ki[k * LS + s * BK_QB + tid] : 0xFFFFFFFFu;
    }
  }
}

// Merge the P partial lists of each query and evaluate KNNModel::estimate (KNNModel.hpp:101-134).
__global__ void __launch_bounds__(128) bk_merge_estimate(int64_t nq, int P, int K, uint64_t M,
                                                         const double* __restrict__ part_d,
                                                         const uint32_t* __restrict__ part_i,
                                                         const double* __restrict__ vals, double prior,
                                                         double prior_weight, double* __restrict__ out) {
  for (int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; qi < nq; qi += (int64_t)gridDim.x * blockDim.x) {
    if (M == 0) {  // KNNModel.hpp:104-106
      out[qi] = prior;
      continue;
    }
    int knn = (uint64_t)K < M ? K : (int)M;  // :108
    unsigned char head[64];
    for (int p = 0; p < P; ++p) head[p] = 0;
    double dot = 0.0, wsum = 0.0;
    for (int k = 0; k < knn; ++k) {
      int best = -1;
      double bd = 0.0;
      uint32_t bi = 0;
      for (int p = 0; p < P; ++p) {
        if (head[p] >= K) continue;
        size_t o = ((size_t)p * nq + qi) * K + head[p];
        uint32_t ci = part_i[o];
        if (ci == 0xFFFFFFFFu) continue;
        double cd = part_d[o];
        if (best < 0 || cd < bd || (cd == bd && ci < bi)) { best = p; bd = cd; bi = ci; }
      }
      if (best < 0) break;
      head[best]++;
      double distance = fmax(0.0001, bd);       // :121
      double w = __ddiv_rn(1.0, distance);      // :122
      double v = vals[bi];                      // :123
      dot = __dadd_rn(dot, __dmul_rn(w, v));    // weights.dot(values), sequential
      wsum = __dadd_rn(wsum, w);                // weights.sum()
    }
    double result = __ddiv_rn(dot, wsum);                                                   // :127
    result = __ddiv_rn(__dadd_rn(result, __dmul_rn(prior_weight, prior)), __dadd_rn(1.0, prior_weight));  // :130
    out[qi] = result;
  }
}

// pfree[e] = prod (1 - est) over the edge's queries, in slot order (BP.cpp:473-488)
__global__ void __launch_bounds__(256) bk_product_kernel(int64_t count, const uint64_t* __restrict__ offs,
                                                         const double* __restrict__ est, double* __restrict__ pfree,
                                                         uint32_t* __restrict__ nlook) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x) {
    uint64_t o = offs[e], c = offs[e + 1] - o;
    double r = 1.0;
    for (uint64_t k = 0; k < c; ++k) r = __dmul_rn(r, __dsub_rn(1.0, est[o + k]));
    pfree[e] = r;
    if (nlook) nlook[e] = (uint32_t)c;
  }
}

// ---------------------------------------------------------------------------
// Small batches (the planner's per-search calls): computeAndSetEdgeFreeProbability of a batch of edges in ONE
// launch.  A block takes an edge: nStates, the query slots 1, 1+step, ... (BP.cpp:472-475), and one warp per
// query: every lane scans the belief points lane, lane+32, ... into its own (distance, index)-ordered list of
// the k best (shared memory), the 32 lists are merged k times by a warp arg-min on (distance, index) — the
// same k nearest, in the same order, as the sequential scan of KNNModel::estimate (KNNModel.hpp:108-127) —
// lane 0 accumulates the weights in that order; thread 0 forms the product over the queries.
// ---------------------------------------------------------------------------
#define BKF_WARPS 4
#define BKF_QMAX 16  // queries per edge: (n-2)/step + 1 <= 12 for n <= 65535

template <int D>
__device__ __forceinline__ double bkf_estimate_warp(const double* x, const double* __restrict__ pts,
                                                    const double* __restrict__ vals, uint64_t M, int K, double prior,
                                                    double prior_weight, double* kd, uint32_t* ki, int lane) {
  if (M == 0) return prior;  // KNNModel.hpp:104-106
  int cnt = 0;
  double worst = DBL_MAX, thr = DBL_MAX;
  for (uint64_t p = (uint64_t)lane; p < M; p += 32) {
    const double* pt = pts + p * D;
    double d2 = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const double diff = __dsub_rn(x[j], __ldg(pt + j));
      d2 = __dadd_rn(d2, __dmul_rn(diff, diff));
    }
    if (cnt == K && d2 > thr) continue;
    const double dist = __dsqrt_rn(d2);
    if (cnt == K && !(dist < worst)) continue;  // an equidistant later point never displaces
    int pos = cnt < K ? cnt : K - 1;
    while (pos > 0 && kd[(pos - 1) * 32] > dist) {
      kd[pos * 32] = kd[(pos - 1) * 32];
      ki[pos * 32] = ki[(pos - 1) * 32];
      --pos;
    }
    kd[pos * 32] = dist;
    ki[pos * 32] = (uint32_t)p;
    if (cnt < K) ++cnt;
    if (cnt == K) {
      worst = kd[(K - 1) * 32];
      thr = __dmul_rn(__dmul_rn(worst, worst), 1.0 + 9.094947017729282e-13);  // conservative (2^-40)
    }
  }
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
    if (mine == ci && ci != 0xFFFFFFFFu) ++head;  // the winner's list advances (indices are unique)
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

template <int D>
__global__ void __launch_bounds__(BKF_WARPS * 32) bk_pfree_fused(EdgeSrc src, int64_t count, double L, PermView pv,
                                                                 const double* __restrict__ pts,
                                                                 const double* __restrict__ vals, uint64_t M, int K,
                                                                 double prior, double prior_weight,
                                                                 double* __restrict__ pfree, uint32_t* __restrict__ nlook,
                                                                 BelGridView bg) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* kd = reinterpret_cast<double*>(smem) + (size_t)warp * K * 32 + lane;                       // [K][32] per warp
  uint32_t* ki = reinterpret_cast<uint32_t*>(smem + (size_t)BKF_WARPS * K * 32 * sizeof(double)) + (size_t)warp * K * 32 + lane;
  __shared__ double s_est[BKF_QMAX];
  for (int64_t e = blockIdx.x; e < count; e += gridDim.x) {
    double f[D], d[D];
    bk_load<D>(src, e, f, d);
    uint32_t n = bp_nstates(__dsqrt_rn(bp_dist2<D>(f, d)), L);
#pragma unroll
    for (int j = 0; j < D; ++j) d[j] = __dsub_rn(d[j], f[j]);
    if (n == 0u) {  // more than BPOMP_NSTATES_MAX states
      if (threadIdx.x == 0) atomicAdd(pv.flags + 1, 1);
      n = 2u;
    }
    uint32_t nqr = 0, step = 1, off = 0;
    if (n > 2u) {
      off = __ldg(pv.off + n);
      if (off == BP_ABSENT) {  // sample order not resident: the host uploads it and redoes the batch
        if (threadIdx.x == 0) {
          pv.need[n] = 1;
          atomicAdd(pv.flags, 1);
        }
      } else {
        step = __ldg(pv.step + n);       // (unsigned)(n / log(n)), BP.cpp:472
        nqr = (n - 2u + step - 1u) / step;  // i = 1, 1+step, ... < n-1
      }
    }
    for (uint32_t qi = (uint32_t)warp; qi < nqr; qi += BKF_WARPS) {
      const double tt = __ldg(pv.tfrac + off + 1u + qi * step);
      double x[D];
#pragma unroll
      for (int j = 0; j < D; ++j) x[j] = __dadd_rn(f[j], __dmul_rn(d[j], tt));
      double est;
      if constexpr (D <= 3) {
        est = bg.Mg > 0 ? bg_estimate_warp<D>(x, bg, pts, vals, M, K, prior, prior_weight, lane)
                        : bkf_estimate_warp<D>(x, pts, vals, M, K, prior, prior_weight, kd, ki, lane);
      } else {
        est = bkf_estimate_warp<D>(x, pts, vals, M, K, prior, prior_weight, kd, ki, lane);
      }
      if (lane == 0) s_est[qi] = est;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double r = 1.0;
      for (uint32_t k = 0; k < nqr; ++k) r = __dmul_rn(r, __dsub_rn(1.0, s_est[k]));  // BP.cpp:473-488, slot order
      pfree[e] = r;
      if (nlook) nlook[e] = nqr;
    }
    __syncthreads();
  }
}

// checkAndSetEdgeBlocked's belief inserts for already-checked edges whose nStates the caller knows: the states
// of slots 1..c (c = first invalid slot, or n-2 for a free edge) go to out + offs[e]*D, values 0.0 and a final
// 1.0 for a blocked edge (BP.cpp:523-537).  offs is the caller's exclusive prefix of the c's.
template <int D>
__global__ void __launch_bounds__(128) bk_append_known(EdgeSrc src, int64_t count, PermView pv,
                                                       const int32_t* __restrict__ first_invalid,
                                                       const uint32_t* __restrict__ nst,
                                                       const uint64_t* __restrict__ offs, double* __restrict__ out,
                                                       double* __restrict__ out_vals) {
  // a warp per edge, lane = slot
  const int lane = threadIdx.x & 31;
  for (int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < count;
       e += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const uint32_t n = nst[e];
    if (n <= 2u) continue;
    const int32_t fi = first_invalid[e];
    const uint32_t c = fi >= 1 ? (uint32_t)fi : n - 2u;
    const uint64_t o = offs[e];
    double f[D], t[D];
    bk_load<D>(src, e, f, t);
    const double* tf = pv.tfrac + __ldg(pv.off + n);
    for (uint32_t k = (uint32_t)lane; k < c; k += 32) {
      const double tt = __ldg(tf + 1u + k);
#pragma unroll
      for (int j = 0; j < D; ++j) out[(o + k) * D + j] = __dadd_rn(f[j], __dmul_rn(__dsub_rn(t[j], f[j]), tt));
      out_vals[o + k] = (fi >= 1 && k + 1 == c) ? 1.0 : 0.0;
    }
  }
}

// KNNModel::estimate of a query array through the index: a warp per query, the k best in registers.
#define BKG_WARPS 8
template <int D>
__global__ void __launch_bounds__(BKG_WARPS * 32) bk_estimate_grid(const double* __restrict__ queries, int64_t nq,
                                                                   const double* __restrict__ pts,
                                                                   const double* __restrict__ vals, uint64_t M, int K,
                                                                   double prior, double prior_weight,
                                                                   double* __restrict__ out, BelGridView bg) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t qi = (int64_t)blockIdx.x * BKG_WARPS + warp; qi < nq; qi += (int64_t)gridDim.x * BKG_WARPS) {
    double x[D];
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = __ldg(queries + qi * D + j);
    const double est = bg_estimate_warp<D>(x, bg, pts, vals, M, K, prior, prior_weight, lane);
    if (lane == 0) out[qi] = est;
  }
}

// Brings the grid index of the belief model up to date (D <= 3) and describes it in `g` (g.Mg = 0: no index,
// the callers scan).  The index covers the prefix [0, bg_count); it is rebuilt once the tail of newer points
// exceeds 512 + bg_count/16, i.e. after a constant fraction of appends, all on the context's stream.
template <int D>
static int bel_grid_prepare(bpomp_ctx* ctx, BelGridView& g) {
  memset(&g, 0, sizeof(g));
  if constexpr (D > 3) {
    return BPOMP_OK;
  } else {
    const uint64_t M = ctx->bel_count;
    if (!ctx->belief_grid || M < 512) return BPOMP_OK;
    if (ctx->bg_count > M) ctx->bg_count = 0;
    if (ctx->bg_count == 0 || M - ctx->bg_count > 512 + ctx->bg_count / 16) {
      // about two points per cell, 2^bits cells per dimension
      const uint32_t bmax = D == 1 ? 20u : D == 2 ? 10u : 6u;
      uint32_t bits = 1;
      while (bits < bmax && std::pow(2.0, (double)(D * bits)) * 2.0 < (double)M) ++bits;
      const uint32_t G = 1u << bits;
      const uint64_t cells = 1ull << (D * bits);
      BP_TRY(bp_reserve(ctx, ctx->d_bg_start, (cells + 1) * sizeof(unsigned long long), false));
      BP_TRY(bp_reserve(ctx, ctx->d_bg_cnt, cells * sizeof(uint32_t), false));
      BP_TRY(bp_reserve(ctx, ctx->d_bg_cell, M * sizeof(uint32_t), false));
      BP_TRY(bp_reserve(ctx, ctx->d_bg_pts, M * D * sizeof(double), false));
      BP_TRY(bp_reserve(ctx, ctx->d_bg_idx, M * sizeof(uint32_t), false));
      BP_TRY(bp_reserve(ctx, ctx->d_bg_geom, 9 * sizeof(double), false));
      const double* pts = ctx->d_bel_pts.as<double>();
      BP_CUDA(ctx, cudaMemsetAsync(ctx->d_bg_cnt.p, 0, cells * sizeof(uint32_t), ctx->stream));
      bg_bbox_kernel<D><<<1, 1024, 0, ctx->stream>>>(pts, M, G, ctx->d_bg_geom.as<double>());
      const int grid = (int)std::min<uint64_t>((M + 255) / 256, (uint64_t)ctx->num_sms * 8);
      bg_hist_kernel<D><<<grid, 256, 0, ctx->stream>>>(pts, M, bits, ctx->d_bg_geom.as<double>(), ctx->d_bg_cell.as<uint32_t>(),
                                                       ctx->d_bg_cnt.as<uint32_t>());
      ctx->launches += 2;
      BP_TRY(bp_exclusive_scan_u32(ctx, ctx->d_bg_cnt.as<uint32_t>(), cells, ctx->d_bg_start.as<uint64_t>(), ctx->s_scan));
      bg_scatter_kernel<D><<<grid, 256, 0, ctx->stream>>>(pts, M, ctx->d_bg_cell.as<uint32_t>(), ctx->d_bg_cnt.as<uint32_t>(),
                                                          ctx->d_bg_start.as<unsigned long long>(), ctx->d_bg_pts.as<double>(),
                                                          ctx->d_bg_idx.as<uint32_t>());
      ctx->launches++;
      BP_CUDA(ctx, cudaGetLastError());
      ctx->bg_count = M;
      ctx->bg_bits = bits;
      ctx->bg_builds++;
    }
    g.start = ctx->d_bg_start.as<unsigned long long>();
    g.pts = ctx->d_bg_pts.as<double>();
    g.idx = ctx->d_bg_idx.as<uint32_t>();
    g.geom = ctx->d_bg_geom.as<double>();
    g.bits = ctx->bg_bits;
    g.Mg = ctx->bg_count;
    return BPOMP_OK;
  }
}

// ---------------------------------------------------------------------------
// Host drivers
// ---------------------------------------------------------------------------
template <int D>
static int knn_estimate_dev(bpomp_ctx* ctx, const double* d_queries, int64_t nq, double* d_out, DevBuf& bd, DevBuf& bi) {
  if (nq <= 0) return BPOMP_OK;
  const uint64_t M = ctx->bel_count;
  const int K = ctx->knn;
  if constexpr (D <= 3) {
    BelGridView bg;
    BP_TRY(bel_grid_prepare<D>(ctx, bg));
    if (bg.Mg > 0) {
      const int grid = (int)std::min<int64_t>((nq + BKG_WARPS - 1) / BKG_WARPS, (int64_t)ctx->num_sms * 16);
      bk_estimate_grid<D><<<grid, BKG_WARPS * 32, 0, ctx->stream>>>(d_queries, nq, ctx->d_bel_pts.as<double>(),
                                                                    ctx->d_bel_vals.as<double>(), M, K, ctx->prior,
                                                                    ctx->prior_weight, d_out, bg);
      ctx->launches++;
      BP_CUDA(ctx, cudaGetLastError());
      return BPOMP_OK;
    }
  }
  int P = 1;
  uint64_t chunk = M ? M : 1;
  if (M > 0) {
    // dimensions >= 4: prefilter on the FP32 pipe, BKS_QT queries per thread; else the FP64 prefilter
    const bool f32 = D >= 4 && ctx->prefilter_f32;
    const int qt = f32 ? BKS_QT : BK_QT;
    int64_t qblocks = (nq + BK_QB * qt - 1) / (BK_QB * qt);
    int64_t want = 2 * (int64_t)ctx->num_sms;
    while (qblocks * P < want && P < 64 && (M + (2 * P) - 1) / (2 * P) >= 1024) P *= 2;
    chunk = (M + P - 1) / P;
    chunk = (chunk + BK_TILE - 1) / BK_TILE * BK_TILE;
    P = (int)((M + chunk - 1) / chunk);
    BP_TRY(bp_reserve(ctx, bd, (size_t)P * nq * K * sizeof(double), false));
    BP_TRY(bp_reserve(ctx, bi, (size_t)P * nq * K * sizeof(uint32_t), false));
    dim3 grid((unsigned)qblocks, (unsigned)P);
    if (f32) {
      size_t sm = bks_smem(D, K);
      auto kern = bk_knn_partial_f32<D>;
      BP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      kern<<<grid, BK_QB, sm, ctx->stream>>>(d_queries, nq, ctx->d_bel_pts.as<double>(), M, chunk, K, bd.as<double>(),
                                             bi.as<uint32_t>());
    } else {
      size_t sm = (size_t)BK_TILE * bk_ds(D) * sizeof(double) + (size_t)K * BK_QT * BK_QB * (sizeof(double) + sizeof(uint32_t));
      auto kern = bk_knn_partial<D>;
      BP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      kern<<<grid, BK_QB, sm, ctx->stream>>>(d_queries, nq, ctx->d_bel_pts.as<double>(), M, chunk, K, bd.as<double>(),
                                             bi.as<uint32_t>());
    }
    ctx->launches++;
    BP_CUDA(ctx, cudaGetLastError());
  }
  int gm = (int)std::min<int64_t>((nq + 127) / 128, (int64_t)ctx->num_sms * 8);
  bk_merge_estimate<<<gm, 128, 0, ctx->stream>>>(nq, P, K, M, bd.as<double>(), bi.as<uint32_t>(),
                                                 ctx->d_bel_vals.as<double>(), ctx->prior, ctx->prior_weight, d_out);
  ctx->launches++;
  BP_CUDA(ctx, cudaGetLastError());
  return BPOMP_OK;
}

static int knn_estimate_dispatch(bpomp_ctx* ctx, const double* d_q, int64_t nq, double* d_out, DevBuf& bd, DevBuf& bi) {
  switch (ctx->dim) {
    case 1: return knn_estimate_dev<1>(ctx, d_q, nq, d_out, bd, bi);
    case 2: return knn_estimate_dev<2>(ctx, d_q, nq, d_out, bd, bi);
    case 3: return knn_estimate_dev<3>(ctx, d_q, nq, d_out, bd, bi);
    case 4: return knn_estimate_dev<4>(ctx, d_q, nq, d_out, bd, bi);
    case 5: return knn_estimate_dev<5>(ctx, d_q, nq, d_out, bd, bi);
    case 6: return knn_estimate_dev<6>(ctx, d_q, nq, d_out, bd, bi);
    case 7: return knn_estimate_dev<7>(ctx, d_q, nq, d_out, bd, bi);
    case 8: return knn_estimate_dev<8>(ctx, d_q, nq, d_out, bd, bi);
  }
  return bp_fail(ctx, BPOMP_ERR_INVALID, "unsupported dim %d", ctx->dim);
}

extern "C" int bpomp_belief_estimate(bpomp_ctx* ctx, const double* queries, int64_t nq, double* out) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (nq < 0 || (nq > 0 && (!queries || !out))) return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  if (nq == 0) return BPOMP_OK;
  size_t bytes = (size_t)nq * ctx->dim * sizeof(double);
  BP_TRY(bp_reserve(ctx, ctx->s_a, bytes, false));
  BP_TRY(bp_reserve(ctx, ctx->s_b, (size_t)nq * sizeof(double), false));
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->s_a.p, queries, bytes, cudaMemcpyHostToDevice, ctx->stream));
  BP_TRY(knn_estimate_dispatch(ctx, ctx->s_a.as<double>(), nq, ctx->s_b.as<double>(), ctx->s_c, ctx->s_d));
  BP_CUDA(ctx, cudaMemcpyAsync(out, ctx->s_b.p, (size_t)nq * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BPOMP_OK;
}

extern "C" int bpomp_belief_estimate_dev(bpomp_ctx* ctx, const double* d_queries, int64_t nq, double* d_out) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (nq < 0 || (nq > 0 && (!d_queries || !d_out))) return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  if (nq == 0) return BPOMP_OK;
  return knn_estimate_dispatch(ctx, d_queries, nq, d_out, ctx->s_c, ctx->s_d);
}

// count pass + scan + (resolve missing sample orders); leaves nst in s_e, offsets in s_g
template <int D>
static int edges_count_scan(bpomp_ctx* ctx, const EdgeSrc& src, int64_t count, int mode, const int32_t* d_first,
                            uint64_t* total) {
  BP_TRY(bp_reserve(ctx, ctx->s_e, (size_t)count * sizeof(uint32_t), false));
  BP_TRY(bp_reserve(ctx, ctx->s_f, (size_t)count * sizeof(uint32_t), false));
  BP_TRY(bp_reserve(ctx, ctx->s_g, ((size_t)count + 1) * sizeof(uint64_t), false));
  int grid = (int)std::min<int64_t>((count + 255) / 256, (int64_t)ctx->num_sms * 8);
  bk_count_kernel<D><<<grid, 256, 0, ctx->stream>>>(src, count, ctx->L, ctx->perm_view(), mode, d_first,
                                                   ctx->s_e.as<uint32_t>(), ctx->s_f.as<uint32_t>());
  ctx->launches++;
  BP_CUDA(ctx, cudaGetLastError());
  BP_TRY(bp_exclusive_scan_u32(ctx, ctx->s_f.as<uint32_t>(), (uint64_t)count, ctx->s_g.as<uint64_t>(), ctx->s_scan));
  BP_CUDA(ctx, cudaMemcpyAsync(total, ctx->s_g.as<uint64_t>() + count, sizeof(uint64_t), cudaMemcpyDeviceToHost,
                               ctx->stream));
  BP_TRY(bp_read_flags(ctx));
  if (ctx->h_flags[1]) {
    ctx->h_flags[1] = 0;
    ctx->h_flags[0] = 0;
    return bp_fail(ctx, BPOMP_ERR_RANGE, "an edge has more than %u states (segment_length too small)",
                   BPOMP_NSTATES_MAX);
  }
  if (ctx->h_flags[0]) {
    ctx->h_flags[0] = 0;
    std::vector<uint8_t> need((size_t)BPOMP_NSTATES_MAX + 1);
    BP_CUDA(ctx, cudaMemcpyAsync(need.data(), ctx->d_need.p, need.size(), cudaMemcpyDeviceToHost, ctx->stream));
    BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    BP_CUDA(ctx, cudaMemsetAsync(ctx->d_need.p, 0, need.size(), ctx->stream));
    std::vector<uint32_t> ns;
    for (uint32_t n = 2; n <= BPOMP_NSTATES_MAX; ++n)
      if (need[n]) ns.push_back(n);
    BP_TRY(bp_perm_ensure_rows(ctx, ns));
  }
  return BPOMP_OK;
}

template <int D>
static int edge_pfree_impl(bpomp_ctx* ctx, const EdgeSrc& src, int64_t count, double* pfree, uint32_t* nlookups) {
  // small batches (a search's worth of edges against a model of moderate size): one launch, one wait
  BelGridView bg;
  BP_TRY(bel_grid_prepare<D>(ctx, bg));
  // with the index a query costs the same whatever the size of the model
  if (ctx->fused_pfree && (bg.Mg > 0 || (double)count * 4.0 * (double)std::max<uint64_t>(ctx->bel_count, 1) <= 4e8)) {
    const int K = ctx->knn;
    BP_TRY(bp_reserve(ctx, ctx->s_h, (size_t)count * (sizeof(double) + sizeof(uint32_t)), false));
    double* d_pf = ctx->s_h.as<double>();
    uint32_t* d_nl = reinterpret_cast<uint32_t*>(d_pf + count);
    const size_t sm = (size_t)BKF_WARPS * K * 32 * (sizeof(double) + sizeof(uint32_t));
    auto kern = bk_pfree_fused<D>;
    BP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const int grid = (int)std::min<int64_t>(count, (int64_t)ctx->num_sms * 16);
    kern<<<grid, BKF_WARPS * 32, sm, ctx->stream>>>(src, count, ctx->L, ctx->perm_view(), ctx->d_bel_pts.as<double>(),
                                                    ctx->d_bel_vals.as<double>(), ctx->bel_count, K, ctx->prior,
                                                    ctx->prior_weight, d_pf, d_nl, bg);
    ctx->launches++;
    BP_CUDA(ctx, cudaGetLastError());
    BP_CUDA(ctx, cudaMemcpyAsync(pfree, d_pf, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (nlookups)
      BP_CUDA(ctx, cudaMemcpyAsync(nlookups, d_nl, (size_t)count * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    BP_TRY(bp_read_flags(ctx));  // one wait for the results and the flags
    if (ctx->h_flags[1]) {
      ctx->h_flags[1] = 0;
      ctx->h_flags[0] = 0;
      return bp_fail(ctx, BPOMP_ERR_RANGE, "an edge has more than %u states (segment_length too small)", BPOMP_NSTATES_MAX);
    }
    if (!ctx->h_flags[0]) return BPOMP_OK;
    // some sample orders were missing: the general path below uploads them and redoes the batch
  }
  uint64_t nq = 0;
  BP_TRY(edges_count_scan<D>(ctx, src, count, 0, nullptr, &nq));
  int grid = (int)std::min<int64_t>((count + 255) / 256, (int64_t)ctx->num_sms * 8);
  // query states -> s_c, estimates -> s_d, kNN partial lists -> s_i / s_j, results -> s_h
  BP_TRY(bp_reserve(ctx, ctx->s_c, (size_t)(nq + 1) * D * sizeof(double), false));
  BP_TRY(bp_reserve(ctx, ctx->s_d, (size_t)(nq + 1) * sizeof(double), false));
  if (nq > 0) {
    bk_expand_kernel<D><<<grid, 256, 0, ctx->stream>>>(src, count, ctx->perm_view(), 0, nullptr, ctx->s_e.as<uint32_t>(),
                                                      ctx->s_g.as<uint64_t>(), ctx->s_c.as<double>(), nullptr);
    ctx->launches++;
    BP_CUDA(ctx, cudaGetLastError());
    BP_TRY(knn_estimate_dispatch(ctx, ctx->s_c.as<double>(), (int64_t)nq, ctx->s_d.as<double>(), ctx->s_i, ctx->s_j));
  }
  BP_TRY(bp_reserve(ctx, ctx->s_h, (size_t)count * (sizeof(double) + sizeof(uint32_t)), false));
  double* d_pf = ctx->s_h.as<double>();
  uint32_t* d_nl = reinterpret_cast<uint32_t*>(d_pf + count);
  bk_product_kernel<<<grid, 256, 0, ctx->stream>>>(count, ctx->s_g.as<uint64_t>(), ctx->s_d.as<double>(), d_pf, d_nl);
  ctx->launches++;
  BP_CUDA(ctx, cudaGetLastError());
  BP_CUDA(ctx, cudaMemcpyAsync(pfree, d_pf, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (nlookups)
    BP_CUDA(ctx, cudaMemcpyAsync(nlookups, d_nl, (size_t)count * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BPOMP_OK;
}

template <int D>
static int append_checked_impl(bpomp_ctx* ctx, const EdgeSrc& src, int64_t count, const int32_t* d_first) {
  uint64_t total = 0;
  BP_TRY(edges_count_scan<D>(ctx, src, count, 1, d_first, &total));
  if (total == 0) return BPOMP_OK;
  BP_TRY(belief_grow(ctx, total));
  int grid = (int)std::min<int64_t>((count + 255) / 256, (int64_t)ctx->num_sms * 8);
  bk_expand_kernel<D><<<grid, 256, 0, ctx->stream>>>(src, count, ctx->perm_view(), 1, d_first, ctx->s_e.as<uint32_t>(),
                                                    ctx->s_g.as<uint64_t>(),
                                                    ctx->d_bel_pts.as<double>() + ctx->bel_count * D,
                                                    ctx->d_bel_vals.as<double>() + ctx->bel_count);
  ctx->launches++;
  BP_CUDA(ctx, cudaGetLastError());
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->bel_count += total;
  return BPOMP_OK;
}

#define BK_DISPATCH(fn, ...)                                           \
  switch (ctx->dim) {                                                  \
    case 1: return fn<1>(__VA_ARGS__);                                 \
    case 2: return fn<2>(__VA_ARGS__);                                 \
    case 3: return fn<3>(__VA_ARGS__);                                 \
    case 4: return fn<4>(__VA_ARGS__);                                 \
    case 5: return fn<5>(__VA_ARGS__);                                 \
    case 6: return fn<6>(__VA_ARGS__);                                 \
    case 7: return fn<7>(__VA_ARGS__);                                 \
    case 8: return fn<8>(__VA_ARGS__);                                 \
  }                                                                    \
  return bp_fail(ctx, BPOMP_ERR_INVALID, "unsupported dim %d", ctx->dim)

static int upload_edges(bpomp_ctx* ctx, const double* from, const double* to, const uint32_t* fi, const uint32_t* ti,
                        int64_t count, EdgeSrc& src) {
  memset(&src, 0, sizeof(src));
  src.verts = ctx->d_verts.as<double>();
  if (fi) {
    for (int64_t e = 0; e < count; ++e)
      if (fi[e] >= ctx->nverts || ti[e] >= ctx->nverts)
        return bp_fail(ctx, BPOMP_ERR_RANGE, "edge %lld references a vertex id out of range", (long long)e);
    size_t b = (size_t)count * sizeof(uint32_t);
    BP_TRY(bp_reserve(ctx, ctx->s_a, b, false));
    BP_TRY(bp_reserve(ctx, ctx->s_b, b, false));
    BP_CUDA(ctx, cudaMemcpyAsync(ctx->s_a.p, fi, b, cudaMemcpyHostToDevice, ctx->stream));
    BP_CUDA(ctx, cudaMemcpyAsync(ctx->s_b.p, ti, b, cudaMemcpyHostToDevice, ctx->stream));
    src.from_ids = ctx->s_a.as<uint32_t>();
    src.to_ids = ctx->s_b.as<uint32_t>();
  } else {
    size_t b = (size_t)count * ctx->dim * sizeof(double);
    BP_TRY(bp_reserve(ctx, ctx->s_a, b, false));
    BP_TRY(bp_reserve(ctx, ctx->s_b, b, false));
    BP_CUDA(ctx, cudaMemcpyAsync(ctx->s_a.p, from, b, cudaMemcpyHostToDevice, ctx->stream));
    BP_CUDA(ctx, cudaMemcpyAsync(ctx->s_b.p, to, b, cudaMemcpyHostToDevice, ctx->stream));
    src.from = ctx->s_a.as<double>();
    src.to = ctx->s_b.as<double>();
  }
  return BPOMP_OK;
}

static int edge_pfree_dispatch(bpomp_ctx* ctx, const EdgeSrc& src, int64_t count, double* pfree, uint32_t* nl) {
  BK_DISPATCH(edge_pfree_impl, ctx, src, count, pfree, nl);
}
static int append_checked_dispatch(bpomp_ctx* ctx, const EdgeSrc& src, int64_t count, const int32_t* d_first) {
  BK_DISPATCH(append_checked_impl, ctx, src, count, d_first);
}

extern "C" int bpomp_belief_edge_pfree(bpomp_ctx* ctx, const double* from, const double* to, int64_t count,
                                       double* pfree, uint32_t* nlookups) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (count < 0 || (count > 0 && (!from || !to || !pfree))) return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  if (count == 0) return BPOMP_OK;
  EdgeSrc src;
  BP_TRY(upload_edges(ctx, from, to, nullptr, nullptr, count, src));
  return edge_pfree_dispatch(ctx, src, count, pfree, nlookups);
}

extern "C" int bpomp_belief_edge_pfree_indexed(bpomp_ctx* ctx, const uint32_t* from_ids, const uint32_t* to_ids,
                                               int64_t count, double* pfree, uint32_t* nlookups) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (count < 0 || (count > 0 && (!from_ids || !to_ids || !pfree)))
    return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  if (count == 0) return BPOMP_OK;
  EdgeSrc src;
  BP_TRY(upload_edges(ctx, nullptr, nullptr, from_ids, to_ids, count, src));
  return edge_pfree_dispatch(ctx, src, count, pfree, nlookups);
}

template <int D>
static int append_known_impl(bpomp_ctx* ctx, const EdgeSrc& src, int64_t count, const int32_t* d_first,
                             const uint32_t* d_nst, const uint64_t* d_offs, uint64_t total) {
  BP_TRY(belief_grow(ctx, total));
  const int grid = (int)std::min<int64_t>((count + 3) / 4, (int64_t)ctx->num_sms * 8);
  bk_append_known<D><<<grid, 128, 0, ctx->stream>>>(src, count, ctx->perm_view(), d_first, d_nst, d_offs,
                                                    ctx->d_bel_pts.as<double>() + ctx->bel_count * D,
                                                    ctx->d_bel_vals.as<double>() + ctx->bel_count);
  ctx->launches++;
  BP_CUDA(ctx, cudaGetLastError());
  ctx->bel_count += total;
  return BPOMP_OK;
}
static int append_known_dispatch(bpomp_ctx* ctx, const EdgeSrc& src, int64_t count, const int32_t* d_first,
                                 const uint32_t* d_nst, const uint64_t* d_offs, uint64_t total) {
  BK_DISPATCH(append_known_impl, ctx, src, count, d_first, d_nst, d_offs, total);
}

extern "C" int bpomp_belief_append_checked_indexed(bpomp_ctx* ctx, const uint32_t* from_ids, const uint32_t* to_ids,
                                                   const int32_t* first_invalid_slot, const uint32_t* nstates,
                                                   int64_t count) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (count < 0 || (count > 0 && (!from_ids || !to_ids || !first_invalid_slot || !nstates)))
    return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  if (count == 0) return BPOMP_OK;
  std::vector<uint64_t> offs((size_t)count + 1, 0);
  uint32_t nmax = 0;
  for (int64_t e = 0; e < count; ++e) {
    const uint32_t n = nstates[e];
    if (n < 2u || n > BPOMP_NSTATES_MAX) return bp_fail(ctx, BPOMP_ERR_RANGE, "nstates[%lld] = %u is out of range", (long long)e, n);
    const int32_t fi = first_invalid_slot[e];
    if (fi >= 1 && (uint32_t)fi > n - 2u) return bp_fail(ctx, BPOMP_ERR_RANGE, "first_invalid_slot[%lld] exceeds nstates - 2", (long long)e);
    offs[(size_t)e + 1] = offs[(size_t)e] + (n <= 2u ? 0u : (fi >= 1 ? (uint32_t)fi : n - 2u));
    nmax = std::max(nmax, n);
  }
  const uint64_t total = offs[(size_t)count];
  if (total == 0) return BPOMP_OK;
  if (nmax > ctx->perm_reserved) {  // make the sample orders of every nStates in the batch resident
    std::vector<uint32_t> ns(nstates, nstates + count);
    std::sort(ns.begin(), ns.end());
    ns.erase(std::unique(ns.begin(), ns.end()), ns.end());
    BP_TRY(bp_perm_ensure_rows(ctx, ns));
  }
  EdgeSrc src;
  BP_TRY(upload_edges(ctx, nullptr, nullptr, from_ids, to_ids, count, src));
  BP_TRY(bp_reserve(ctx, ctx->s_c, (size_t)count * sizeof(int32_t), false));
  BP_TRY(bp_reserve(ctx, ctx->s_e, (size_t)count * sizeof(uint32_t), false));
  BP_TRY(bp_reserve(ctx, ctx->s_g, ((size_t)count + 1) * sizeof(uint64_t), false));
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->s_c.p, first_invalid_slot, (size_t)count * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->s_e.p, nstates, (size_t)count * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->s_g.p, offs.data(), ((size_t)count + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  BP_TRY(append_known_dispatch(ctx, src, count, ctx->s_c.as<int32_t>(), ctx->s_e.as<uint32_t>(), ctx->s_g.as<uint64_t>(), total));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `offs` is a host temporary
  return BPOMP_OK;
}

extern "C" int bpomp_belief_append_checked(bpomp_ctx* ctx, const double* from, const double* to,
                                           const int32_t* first_invalid_slot, int64_t count) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (count < 0 || (count > 0 && (!from || !to || !first_invalid_slot)))
    return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  if (count == 0) return BPOMP_OK;
  EdgeSrc src;
  BP_TRY(upload_edges(ctx, from, to, nullptr, nullptr, count, src));
  BP_TRY(bp_reserve(ctx, ctx->s_c, (size_t)count * sizeof(int32_t), false));
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->s_c.p, first_invalid_slot, (size_t)count * sizeof(int32_t), cudaMemcpyHostToDevice,
                               ctx->stream));
  return append_checked_dispatch(ctx, src, count, ctx->s_c.as<int32_t>());
}
