// neighbors.cu — r-disk neighbour generation ("densification"): replaces
// ompl::NearestNeighborsGNAT<Vertex>::nearestR at /root/reference/src/BatchingPOMP.cpp:150.
//
// Grid-binned radius search emitting a device CSR:
//   1. bin the ACTIVE vertices on a uniform grid over the first <=3 dimensions (cell >= r where
//      the cell budget allows; one cell == brute force when r is huge, e.g. vertex batching's
//      r = DBL_MAX, VertexBatching.hpp:71), points stored SoA in cell order;
//   2. count pass  (S warps per query, ballot/popc),   3. exclusive scan -> row offsets,
//   4. fill pass   (same traversal, ballot-ordered compaction into the row),
//   5. per-row sort by (distance, id): GNAT returns rows ascending by distance; ties by id is this
//      project's canonical rule.
// Distances are RealVectorStateSpace::distance bit-for-bit (sequential unfused FP64 + IEEE sqrt).
#include <cfloat>
#include <cmath>
#include <cstring>

#include "common.cuh"

#ifndef NB_U
#define NB_U(D) ((D) <= 3 ? 1 : 2)  // point groups in flight per lane
#endif


__host__ __device__ __forceinline__ int nb_cell(double x, double lo, double inv, int G) {
#ifdef __CUDA_ARCH__
  double v = __dmul_rn(__dsub_rn(x, lo), inv);
#else
  volatile double t = x - lo;
  double v = t * inv;
#endif
  if (!(v > 0.0)) return 0;
  if (v >= (double)G) return G - 1;
  return (int)v;
}

// ---------------------------------------------------------------------------
// Grid build
// ---------------------------------------------------------------------------
template <int D>
__global__ void nb_bbox_kernel(const double* __restrict__ verts, const uint8_t* __restrict__ active, uint32_t n,
                               double* __restrict__ partial /* [blocks][2*NB_MAXG] */) {
  double mn[NB_MAXG], mx[NB_MAXG];
  constexpr int GD = D < NB_MAXG ? D : NB_MAXG;
#pragma unroll
  for (int j = 0; j < GD; ++j) { mn[j] = DBL_MAX; mx[j] = -DBL_MAX; }
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (!active[i]) continue;
#pragma unroll
    for (int j = 0; j < GD; ++j) {
      double x = verts[(size_t)i * D + j];
      mn[j] = fmin(mn[j], x);
      mx[j] = fmax(mx[j], x);
    }
  }
  __shared__ double s[256][2 * NB_MAXG];
#pragma unroll
  for (int j = 0; j < GD; ++j) { s[threadIdx.x][2 * j] = mn[j]; s[threadIdx.x][2 * j + 1] = mx[j]; }
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st)
      for (int j = 0; j < GD; ++j) {
        s[threadIdx.x][2 * j] = fmin(s[threadIdx.x][2 * j], s[threadIdx.x + st][2 * j]);
        s[threadIdx.x][2 * j + 1] = fmax(s[threadIdx.x][2 * j + 1], s[threadIdx.x + st][2 * j + 1]);
      }
    __syncthreads();
  }
  if (threadIdx.x == 0)
    for (int j = 0; j < 2 * GD; ++j) partial[(size_t)blockIdx.x * 2 * NB_MAXG + j] = s[0][j];
}

template <int D>
__device__ __forceinline__ uint32_t nb_cell_of(const double* v, const GridView& g) {
  uint32_t c = 0;
#pragma unroll
  for (int j = 0; j < NB_MAXG; ++j)
    if (j < g.gd && j < D) c = c * g.G[j] + nb_cell(v[j], g.lo[j], g.inv[j], g.G[j]);
  return c;
}

template <int D>
__global__ void nb_hist_kernel(const double* __restrict__ verts, const uint8_t* __restrict__ active, uint32_t n,
                               GridView g, uint32_t* __restrict__ counts, uint32_t* __restrict__ cell_of) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (!active[i]) { cell_of[i] = BP_ABSENT; continue; }
    double v[NB_MAXG];
#pragma unroll
    for (int j = 0; j < NB_MAXG; ++j) v[j] = j < D ? verts[(size_t)i * D + j] : 0.0;
    uint32_t c = nb_cell_of<D>(v, g);
    cell_of[i] = c;
    atomicAdd(counts + c, 1u);
  }
}

template <int D>
__global__ void nb_scatter_kernel(const double* __restrict__ verts, uint32_t n, const uint32_t* __restrict__ cell_of,
                                  const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ fill,
                                  uint32_t* __restrict__ ids, double* __restrict__ pts, uint32_t nactive) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t c = cell_of[i];
    if (c == BP_ABSENT) continue;
    uint32_t p = cell_start[c] + atomicAdd(fill + c, 1u);
    ids[p] = i;
#pragma unroll
    for (int j = 0; j < D; ++j) pts[(size_t)j * nactive + p] = verts[(size_t)i * D + j];
  }
}

// ---------------------------------------------------------------------------
// Exclusive scan of u32 counts -> u64 offsets (three small kernels)
// ---------------------------------------------------------------------------
#define SCAN_BLOCK 1024
__global__ void scan_block_sums(const uint32_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ bsum) {
  __shared__ unsigned long long s[32];
  uint64_t i = (uint64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  unsigned long long v = i < n ? in[i] : 0;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = s[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) bsum[blockIdx.x] = v;
  }
}
__global__ void scan_of_sums(uint64_t* bsum, uint64_t nb, uint64_t* total) {
  // single thread block; sequential over chunks of 1024 with a block scan each
  __shared__ unsigned long long s[32];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint64_t base = 0; base < nb; base += SCAN_BLOCK) {
    uint64_t i = base + threadIdx.x;
    unsigned long long v = i < nb ? bsum[i] : 0, x = v;
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) s[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned long long w = s[threadIdx.x], z = w;
      for (int o = 1; o < 32; o <<= 1) {
        unsigned long long y = __shfl_up_sync(0xffffffffu, z, o);
        if (threadIdx.x >= o) z += y;
      }
      s[threadIdx.x] = z - w;  // exclusive warp offsets
    }
    __syncthreads();
    unsigned long long excl = x - v + s[threadIdx.x >> 5] + carry;
    if (i < nb) bsum[i] = excl;
    __syncthreads();
    if (threadIdx.x == SCAN_BLOCK - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void scan_finish(const uint32_t* __restrict__ in, uint64_t n, const uint64_t* __restrict__ bsum,
                            uint64_t* __restrict__ out) {
  __shared__ unsigned long long s[32];
  uint64_t i = (uint64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  unsigned long long v = i < n ? in[i] : 0, x = v;
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) s[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned long long w = s[threadIdx.x], z = w;
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long y = __shfl_up_sync(0xffffffffu, z, o);
      if (threadIdx.x >= o) z += y;
    }
    s[threadIdx.x] = z - w;
  }
  __syncthreads();
  if (i < n) out[i] = x - v + s[threadIdx.x >> 5] + bsum[blockIdx.x];
}

// out[n] receives the total.
int bp_exclusive_scan_u32(bpomp_ctx* ctx, const uint32_t* d_in, uint64_t n, uint64_t* d_out, DevBuf& tmp) {
  uint64_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  if (nb == 0) nb = 1;
  BP_TRY(bp_reserve(ctx, tmp, nb * sizeof(uint64_t), false));
  uint64_t* bsum = tmp.as<uint64_t>();
  scan_block_sums<<<(unsigned)nb, SCAN_BLOCK, 0, ctx->stream>>>(d_in, n, bsum);
  scan_of_sums<<<1, SCAN_BLOCK, 0, ctx->stream>>>(bsum, nb, d_out + n);
  scan_finish<<<(unsigned)nb, SCAN_BLOCK, 0, ctx->stream>>>(d_in, n, bsum, d_out);
  ctx->launches += 3;
  BP_CUDA(ctx, cudaGetLastError());
  return BPOMP_OK;
}

__global__ void scan_to_u32(const uint64_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = (uint32_t)in[i];
}

// ---------------------------------------------------------------------------
// Count / fill traversal: S warps per query.
// ---------------------------------------------------------------------------
struct QueryArgs {
  const double* verts;      // AoS vertex table
  const uint8_t* active;
  const uint32_t* query;    // query ids, or null = identity (neighbors_all)
  uint32_t nq;
  int S;                    // slices (warps) per query
  double r, r2pad;
  int skip_inactive_query;  // neighbors_all: inactive vertices get empty rows
  uint32_t* counts;         // [nq*S]      (count pass)
  const uint64_t* offs;     // [nq*S+1]    (fill pass)
  double* out_dist;         // unsorted pairs
  uint32_t* out_id;
};

template <int D, bool FILL>
__global__ void __launch_bounds__(256) nb_query_kernel(QueryArgs a, GridView g) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t total = (uint64_t)a.nq * a.S;
  for (uint64_t w = warp; w < total; w += nwarps) {
    uint32_t qi = (uint32_t)(w / a.S);
    int s = (int)(w % a.S);
    uint32_t qid = a.query ? __ldg(a.query + qi) : qi;
    uint32_t cnt = 0;
    uint64_t base = FILL ? a.offs[w] : 0;
    bool skip = a.skip_inactive_query && !a.active[qid];
    if (!skip) {
      double q[D];
#pragma unroll
      for (int j = 0; j < D; ++j) q[j] = __ldg(a.verts + (size_t)qid * D + j);
      // cell range per gridded dim, padded so that rounding can never exclude a true neighbour
      int c0[NB_MAXG], c1[NB_MAXG];
#pragma unroll
      for (int j = 0; j < NB_MAXG; ++j) {
        c0[j] = 0; c1[j] = 0;
        if (j < g.gd && j < D) {
          double pad = (fabs(q[j]) + a.r) * 9.094947017729282e-13;  // 2^-40 relative
          c0[j] = nb_cell(q[j] - a.r - pad, g.lo[j], g.inv[j], g.G[j]);
          c1[j] = nb_cell(q[j] + a.r + pad, g.lo[j], g.inv[j], g.G[j]);
        }
      }
      // cell id = ((c[0])*G[1] + c[1])*G[2] + c[2] (gd = 3); the last gridded dim is contiguous in
      // memory, so every combination of the outer dims contributes one run of cells.
      const int last = g.gd - 1;
      const int GL = g.G[last];
      const int A0 = g.gd >= 2 ? c0[0] : 0, A1 = g.gd >= 2 ? c1[0] : 0;
      const int B0 = g.gd >= 3 ? c0[1] : 0, B1 = g.gd >= 3 ? c1[1] : 0;
      for (int a0 = A0; a0 <= A1; ++a0) {
        for (int b0 = B0; b0 <= B1; ++b0) {
          const int prefix = g.gd == 1 ? 0 : (g.gd == 2 ? a0 : a0 * g.G[1] + b0);
          const uint32_t first_cell = (uint32_t)(prefix * GL + c0[last]);
          const uint32_t last_cell = (uint32_t)(prefix * GL + c1[last]);
          uint32_t p0 = __ldg(g.cell_start + first_cell), p1 = __ldg(g.cell_start + last_cell + 1);
          // U point groups per iteration: all coordinate loads are issued before the first use, so each
          // lane has U*D independent loads in flight against the (L2-resident) point table
          constexpr int U = NB_U(D);  // measured: 7-D rows 23.6 -> 15.4 ms with 2; no gain for 2-D rows
          const uint32_t stride = (uint32_t)a.S * 32u;
          for (uint32_t pb = p0 + (uint32_t)s * 32u; pb < p1; pb += U * stride) {
            double c[U][D];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const uint32_t p = pb + u * stride + lane;
#pragma unroll
              for (int j = 0; j < D; ++j) c[u][j] = p < p1 ? __ldg(g.pts + (size_t)j * g.nactive + p) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (pb + u * stride >= p1) continue;  // warp-uniform
              const uint32_t p = pb + u * stride + lane;
              bool hit = false;
              double dist = 0.0;
              if (p < p1) {
                double d2 = 0.0;
#pragma unroll
                for (int j = 0; j < D; ++j) {
                  double diff = __dsub_rn(q[j], c[u][j]);
                  d2 = __dadd_rn(d2, __dmul_rn(diff, diff));
                }
                if (d2 <= a.r2pad) {  // cheap conservative prefilter, exact test below
                  dist = __dsqrt_rn(d2);
                  hit = dist <= a.r;  // inclusive, as GNAT nearestR
                }
              }
              unsigned m = __ballot_sync(0xffffffffu, hit);
              if (FILL && hit) {
                uint64_t o = base + cnt + __popc(m & ((1u << lane) - 1u));
                a.out_dist[o] = dist;
                a.out_id[o] = __ldg(g.ids + p);
              }
              cnt += __popc(m);
            }
          }
        }
      }
    }
    if (!FILL && lane == 0) a.counts[w] = cnt;
  }
}

// Brute-force regime (the grid has a handful of cells because r is of the order of the domain, e.g.
// 7-D roadmaps or vertex batching's unbounded radius): every query meets most points, so the scan is
// organised as tiled all-pairs instead — thread = query, a block of NB_TQ queries walks its chunk of the
// cell-ordered point table in tiles staged through shared memory (coalesced tile loads, broadcast reads),
// exactly like the belief kNN kernel.  Slice index = point chunk, so counts / offsets / fill keep the
// (query, slice) layout of the warp kernel and the rest of the pipeline is unchanged.
#define NB_TQ 128   // threads per block
#define NB_TP 128   // points per tile
#define NB_QT 2     // queries per thread: a point read from the tile (shared-memory broadcast) serves both
__host__ __device__ constexpr int nb_ds(int D) { return (D + 2) & ~1; }  // coordinates, |p|^2/2, pad to 16 B

template <int D, bool FILL>
__global__ void __launch_bounds__(NB_TQ, 4) nb_tile_kernel(QueryArgs a, GridView g, uint32_t chunk) {
  constexpr int DS = nb_ds(D);
  __shared__ __align__(16) double s_pts[NB_TP * DS];
  __shared__ uint32_t s_ids[NB_TP];
  __shared__ unsigned int s_pnhi;  // high word of the tile's largest |p|^2 (non-negative doubles order as integers)
  const int tid = threadIdx.x;
  const uint32_t c = blockIdx.y;
  const uint32_t p_begin = c * chunk;
  const uint32_t p_end = min(p_begin + chunk, g.nactive);
  bool have[NB_QT], skip[NB_QT];
  double nq[NB_QT][D];  // -query (negation is exact)
  double qn[NB_QT];     // |q|^2, prefilter only
  uint64_t w[NB_QT], o[NB_QT];
  uint32_t cnt[NB_QT];
  bool skip_all = true;
#pragma unroll
  for (int s = 0; s < NB_QT; ++s) {
    const uint32_t qi = (blockIdx.x * NB_QT + s) * NB_TQ + tid;
    have[s] = qi < a.nq;
    const uint32_t qid = have[s] ? (a.query ? __ldg(a.query + qi) : qi) : 0u;
    skip[s] = !have[s] || (a.skip_inactive_query && !a.active[qid]);
    skip_all = skip_all && skip[s];
    qn[s] = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const double v = have[s] ? __ldg(a.verts + (size_t)qid * D + j) : 0.0;
      nq[s][j] = -v;
      qn[s] = fma(v, v, qn[s]);
    }
    w[s] = (uint64_t)qi * a.S + c;
    o[s] = (FILL && have[s]) ? a.offs[w[s]] : 0;
    cnt[s] = 0;
  }
  // Prefilter: |q-p|^2/2 = |p|^2/2 - q.p + |q|^2/2 as D fused multiply-adds on the stored |p|^2/2; absolute
  // error below 2^-49 (|q|^2 + |p|^2), bound padded by 2^-46 (|q|^2 + an upper bound of the tile's max |p|^2).
  // Survivors are recomputed with the reference's unfused arithmetic.  NaN compares false -> recomputed.
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
  auto exact = [&](int s, uint32_t t) {
    double d2 = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const double diff = __dsub_rn(-nq[s][j], s_pts[t * DS + j]);
      d2 = __dadd_rn(d2, __dmul_rn(diff, diff));
    }
    const double dist = __dsqrt_rn(d2);
    if (dist <= a.r) {  // inclusive, as GNAT nearestR
      if (FILL) {
        a.out_dist[o[s] + cnt[s]] = dist;
        a.out_id[o[s] + cnt[s]] = s_ids[t];
      }
      ++cnt[s];
    }
  };
  for (uint32_t base = p_begin; base < p_end; base += NB_TP) {
    const uint32_t tn = min((uint32_t)NB_TP, p_end - base);
    __syncthreads();
    if (tid == 0) s_pnhi = 0u;
    __syncthreads();
    {
      double pn = 0.0;
      if ((uint32_t)tid < tn) {
#pragma unroll
        for (int j = 0; j < D; ++j) {
          const double v = __ldg(g.pts + (size_t)j * g.nactive + base + tid);
          s_pts[tid * DS + j] = v;
          pn = fma(v, v, pn);
        }
        s_pts[tid * DS + D] = 0.5 * pn;
        if (FILL) s_ids[tid] = __ldg(g.ids + base + tid);
      }
      const unsigned int hi = __reduce_max_sync(0xffffffffu, (unsigned int)__double2hiint(pn));
      if ((tid & 31) == 0) atomicMax(&s_pnhi, hi);
    }
    __syncthreads();
    if (skip_all) continue;
    // upper bound of max |p|^2: next high word (inf / NaN high words stay non-finite -> every point is recomputed)
    const double pnmax = __hiloint2double((int)(s_pnhi < 0x7ff00000u ? s_pnhi + 1u : 0x7ff80000u), 0);
    double bound[NB_QT];
#pragma unroll
    for (int s = 0; s < NB_QT; ++s)
      bound[s] = skip[s] ? -DBL_MAX : 0.5 * ((a.r2pad - qn[s]) + 2.842170943040401e-14 * (qn[s] + pnmax));
    uint32_t t = 0;
    for (; t + 4 <= tn; t += 4) {
      double acc[4][NB_QT];
      bool far = true;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double2* pp = reinterpret_cast<const double2*>(s_pts + (t + u) * DS);
#pragma unroll
        for (int s = 0; s < NB_QT; ++s) {
          acc[u][s] = approx(pp, s);
          far = far && acc[u][s] > bound[s];
        }
      }
      if (far) continue;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int s = 0; s < NB_QT; ++s)
          if (!skip[s] && !(acc[u][s] > bound[s])) exact(s, t + u);
      }
    }
    for (; t < tn; ++t) {
      const double2* pp = reinterpret_cast<const double2*>(s_pts + t * DS);
#pragma unroll
      for (int s = 0; s < NB_QT; ++s)
        if (!skip[s] && !(approx(pp, s) > bound[s])) exact(s, t);
    }
  }
  if (!FILL) {
#pragma unroll
    for (int s = 0; s < NB_QT; ++s)
      if (have[s]) a.counts[w[s]] = cnt[s];
  }
}

// The tiled all-pairs scan for dimensions >= 4 and a finite radius, with the prefilter on the FP32 pipe — the
// FP64 version above runs at 57 % of the FP64 pipe (profiles/r1_neighbors7d_tiled_v2_summary.md).  Same scheme as
// the belief kNN kernel (belief.cu: bk_knn_partial_f32): the tile holds the points rounded to float, two points
// interleaved per record plus their |p|^2/2, so a 16-byte broadcast load feeds two packed fused multiply-adds
// (FFMA2);  a = |p|^2/2 - q.p - t  with  t = (r2pad - |q|^2)/2 + 2^-20 (|q|^2 + max|p|^2)  rounded up rejects a
// point iff a > 0 (the pad bounds the FP32 error: inputs rounded to float, D + 1 <= 9 roundings of the chain, the
// rounding of |p|^2/2), tested for 16 points at once by OR-ing the accumulators' sign bits.  Survivors go to a
// small per-thread queue and are evaluated together — the reference's unfused FP64 distance on the original
// coordinates, `dist <= r` inclusive as GNAT's nearestR — so that path runs with most lanes of a warp busy.
#define NBS_G 16
#define NBS_QCAP 8
__host__ __device__ constexpr int nbs_pp(int D) { return (D + 2) & ~1; }  // float2 entries per pair record

template <int D, bool FILL>
__global__ void __launch_bounds__(NB_TQ, 4) nb_tile_kernel_f32(QueryArgs a, GridView g, uint32_t chunk) {
  constexpr int PP = nbs_pp(D);
  static_assert(NBS_G * NB_QT <= 32, "hit mask is 32 bits");
  __shared__ __align__(16) float s_pf[(NB_TP / 2) * PP * 2];
  __shared__ double s_pd[NB_TP * D];
  __shared__ uint32_t s_ids[NB_TP];
  __shared__ uint16_t s_queue[NBS_QCAP * NB_TQ];
  __shared__ unsigned int s_pnhi;
  const int tid = threadIdx.x;
  uint16_t* s_q = s_queue + tid;
  const uint32_t c = blockIdx.y;
  const uint32_t p_begin = c * chunk;
  const uint32_t p_end = min(p_begin + chunk, g.nactive);
  bool have[NB_QT], skip[NB_QT];
  float2 nq2[NB_QT][D];
  float2 nthr2[NB_QT];
  double qn[NB_QT];
  const double* qptr[NB_QT];
  uint64_t w[NB_QT], o[NB_QT];
  uint32_t cnt[NB_QT];
  bool skip_all = true;
#pragma unroll
  for (int s = 0; s < NB_QT; ++s) {
    const uint32_t qi = (blockIdx.x * NB_QT + s) * NB_TQ + tid;
    have[s] = qi < a.nq;
    const uint32_t qid = have[s] ? (a.query ? __ldg(a.query + qi) : qi) : 0u;
    skip[s] = !have[s] || (a.skip_inactive_query && !a.active[qid]);
    skip_all = skip_all && skip[s];
    qptr[s] = a.verts + (size_t)qid * D;
    qn[s] = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const double v = have[s] ? __ldg(qptr[s] + j) : 0.0;
      const float f = -(float)v;
      nq2[s][j] = make_float2(f, f);
      qn[s] = fma(v, v, qn[s]);
    }
    w[s] = (uint64_t)qi * a.S + c;
    o[s] = (FILL && have[s]) ? a.offs[w[s]] : 0;
    cnt[s] = 0;
  }
  // a warp leaves a tile early only as a whole (a lane without queries keeps voting; its slots never survive)
  const bool skip_warp = __all_sync(0xffffffffu, skip_all);
  int qlen = 0;
  auto exact = [&](int s, uint32_t t) {
    double d2 = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const double diff = __dsub_rn(__ldg(qptr[s] + j), s_pd[t * D + j]);
      d2 = __dadd_rn(d2, __dmul_rn(diff, diff));
    }
    const double dist = __dsqrt_rn(d2);
    if (dist <= a.r) {  // inclusive, as GNAT nearestR
      if (FILL) {
        a.out_dist[o[s] + cnt[s]] = dist;
        a.out_id[o[s] + cnt[s]] = s_ids[t];
      }
      ++cnt[s];
    }
  };
  auto exact_slot = [&](int s, uint32_t t) {
#pragma unroll
    for (int k = 0; k < NB_QT; ++k)
      if (k == s) exact(k, t);
  };
  auto drain = [&]() {
    int i = 0;
    while (__any_sync(0xffffffffu, i < qlen)) {
      if (i < qlen) {
        const uint32_t code = s_q[i * NB_TQ];
        exact_slot((int)(code >> 12), code & 0xfffu);
      }
      ++i;
    }
    qlen = 0;
  };
  for (uint32_t base = p_begin; base < p_end; base += NB_TP) {
    const uint32_t tn = min((uint32_t)NB_TP, p_end - base);
    const uint32_t tn_pad = (tn + NBS_G - 1) & ~(uint32_t)(NBS_G - 1);
    __syncthreads();  // the previous tile has been drained by every thread
    if (tid == 0) s_pnhi = 0u;
    __syncthreads();
    {
      double pn = 0.0;
      float* rec = s_pf + (size_t)(tid >> 1) * 2 * PP + (tid & 1);
      if ((uint32_t)tid < tn) {
#pragma unroll
        for (int j = 0; j < D; ++j) {
          const double v = __ldg(g.pts + (size_t)j * g.nactive + base + tid);
          s_pd[tid * D + j] = v;
          rec[2 * j] = (float)v;
          pn = fma(v, v, pn);
        }
        rec[2 * D] = (float)(0.5 * pn);
        if (FILL) s_ids[tid] = __ldg(g.ids + base + tid);
      } else if ((uint32_t)tid < tn_pad) {  // padding up to a whole group: never a survivor
#pragma unroll
        for (int j = 0; j < D; ++j) rec[2 * j] = 0.f;
        rec[2 * D] = __int_as_float(0x7f800000);
      }
      const unsigned int hi = __reduce_max_sync(0xffffffffu, (unsigned int)__double2hiint(pn));
      if ((tid & 31) == 0) atomicMax(&s_pnhi, hi);
    }
    __syncthreads();
    if (skip_warp) continue;  // warp-uniform: the group loop below uses warp-wide votes
    // upper bound of max |p|^2: next high word (inf / NaN high words stay non-finite -> nothing is rejected)
    const double pnmax = __hiloint2double((int)(s_pnhi < 0x7ff00000u ? s_pnhi + 1u : 0x7ff80000u), 0);
#pragma unroll
    for (int s = 0; s < NB_QT; ++s) {
      // a skipped slot never survives (t = -inf); a NaN bound makes every comparison fail -> everything survives
      const float t = skip[s] ? __int_as_float(0x7f800000)
                              : -__double2float_ru(0.5 * (a.r2pad - qn[s]) + 9.5367431640625e-07 * (qn[s] + pnmax));
      nthr2[s] = make_float2(t, t);
    }
    for (uint32_t p = 0; p < tn; p += NBS_G) {
      const float4* g4 = reinterpret_cast<const float4*>(s_pf) + (size_t)(p >> 1) * (PP / 2);
      float2 acc[NBS_G / 2][NB_QT];
      uint32_t sign = 0;
#pragma unroll
      for (int u = 0; u < NBS_G / 2; ++u) {
        float2 x2[PP];
#pragma unroll
        for (int h = 0; h < PP / 2; ++h) {
          const float4 t = g4[u * (PP / 2) + h];
          x2[2 * h] = make_float2(t.x, t.y);
          x2[2 * h + 1] = make_float2(t.z, t.w);
        }
#pragma unroll
        for (int s = 0; s < NB_QT; ++s) {
          float2 v = __fadd2_rn(x2[D], nthr2[s]);
#pragma unroll
          for (int j = 0; j < D; ++j) v = __ffma2_rn(nq2[s][j], x2[j], v);
          acc[u][s] = v;
          sign |= __float_as_uint(v.x) | __float_as_uint(v.y);
        }
      }
      uint32_t hits = 0;
      if (sign >> 31) {
#pragma unroll
        for (int u = 0; u < NBS_G / 2; ++u)
#pragma unroll
          for (int s = 0; s < NB_QT; ++s) {
            if (!(acc[u][s].x > 0.f)) hits |= 1u << ((2 * u) * NB_QT + s);
            if (!(acc[u][s].y > 0.f)) hits |= 1u << ((2 * u + 1) * NB_QT + s);
          }
      }
      if (__any_sync(0xffffffffu, hits != 0u)) {
        const int nh = __popc(hits);
        if (__any_sync(0xffffffffu, qlen + nh > NBS_QCAP)) drain();
        if (nh > NBS_QCAP) {  // more than a queue holds: evaluate this group in place
          while (hits) {
            const int b = __ffs(hits) - 1;
            hits &= hits - 1;
            if (p + (uint32_t)(b / NB_QT) < tn) exact_slot(b % NB_QT, p + (uint32_t)(b / NB_QT));
          }
        } else {
          while (hits) {
            const int b = __ffs(hits) - 1;
            hits &= hits - 1;
            if (p + (uint32_t)(b / NB_QT) < tn) {
              s_q[qlen * NB_TQ] = (uint16_t)(((uint32_t)(b % NB_QT) << 12) | (p + (uint32_t)(b / NB_QT)));
              ++qlen;
            }
          }
        }
      }
    }
    drain();
  }
  if (!FILL) {
#pragma unroll
    for (int s = 0; s < NB_QT; ++s)
      if (have[s]) a.counts[w[s]] = cnt[s];
  }
}

// row_ptr[q] = offs[q*S]
__global__ void nb_rowptr_kernel(const uint64_t* __restrict__ offs, uint32_t nq, int S, uint64_t* __restrict__ row_ptr,
                                 unsigned long long* __restrict__ maxrow) {
  unsigned long long mx = 0;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q <= nq; q += gridDim.x * blockDim.x) {
    uint64_t o = offs[(uint64_t)q * S];
    row_ptr[q] = o;
    if (q < nq) {
      unsigned long long len = offs[(uint64_t)(q + 1) * S] - o;
      mx = len > mx ? len : mx;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long y = __shfl_down_sync(0xffffffffu, mx, o);
    mx = y > mx ? y : mx;
  }
  if ((threadIdx.x & 31) == 0 && mx) atomicMax(maxrow, mx);
}

// ---------------------------------------------------------------------------
// Row sort by (dist, id).  Small rows: rank sort, one warp each.  Medium rows: bitonic network in
// shared memory, one block each.  Rows beyond NB_MEDIUM: global bitonic network (huge radius only).
// ---------------------------------------------------------------------------
#define NB_SMALL 128
#define NB_MEDIUM 8192

__device__ __forceinline__ bool nb_less(double da, uint32_t ia, double db, uint32_t ib) {
  return da < db || (da == db && ia < ib);
}

__global__ void __launch_bounds__(256) nb_sort_small(const uint64_t* __restrict__ row_ptr, uint32_t nq,
                                                     const double* __restrict__ in_d, const uint32_t* __restrict__ in_i,
                                                     double* __restrict__ out_d, uint32_t* __restrict__ out_i) {
  __shared__ double sd[8][NB_SMALL];
  __shared__ uint32_t si[8][NB_SMALL];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nq; q += nw) {
    uint64_t o = row_ptr[q];
    uint32_t len = (uint32_t)min((unsigned long long)(row_ptr[q + 1] - o), 0xFFFFFFFFull);
    if (len == 0 || len > NB_SMALL) continue;
    for (uint32_t i = lane; i < len; i += 32) { sd[wib][i] = in_d[o + i]; si[wib][i] = in_i[o + i]; }
    __syncwarp();
    for (uint32_t i = lane; i < len; i += 32) {
      double d = sd[wib][i];
      uint32_t id = si[wib][i];
      uint32_t rank = 0;
      for (uint32_t k = 0; k < len; ++k) rank += nb_less(sd[wib][k], si[wib][k], d, id) ? 1u : 0u;
      out_d[o + rank] = d;
      out_i[o + rank] = id;
    }
    __syncwarp();
  }
}

// Medium rows: one block per row, ascending-only bitonic network in shared memory (virtual
// power-of-two length; slots >= len act as +infinity and are never touched).
__global__ void __launch_bounds__(512) nb_sort_medium(const uint64_t* __restrict__ row_ptr, uint32_t nq,
                                                      const double* __restrict__ in_d, const uint32_t* __restrict__ in_i,
                                                      double* __restrict__ out_d, uint32_t* __restrict__ out_i) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* sd = reinterpret_cast<double*>(smem);
  uint32_t* si = reinterpret_cast<uint32_t*>(smem + NB_MEDIUM * sizeof(double));
  for (uint32_t q = blockIdx.x; q < nq; q += gridDim.x) {
    uint64_t o = row_ptr[q];
    unsigned long long len64 = row_ptr[q + 1] - o;
    if (len64 <= NB_SMALL || len64 > NB_MEDIUM) continue;
    const uint32_t len = (uint32_t)len64;
    uint32_t npad = 1;
    while (npad < len) npad <<= 1;
    for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) { sd[i] = in_d[o + i]; si[i] = in_i[o + i]; }
    __syncthreads();
    for (uint32_t k = 2; k <= npad; k <<= 1) {
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        const uint32_t mask = (j == (k >> 1)) ? k - 1 : j;  // first step of a merge mirrors, the rest shift
        for (uint32_t i = threadIdx.x; i < npad; i += blockDim.x) {
          uint32_t l = i ^ mask;
          if (l > i && l < len) {
            double di = sd[i], dl = sd[l];
            uint32_t ii = si[i], il = si[l];
            if (nb_less(dl, il, di, ii)) { sd[i] = dl; sd[l] = di; si[i] = il; si[l] = ii; }
          }
        }
        __syncthreads();
      }
    }
    for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) { out_d[o + i] = sd[i]; out_i[o + i] = si[i]; }
    __syncthreads();
  }
}

// One compare-exchange step of an ascending-only bitonic network over a virtual length npad
// (power of two): partner = i ^ mask, smaller key goes to the lower index.  The first step of
// each merge uses mask = k-1 (mirror), the others mask = j.  Indices >= len act as +infinity
// and are never moved, so no padding storage is needed.
__global__ void nb_bitonic_step(double* __restrict__ d, uint32_t* __restrict__ ids, uint64_t len, uint64_t mask,
                                uint64_t npad) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npad; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t l = i ^ mask;
    if (l <= i || l >= len) continue;
    double di = d[i], dl = d[l];
    uint32_t ii = ids[i], il = ids[l];
    if (nb_less(dl, il, di, ii)) { d[i] = dl; d[l] = di; ids[i] = il; ids[l] = ii; }
  }
}

// ---------------------------------------------------------------------------
// Whole-roadmap densification when the grid covers every dimension (D <= 3) and rows are short: one
// THREAD per vertex, taken in cell order, so that the threads of a warp are spatial neighbours and walk
// (nearly) the same candidate runs out of L1.  Count pass, scan over the vertex ids, fill pass, then the
// warp-per-row rank sort by (distance, id).
// ---------------------------------------------------------------------------
#define NB_DENSE_CAP NB_SMALL  // rows longer than the warp rank sort handles take the general path

template <int D, bool FILL>
__global__ void __launch_bounds__(128) nb_dense_kernel(GridView g, double r, double r2pad,
                                                       uint32_t* __restrict__ counts,
                                                       const uint64_t* __restrict__ row_ptr,
                                                       uint32_t* __restrict__ col, double* __restrict__ dist_out) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.nactive) return;
  double q[D];
#pragma unroll
  for (int j = 0; j < D; ++j) q[j] = __ldg(g.pts + (size_t)j * g.nactive + p);
  int c0[D], c1[D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    double pad = (fabs(q[j]) + r) * 9.094947017729282e-13;
    c0[j] = nb_cell(q[j] - r - pad, g.lo[j], g.inv[j], g.G[j]);
    c1[j] = nb_cell(q[j] + r + pad, g.lo[j], g.inv[j], g.G[j]);
  }
  uint32_t cnt = 0;
  const uint32_t vid = __ldg(g.ids + p);
  const uint64_t o = FILL ? row_ptr[vid] : 0;
  const int A0 = D >= 2 ? c0[0] : 0, A1 = D >= 2 ? c1[0] : 0;
  const int B0 = D >= 3 ? c0[1] : 0, B1 = D >= 3 ? c1[1] : 0;
  const int GL = g.G[D - 1];
  for (int a0 = A0; a0 <= A1; ++a0)
    for (int b0 = B0; b0 <= B1; ++b0) {
      const int prefix = D == 1 ? 0 : (D == 2 ? a0 : a0 * g.G[1] + b0);
      const uint32_t i0 = __ldg(g.cell_start + (uint32_t)(prefix * GL + c0[D - 1]));
      const uint32_t i1 = __ldg(g.cell_start + (uint32_t)(prefix * GL + c1[D - 1]) + 1);
      for (uint32_t i = i0; i < i1; ++i) {
        double d2 = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) {
          const double diff = __dsub_rn(q[j], __ldg(g.pts + (size_t)j * g.nactive + i));
          d2 = __dadd_rn(d2, __dmul_rn(diff, diff));
        }
        if (d2 > r2pad) continue;
        const double dd = __dsqrt_rn(d2);
        if (!(dd <= r)) continue;  // inclusive, as GNAT nearestR
        if (FILL) {  // unsorted; the row sort kernel orders it by (distance, id)
          col[o + cnt] = __ldg(g.ids + i);
          dist_out[o + cnt] = dd;
        }
        ++cnt;
      }
    }
  if (!FILL) counts[vid] = cnt;
}

template <int D>
static int neighbors_all_dense(bpomp_ctx* ctx, const GridView& g, double r, double r2pad, uint64_t* nnz_out, bool* done) {
  *done = false;
  const uint32_t nv = ctx->nverts;
  BP_TRY(bp_reserve(ctx, ctx->s_f, ((size_t)nv + 2) * sizeof(uint32_t), false));
  BP_TRY(bp_reserve(ctx, ctx->d_nb_rowptr, ((size_t)nv + 2) * sizeof(uint64_t), false));
  uint32_t* counts = ctx->s_f.as<uint32_t>();
  uint64_t* rp = ctx->d_nb_rowptr.as<uint64_t>();
  BP_CUDA(ctx, cudaMemsetAsync(counts, 0, ((size_t)nv + 1) * sizeof(uint32_t), ctx->stream));  // inactive: empty rows
  const int grid = (int)((g.nactive + 127) / 128);
  if (grid > 0) nb_dense_kernel<D, false><<<grid, 128, 0, ctx->stream>>>(g, r, r2pad, counts, nullptr, nullptr, nullptr);
  ctx->launches++;
  BP_TRY(bp_exclusive_scan_u32(ctx, counts, nv, rp, ctx->d_nb_tmp2));
  // longest row: reuse the row-pointer kernel's reduction (S = 1 maps offsets onto themselves)
  BP_TRY(bp_reserve(ctx, ctx->s_g, 2 * sizeof(uint64_t), false));
  unsigned long long* d_max = ctx->s_g.as<unsigned long long>();
  BP_CUDA(ctx, cudaMemsetAsync(d_max, 0, sizeof(unsigned long long), ctx->stream));
  BP_TRY(bp_reserve(ctx, ctx->d_nb_query, ((size_t)nv + 2) * sizeof(uint64_t), false));
  nb_rowptr_kernel<<<(int)std::min<uint32_t>((nv + 256) / 256, 1024u), 256, 0, ctx->stream>>>(
      rp, nv, 1, ctx->d_nb_query.as<uint64_t>(), d_max);
  ctx->launches++;
  uint64_t hostv[2];
  BP_CUDA(ctx, cudaMemcpyAsync(&hostv[0], rp + nv, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaMemcpyAsync(&hostv[1], d_max, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (hostv[1] > NB_DENSE_CAP) return BPOMP_OK;  // some row is too long for the warp rank sort: general path
  const uint64_t nnz = hostv[0];
  BP_TRY(bp_reserve(ctx, ctx->d_nb_dist, (nnz + 1) * sizeof(double), false));
  BP_TRY(bp_reserve(ctx, ctx->d_nb_col, (nnz + 1) * sizeof(uint32_t), false));
  BP_TRY(bp_reserve(ctx, ctx->d_nb_tmp, (nnz + 1) * sizeof(double), false));
  BP_TRY(bp_reserve(ctx, ctx->s_e, (nnz + 1) * sizeof(uint32_t), false));
  if (nnz > 0 && grid > 0) {
    nb_dense_kernel<D, true><<<grid, 128, 0, ctx->stream>>>(g, r, r2pad, nullptr, rp, ctx->s_e.as<uint32_t>(),
                                                           ctx->d_nb_tmp.as<double>());
    int gs = (int)std::min<uint64_t>(((uint64_t)nv * 32 + 255) / 256, (uint64_t)ctx->num_sms * 16);
    nb_sort_small<<<std::max(gs, 1), 256, 0, ctx->stream>>>(rp, nv, ctx->d_nb_tmp.as<double>(), ctx->s_e.as<uint32_t>(),
                                                           ctx->d_nb_dist.as<double>(), ctx->d_nb_col.as<uint32_t>());
    ctx->launches += 2;
    BP_CUDA(ctx, cudaGetLastError());
  }
  ctx->nb_nnz = nnz;
  ctx->nb_nq = nv;
  if (nnz_out) *nnz_out = nnz;
  *done = true;
  return BPOMP_OK;
}

// ---------------------------------------------------------------------------
// Host driver
// ---------------------------------------------------------------------------
template <int D>
static int build_grid(bpomp_ctx* ctx, double r, GridView& g) {
  const uint32_t n = ctx->nverts;
  // The grid of the last call stands while the vertex table, the active flags and the radius (which sets the cell
  // size) are unchanged: the planner asks for one row per expanded vertex when the whole CSR exceeds its budget,
  // and must not pay a rebuild (bounding box, histogram, scan, scatter and a wait) per expansion.
  if (ctx->grid_valid && ctx->grid_r == r) {
    g = ctx->grid_view;
    return BPOMP_OK;
  }
  memset(&g, 0, sizeof(g));
  // bounding box of the active vertices over the gridded dims
  const int blocks = 64;
  BP_TRY(bp_reserve(ctx, ctx->d_nb_tmp, blocks * 2 * NB_MAXG * sizeof(double), false));
  nb_bbox_kernel<D><<<blocks, 256, 0, ctx->stream>>>(ctx->d_verts.as<double>(), ctx->d_active.as<uint8_t>(), n,
                                                    ctx->d_nb_tmp.as<double>());
  ctx->launches++;
  std::vector<double> part((size_t)blocks * 2 * NB_MAXG);
  BP_CUDA(ctx, cudaMemcpyAsync(part.data(), ctx->d_nb_tmp.p, part.size() * sizeof(double), cudaMemcpyDeviceToHost,
                               ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const uint32_t nactive = ctx->nactive;  // kept by the vertex-table calls (ctx.cu)
  g.gd = D < NB_MAXG ? D : NB_MAXG;
  // cell budget: ~2^21 cells in total, and never more cells than ~4x the points
  double budget = std::min(2097152.0, std::max(1.0, 4.0 * (double)nactive));
  int gcap = (int)std::floor(std::pow(budget, 1.0 / g.gd));
  if (gcap < 1) gcap = 1;
  uint32_t ncells = 1;
  for (int j = 0; j < g.gd; ++j) {
    double mn = DBL_MAX, mx = -DBL_MAX;
    for (int b = 0; b < blocks; ++b) {
      mn = std::min(mn, part[(size_t)b * 2 * NB_MAXG + 2 * j]);
      mx = std::max(mx, part[(size_t)b * 2 * NB_MAXG + 2 * j + 1]);
    }
    double span = mx - mn;
    int G = 1;
    if (nactive > 0 && std::isfinite(span) && span > 0.0 && std::isfinite(r) && r > 0.0) {
      double want = std::floor(span / r);
      G = want >= (double)gcap ? gcap : (want < 1.0 ? 1 : (int)want);
    }
    g.G[j] = G;
    g.lo[j] = (nactive > 0 && std::isfinite(mn)) ? mn : 0.0;
    g.inv[j] = (G > 1) ? (double)G / span : 0.0;
    ncells *= (uint32_t)G;
  }
  g.ncells = ncells;
  g.nactive = nactive;
  BP_TRY(bp_reserve(ctx, ctx->d_grid_cell_start, ((size_t)ncells + 2) * sizeof(uint32_t), false));
  BP_TRY(bp_reserve(ctx, ctx->d_grid_keys, ((size_t)n + 1) * sizeof(uint32_t), false));
  BP_TRY(bp_reserve(ctx, ctx->d_grid_ids, ((size_t)nactive + 1) * sizeof(uint32_t), false));
  BP_TRY(bp_reserve(ctx, ctx->d_grid_pts, ((size_t)nactive + 1) * D * sizeof(double), false));
  const size_t counts_bytes = (((size_t)ncells + 2) * sizeof(uint32_t) + 7) / 8 * 8;
  BP_TRY(bp_reserve(ctx, ctx->d_nb_tmp, counts_bytes + ((size_t)ncells + 2) * sizeof(uint64_t), false));
  uint32_t* counts = ctx->d_nb_tmp.as<uint32_t>();
  uint64_t* offs64 = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(ctx->d_nb_tmp.p) + counts_bytes);
  BP_CUDA(ctx, cudaMemsetAsync(counts, 0, ((size_t)ncells + 2) * sizeof(uint32_t), ctx->stream));
  int grid = (int)std::min<uint64_t>(((uint64_t)n + 255) / 256, (uint64_t)ctx->num_sms * 8);
  if (grid < 1) grid = 1;
  nb_hist_kernel<D><<<grid, 256, 0, ctx->stream>>>(ctx->d_verts.as<double>(), ctx->d_active.as<uint8_t>(), n, g, counts,
                                                  ctx->d_grid_keys.as<uint32_t>());
  ctx->launches++;
  BP_TRY(bp_exclusive_scan_u32(ctx, counts, ncells, offs64, ctx->d_nb_tmp2));
  scan_to_u32<<<grid, 256, 0, ctx->stream>>>(offs64, ncells + 1, ctx->d_grid_cell_start.as<uint32_t>());
  BP_CUDA(ctx, cudaMemsetAsync(counts, 0, ((size_t)ncells + 2) * sizeof(uint32_t), ctx->stream));
  nb_scatter_kernel<D><<<grid, 256, 0, ctx->stream>>>(ctx->d_verts.as<double>(), n, ctx->d_grid_keys.as<uint32_t>(),
                                                     ctx->d_grid_cell_start.as<uint32_t>(), counts,
                                                     ctx->d_grid_ids.as<uint32_t>(), ctx->d_grid_pts.as<double>(),
                                                     nactive);
  ctx->launches += 2;
  BP_CUDA(ctx, cudaGetLastError());
  g.cell_start = ctx->d_grid_cell_start.as<uint32_t>();
  g.ids = ctx->d_grid_ids.as<uint32_t>();
  g.pts = ctx->d_grid_pts.as<double>();
  ctx->grid_view = g;
  ctx->grid_r = r;
  ctx->grid_valid = true;
  return BPOMP_OK;
}

template <int D>
static int neighbors_impl(bpomp_ctx* ctx, const uint32_t* d_query, uint32_t nq, double r, bool all, uint64_t* nnz_out) {
  GridView g;
  BP_TRY(build_grid<D>(ctx, r, g));
  if (all && D <= NB_MAXG && g.gd == D && g.ncells >= 64 && std::isfinite(r)) {
    // whole-roadmap densification on a grid that covers every dimension: thread-per-vertex path
    double r2p = (r * r) * (1.0 + 9.094947017729282e-13);
    bool done = false;
    BP_TRY((neighbors_all_dense<(D <= NB_MAXG ? D : 1)>(ctx, g, r, r2p, nnz_out, &done)));
    if (done) return BPOMP_OK;
  }
  // Brute-force regime with enough queries to fill thread-per-query blocks -> tiled all-pairs kernel
  const bool tiled = g.ncells <= 8 && nq >= 4 * NB_TQ && g.nactive >= 4 * NB_TP;
  // slices per query: enough warps to fill the machine when there are few queries
  int S = 1;
  uint32_t chunk = 0;
  if (tiled) {
    uint64_t qblocks = ((uint64_t)nq + NB_TQ * NB_QT - 1) / (NB_TQ * NB_QT);
    uint64_t want = (uint64_t)ctx->num_sms * 8;  // resident blocks
    while (qblocks * S < want && (uint64_t)g.nactive / (2 * S) >= 8 * NB_TP) S *= 2;
    chunk = (uint32_t)(((uint64_t)g.nactive + S - 1) / S);
    chunk = (chunk + NB_TP - 1) / NB_TP * NB_TP;
    S = (int)(((uint64_t)g.nactive + chunk - 1) / chunk);
  } else {
    uint64_t want = (uint64_t)ctx->num_sms * 64;  // resident warps
    while ((uint64_t)nq * S < want && S < 1024 && (uint64_t)S * 32 * 4 < (uint64_t)g.nactive + 1) S *= 2;
  }
  uint64_t nw = (uint64_t)nq * S;
  QueryArgs a;
  memset(&a, 0, sizeof(a));
  a.verts = ctx->d_verts.as<double>();
  a.active = ctx->d_active.as<uint8_t>();
  a.query = d_query;
  a.nq = nq;
  a.S = S;
  a.r = r;
  a.r2pad = (r * r) * (1.0 + 9.094947017729282e-13);  // 2^-40: covers the sqrt rounding and the fused prefilter
  if (!(a.r2pad >= 0.0)) a.r2pad = DBL_MAX;           // NaN guard
  if (std::isinf(a.r2pad)) a.r2pad = DBL_MAX;
  a.skip_inactive_query = all ? 1 : 0;
  BP_TRY(bp_reserve(ctx, ctx->s_f, (nw + 2) * sizeof(uint32_t), false));
  BP_TRY(bp_reserve(ctx, ctx->s_g, (nw + 2) * sizeof(uint64_t), false));
  a.counts = ctx->s_f.as<uint32_t>();
  int grid = (int)std::min<uint64_t>((nw * 32 + 255) / 256, (uint64_t)ctx->num_sms * 8);
  if (grid < 1) grid = 1;
  const dim3 tgrid((unsigned)(((uint64_t)nq + NB_TQ * NB_QT - 1) / (NB_TQ * NB_QT)), (unsigned)S);
  // dimensions >= 4 with a finite radius: prefilter on the FP32 pipe; else (every pair passes / low dimension) FP64
  const bool tiled32 = tiled && D >= 4 && a.r2pad < DBL_MAX && ctx->prefilter_f32;
  if (tiled32) nb_tile_kernel_f32<D, false><<<tgrid, NB_TQ, 0, ctx->stream>>>(a, g, chunk);
  else if (tiled) nb_tile_kernel<D, false><<<tgrid, NB_TQ, 0, ctx->stream>>>(a, g, chunk);
  else nb_query_kernel<D, false><<<grid, 256, 0, ctx->stream>>>(a, g);
  ctx->launches++;
  BP_CUDA(ctx, cudaGetLastError());
  uint64_t* offs = ctx->s_g.as<uint64_t>();
  BP_TRY(bp_exclusive_scan_u32(ctx, a.counts, nw, offs, ctx->d_nb_tmp2));
  BP_TRY(bp_reserve(ctx, ctx->d_nb_rowptr, ((size_t)nq + 2) * sizeof(uint64_t), false));
  unsigned long long* d_max = reinterpret_cast<unsigned long long*>(ctx->d_nb_rowptr.as<uint64_t>() + nq + 1);
  BP_CUDA(ctx, cudaMemsetAsync(d_max, 0, sizeof(unsigned long long), ctx->stream));
  nb_rowptr_kernel<<<(int)std::min<uint32_t>((nq + 256) / 256, 1024u), 256, 0, ctx->stream>>>(
      offs, nq, S, ctx->d_nb_rowptr.as<uint64_t>(), d_max);
  ctx->launches++;
  uint64_t hostv[2];
  BP_CUDA(ctx, cudaMemcpyAsync(&hostv[0], offs + nw, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaMemcpyAsync(&hostv[1], d_max, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  uint64_t nnz = hostv[0], maxrow = hostv[1];
  BP_TRY(bp_reserve(ctx, ctx->d_nb_tmp, (nnz + 1) * sizeof(double), false));
  BP_TRY(bp_reserve(ctx, ctx->s_f, std::max((nw + 2) * sizeof(uint32_t), (nnz + 1) * sizeof(uint32_t)), false));
  BP_TRY(bp_reserve(ctx, ctx->d_nb_dist, (nnz + 1) * sizeof(double), false));
  BP_TRY(bp_reserve(ctx, ctx->d_nb_col, (nnz + 1) * sizeof(uint32_t), false));
  // s_f is reused for the unsorted ids: the counts are no longer needed after the scan
  a.counts = nullptr;
  a.offs = offs;
  a.out_dist = ctx->d_nb_tmp.as<double>();
  a.out_id = ctx->s_f.as<uint32_t>();
  if (nnz > 0) {
    if (tiled32) nb_tile_kernel_f32<D, true><<<tgrid, NB_TQ, 0, ctx->stream>>>(a, g, chunk);
    else if (tiled) nb_tile_kernel<D, true><<<tgrid, NB_TQ, 0, ctx->stream>>>(a, g, chunk);
    else nb_query_kernel<D, true><<<grid, 256, 0, ctx->stream>>>(a, g);
    ctx->launches++;
    BP_CUDA(ctx, cudaGetLastError());
    const uint64_t* rp = ctx->d_nb_rowptr.as<uint64_t>();
    if (maxrow > NB_MEDIUM) {
      // Huge rows: sort them in place in the unsorted buffers with a global bitonic network, then
      // the copy below moves them (rows <= NB_MEDIUM are rank-sorted out of place).
      std::vector<uint64_t> h_rp((size_t)nq + 1);
      BP_CUDA(ctx, cudaMemcpyAsync(h_rp.data(), rp, h_rp.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
      BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      for (uint32_t q = 0; q < nq; ++q) {
        uint64_t len = h_rp[q + 1] - h_rp[q];
        if (len <= NB_MEDIUM) continue;
        uint64_t npad = 1;
        while (npad < len) npad <<= 1;
        double* dd = a.out_dist + h_rp[q];
        uint32_t* di = a.out_id + h_rp[q];
        int gb = (int)std::min<uint64_t>((npad + 255) / 256, (uint64_t)ctx->num_sms * 8);
        for (uint64_t k = 2; k <= npad; k <<= 1) {
          nb_bitonic_step<<<gb, 256, 0, ctx->stream>>>(dd, di, len, k - 1, npad);
          ctx->launches++;
          for (uint64_t j = k >> 2; j > 0; j >>= 1) {
            nb_bitonic_step<<<gb, 256, 0, ctx->stream>>>(dd, di, len, j, npad);
            ctx->launches++;
          }
        }
        BP_CUDA(ctx, cudaMemcpyAsync(ctx->d_nb_dist.as<double>() + h_rp[q], dd, len * sizeof(double),
                                     cudaMemcpyDeviceToDevice, ctx->stream));
        BP_CUDA(ctx, cudaMemcpyAsync(ctx->d_nb_col.as<uint32_t>() + h_rp[q], di, len * sizeof(uint32_t),
                                     cudaMemcpyDeviceToDevice, ctx->stream));
      }
    }
    int gs = (int)std::min<uint64_t>(((uint64_t)nq * 32 + 255) / 256, (uint64_t)ctx->num_sms * 16);
    nb_sort_small<<<std::max(gs, 1), 256, 0, ctx->stream>>>(rp, nq, a.out_dist, a.out_id, ctx->d_nb_dist.as<double>(),
                                                           ctx->d_nb_col.as<uint32_t>());
    ctx->launches++;
    if (maxrow > NB_SMALL) {
      size_t sm = NB_MEDIUM * (sizeof(double) + sizeof(uint32_t));
      BP_CUDA(ctx, cudaFuncSetAttribute(nb_sort_medium, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      nb_sort_medium<<<(int)std::min<uint32_t>(nq, (uint32_t)ctx->num_sms * 2), 512, sm, ctx->stream>>>(
          rp, nq, a.out_dist, a.out_id, ctx->d_nb_dist.as<double>(), ctx->d_nb_col.as<uint32_t>());
      ctx->launches++;
    }
    BP_CUDA(ctx, cudaGetLastError());
  }
  ctx->nb_nnz = nnz;
  ctx->nb_nq = nq;
  if (nnz_out) *nnz_out = nnz;
  return BPOMP_OK;
}

static int neighbors_dispatch(bpomp_ctx* ctx, const uint32_t* d_query, uint32_t nq, double r, bool all, uint64_t* nnz) {
  switch (ctx->dim) {
    case 1: return neighbors_impl<1>(ctx, d_query, nq, r, all, nnz);
    case 2: return neighbors_impl<2>(ctx, d_query, nq, r, all, nnz);
    case 3: return neighbors_impl<3>(ctx, d_query, nq, r, all, nnz);
    case 4: return neighbors_impl<4>(ctx, d_query, nq, r, all, nnz);
    case 5: return neighbors_impl<5>(ctx, d_query, nq, r, all, nnz);
    case 6: return neighbors_impl<6>(ctx, d_query, nq, r, all, nnz);
    case 7: return neighbors_impl<7>(ctx, d_query, nq, r, all, nnz);
    case 8: return neighbors_impl<8>(ctx, d_query, nq, r, all, nnz);
  }
  return bp_fail(ctx, BPOMP_ERR_INVALID, "unsupported dim %d", ctx->dim);
}

extern "C" int bpomp_neighbors_radius(bpomp_ctx* ctx, const uint32_t* query_ids, uint32_t nq, double r, uint64_t* nnz) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (nq > 0 && !query_ids) return bp_fail(ctx, BPOMP_ERR_INVALID, "query_ids is NULL");
  if (!(r >= 0.0)) return bp_fail(ctx, BPOMP_ERR_INVALID, "radius must be >= 0");
  for (uint32_t i = 0; i < nq; ++i)
    if (query_ids[i] >= ctx->nverts) return bp_fail(ctx, BPOMP_ERR_RANGE, "query id %u out of range", query_ids[i]);
  if (nq == 0) {
    ctx->nb_nnz = 0;
    ctx->nb_nq = 0;
    BP_TRY(bp_reserve(ctx, ctx->d_nb_rowptr, 2 * sizeof(uint64_t), false));
    BP_CUDA(ctx, cudaMemsetAsync(ctx->d_nb_rowptr.p, 0, 2 * sizeof(uint64_t), ctx->stream));
    if (nnz) *nnz = 0;
    return BPOMP_OK;
  }
  BP_TRY(bp_reserve(ctx, ctx->d_nb_query, (size_t)nq * sizeof(uint32_t), false));
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->d_nb_query.p, query_ids, (size_t)nq * sizeof(uint32_t), cudaMemcpyHostToDevice,
                               ctx->stream));
  return neighbors_dispatch(ctx, ctx->d_nb_query.as<uint32_t>(), nq, r, false, nnz);
}

extern "C" int bpomp_neighbors_all(bpomp_ctx* ctx, double r, uint64_t* nnz) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (!(r >= 0.0)) return bp_fail(ctx, BPOMP_ERR_INVALID, "radius must be >= 0");
  if (ctx->nverts == 0) {
    ctx->nb_nnz = 0;
    ctx->nb_nq = 0;
    BP_TRY(bp_reserve(ctx, ctx->d_nb_rowptr, 2 * sizeof(uint64_t), false));
    BP_CUDA(ctx, cudaMemsetAsync(ctx->d_nb_rowptr.p, 0, 2 * sizeof(uint64_t), ctx->stream));
    if (nnz) *nnz = 0;
    return BPOMP_OK;
  }
  return neighbors_dispatch(ctx, nullptr, ctx->nverts, r, true, nnz);
}

extern "C" int bpomp_neighbors_fetch(bpomp_ctx* ctx, uint64_t* row_ptr, uint32_t* col, double* dist) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (row_ptr)
    BP_CUDA(ctx, cudaMemcpyAsync(row_ptr, ctx->d_nb_rowptr.p, ((size_t)ctx->nb_nq + 1) * sizeof(uint64_t),
                                 cudaMemcpyDeviceToHost, ctx->stream));
  if (col && ctx->nb_nnz)
    BP_CUDA(ctx, cudaMemcpyAsync(col, ctx->d_nb_col.p, ctx->nb_nnz * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                 ctx->stream));
  if (dist && ctx->nb_nnz)
    BP_CUDA(ctx, cudaMemcpyAsync(dist, ctx->d_nb_dist.p, ctx->nb_nnz * sizeof(double), cudaMemcpyDeviceToHost,
                                 ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BPOMP_OK;
}

extern "C" int bpomp_neighbors_device(bpomp_ctx* ctx, const uint64_t** d_row_ptr, const uint32_t** d_col,
                                      const double** d_dist) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (d_row_ptr) *d_row_ptr = ctx->d_nb_rowptr.as<uint64_t>();
  if (d_col) *d_col = ctx->d_nb_col.as<uint32_t>();
  if (d_dist) *d_dist = ctx->d_nb_dist.as<double>();
  return BPOMP_OK;
}
