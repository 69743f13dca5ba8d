// common.cuh — internal definitions shared by the sm_100a kernels and the C-ABI glue.
// Public interface: include/bpomp_cuda.h.  Reference citations are relative to /root/reference/.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "bpomp_cuda.h"

#define BP_ABSENT 0xFFFFFFFFu
#define BP_GRID_C 32   // broad-phase cells per dimension
#define BP_SLOW_LIST_CAP (1 << 20)  // entries of the fast kernel's left-over list (beyond it: status scan)
#define BP_SPAN 8      // cell ranges [c0, c0+span], span < BP_SPAN, have a combined mask row

// ---------------------------------------------------------------------------
// Error handling: no exception crosses the ABI.
// ---------------------------------------------------------------------------
int bp_fail(bpomp_ctx* ctx, int code, const char* fmt, ...);

#define BP_CUDA(ctx, call)                                                                        \
  do {                                                                                            \
    cudaError_t _e = (call);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return bp_fail((ctx), BPOMP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                         \
  } while (0)

#define BP_TRY(expr)            \
  do {                          \
    int _r = (expr);            \
    if (_r != BPOMP_OK) return _r; \
  } while (0)

// Growable device buffer (library-owned device memory).
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};
int bp_reserve(bpomp_ctx* ctx, DevBuf& b, size_t bytes, bool keep);  // grows geometrically
void bp_free(DevBuf& b);

// ---------------------------------------------------------------------------
// Device views passed to kernels by value.
// ---------------------------------------------------------------------------
struct PermView {               // per-nStates sample order (util/BisectPerm.hpp + BP.cpp:459)
  const uint32_t* off;          // [BPOMP_NSTATES_MAX+1] offset of row n in tfrac, BP_ABSENT if not resident
  const double* tfrac;          // tfrac[off[n]+i] = double(1+perm_n[i]) / double(n+1)
  const uint32_t* step;         // [BPOMP_NSTATES_MAX+1] (unsigned)(n / log(n)), host libm (BP.cpp:472)
  uint8_t* need;                // [BPOMP_NSTATES_MAX+1] set to 1 for a missing row
  int* flags;                   // device: [0] deferred edges, [1] range errors
};

struct BoxWorldView {
  int nboxes;                   // B
  int w4;                       // candidate-mask width in uint4 units (power of two >= ceil(B/128))
  int lohi_stride;              // row stride of lohi in double2 units (dim | 1)
  const double2* lohi;          // [B][lohi_stride]  (.x = lo, .y = hi)
  const uint4* rmask;           // [w4][rows] range masks: row bp_rrow(j,c0,span) = boxes whose cell interval in
                                // dim j meets [c0, c0+span]; then (for longer ranges) [w4][rows_ab] prefix masks:
                                // row bp_row(j,0,c) = lowest cell <= c, bp_row(j,1,c) = highest cell >= c
  int rows;                     // dim*BP_GRID_C*BP_SPAN
  int rows_ab;                  // dim*2*BP_GRID_C
  size_t ab_off;                // uint4 offset of the prefix masks inside rmask (kept in global memory)
  double g0[BPOMP_DIM_MAX];     // grid origin per dim
  double gs[BPOMP_DIM_MAX];     // cells per unit length per dim
  size_t rmask_bytes, lohi_bytes;
};

// Tables of the fast box-world edge kernel (fastpath.cuh, edge_check_fast.cu).
struct FastWorldView {
  int enabled;                  // 0: this world is evaluated by the general kernel only
  int nps;                      // pairs of dimension slots = (dim+1)/2
  int nboxes;
  double qs[BPOMP_DIM_MAX];     // q_j(x) = fma(x, qs_j, qo_j)
  double qo[BPOMP_DIM_MAX];
  const uint4* tmask;           // broad-phase mask rows, fp_row_offset() layout (fastpath.cuh)
  const uint4* qbox;            // [nboxes][2] x 16 B: lo_q[8] then hi_q[8], u16 each
  const float* perm_s;          // float(1 + perm_n[slot]) rows for n = 2..FP_NMAX at fp_perm_off(n)
  const float* perm4;           // the rows of n <= FP_PERM4_NMAX in the padded layout staged into shared memory
  size_t tmask_bytes, qbox_bytes;
};

struct ImageWorldView {
  int W, H;
  int tiles_x;                  // number of 8x8 tiles per row
  const unsigned long long* bits;  // [tiles_y][tiles_x] 8x8 tile, bit (r&7)*8+(c&7) = pixel >= threshold
};

// Neighbour grid over the first NB_MAXG dimensions of the active vertices (neighbors.cu).
#define NB_MAXG 3
struct GridView {
  int gd;                      // gridded dimensions (<= NB_MAXG)
  int G[NB_MAXG];              // cells per gridded dim
  double lo[NB_MAXG], inv[NB_MAXG];
  const uint32_t* cell_start;  // [ncells+1]
  const uint32_t* ids;         // [nactive] vertex ids in cell order
  const double* pts;           // SoA [D][nactive] in cell order
  uint32_t nactive;
  uint32_t ncells;
};

enum WorldKind { WORLD_NONE = 0, WORLD_BOXES = 1, WORLD_IMAGE = 2 };

// ---------------------------------------------------------------------------
// The context.
// ---------------------------------------------------------------------------
struct bpomp_comm;  // comm.cu

struct bpomp_ctx {
  int device = 0;
  bpomp_comm* comm = nullptr;  // multi-GPU communicator (comm.cu), null for a single GPU
  int dim = 0;
  double L = 0.0;
  int num_sms = 148;
  int smem_optin = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;

  // sample-order tables
  std::vector<uint32_t> h_perm_off;  // host mirror of PermView::off
  uint32_t perm_reserved = 0;        // all n <= perm_reserved are resident
  size_t tfrac_used = 0;             // doubles used in d_tfrac
  DevBuf d_perm_off, d_tfrac, d_step, d_need;
  int* h_flags = nullptr;            // pinned host copy of d_flags
  int* d_flags = nullptr;            // device status flags written by kernels

  // world
  int world = WORLD_NONE;
  BoxWorldView boxes{};
  ImageWorldView image{};
  FastWorldView fast{};
  DevBuf d_lohi, d_rmask, d_bits, d_work, d_tmask, d_qbox, d_perm_s, d_perm4;
  std::vector<double> h_box_lo, h_box_hi;   // host copy of the boxes (tables are rebuilt when bounds change)
  bool have_bounds = false;
  double bounds_lo[BPOMP_DIM_MAX] = {0}, bounds_hi[BPOMP_DIM_MAX] = {0};
  // d_work: 80 x u64.  Two alternating sets s of per-lane counters, so that no launch needs a memset: entry
  // s*32 + l = tile counter of the general kernel, s*32 + 8 + l = tile counter of the fast kernel, s*32 + 16 + l
  // = number of edges the fast kernel left over.  The fast kernel of a call zeroes the general kernel's counter
  // of its set; the general kernel zeroes the other set's fast entries for the next call.  64 + l: tile counter
  // of a general-kernel launch without the fast kernel (memset per launch).
  int work_epoch[4] = {0, 0, 0, 0};
  bool work_dirty[4] = {true, true, true, true};
  const int* last_slow_count = nullptr;
  DevBuf d_slow_list[4];                    // per pipeline lane: indices of the edges the fast kernel left over
  int nccl_max_ctas = -1;                   // CTAs NCCL may use: -1 = as many as SMs are left free, 0 = no limit
  int comm_nranks = 0;                      // ranks of the communicator (0: none)
  int fused_pfree = 1;                      // 0: never use the one-launch free-probability kernel (testing)
  int fast_mode = 1;                        // 0: never use the fast kernel (testing / tuning)
  int reserve_sms = -1;                     // SMs the fast kernel leaves free (-1: bp_reserved_sms' default)
  int fast_warps = 0;                       // > 0: cap on the warps per CTA of the fast kernel (tuning)
  int64_t fast_min_edges = 4096;            // smaller batches go to the general kernel alone (one launch)

  // vertex table (elements of mVertexNN)
  DevBuf d_verts, d_active;
  uint32_t nverts = 0;
  std::vector<uint8_t> h_active;

  // neighbour results
  DevBuf d_nb_rowptr, d_nb_col, d_nb_dist, d_nb_query, d_nb_tmp, d_nb_tmp2;
  uint64_t nb_nnz = 0;
  uint32_t nb_nq = 0;
  // neighbour grid
  DevBuf d_grid_cell_start, d_grid_ids, d_grid_pts, d_grid_keys;
  bool grid_valid = false;   // grid_view describes the current vertex table / active flags for radius grid_r
  GridView grid_view{};
  double grid_r = 0.0;
  uint32_t nactive = 0;      // number of active vertices

  // belief model
  int knn = 15;
  double prior = 0.5, prior_weight = 0.25;
  DevBuf d_bel_pts, d_bel_vals;
  uint64_t bel_count = 0;
  // exact grid index of the belief points for dim <= 3 (belief_grid.cuh): prefix [0, bg_count) indexed
  DevBuf d_bg_start, d_bg_cnt, d_bg_cell, d_bg_pts, d_bg_idx, d_bg_geom;
  uint64_t bg_count = 0;
  uint32_t bg_bits = 0;
  uint64_t bg_builds = 0;
  int prefilter_f32 = 1;                    // 0: FP64 prefilter in the tiled kNN / neighbour kernels for every dimension (testing)
  int belief_grid = 1;                      // 0: always scan the whole model (testing)

  // pipelined host-buffer edge checks (edge_check.cu)
  cudaStream_t lane_stream[3] = {nullptr, nullptr, nullptr};
  DevBuf lane_a[3], lane_b[3], lane_o[3];
  int work_slot = 0;

  // staging / scratch
  DevBuf s_a, s_b, s_c, s_d, s_e, s_f, s_g, s_h, s_i, s_j, s_scan;

  PermView perm_view() const;
};

// Every entry point makes the context's device current first: a process may drive several contexts (one per
// GPU) from one or several threads, and the CUDA current device is per-thread state.
inline void bp_enter(const bpomp_ctx* ctx) { cudaSetDevice(ctx->device); }

// SMs the fast edge kernel leaves free.  Its CTAs take a whole SM each (226 KB of shared memory), so under a
// communicator a few SMs stay empty for NCCL's kernel, which merges the previous batch meanwhile — and NCCL is
// asked for no more CTAs than that (comm.cu), else part of the collective would wait for an edge CTA to exit.
// Measured, 10^7 edges per GPU (ms per step; single GPU 0.398): 2 GPUs 0.435 with 4 SMs, 0.463 with 8;
// 8 GPUs 0.682 with 4 SMs, 0.527 with 2, 0.462 with 8 (the all-gather moves 8 x 2.5 MB there).
inline int bp_reserved_sms(const bpomp_ctx* ctx, int nranks = -1) {
  if (ctx->reserve_sms >= 0) return ctx->reserve_sms;
  if (nranks < 0) nranks = ctx->comm_nranks;
  return nranks <= 1 ? 0 : (nranks <= 2 ? 4 : 8);
}

// internal cross-file helpers
int bp_read_flags(bpomp_ctx* ctx);
int bp_perm_ensure_rows(bpomp_ctx* ctx, const std::vector<uint32_t>& ns);
void bp_bisect_perm_host(uint32_t n, std::vector<int32_t>& first);

// exclusive scan of n u32 counts into n+1 u64 offsets (d_out[n] = total); neighbors.cu
int bp_exclusive_scan_u32(bpomp_ctx* ctx, const uint32_t* d_in, uint64_t n, uint64_t* d_out, DevBuf& tmp);

// kernels' launchers (each bumps ctx->launches)
int bp_launch_edge_check(bpomp_ctx* ctx, const double* d_from, const double* d_to, const uint32_t* d_from_ids,
                         const uint32_t* d_to_ids, int64_t count, uint8_t* d_valid, int32_t* d_first,
                         uint32_t* d_nstates, bool deferred_only);
int bp_launch_states_valid(bpomp_ctx* ctx, const double* d_states, int64_t count, uint8_t* d_out);
// fast box-world kernel (edge_check_fast.cu): evaluates the edges it can and marks the rest BPOMP_EDGE_SLOW;
// returns 1 if it was launched, 0 if it does not apply to this call, < 0 on error
int bp_launch_edge_check_fast(bpomp_ctx* ctx, const double* d_from, const double* d_to, const uint32_t* d_from_ids,
                              const uint32_t* d_to_ids, int64_t count, uint8_t* d_valid, int32_t* d_first,
                              uint32_t* d_nstates);
int bp_build_fast_world(bpomp_ctx* ctx);
// device buffers in and out; uploads missing sample orders and re-runs the affected edges (waits for the stream)
int bp_edges_check_resolved(bpomp_ctx* ctx, const double* d_from, const double* d_to, const uint32_t* d_from_ids,
                            const uint32_t* d_to_ids, int64_t count, uint8_t* d_valid, int32_t* d_first,
                            uint32_t* d_nstates);

// ---------------------------------------------------------------------------
// Device arithmetic.  Everything that must be bit-equal to the reference's unfused x86-64
// FP64 code uses explicit round-to-nearest intrinsics, so it stays exact even if a file
// were compiled without --fmad=false.
// ---------------------------------------------------------------------------
#ifdef __CUDACC__

// RealVectorStateSpace::distance (OMPL; called at BP.cpp:431): sequential sum of squares, sqrt.
template <int D>
__device__ __forceinline__ double bp_dist2(const double* a, const double* b) {
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    double diff = __dsub_rn(a[j], b[j]);
    s = __dadd_rn(s, __dmul_rn(diff, diff));
  }
  return s;
}

// nStates of initializeEdgePoints, BP.cpp:440-445.  Returns 0 if the count is out of range.
__device__ __forceinline__ uint32_t bp_nstates(double dist, double L) {
  double q = floor(__ddiv_rn(dist, L));
  if (!(q >= 2.0)) return 2u;  // also NaN -> 2 is never reached by the reference; keep defined
  if (q > (double)BPOMP_NSTATES_MAX) return 0u;
  return (uint32_t)q;
}

// Monotone cell index used by the broad phase; the host builds the masks with the same function.
__host__ __device__ __forceinline__ int bp_cell(double x, double g0, double gs) {
#ifdef __CUDA_ARCH__
  // the conversion saturates and maps NaN to 0, so clamping the integer gives the same cells as the
  // comparisons of the host branch
  return min(max(__double2int_rz(__dmul_rn(__dsub_rn(x, g0), gs)), 0), BP_GRID_C - 1);
#else
  volatile double t = x - g0;
  double v = t * gs;
  if (!(v > 0.0)) return 0;
  if (v >= (double)BP_GRID_C) return BP_GRID_C - 1;
  return (int)v;
#endif
}

// Row of the range-mask table for cells [c0, c0+span] of dimension j (span < BP_SPAN).  A row is 16 B per
// group, so the shared-memory bank group of a lookup is (row mod 8).  The minor index is c0 rotated by 3*span:
// random segments spread evenly over the bank groups (c0 is uniform; span is 0-1 for most segments), and
// the lanes of a warp that walk one CSR row (same source, nearby targets: neighbouring (c0, span) pairs)
// also land in different bank groups.  Measured: span-minor 0.877 ms, c0-minor 0.795 ms per 1e7 random
// edges; rotation 1.435 -> 1.272 ms on the 9.6e6 CSR edges of config 3.
__host__ __device__ __forceinline__ int bp_rrow(int j, int c0, int span) {
  return (j * BP_SPAN + span) * BP_GRID_C + ((c0 + 3 * span) & (BP_GRID_C - 1));
}
// Row of the prefix-mask table: which = 0 -> "lowest cell <= c", which = 1 -> "highest cell >= c".
__host__ __device__ __forceinline__ int bp_row(int j, int which, int c) { return (j * 2 + which) * BP_GRID_C + c; }

// Index of uint4 group g (boxes 128g .. 128g+127) of mask row `row`: the table is group-major,
// [w4][rows], so the lanes of a warp (different rows, same g) spread over all shared-memory banks.
__host__ __device__ __forceinline__ size_t bp_midx(int row, int g, int rows) { return (size_t)g * rows + row; }

#endif  // __CUDACC__
