// edge_check_fast.cu — the fast box-world edge kernel (the isValid loop of checkAndSetEdgeBlocked,
// /root/reference/src/BatchingPOMP.cpp:496-544, over edges discretised as initializeEdgePoints does,
// :435-461).  fastpath.cuh holds the arithmetic and its error analysis; this file is the data movement.
//
// One CTA per SM.  Shared memory: the broad-phase mask table in the team layout (128 KB at 7-DoF), the
// 16-bit box bounds (32 B per box), and per warp a staging buffer for one tile of 32 edges plus a pair list.
// A warp processes tiles of 32 consecutive edges:
//   load   the 2 x 32 x D doubles of the tile arrive by cp.async.bulk (TMA, 1-D) on the warp's mbarrier; the
//          copy of the next tile is issued as soon as phase A has read the current one, so it overlaps
//          phases B and C.  (Indexed endpoints are gathered with 8-byte cp.async instead.)
//   A      lane = edge.  Endpoints from shared memory (stride D doubles: conflict-free for odd D), distance,
//          nStates, the cell range per dimension -> one byte per dimension (span class * 32 + first cell), the
//          FP32 slab data (inv, c per dimension) kept in the lane's registers.
//   B      lane = (edge, group of 128 boxes): a team of 4 lanes ANDs the 64-byte mask rows of its edge, one
//          16-byte group per lane.  Even and odd teams walk the dimensions in swapped order and the table
//          interleaves two dimensions per 128-byte line, so the two teams of a quarter-warp always read
//          different halves of the banks: every 16-byte load of the warp is 4 conflict-free wavefronts.
//          Surviving bits are appended as (edge lane, box) pairs to the warp's list.
//   C      lane = pair: FP32 slab test against the 16-bit bounds (two 16-byte loads per box), edge data by
//          warp shuffles from the owner lane; the first slot in the outer range is taken if it is also in the
//          inner range, else verified in FP64 (rare).  atomicMin on the edge's first invalid slot.
// Edges the fast path does not cover are marked BPOMP_EDGE_SLOW for the general kernel (edge_check.cu), which
// runs right after on exactly those edges.
#include <cfloat>

#include "fastpath.cuh"
#include "world.cuh"

#define FE_SLOTCAP 40    // non-zero (edge, group) masks collected before they are expanded to pairs
#define FE_LISTCAP 64    // pairs: fewer than 32 waiting plus at most 32 appended per step
#ifndef FE_MAXWARPS
#define FE_MAXWARPS 20
#endif
#ifndef FE_CHUNK
#define FE_CHUNK 16      // consecutive tiles per static claim (long contiguous runs are kinder to DRAM)
#endif
#ifndef FE_CHUNK_DYN
#define FE_CHUNK_DYN 4   // tiles per dynamic claim of the indexed form
#endif
#ifndef FE_TAIL8
#define FE_TAIL8 7       // eighths of the batch claimed in chunks; the rest in single tiles
#endif

struct FastArgs {
  const double* from;        // [count][D] or null (indexed)
  const double* to;
  const uint32_t* from_ids;  // [count] or null
  const uint32_t* to_ids;
  const double* verts;
  int64_t count;
  uint8_t* valid;
  int32_t* first;
  uint32_t* nstates;         // may be null
  double L;
  float inv_L;               // 1/L as a float (nStates estimate, fp_nstates_thr)
  uint32_t magic;            // 0x4B000000 (2^23 as float bits), an operand of the byte permutes
  unsigned long long* work;  // dynamic tile counter (zero at launch)
  unsigned long long* reset_general;  // tile counter of the general kernel that follows: zeroed here
  int* flags;                // [1] range errors
  int* slow_count;           // number of edges left to the general kernel (its gate)
  uint32_t* slow_list;       // their indices (the first BP_SLOW_LIST_CAP of them)
  int tma_ok;                // from/to are 16-byte aligned: tiles are fetched with cp.async.bulk
  int nwarps;
  // FP64 tables of the exact test (rare path)
  const double2* lohi;
  int lohi_stride;
  const uint32_t* perm_off;
  const double* tfrac;
};

__device__ __forceinline__ uint32_t fe_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fe_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fe_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fe_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fe_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fe_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   fe_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(fe_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fe_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(fe_smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ uint4 fe_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void fe_cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(fe_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void fe_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Per-warp shared-memory block.
template <int D>
struct FeWarp {
  double stage[2 * 32 * D];        // from rows then to rows of the current / next tile
  uint4 slots[FE_SLOTCAP];         // non-zero 128-box masks of (edge, group) slots
  uint16_t smeta[FE_SLOTCAP];      // their edge lane | group << 8
  uint32_t list[FE_LISTCAP];       // (edge lane << 16) | box
  int32_t first[32];               // first invalid slot per edge of the tile
  uint64_t bar;                    // mbarrier of the staging buffer
  uint64_t pad;
};

// Rare path: the search for the first slot >= from_slot (< lim) of edge e inside box b when a sample lies too
// close to a face for the FP32 ranges (or the edge is longer than the shared-memory table of sample orders):
// slots in the OUTER range are taken if they are in the INNER range, else tested in FP64 with the reference's
// arithmetic (RealVectorStateSpace::interpolate, closed-box compare).  Returns the slot or -1.
template <int D, bool INDEXED>
__device__ __noinline__ int fe_search_exact(const FastArgs* a, const float* perm_s, int64_t e, uint32_t n,
                                            int from_slot, int lim, int b, float smin, float smax, float imin,
                                            float imax) {
  const float* ps = perm_s + fp_perm_off(n);
  const double* pf = nullptr;
  const double* pt = nullptr;
  for (int i = from_slot; i < lim; ++i) {
    const float s = __ldg(ps + i);
    if (!(s >= smin && s <= smax)) continue;
    if (s >= imin && s <= imax) return i;
    if (!pf) {
      if (INDEXED) {
        pf = a->verts + (size_t)__ldg(a->from_ids + e) * D;
        pt = a->verts + (size_t)__ldg(a->to_ids + e) * D;
      } else {
        pf = a->from + (size_t)e * D;
        pt = a->to + (size_t)e * D;
      }
    }
    const double tt = __ldg(a->tfrac + __ldg(a->perm_off + n) + i);  // (1+perm[slot])/(n+1), BP.cpp:459
    const double2* row = a->lohi + (size_t)b * a->lohi_stride;
    bool inside = true;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const double fj = __ldg(pf + j), tj = __ldg(pt + j);
      const double x = __dadd_rn(fj, __dmul_rn(__dsub_rn(tj, fj), tt));  // RealVectorStateSpace::interpolate
      const double2 lh = row[j];
      if (!(lh.x <= x && x <= lh.y)) inside = false;
    }
    if (inside) return i;
  }
  return -1;
}

template <int D, bool INDEXED>
__global__ void __launch_bounds__(FE_MAXWARPS * 32, 1)
    edge_check_fast_kernel(const __grid_constant__ FastArgs a, const __grid_constant__ FastWorldView w) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint4* s_tmask = reinterpret_cast<uint4*>(smem);
  uint4* s_qbox = reinterpret_cast<uint4*>(smem + w.tmask_bytes);
  float* s_perm = reinterpret_cast<float*>(smem + w.tmask_bytes + w.qbox_bytes);
  const double* s_thr = reinterpret_cast<const double*>(smem + w.tmask_bytes + w.qbox_bytes + FP_PERM4_BYTES);
  const size_t tables = (w.tmask_bytes + w.qbox_bytes + FP_PERM4_SMEM_BYTES + 127) / 128 * 128;
  FeWarp<D>* ws = reinterpret_cast<FeWarp<D>*>(smem + tables) + (threadIdx.x >> 5);
  __shared__ uint64_t s_tbar;
  const int lane = threadIdx.x & 31;

  // ---- tables -> shared memory, once per CTA (TMA bulk copies on one mbarrier) -----------------------------
  if (threadIdx.x == 0) {
    fe_mbar_init(&s_tbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fe_mbar_expect_tx(&s_tbar, (uint32_t)(w.tmask_bytes + w.qbox_bytes + FP_PERM4_SMEM_BYTES));
    for (size_t o = 0; o < w.tmask_bytes; o += 32768)
      fe_bulk_g2s(reinterpret_cast<unsigned char*>(s_tmask) + o, reinterpret_cast<const unsigned char*>(w.tmask) + o,
                  (uint32_t)min((size_t)32768, w.tmask_bytes - o), &s_tbar);
    for (size_t o = 0; o < w.qbox_bytes; o += 32768)
      fe_bulk_g2s(reinterpret_cast<unsigned char*>(s_qbox) + o, reinterpret_cast<const unsigned char*>(w.qbox) + o,
                  (uint32_t)min((size_t)32768, w.qbox_bytes - o), &s_tbar);
    fe_bulk_g2s(s_perm, w.perm4, FP_PERM4_SMEM_BYTES, &s_tbar);
  }
  if (lane == 0) {
    fe_mbar_init(&ws->bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t tmask_addr = fe_smem_u32(s_tmask);
  const uint32_t qbox_addr = fe_smem_u32(s_qbox);
  const uint32_t perm_addr = fe_smem_u32(s_perm);
  const int64_t ntiles = (a.count + 31) / 32;
  constexpr uint32_t kTileBytes = 32u * D * 8u;  // one endpoint array of a full tile

  // Fetch of tile `tile` into the warp's staging buffer.  Full tiles of 16-byte aligned explicit arrays go
  // through the TMA engine; everything else (indexed endpoints, the last partial tile, unaligned arrays) is
  // gathered with 8-byte cp.async.  Returns the mechanism used (the wait differs).
  auto issue_tile = [&](int64_t tile) -> int {
    const int64_t base = tile * 32;
    const int m = (int)min((int64_t)32, a.count - base);
    if (!INDEXED && a.tma_ok && m == 32) {
      if (lane == 0) {
        fe_mbar_expect_tx(&ws->bar, 2u * kTileBytes);
        fe_bulk_g2s(ws->stage, a.from + (size_t)base * D, kTileBytes, &ws->bar);
        fe_bulk_g2s(ws->stage + 32 * D, a.to + (size_t)base * D, kTileBytes, &ws->bar);
      }
      return 1;
    }
    if (INDEXED) {
      uint32_t idf = 0, idt = 0;
      if (lane < m) {
        idf = __ldg(a.from_ids + base + lane);
        idt = __ldg(a.to_ids + base + lane);
      }
#pragma unroll
      for (int r = 0; r < D; ++r) {  // element i = lane + 32 r of the 32 x D tile: row i / D, column i % D
        const int i = lane + 32 * r;
        const int row = i / D, col = i - row * D;
        const uint32_t vf = __shfl_sync(0xffffffffu, idf, row), vt = __shfl_sync(0xffffffffu, idt, row);
        if (row < m) {
          fe_cp_async8(ws->stage + i, a.verts + (size_t)vf * D + col);
          fe_cp_async8(ws->stage + 32 * D + i, a.verts + (size_t)vt * D + col);
        }
      }
    } else {
      for (int i = lane; i < m * D; i += 32) {
        fe_cp_async8(ws->stage + i, a.from + (size_t)base * D + i);
        fe_cp_async8(ws->stage + 32 * D + i, a.to + (size_t)base * D + i);
      }
    }
    return 2;
  };

  if (blockIdx.x == 0 && threadIdx.x == 0) *a.reset_general = 0ull;

  uint32_t bar_parity = 0;
  // dynamic tile distribution, FE_CHUNK consecutive tiles per claim; the claim for the next chunk is made
  // while the last tile of the current one is processed so that its first tile can be prefetched
  int64_t cur = -1, cur_end = -1;  // current tile and end of its chunk
  int64_t nxt = -1;                // prefetched tile (already issued), -1 none
  int nxt_how = 0;
  // Tile distribution.  The first FE_TAIL8/8 of the batch is dealt statically, FE_CHUNK consecutive tiles at a
  // time, chunk c to warp (c mod #warps) — no atomics (one atomic per tile on a single address would saturate
  // the L2 atomic unit); the rest is claimed dynamically, one tile per atomic, which balances the tail and
  // absorbs SMs that are shared with another kernel (e.g. NCCL's).  A warp always knows its next claim one chunk
  // ahead, so the atomic's latency is not exposed.
  // (indexed endpoints are gathered from all over the vertex table anyway and their tiles differ a lot in cost —
  // the rows of a CSR share a source — so there everything is claimed dynamically, FE_CHUNK_DYN tiles at a time)
  const int64_t big_claims = INDEXED ? 0 : ntiles / 8 * FE_TAIL8 / FE_CHUNK;
  const int64_t nwarps_grid = (int64_t)gridDim.x * (blockDim.x >> 5);
  int64_t stat_next = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  unsigned long long ahead = 0;  // the claim made ahead (lane 0's copy counts)
  auto claim_issue = [&]() {
    if (stat_next < big_claims) {
      ahead = (unsigned long long)stat_next;
      stat_next += nwarps_grid;
    } else if (lane == 0) {
      ahead = (unsigned long long)big_claims + atomicAdd(a.work, 1ull);
    }
  };
  auto claim_take = [&](int64_t& first, int64_t& end) {
    const int64_t c = (int64_t)__shfl_sync(0xffffffffu, ahead, 0);
    if (c < big_claims) {
      first = c * FE_CHUNK;
      end = first + FE_CHUNK;
    } else {
      constexpr int64_t kDyn = INDEXED ? FE_CHUNK_DYN : 1;
      first = big_claims * FE_CHUNK + (c - big_claims) * kDyn;
      end = first + kDyn;
    }
    end = min(end, ntiles);
  };
  {
    int64_t c0, c1;
    claim_issue();
    claim_take(c0, c1);
    claim_issue();
    if (c0 < ntiles) {  // the first tile is on its way while the tables arrive
      cur = c0;
      cur_end = c1;
      nxt = cur;
      nxt_how = issue_tile(cur);
    }
    fe_mbar_wait(&s_tbar, 0);  // tables resident
    if (c0 >= ntiles) return;
  }
  // lanes 4..7 of every quarter-warp walk the dimension pairs in swapped order (and start from another group),
  // so that the eight 16-byte loads of a quarter-warp always fall into eight different bank groups
  const bool oddq = (lane & 4) != 0;
  const uint32_t lt = (1u << lane) - 1u;


  for (;;) {
    // ---- wait for the current tile ------------------------------------------------------------------------
    if (nxt_how == 1) {
      fe_mbar_wait(&ws->bar, bar_parity);
      bar_parity ^= 1u;
    } else {
      fe_cp_async_wait_all();
      __syncwarp();
    }
    const int64_t base = cur * 32;
    const int64_t e = base + lane;
    const bool live = e < a.count;

    // ---- A. lane = edge ------------------------------------------------------------------------------------
    double f[D], d[D];
    {
      const double* sf = ws->stage + lane * D;
      const double* st = ws->stage + 32 * D + lane * D;
#pragma unroll
      for (int j = 0; j < D; ++j) {  // lanes beyond the batch read stale rows; nothing of theirs is stored
        f[j] = sf[j];
        d[j] = st[j];  // holds `to` for now
      }
    }
    __syncwarp();  // every lane has read the staging buffer: the next tile may overwrite it
    {
      int64_t t2 = cur + 1;
      if (t2 >= cur_end) {
        claim_take(t2, cur_end);
        claim_issue();
      }
      if (t2 < ntiles) {
        nxt = t2;
        nxt_how = issue_tile(t2);
      } else {
        nxt = -1;
      }
    }
    // q coordinates of both endpoints: cells, safe zone, slab data
    uint32_t roff[D];  // byte offset of the mask row of each dimension (home half included)
    float inv[D], cc[D];
    bool has_static = false;
    int qmin = 0x7fffffff, qmax = -0x7fffffff, spanmax = 0;
    double dist2 = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const double F = fp_q(f[j], w.qs[j], w.qo[j]);
      const double T = fp_q(d[j], w.qs[j], w.qo[j]);
      // inside the safe zone the truncation equals fp_q16 (the host's cell function); outside it the edge is
      // not evaluated here
      const int qf = __double2int_rz(F), qt = __double2int_rz(T);
      qmin = min(qmin, min(qf, qt));
      qmax = max(qmax, max(qf, qt));
      const int cf = qf >> FP_CELL_SHIFT, ct = qt >> FP_CELL_SHIFT;
      const int c0 = min(cf, ct), span = max(cf, ct) - c0;
      spanmax = max(spanmax, span);
      roff[j] = fp_row_offset(D, j, (uint32_t)(min(span, FP_CLASSES - 1) * FP_CELLS + c0));
      d[j] = __dsub_rn(d[j], f[j]);                     // interpolate: (to - from)
      dist2 = __dadd_rn(dist2, __dmul_rn(d[j], d[j]));  // distance: (from-to)^2 == (to-from)^2 exactly
      // FP32 slab data (fastpath.cuh: fp_edge_dim, with n applied below)
      const float Dqf = (float)(T - F);
      const bool st = fabsf(Dqf) < 1.0f;
      has_static = has_static || st;
      float r;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(Dqf));
      inv[j] = st ? FP_BIG : r;                                   // fastpath.cuh: fp_edge_dim
      cc[j] = fmaf(-inv[j], (float)F, -8388608.0f * inv[j]);
    }
    const bool safe = qmin >= (int)FP_QOFF - 1 && qmax <= (int)(FP_QOFF + FP_QSCALE);
    uint32_t n = fp_nstates_thr(dist2, a.inv_L, s_thr);      // BP.cpp:440-445 through exact thresholds
    if (n == 0u) n = bp_nstates(__dsqrt_rn(dist2), a.L);     // long edge (or NaN): the FP64 formula
    bool need = false;                                       // has slots 1..n-2 to test
    bool slow = false;
    if (live) {
      if (a.nstates) a.nstates[e] = n;
      if (n == 0u) {  // more than BPOMP_NSTATES_MAX states: range error, as the general kernel reports it
        atomicAdd(a.flags + 1, 1);
        a.valid[e] = BPOMP_EDGE_BLOCKED;
        a.first[e] = -1;
      } else if (n <= 2u) {  // BP.cpp:510: no slot in 1..n-2 -> FREE unchecked
        a.valid[e] = BPOMP_EDGE_FREE;
        a.first[e] = -1;
      } else {
        need = true;
        slow = !safe || n > FP_NMAX || spanmax > FP_SPAN_MAX;
      }
    }
    const uint32_t meta = n | (has_static ? 0x10000u : 0u);
    const bool fastedge = need && !slow;
    if (oddq) {  // swapped walk of the dimension pairs
#pragma unroll
      for (int j = 0; j + 1 < D; j += 2) {
        const uint32_t t0 = roff[j];
        roff[j] = roff[j + 1];
        roff[j + 1] = t0;
      }
    }
    ws->first[lane] = 0x7fffffff;
    __syncwarp();

    // ---- C: one round over list[lbase .. lbase+cntr) (lane = pair) -----------------------------------------
    auto pair_round = [&](int lbase, int cntr) {
      const bool pv = lane < cntr;
      const uint32_t entry = pv ? ws->list[lbase + lane] : 0u;
      const int el = (int)(entry >> 16), b = (int)(entry & 0xFFFFu);
      const uint4 qlo = fe_lds128(qbox_addr + (uint32_t)b * 32u);
      const uint4 qhi = fe_lds128(qbox_addr + (uint32_t)b * 32u + 16u);
      const uint32_t m_el = __shfl_sync(0xffffffffu, meta, el);
      const uint32_t n_el = m_el & 0xFFFFu;
      float smin = -3.0e38f, smax = 3.0e38f, imin = smin, imax = smax;
      const uint32_t ql[4] = {qlo.x, qlo.y, qlo.z, qlo.w};
      const uint32_t qh[4] = {qhi.x, qhi.y, qhi.z, qhi.w};
#pragma unroll
      for (int j = 0; j < D; ++j) {
        const float iv = __shfl_sync(0xffffffffu, inv[j], el);
        const float cv = __shfl_sync(0xffffffffu, cc[j], el);
        const float lo_m = __uint_as_float(__byte_perm(ql[j >> 1], a.magic, (j & 1) ? 0x7432 : 0x7410));
        const float hi_m = __uint_as_float(__byte_perm(qh[j >> 1], a.magic, (j & 1) ? 0x7432 : 0x7410));
        fp_slab(lo_m, hi_m, iv, cv, smin, smax, imin, imax);
      }
      fp_range_to_s(n_el, smin, smax, imin, imax);
      // some integer s in [smin, smax]?
      if (pv && smax >= smin && floorf(smax) >= smin) {
        const int curfirst = ws->first[el];
        const int lim = min((int)n_el - 1, curfirst);  // slots 1..n-2, and only those before a known hit
        if (m_el & 0x10000u) imin = 3.0e38f;            // an edge with a static dimension verifies every hit
        int hit = -1, from_slot = 1;  // from_slot > 0: continue with the exact search from there
        if (n_el <= FP_PERM4_NMAX) {
          // s = 1 + perm[slot] of four slots per 16-byte load; rows start at slot 1 and are padded with -1
          const uint32_t rowa = perm_addr + fp_perm4_off(n_el) * 4u;
          from_slot = 0;
          for (int i0 = 1; i0 < lim; i0 += 4) {
            float4 s4;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(s4.x), "=f"(s4.y), "=f"(s4.z), "=f"(s4.w)
                         : "r"(rowa + (uint32_t)(i0 - 1) * 4u));
            const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
            bool found = false;  // the first slot in the OUTER range decides
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float sq = sv[q];
              const bool ok = !found && sq >= smin && sq <= smax && i0 + q < lim;
              if (ok) {
                found = true;
                if (sq >= imin && sq <= imax) hit = i0 + q;
                else from_slot = i0 + q;  // within a few 1e-5 of a face: verified below
              }
            }
            if (found) break;
          }
        }
        if (from_slot > 0)
          hit = fe_search_exact<D, INDEXED>(&a, w.perm_s, base + el, n_el, from_slot, lim, b, smin, smax, imin, imax);
        if (hit >= 0) atomicMin(&ws->first[el], hit);
      }
    };

    // ---- B. lane = edge: AND of the D mask rows, one 128-box group per round (lane l takes group (l + r) & 3);
    // the non-zero (edge, group) masks are collected in ws->slots, then expanded to (edge, box) pairs that go
    // through phase C 32 at a time.  (One copy of each stage in the code: the kernel fits the instruction cache.)
    if (__any_sync(0xffffffffu, fastedge)) {
      int r = 0;
      for (;;) {
        int cnts = 0;  // non-zero slots in ws->slots
#pragma unroll 1
        while (r < 4) {
          const int g = (lane + r) & 3;
          uint4 m = make_uint4(0u, 0u, 0u, 0u);
          if (fastedge) {
            const uint32_t gaddr = tmask_addr + (uint32_t)g * 16u;
            m = fe_lds128(gaddr + roff[0]);
#pragma unroll
            for (int j = 1; j < D; ++j) {
              const uint4 v = fe_lds128(gaddr + roff[j]);
              m.x &= v.x; m.y &= v.y; m.z &= v.z; m.w &= v.w;
            }
          }
          const bool nz = (m.x | m.y | m.z | m.w) != 0u;
          const uint32_t act = __ballot_sync(0xffffffffu, nz);
          const int add = __popc(act);
          if (cnts + add > FE_SLOTCAP) break;  // rare (dense worlds): expand what is there, then redo this round
          if (nz) {
            const int p = cnts + __popc(act & lt);
            ws->slots[p] = m;
            ws->smeta[p] = (uint16_t)(lane | (g << 8));
          }
          cnts += add;
          ++r;
        }
        __syncwarp();
#pragma unroll 1
        for (int s0 = 0; s0 < cnts; s0 += 32) {
          uint4 m = make_uint4(0u, 0u, 0u, 0u);
          uint32_t sm = 0;
          if (s0 + lane < cnts) {
            m = ws->slots[s0 + lane];
            sm = ws->smeta[s0 + lane];
          }
          const uint32_t ebase = ((sm & 0xFFu) << 16) | ((sm >> 8) * 128u);  // (edge lane << 16) | first box of the group
          unsigned long long m01 = (unsigned long long)m.x | ((unsigned long long)m.y << 32);
          unsigned long long m23 = (unsigned long long)m.z | ((unsigned long long)m.w << 32);
          // every lane that still has bits appends one pair per step
          int cntl = 0;
          for (;;) {
            const uint32_t act = __ballot_sync(0xffffffffu, (m01 | m23) != 0ull);
            if ((m01 | m23) != 0ull) {
              int bit;
              if (m01) {
                bit = __ffsll((long long)m01) - 1;
                m01 &= m01 - 1;
              } else {
                bit = 64 + __ffsll((long long)m23) - 1;
                m23 &= m23 - 1;
              }
              ws->list[cntl + __popc(act & lt)] = ebase + (uint32_t)bit;
            }
            cntl += __popc(act);
            if (cntl >= 32 || (act == 0u && cntl > 0)) {  // a full round, or the rest
              __syncwarp();
              const int take = min(cntl, 32);
              pair_round(cntl - take, take);
              cntl -= take;
              __syncwarp();
            }
            if (act == 0u && cntl == 0) break;
          }
        }
        if (r >= 4) break;
      }
    }

    // ---- results (lane = edge) -----------------------------------------------------------------------------
    const bool to_slow = need && slow;
    {
      const uint32_t sm = __ballot_sync(0xffffffffu, to_slow);
      if (sm) {
        int at = 0;
        if (lane == 0) at = atomicAdd(a.slow_count, __popc(sm));
        at = __shfl_sync(0xffffffffu, at, 0) + __popc(sm & lt);
        if (to_slow && at < BP_SLOW_LIST_CAP) a.slow_list[at] = (uint32_t)e;
      }
    }
    if (need) {
      if (to_slow) {
        a.valid[e] = BPOMP_EDGE_SLOW;
        a.first[e] = -2;
      } else {
        const int32_t r = ws->first[lane];
        const int32_t bad = r == 0x7fffffff ? -1 : r;
        a.valid[e] = bad < 0 ? BPOMP_EDGE_FREE : BPOMP_EDGE_BLOCKED;
        a.first[e] = bad;
      }
    }
    __syncwarp();
    if (nxt < 0) break;
    cur = nxt;
  }
}

// ---------------------------------------------------------------------------
// Launch
// ---------------------------------------------------------------------------
template <int D, bool INDEXED>
static int launch_fast(bpomp_ctx* ctx, FastArgs& a) {
  const FastWorldView& w = ctx->fast;
  const size_t tables = (w.tmask_bytes + w.qbox_bytes + FP_PERM4_SMEM_BYTES + 127) / 128 * 128;
  const size_t per_warp = sizeof(FeWarp<D>);
  const size_t avail = (size_t)ctx->smem_optin - 64;  // the static mbarrier
  if (tables + 4 * per_warp > avail) return 0;
  int nwarps = (int)std::min<size_t>(FE_MAXWARPS, (avail - tables) / per_warp);
  nwarps &= ~3;  // equal load on the four schedulers
  if (ctx->fast_warps > 0) nwarps = std::min(nwarps, ctx->fast_warps);
  if (nwarps < 4) return 0;
  a.nwarps = nwarps;
  const size_t smem = tables + (size_t)nwarps * per_warp;
  auto k = edge_check_fast_kernel<D, INDEXED>;
  BP_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntiles = (a.count + 31) / 32;
  // With a communicator, a few SMs are left free: this kernel's CTAs take a whole SM each (shared memory), and
  // NCCL's all-gather of the previous step's flags could otherwise only start when they exit.
  const int sms = std::max(1, ctx->num_sms - bp_reserved_sms(ctx));
  const int grid = (int)std::min<int64_t>((ntiles + nwarps - 1) / nwarps, sms);
  k<<<grid, nwarps * 32, smem, ctx->stream>>>(a, w);
  return 1;
}

template <bool INDEXED>
static int launch_fast_dim(bpomp_ctx* ctx, FastArgs& a) {
  switch (ctx->dim) {
    case 1: return launch_fast<1, INDEXED>(ctx, a);
    case 2: return launch_fast<2, INDEXED>(ctx, a);
    case 3: return launch_fast<3, INDEXED>(ctx, a);
    case 4: return launch_fast<4, INDEXED>(ctx, a);
    case 5: return launch_fast<5, INDEXED>(ctx, a);
    case 6: return launch_fast<6, INDEXED>(ctx, a);
    case 7: return launch_fast<7, INDEXED>(ctx, a);
    case 8: return launch_fast<8, INDEXED>(ctx, a);
  }
  return 0;
}

int bp_launch_edge_check_fast(bpomp_ctx* ctx, const double* d_from, const double* d_to, const uint32_t* d_from_ids,
                              const uint32_t* d_to_ids, int64_t count, uint8_t* d_valid, int32_t* d_first,
                              uint32_t* d_nstates) {
  if (ctx->world != WORLD_BOXES || !ctx->fast.enabled || !ctx->fast_mode) return 0;
  if (count < ctx->fast_min_edges || count > 0xFFFFFFFFll) return 0;
  FastArgs a;
  a.from = d_from; a.to = d_to; a.from_ids = d_from_ids; a.to_ids = d_to_ids;
  a.verts = ctx->d_verts.as<double>();
  a.count = count; a.valid = d_valid; a.first = d_first; a.nstates = d_nstates;
  a.L = ctx->L;
  a.inv_L = (float)(1.0 / ctx->L);
  a.magic = 0x4B000000u;
  BP_TRY(bp_reserve(ctx, ctx->d_work, 80 * sizeof(unsigned long long), false));
  const int l = ctx->work_slot;
  unsigned long long* W = ctx->d_work.as<unsigned long long>();
  if (ctx->work_dirty[l]) {  // first use of this lane, or a call that did not complete: zero both sets
    for (int s2 = 0; s2 < 2; ++s2)
      for (int k = 0; k < 3; ++k)
        BP_CUDA(ctx, cudaMemsetAsync(W + s2 * 32 + k * 8 + l, 0, sizeof(unsigned long long), ctx->stream));
  }
  ctx->work_dirty[l] = true;  // until the general kernel that cleans up behind this launch is queued
  const int set = ctx->work_epoch[l];
  a.work = W + set * 32 + 8 + l;
  a.reset_general = W + set * 32 + l;
  a.flags = ctx->d_flags;
  a.slow_count = reinterpret_cast<int*>(W + set * 32 + 16 + l);
  ctx->last_slow_count = a.slow_count;
  BP_TRY(bp_reserve(ctx, ctx->d_slow_list[ctx->work_slot], (size_t)std::min<int64_t>(count, BP_SLOW_LIST_CAP) * 4, false));
  a.slow_list = ctx->d_slow_list[ctx->work_slot].as<uint32_t>();
  a.tma_ok = (d_from && ((reinterpret_cast<uintptr_t>(d_from) | reinterpret_cast<uintptr_t>(d_to)) & 15u) == 0) ? 1 : 0;
  a.nwarps = 0;
  a.lohi = ctx->boxes.lohi;
  a.lohi_stride = ctx->boxes.lohi_stride;
  a.perm_off = ctx->d_perm_off.as<uint32_t>();
  a.tfrac = ctx->d_tfrac.as<double>();
  const bool indexed = d_from_ids != nullptr;
  int rc = indexed ? launch_fast_dim<true>(ctx, a) : launch_fast_dim<false>(ctx, a);
  if (rc == 0) ctx->work_dirty[l] = false;  // nothing was launched: the counters are still zero
  if (rc <= 0) return rc;
  ctx->launches++;
  BP_CUDA(ctx, cudaGetLastError());
  return 1;
}
