// edge_check.cu — batched edge evaluation (the isValid loop of checkAndSetEdgeBlocked,
// /root/reference/src/BatchingPOMP.cpp:496-544, over edges discretised as
// initializeEdgePoints does, :435-461) against the device-resident collision world.
//
// See the comment above edge_check_boxes_kernel for the work decomposition (lane = edge for the
// per-edge phases, lane = sample for the collision tests; DESIGN.md explains why neither a plain
// thread-per-edge nor a warp-per-edge mapping fits: an FP64 instruction costs the same issue slots
// for 1 or 32 active lanes, and a 7-DoF edge at the default resolution has only ~5 checked samples).
#include <cfloat>

#include "fastpath.cuh"
#include "world.cuh"

#ifndef EC_THREADS
#define EC_THREADS 768
#endif
#define EC_MAXCAND 16
#define EC_CHUNK 8   // tiles of 32 edges claimed per atomic

struct EdgeArgs {
  const double* from;        // [count][D] or null (indexed)
  const double* to;
  const uint32_t* from_ids;  // [count] or null
  const uint32_t* to_ids;
  const double* verts;       // vertex table for the indexed form
  int64_t count;
  uint8_t* valid;
  int32_t* first;
  uint32_t* nstates;         // may be null
  int only_status;           // 0: every edge; else only edges whose status byte equals it (BPOMP_EDGE_DEFERRED
                             // on the re-run after a sample-order upload, BPOMP_EDGE_SLOW after the fast kernel)
  const int* gate;           // non-null: number of edges the fast kernel left to this pass; the CTA exits at
                             // once when it is 0
  const uint32_t* list;      // with gate: their indices, when *gate <= list_cap (else the status bytes are scanned)
  int list_cap;
  unsigned long long* reset0;  // with gate: counters of the other set, zeroed here for the next call's fast kernel
  unsigned long long* reset1;
  double L;
  unsigned long long* work;  // dynamic tile counter (zeroed before the launch)
};

// ---- bulk async copy global -> shared (TMA engine, 1-D), completion on an mbarrier ----------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

template <int D, bool INDEXED>
__device__ __forceinline__ void load_endpoints(const EdgeArgs& a, int64_t e, double* f, double* t) {
  const double* pf;
  const double* pt;
  if (INDEXED) {
    pf = a.verts + (size_t)__ldg(a.from_ids + e) * D;
    pt = a.verts + (size_t)__ldg(a.to_ids + e) * D;
  } else {
    pf = a.from + (size_t)e * D;
    pt = a.to + (size_t)e * D;
  }
  if ((D & 1) == 0) {
#pragma unroll
    for (int j = 0; j < D; j += 2) {
      double2 vf = __ldg(reinterpret_cast<const double2*>(pf + j));
      double2 vt = __ldg(reinterpret_cast<const double2*>(pt + j));
      f[j] = vf.x; f[j + 1] = vf.y;
      t[j] = vt.x; t[j + 1] = vt.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < D; ++j) {
      f[j] = __ldg(pf + j);
      t[j] = __ldg(pt + j);
    }
  }
}

// Phase A shared by both worlds.  Returns false when the edge is already decided (outputs written).
template <int D>
__device__ __forceinline__ bool edge_prologue_loaded(const EdgeArgs& a, const PermView& pv, int64_t e, const double* f,
                                                     const double* t, uint32_t& n, uint32_t& off, int only_status);

template <int D, bool INDEXED>
__device__ __forceinline__ bool edge_prologue(const EdgeArgs& a, const PermView& pv, int64_t e, double* f, double* t,
                                              uint32_t& n, uint32_t& off, int only_status) {
  if (only_status && a.valid[e] != (uint8_t)only_status) return false;
  load_endpoints<D, INDEXED>(a, e, f, t);
  return edge_prologue_loaded<D>(a, pv, e, f, t, n, off, 0);
}

// The same with the endpoints already in registers (the image kernel fetches them one round ahead).
template <int D>
__device__ __forceinline__ bool edge_prologue_loaded(const EdgeArgs& a, const PermView& pv, int64_t e, const double* f,
                                                     const double* t, uint32_t& n, uint32_t& off, int only_status) {
  if (only_status && a.valid[e] != (uint8_t)only_status) return false;
  double dist = __dsqrt_rn(bp_dist2<D>(f, t));  // BP.cpp:163 distance(source, target)
  n = bp_nstates(dist, a.L);                     // BP.cpp:440-445
  if (n == 0u) {                                 // more than BPOMP_NSTATES_MAX states
    atomicAdd(pv.flags + 1, 1);
    a.valid[e] = BPOMP_EDGE_BLOCKED;
    a.first[e] = -1;
    if (a.nstates) a.nstates[e] = 0u;
    return false;
  }
  if (a.nstates) a.nstates[e] = n;
  if (n <= 2u) {  // BP.cpp:510: no slot in 1..n-2 -> FREE unchecked
    a.valid[e] = BPOMP_EDGE_FREE;
    a.first[e] = -1;
    return false;
  }
  off = __ldg(pv.off + n);
  if (off == BP_ABSENT) {  // sample order for this n not resident: host uploads it and re-runs
    pv.need[n] = 1;
    atomicAdd(pv.flags, 1);
    a.valid[e] = BPOMP_EDGE_DEFERRED;
    a.first[e] = -2;
    return false;
  }
  return true;
}

// ---------------------------------------------------------------------------
// Box world.
//
// A warp owns 32 consecutive edges at a time and runs three phases with uniform control flow:
//   A/B (lane = edge)   endpoints, distance, nStates; broad phase = AND over the dimensions of the
//                       prefix masks of the edge's cell range (shared memory) -> candidate boxes;
//                       edges without candidates are FREE at once.  Endpoint data stays in the
//                       owner lane's registers (phase C fetches it with warp shuffles), which keeps
//                       the per-warp shared-memory scratch at 1.1 KB and lets 24 warps share an SM
//                       with the 113 KB of staged tables.
//   C   (lane = sample) the (edge, slot) pairs of the 32 edges are flattened by a warp prefix sum
//                       and dealt to the lanes 32 at a time, whatever the mix of nStates — a long
//                       edge simply spreads over many lanes.  Each lane interpolates its state with
//                       the reference's arithmetic and tests it against its edge's candidates; the
//                       first invalid slot is a shared-memory atomicMin.
// ---------------------------------------------------------------------------
#define EC_WARPS (EC_THREADS / 32)

struct WarpScratch {
  int32_t first[32];                         // first invalid slot found so far, per edge of the tile
  unsigned short list[EC_MAXCAND * 32];      // candidate boxes [c][lane]
  unsigned char nzlane[32];                  // lanes of the edges that have slots to test, in lane order
};

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// W4 = number of 128-box groups when it is 1, 2, 4 or 8 (compile-time fast path), 0 = any (runtime loop).
template <int D, bool INDEXED, bool STAGE, int W4>
__global__ void __launch_bounds__(EC_THREADS, 1)
    edge_check_boxes_kernel(EdgeArgs a, BoxWorldView w, PermView pv) {
  extern __shared__ __align__(16) unsigned char smem[];
  int only_status = a.only_status;
  int64_t total = a.count;       // edges this launch walks
  const uint32_t* elist = nullptr;  // their indices (null: 0..count-1)
  if (a.gate) {
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the fast kernel has completed and its writes are visible
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      *a.reset0 = 0ull;
      *a.reset1 = 0ull;
    }
    const int left = *reinterpret_cast<const volatile int*>(a.gate);
    if (left == 0) return;
    if (a.list && left <= a.list_cap) {
      total = left;
      elist = a.list;
      only_status = 0;
      // a short list needs few CTAs (one tile per warp); the others leave before staging the tables
      if ((int64_t)blockIdx.x * EC_WARPS * 32 >= total) return;
    }
  }
  // layout: [WarpScratch x EC_WARPS][mbarrier 16 B][prefix masks][lohi]
  WarpScratch* ws = reinterpret_cast<WarpScratch*>(smem) + (threadIdx.x >> 5);
  constexpr size_t kScratch = (sizeof(WarpScratch) * EC_WARPS + 15) / 16 * 16;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + kScratch);
  const uint4* rmask = w.rmask;
  const double2* lohi = w.lohi;
  uint32_t s_rmask_addr = 0;
  if (STAGE) {
    uint4* s_rmask = reinterpret_cast<uint4*>(smem + kScratch + 16);
    double2* s_lohi = reinterpret_cast<double2*>(reinterpret_cast<unsigned char*>(s_rmask) + w.rmask_bytes);
    if (threadIdx.x == 0) {
      mbar_init(s_bar, 1);
      mbar_expect_tx(s_bar, (uint32_t)(w.rmask_bytes + w.lohi_bytes));
      // bulk copies are issued in <= 32 KB pieces
      for (size_t o = 0; o < w.rmask_bytes; o += 32768)
        bulk_g2s(reinterpret_cast<unsigned char*>(s_rmask) + o, reinterpret_cast<const unsigned char*>(w.rmask) + o,
                 (uint32_t)min((size_t)32768, w.rmask_bytes - o), s_bar);
      for (size_t o = 0; o < w.lohi_bytes; o += 32768)
        bulk_g2s(reinterpret_cast<unsigned char*>(s_lohi) + o, reinterpret_cast<const unsigned char*>(w.lohi) + o,
                 (uint32_t)min((size_t)32768, w.lohi_bytes - o), s_bar);
    }
    __syncthreads();
    mbar_wait(s_bar, 0);
    rmask = s_rmask;
    lohi = s_lohi;
    s_rmask_addr = smem_u32(s_rmask);
  }

  const int lane = threadIdx.x & 31;
  volatile int32_t* vfirst = ws->first;

  // Dynamic work distribution: a warp claims EC_CHUNK consecutive tiles of 32 edges at a time from a
  // global counter, so the kernel keeps its throughput when some SMs are busy with other work (e.g. the
  // NCCL kernel that merges the previous step's results) and the tail is balanced.
  for (;;) {
    unsigned long long chunk = 0;
    if (lane == 0) chunk = atomicAdd(a.work, 1ull);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    // a left-over list is short: one tile per claim spreads it over all warps
    const int chunk_tiles = elist ? 1 : EC_CHUNK;
    const int64_t chunk_base = (int64_t)chunk * (chunk_tiles * 32);
    if (chunk_base >= total) break;
  for (int ti = 0; ti < chunk_tiles; ++ti) {
    const int64_t base = chunk_base + (int64_t)ti * 32;
    if (base >= total) break;
    const bool inb = base + lane < total;
    const int64_t e = !inb ? 0 : (elist ? (int64_t)__ldg(elist + base + lane) : base + lane);
    if (!INDEXED && !elist && ti + 1 < chunk_tiles) {  // pull the next tile's endpoint rows (contiguous) towards L2
      const int64_t nb = base + 32;
      constexpr int kLines = (32 * D * 8 + 127) / 128 + 1;
      if (nb < a.count && lane < kLines) {
        const size_t o = (size_t)nb * D * 8 + (size_t)lane * 128;
        if (o < (size_t)a.count * D * 8) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(a.from) + o));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(a.to) + o));
        }
      }
    }
    // ---- A. endpoints, distance, nStates (lane = edge) --------------------------------------
    double f[D], t[D];
#pragma unroll
    for (int j = 0; j < D; ++j) { f[j] = 0.0; t[j] = 0.0; }
    uint32_t n = 0, off = 0;
    bool need = inb && edge_prologue<D, INDEXED>(a, pv, e, f, t, n, off, only_status);

    // ---- B. broad phase ------------------------------------------------------------------------
    int nc = 0;
    if (need && w.nboxes > 0) {
      int rr[D];  // >= 0: combined range row; < 0: -(1 + c0*BP_GRID_C + c1), prefix rows from global memory
      bool longspan = false;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        // bp_cell is monotone: the cells of the two endpoints bound the cells of every sample between them
        const int cf = bp_cell(f[j], w.g0[j], w.gs[j]), ct = bp_cell(t[j], w.g0[j], w.gs[j]);
        const int c0 = min(cf, ct), c1 = max(cf, ct);
        rr[j] = (c1 - c0 < BP_SPAN) ? bp_rrow(j, c0, c1 - c0) : -(1 + c0 * BP_GRID_C + c1);
        longspan = longspan || rr[j] < 0;
      }
      auto emit = [&](int g, uint4 m) {  // append the set bits of one 128-box group to the candidate list
        unsigned long long m01 = (unsigned long long)m.x | ((unsigned long long)m.y << 32);
        unsigned long long m23 = (unsigned long long)m.z | ((unsigned long long)m.w << 32);
        while (m01) {
          int b = g * 128 + (__ffsll((long long)m01) - 1);
          m01 &= m01 - 1;
          if (nc < EC_MAXCAND) ws->list[nc * 32 + lane] = (unsigned short)b;
          ++nc;
        }
        while (m23) {
          int b = g * 128 + 64 + (__ffsll((long long)m23) - 1);
          m23 &= m23 - 1;
          if (nc < EC_MAXCAND) ws->list[nc * 32 + lane] = (unsigned short)b;
          ++nc;
        }
      };
      if (STAGE && W4 > 0 && !longspan) {
        // fast path: table in shared memory, group count and row count known at compile time, so every
        // load is one LDS.128 at (row address + immediate)
        constexpr int kRows = D * BP_GRID_C * BP_SPAN;
        uint32_t ra[D];
#pragma unroll
        for (int j = 0; j < D; ++j) ra[j] = s_rmask_addr + (uint32_t)rr[j] * 16u;
#pragma unroll 1
        for (int g = 0; g < (W4 > 0 ? W4 : 1); ++g) {
          uint4 m = lds128(ra[0] + (uint32_t)(g * kRows * 16));
#pragma unroll
          for (int j = 1; j < D; ++j) {
            const uint4 v = lds128(ra[j] + (uint32_t)(g * kRows * 16));
            m.x &= v.x; m.y &= v.y; m.z &= v.z; m.w &= v.w;
          }
          emit(g, m);
        }
      } else {
        for (int g = 0; g < w.w4; ++g) {
          uint4 m = make_uint4(~0u, ~0u, ~0u, ~0u);
#pragma unroll
          for (int j = 0; j < D; ++j) {
            uint4 v;
            if (rr[j] >= 0) {
              v = rmask[bp_midx(rr[j], g, w.rows)];
            } else {  // long range in this dimension (rare): lowest cell <= c1 and highest cell >= c0
              const int cc = -rr[j] - 1, c0 = cc / BP_GRID_C, c1 = cc % BP_GRID_C;
              const uint4 va = __ldg(w.rmask + w.ab_off + bp_midx(bp_row(j, 0, c1), g, w.rows_ab));
              const uint4 vb = __ldg(w.rmask + w.ab_off + bp_midx(bp_row(j, 1, c0), g, w.rows_ab));
              v = make_uint4(va.x & vb.x, va.y & vb.y, va.z & vb.z, va.w & vb.w);
            }
            m.x &= v.x; m.y &= v.y; m.z &= v.z; m.w &= v.w;
          }
          emit(g, m);
        }
      }
    }
    if (need && nc == 0) {  // no box can touch the edge: FREE without interpolating a state
      a.valid[e] = BPOMP_EDGE_FREE;
      a.first[e] = -1;
      need = false;
    }
#pragma unroll
    for (int j = 0; j < D; ++j) t[j] = __dsub_rn(t[j], f[j]);  // interpolate: (to - from); t now holds it
    const uint32_t cnt = need ? n - 2u : 0u;                    // slots 1..n-2, BP.cpp:510
    // -1: list overflow -> test every box; the list holds 16-bit ids, so worlds with more boxes than that
    // always take the every-box path
    const int ncl = (nc > EC_MAXCAND || w.nboxes > 65536) ? -1 : nc;
    ws->first[lane] = 0x7fffffff;
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    const uint32_t start = incl - cnt;  // exclusive prefix of the slots to test per edge
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    {
      const uint32_t nz = __ballot_sync(0xffffffffu, cnt > 0);
      if (cnt > 0) ws->nzlane[__popc(nz & ((1u << lane) - 1u))] = (unsigned char)lane;
    }
    uint32_t before = 0;  // edges with slots whose first item lies before the current window
    __syncwarp();

    // ---- C. flattened (edge, slot) pairs (lane = sample).  The edge data stays in the owner lane's
    // registers and is fetched with warp shuffles; every lane executes the shuffles of every round. ----
    for (uint32_t ib = 0; ib < total; ib += 32) {
      const uint32_t it = ib + lane;
      // edge of item `it`: the window's edge heads as a bit mask (one warp OR-reduction), then the rank of
      // the last head at or before this lane indexes the list of edges that have slots
      const uint32_t rel = start - ib;
      const uint32_t heads = __reduce_or_sync(0xffffffffu, (cnt > 0 && rel < 32u) ? (1u << rel) : 0u);
      const int el = ws->nzlane[(before + __popc(heads & (0xffffffffu >> (31 - lane))) - 1u) & 31u];
      before += __popc(heads);
      const int32_t slot = (int32_t)(1u + it - __shfl_sync(0xffffffffu, start, el));
      const uint32_t off_el = __shfl_sync(0xffffffffu, off, el);
      const int ncl_el = __shfl_sync(0xffffffffu, ncl, el);
      const bool act = it < total && slot < vfirst[el];  // a smaller invalid slot already known: skip
      double tt = 0.0;
      if (act) tt = __ldg(pv.tfrac + off_el + slot);    // (1+perm[i])/(n+1), BP.cpp:459
      double x[D];
#pragma unroll
      for (int j = 0; j < D; ++j) {
        const double fj = __shfl_sync(0xffffffffu, f[j], el);
        const double dj = __shfl_sync(0xffffffffu, t[j], el);
        x[j] = __dadd_rn(fj, __dmul_rn(dj, tt));
      }
      if (act) {
        bool inside = false;
        if (ncl_el >= 0) {
          int b = ws->list[el];  // the next candidate id is read while the current box is tested
          for (int c = 0; c < ncl_el; ++c) {
            const int bn = ws->list[min(c + 1, EC_MAXCAND - 1) * 32 + el];
            if (bp_inside_box_grouped<D>(x, lohi + (size_t)b * w.lohi_stride)) { inside = true; break; }
            b = bn;
          }
        } else {  // more candidates than the list holds: test every box (rare)
          for (int b = 0; b < w.nboxes; ++b)
            if (bp_inside_box<D>(x, lohi + (size_t)b * w.lohi_stride)) { inside = true; break; }
        }
        if (inside) atomicMin(&ws->first[el], slot);
      }
    }
    __syncwarp();
    if (need) {
      const int32_t bad = vfirst[lane] == 0x7fffffff ? -1 : vfirst[lane];
      a.valid[e] = bad < 0 ? BPOMP_EDGE_FREE : BPOMP_EDGE_BLOCKED;
      a.first[e] = bad;
    }
    __syncwarp();
  }
  }
}

// ---------------------------------------------------------------------------
// Image world (2-D)
// ---------------------------------------------------------------------------
// Thread per edge.  The samples of an edge lie inside the pixel bounding box of its endpoints (floor(x*W) is
// monotone), and a roadmap edge is a few pixels long: when that box spans at most 2 x 2 tiles of 8 x 8 pixels the
// four tiles (32 bytes, one byte per pixel row of a tile) are copied ONCE into the thread's window in shared
// memory (stride 9 words: conflict-free) and every sample is one byte load from that window — no global load
// depends on a sample, so there is no dependent-load chain and no early exit: every lane runs the same
// straight-line code over the slots and the invalid slots are collected in a bit mask (first invalid slot =
// lowest set bit).  The few longer edges of a warp are then evaluated cooperatively, lane = sample.
#define EC_IMG_THREADS 256
template <bool INDEXED>
__global__ void __launch_bounds__(EC_IMG_THREADS) edge_check_image_kernel(EdgeArgs a, ImageWorldView im, PermView pv) {
  if (a.gate && *reinterpret_cast<const volatile int*>(a.gate) == 0) return;  // (never set for the image world)
  __shared__ uint32_t s_win[EC_IMG_THREADS * 9];
  const double Wd = (double)im.W, Hd = (double)im.H;
  const int lane = threadIdx.x & 31;
  uint32_t* win = s_win + threadIdx.x * 9;
  const unsigned char* winb = reinterpret_cast<const unsigned char*>(win);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t rounds = (a.count + stride - 1) / stride;
  const int64_t e0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // Software pipeline over the rounds: the endpoints of round r+1 (and, for indexed endpoints, the ids of round
  // r+2 they are gathered with) are requested before round r is evaluated, so the dependent id -> vertex-table
  // loads are in flight during a whole round of arithmetic instead of being waited for at its start.
  uint32_t idf = 0, idt = 0;  // ids of the round after next (indexed form)
  double nf[2] = {0.0, 0.0}, nt[2] = {0.0, 0.0};  // endpoints of the next round
  auto fetch_ids = [&](int64_t e) {
    if (INDEXED && e < a.count) {
      idf = __ldg(a.from_ids + e);
      idt = __ldg(a.to_ids + e);
    }
  };
  auto fetch_endpoints = [&](int64_t e) {  // uses idf / idt fetched for this very edge
    if (e < a.count) {
      const double* pf = INDEXED ? a.verts + (size_t)idf * 2 : a.from + (size_t)e * 2;
      const double* pt = INDEXED ? a.verts + (size_t)idt * 2 : a.to + (size_t)e * 2;
      const double2 vf = __ldg(reinterpret_cast<const double2*>(pf));
      const double2 vt = __ldg(reinterpret_cast<const double2*>(pt));
      nf[0] = vf.x; nf[1] = vf.y;
      nt[0] = vt.x; nt[1] = vt.y;
    }
  };
  fetch_ids(e0);
  fetch_endpoints(e0);
  fetch_ids(e0 + stride);
  for (int64_t rd = 0; rd < rounds; ++rd) {  // warp-uniform trip count: the lanes of a warp stay together
    const int64_t e = rd * stride + e0;
    const double f[2] = {nf[0], nf[1]}, t[2] = {nt[0], nt[1]};
    fetch_endpoints(e + stride);   // round rd+1, with the ids requested a round ago
    fetch_ids(e + 2 * stride);     // round rd+2
    uint32_t n = 0, off = 0;
    const bool need = e < a.count && edge_prologue_loaded<2>(a, pv, e, f, t, n, off, a.only_status);
    const double d0 = __dsub_rn(t[0], f[0]), d1 = __dsub_rn(t[1], f[1]);
    const double* tf = pv.tfrac + off;
    int32_t bad = 0x7fffffff;
    // strictly inside [0,1): no sample reaches pixel W (x = 1.0) and no clamp is needed
    const bool inside = f[0] >= 0.0 && f[0] < 1.0 && f[1] >= 0.0 && f[1] < 1.0 && t[0] >= 0.0 && t[0] < 1.0 &&
                        t[1] >= 0.0 && t[1] < 1.0;
    bool fastpath = false;
    if (need && inside && n <= 33u) {
      const int cf = bp_floor_int(__dmul_rn(f[0], Wd)), ct = bp_floor_int(__dmul_rn(t[0], Wd));
      const int rf = bp_floor_int(__dmul_rn(f[1], Hd)), rt = bp_floor_int(__dmul_rn(t[1], Hd));
      const int tx0 = min(cf, ct) >> 3, tx1 = max(cf, ct) >> 3, ty0 = min(rf, rt) >> 3, ty1 = max(rf, rt) >> 3;
      if (tx1 - tx0 <= 1 && ty1 - ty0 <= 1) {
        fastpath = true;
        const unsigned long long* r0 = im.bits + (size_t)ty0 * im.tiles_x;
        const unsigned long long* r1 = im.bits + (size_t)ty1 * im.tiles_x;
        const unsigned long long T00 = __ldg(r0 + tx0), T01 = __ldg(r0 + tx1), T10 = __ldg(r1 + tx0), T11 = __ldg(r1 + tx1);
        // window bytes: [(tile row * 2 + tile col) * 8 + pixel row] = the 8 pixels of that row of that tile
        win[0] = (uint32_t)T00; win[1] = (uint32_t)(T00 >> 32); win[2] = (uint32_t)T01; win[3] = (uint32_t)(T01 >> 32);
        win[4] = (uint32_t)T10; win[5] = (uint32_t)(T10 >> 32); win[6] = (uint32_t)T11; win[7] = (uint32_t)(T11 >> 32);
        const int cbase = tx0 << 3, rbase = ty0 << 3;
        uint32_t invalid = 0;  // bit i: slot i is in collision
#pragma unroll 4
        for (uint32_t i = 1; i + 1 < n; ++i) {
          const double tt = __ldg(tf + i);
          const double x0 = __dadd_rn(f[0], __dmul_rn(d0, tt));  // inside the endpoints' box
          const double x1 = __dadd_rn(f[1], __dmul_rn(d1, tt));
          const int col = bp_floor_int(__dmul_rn(x0, Wd)) - cbase;  // 0..15
          const int row = bp_floor_int(__dmul_rn(x1, Hd)) - rbase;
          const uint32_t byte = winb[((row & 8) << 1) | (col & 8) | (row & 7)];
          invalid |= (((byte >> (col & 7)) & 1u) ^ 1u) << i;
        }
        if (invalid) bad = __ffs((int)invalid) - 1;
      }
    }
    // the other edges of the warp (long, or touching the border of the unit square): lane = sample
    uint32_t slow = __ballot_sync(0xffffffffu, need && !fastpath);
    while (slow) {
      const int src = __ffs((int)slow) - 1;
      slow &= slow - 1;
      const double sf0 = __shfl_sync(0xffffffffu, f[0], src), sf1 = __shfl_sync(0xffffffffu, f[1], src);
      const double sd0 = __shfl_sync(0xffffffffu, d0, src), sd1 = __shfl_sync(0xffffffffu, d1, src);
      const uint32_t sn = __shfl_sync(0xffffffffu, n, src), soff = __shfl_sync(0xffffffffu, off, src);
      int32_t first = 0x7fffffff;
      for (uint32_t i0 = 1; i0 + 1 < sn && first == 0x7fffffff; i0 += 32) {
        const uint32_t i = i0 + lane;
        bool hit = false;
        if (i + 1 < sn) {
          const double tt = __ldg(pv.tfrac + soff + i);
          hit = !bp_valid_image(__dadd_rn(sf0, __dmul_rn(sd0, tt)), __dadd_rn(sf1, __dmul_rn(sd1, tt)), im);
        }
        const uint32_t hits = __ballot_sync(0xffffffffu, hit);
        if (hits) first = (int32_t)(i0 + __ffs((int)hits) - 1);
      }
      if (lane == src) bad = first;
    }
    if (need) {
      const bool blocked = bad != 0x7fffffff;
      a.valid[e] = blocked ? BPOMP_EDGE_BLOCKED : BPOMP_EDGE_FREE;
      a.first[e] = blocked ? bad : -1;
    }
  }
}

// ---------------------------------------------------------------------------
// Launch
// ---------------------------------------------------------------------------
template <int D, bool INDEXED>
static int launch_boxes(bpomp_ctx* ctx, const EdgeArgs& a) {
  const BoxWorldView& w = ctx->boxes;
  size_t base = (sizeof(WarpScratch) * EC_WARPS + 15) / 16 * 16 + 16;
  size_t staged = base + w.rmask_bytes + w.lohi_bytes;
  bool stage = staged <= (size_t)ctx->smem_optin && w.rmask_bytes + w.lohi_bytes < (1u << 20) && w.nboxes > 0;
  int64_t tiles = (a.count + EC_THREADS - 1) / EC_THREADS;
  int grid = (int)std::min<int64_t>(tiles, ctx->num_sms);
  if (stage) {
    void (*k)(EdgeArgs, BoxWorldView, PermView);
    switch (w.w4) {
      case 1: k = edge_check_boxes_kernel<D, INDEXED, true, 1>; break;
      case 2: k = edge_check_boxes_kernel<D, INDEXED, true, 2>; break;
      case 4: k = edge_check_boxes_kernel<D, INDEXED, true, 4>; break;
      default: k = edge_check_boxes_kernel<D, INDEXED, true, 0>; break;
    }
    BP_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged));
    if (a.gate) {
      // follow-up of the fast kernel: programmatic dependent launch, so that this kernel's launch latency is
      // hidden behind the fast kernel (it waits for it with griddepcontrol.wait before touching its results)
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(EC_THREADS);
      cfg.dynamicSmemBytes = staged;
      cfg.stream = ctx->stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      BP_CUDA(ctx, cudaLaunchKernelEx(&cfg, k, a, w, ctx->perm_view()));
    } else {
      k<<<grid, EC_THREADS, staged, ctx->stream>>>(a, w, ctx->perm_view());
    }
  } else {
    auto k = edge_check_boxes_kernel<D, INDEXED, false, 0>;
    BP_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)base));
    k<<<grid, EC_THREADS, base, ctx->stream>>>(a, w, ctx->perm_view());
  }
  return BPOMP_OK;
}

template <bool INDEXED>
static int launch_boxes_dim(bpomp_ctx* ctx, const EdgeArgs& a) {
  switch (ctx->dim) {
    case 1: return launch_boxes<1, INDEXED>(ctx, a);
    case 2: return launch_boxes<2, INDEXED>(ctx, a);
    case 3: return launch_boxes<3, INDEXED>(ctx, a);
    case 4: return launch_boxes<4, INDEXED>(ctx, a);
    case 5: return launch_boxes<5, INDEXED>(ctx, a);
    case 6: return launch_boxes<6, INDEXED>(ctx, a);
    case 7: return launch_boxes<7, INDEXED>(ctx, a);
    case 8: return launch_boxes<8, INDEXED>(ctx, a);
  }
  return bp_fail(ctx, BPOMP_ERR_INVALID, "unsupported dim %d", ctx->dim);
}

int bp_launch_edge_check(bpomp_ctx* ctx, const double* d_from, const double* d_to, const uint32_t* d_from_ids,
                         const uint32_t* d_to_ids, int64_t count, uint8_t* d_valid, int32_t* d_first,
                         uint32_t* d_nstates, bool deferred_only) {
  if (ctx->world == WORLD_NONE) return bp_fail(ctx, BPOMP_ERR_NO_WORLD, "no collision world set");
  if (count <= 0) return BPOMP_OK;
  // Box worlds the fast kernel covers: it evaluates the edges it can and marks the rest BPOMP_EDGE_SLOW; the
  // general kernel below then runs on exactly those (its CTAs exit at once when there are none).
  int fast = 0;
  if (!deferred_only) {
    fast = bp_launch_edge_check_fast(ctx, d_from, d_to, d_from_ids, d_to_ids, count, d_valid, d_first, d_nstates);
    if (fast < 0) return fast;
  }
  EdgeArgs a;
  a.from = d_from; a.to = d_to; a.from_ids = d_from_ids; a.to_ids = d_to_ids;
  a.verts = ctx->d_verts.as<double>();
  a.count = count; a.valid = d_valid; a.first = d_first; a.nstates = d_nstates;
  a.only_status = deferred_only ? (int)BPOMP_EDGE_DEFERRED : (fast ? (int)BPOMP_EDGE_SLOW : 0);
  a.L = ctx->L;
  BP_TRY(bp_reserve(ctx, ctx->d_work, 80 * sizeof(unsigned long long), false));
  const int l = ctx->work_slot;
  unsigned long long* W = ctx->d_work.as<unsigned long long>();
  if (fast) {  // counters of this call's set; the fast kernel has zeroed the tile counter (common.cuh: d_work)
    const int set = ctx->work_epoch[l];
    a.work = W + set * 32 + l;
    a.gate = reinterpret_cast<const int*>(W + set * 32 + 16 + l);
    a.reset0 = W + (1 - set) * 32 + 8 + l;
    a.reset1 = W + (1 - set) * 32 + 16 + l;
    a.list = ctx->d_slow_list[l].as<uint32_t>();
    a.list_cap = BP_SLOW_LIST_CAP;
  } else {
    a.work = W + 64 + l;
    a.gate = nullptr;
    a.reset0 = a.reset1 = nullptr;
    a.list = nullptr;
    a.list_cap = 0;
    BP_CUDA(ctx, cudaMemsetAsync(a.work, 0, sizeof(unsigned long long), ctx->stream));
  }
  const bool indexed = d_from_ids != nullptr;
  if (ctx->world == WORLD_IMAGE) {
    int grid = (int)std::min<int64_t>((count + EC_IMG_THREADS - 1) / EC_IMG_THREADS, (int64_t)ctx->num_sms * 8);
    if (indexed) edge_check_image_kernel<true><<<grid, EC_IMG_THREADS, 0, ctx->stream>>>(a, ctx->image, ctx->perm_view());
    else edge_check_image_kernel<false><<<grid, EC_IMG_THREADS, 0, ctx->stream>>>(a, ctx->image, ctx->perm_view());
  } else {
    BP_TRY(indexed ? launch_boxes_dim<true>(ctx, a) : launch_boxes_dim<false>(ctx, a));
  }
  ctx->launches++;
  BP_CUDA(ctx, cudaGetLastError());
  if (fast) {  // this call's sets are consistent again: the next call on this lane uses the other set
    ctx->work_dirty[l] = false;
    ctx->work_epoch[l] ^= 1;
  }
  return BPOMP_OK;
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" int bpomp_edges_check_dev(bpomp_ctx* ctx, const double* d_from, const double* d_to, int64_t count,
                                     uint8_t* d_valid, int32_t* d_first, uint32_t* d_nstates) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (count < 0 || (count > 0 && (!d_from || !d_to || !d_valid || !d_first)))
    return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  return bp_launch_edge_check(ctx, d_from, d_to, nullptr, nullptr, count, d_valid, d_first, d_nstates, false);
}

extern "C" int bpomp_edges_check_indexed_dev(bpomp_ctx* ctx, const uint32_t* d_from_ids, const uint32_t* d_to_ids,
                                             int64_t count, uint8_t* d_valid, int32_t* d_first,
                                             uint32_t* d_nstates) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (count < 0 || (count > 0 && (!d_from_ids || !d_to_ids || !d_valid || !d_first)))
    return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  return bp_launch_edge_check(ctx, nullptr, nullptr, d_from_ids, d_to_ids, count, d_valid, d_first, d_nstates, false);
}

// After a pass: if some edges met a non-resident sample order, upload the missing rows and
// re-run the kernel on exactly those edges.
// Vertex ids of a host-buffer call are validated on the device, on the staged copies: an id beyond the vertex
// table is replaced by 0 (so that the edge kernel never reads out of bounds) and counted in flags[2]; the call
// then fails with BPOMP_ERR_RANGE.  (A host loop over the ids cost more than the H2D copy of a large batch.)
__global__ void __launch_bounds__(256) ids_validate_kernel(uint32_t* __restrict__ a, uint32_t* __restrict__ b,
                                                           int64_t count, uint32_t nverts, int* __restrict__ flags) {
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    if (a[i] >= nverts) { a[i] = 0; bad = true; }
    if (b[i] >= nverts) { b[i] = 0; bad = true; }
  }
  if (bad) atomicAdd(flags + 2, 1);
}
static int launch_ids_validate(bpomp_ctx* ctx, uint32_t* d_a, uint32_t* d_b, int64_t count, cudaStream_t s) {
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((count + 255) / 256, (int64_t)ctx->num_sms * 8));
  ids_validate_kernel<<<grid, 256, 0, s>>>(d_a, d_b, count, ctx->nverts, ctx->d_flags);
  ctx->launches++;
  BP_CUDA(ctx, cudaGetLastError());
  return BPOMP_OK;
}
static int ids_flag_result(bpomp_ctx* ctx) {  // after bp_read_flags
  if (!ctx->h_flags[2]) return BPOMP_OK;
  ctx->h_flags[2] = 0;
  return bp_fail(ctx, BPOMP_ERR_RANGE, "an edge references a vertex id out of range");
}

static int resolve_deferred(bpomp_ctx* ctx, const double* d_from, const double* d_to, const uint32_t* d_fi,
                            const uint32_t* d_ti, int64_t count, uint8_t* d_valid, int32_t* d_first,
                            uint32_t* d_nstates) {
  BP_TRY(bp_read_flags(ctx));
  if (ctx->h_flags[1]) {
    ctx->h_flags[1] = 0;
    return bp_fail(ctx, BPOMP_ERR_RANGE, "an edge has more than %u states (segment_length too small)",
                   BPOMP_NSTATES_MAX);
  }
  if (!ctx->h_flags[0]) return BPOMP_OK;
  ctx->h_flags[0] = 0;
  std::vector<uint8_t> need((size_t)BPOMP_NSTATES_MAX + 1);
  BP_CUDA(ctx, cudaMemcpyAsync(need.data(), ctx->d_need.p, need.size(), cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  BP_CUDA(ctx, cudaMemsetAsync(ctx->d_need.p, 0, need.size(), ctx->stream));
  std::vector<uint32_t> ns;
  for (uint32_t n = 2; n <= BPOMP_NSTATES_MAX; ++n)
    if (need[n]) ns.push_back(n);
  BP_TRY(bp_perm_ensure_rows(ctx, ns));
  BP_TRY(bp_launch_edge_check(ctx, d_from, d_to, d_fi, d_ti, count, d_valid, d_first, d_nstates, true));
  BP_TRY(bp_read_flags(ctx));
  if (ctx->h_flags[0] || ctx->h_flags[1]) {
    ctx->h_flags[0] = ctx->h_flags[1] = 0;
    return bp_fail(ctx, BPOMP_ERR_RANGE, "internal: deferred edges remain after re-run");
  }
  return BPOMP_OK;
}

int bp_edges_check_resolved(bpomp_ctx* ctx, const double* d_from, const double* d_to, const uint32_t* d_from_ids,
                            const uint32_t* d_to_ids, int64_t count, uint8_t* d_valid, int32_t* d_first,
                            uint32_t* d_nstates) {
  BP_TRY(bp_launch_edge_check(ctx, d_from, d_to, d_from_ids, d_to_ids, count, d_valid, d_first, d_nstates, false));
  return resolve_deferred(ctx, d_from, d_to, d_from_ids, d_to_ids, count, d_valid, d_first, d_nstates);
}

// Large host-buffer batches are pipelined: chunks of EC_PIPE_CHUNK edges alternate over EC_PIPE_LANES
// streams, each doing H2D -> kernel -> D2H on its own staging buffers, so the copies of one chunk overlap
// the kernel and the copies of the others (the call is PCIe-bound: 112 B in + 5..9 B out per 7-DoF edge).
#define EC_PIPE_LANES 3
#define EC_PIPE_CHUNK (1 << 20)

static int edges_check_pipelined(bpomp_ctx* ctx, const double* from, const double* to, const uint32_t* fi,
                                 const uint32_t* ti, int64_t count, uint8_t* valid, int32_t* first,
                                 uint32_t* nstates) {
  const bool indexed = fi != nullptr;
  const size_t in_elem = indexed ? sizeof(uint32_t) : (size_t)ctx->dim * sizeof(double);
  for (int l = 0; l < EC_PIPE_LANES; ++l) {
    if (!ctx->lane_stream[l]) BP_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->lane_stream[l], cudaStreamNonBlocking));
    BP_TRY(bp_reserve(ctx, ctx->lane_a[l], EC_PIPE_CHUNK * in_elem, false));
    BP_TRY(bp_reserve(ctx, ctx->lane_b[l], EC_PIPE_CHUNK * in_elem, false));
    BP_TRY(bp_reserve(ctx, ctx->lane_o[l], (size_t)EC_PIPE_CHUNK * 9, false));
  }
  cudaEvent_t ev;
  BP_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  BP_CUDA(ctx, cudaEventRecord(ev, ctx->stream));  // lanes start after whatever is queued on the ctx stream
  cudaStream_t saved = ctx->stream;
  int rc = BPOMP_OK;
  for (int64_t c0 = 0, c = 0; c0 < count && rc == BPOMP_OK; c0 += EC_PIPE_CHUNK, ++c) {
    const int l = (int)(c % EC_PIPE_LANES);
    const int64_t m = std::min<int64_t>(EC_PIPE_CHUNK, count - c0);
    cudaStream_t s = ctx->lane_stream[l];
    if (c < EC_PIPE_LANES) cudaStreamWaitEvent(s, ev, 0);
    const char* hf = indexed ? (const char*)(fi + c0) : (const char*)(from + c0 * ctx->dim);
    const char* ht = indexed ? (const char*)(ti + c0) : (const char*)(to + c0 * ctx->dim);
    cudaError_t ce = cudaMemcpyAsync(ctx->lane_a[l].p, hf, (size_t)m * in_elem, cudaMemcpyHostToDevice, s);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(ctx->lane_b[l].p, ht, (size_t)m * in_elem, cudaMemcpyHostToDevice, s);
    if (ce != cudaSuccess) {
      rc = bp_fail(ctx, BPOMP_ERR_CUDA, "pipelined edge check: H2D copy failed: %s", cudaGetErrorString(ce));
      break;
    }
    uint8_t* dv = ctx->lane_o[l].as<uint8_t>();
    int32_t* dfirst = reinterpret_cast<int32_t*>(dv + EC_PIPE_CHUNK);
    uint32_t* dn = reinterpret_cast<uint32_t*>(dv + (size_t)EC_PIPE_CHUNK * 5);
    ctx->stream = s;
    ctx->work_slot = 1 + l;
    if (indexed) rc = launch_ids_validate(ctx, ctx->lane_a[l].as<uint32_t>(), ctx->lane_b[l].as<uint32_t>(), m, s);
    if (rc == BPOMP_OK)
      rc = bp_launch_edge_check(ctx, indexed ? nullptr : ctx->lane_a[l].as<double>(),
                              indexed ? nullptr : ctx->lane_b[l].as<double>(),
                              indexed ? ctx->lane_a[l].as<uint32_t>() : nullptr,
                              indexed ? ctx->lane_b[l].as<uint32_t>() : nullptr, m, dv, dfirst, dn, false);
    ctx->stream = saved;
    ctx->work_slot = 0;
    if (rc != BPOMP_OK) break;
    ce = cudaMemcpyAsync(valid + c0, dv, (size_t)m, cudaMemcpyDeviceToHost, s);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(first + c0, dfirst, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, s);
    if (ce == cudaSuccess && nstates)
      ce = cudaMemcpyAsync(nstates + c0, dn, (size_t)m * sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
    if (ce != cudaSuccess) {
      rc = bp_fail(ctx, BPOMP_ERR_CUDA, "pipelined edge check: D2H copy failed: %s", cudaGetErrorString(ce));
      break;
    }
  }
  for (int l = 0; l < EC_PIPE_LANES; ++l) {
    cudaError_t e2 = cudaStreamSynchronize(ctx->lane_stream[l]);
    if (e2 != cudaSuccess && rc == BPOMP_OK)
      rc = bp_fail(ctx, BPOMP_ERR_CUDA, "pipelined edge check failed: %s", cudaGetErrorString(e2));
  }
  cudaEventDestroy(ev);
  BP_TRY(rc);
  BP_TRY(bp_read_flags(ctx));
  BP_TRY(ids_flag_result(ctx));
  if (ctx->h_flags[1]) {
    ctx->h_flags[1] = 0;
    ctx->h_flags[0] = 0;
    return bp_fail(ctx, BPOMP_ERR_RANGE, "an edge has more than %u states (segment_length too small)",
                   BPOMP_NSTATES_MAX);
  }
  return ctx->h_flags[0] ? 1 : BPOMP_OK;  // 1: some sample orders were missing -> caller re-runs unpipelined
}

static int edges_check_host(bpomp_ctx* ctx, const double* from, const double* to, const uint32_t* fi,
                            const uint32_t* ti, int64_t count, uint8_t* valid, int32_t* first, uint32_t* nstates) {
  if (ctx->world == WORLD_NONE) return bp_fail(ctx, BPOMP_ERR_NO_WORLD, "no collision world set");
  if (count == 0) return BPOMP_OK;
  const bool indexed = fi != nullptr;
  if (indexed && ctx->nverts == 0) return bp_fail(ctx, BPOMP_ERR_RANGE, "an edge references a vertex id out of range");
  if (count >= 2 * (int64_t)EC_PIPE_CHUNK) {
    int rc = edges_check_pipelined(ctx, from, to, fi, ti, count, valid, first, nstates);
    if (rc <= 0) return rc;
    // rc == 1: the missing sample orders get uploaded by the plain path below, which redoes the batch
    ctx->h_flags[0] = 0;
  }
  size_t in_bytes = indexed ? (size_t)count * sizeof(uint32_t) : (size_t)count * ctx->dim * sizeof(double);
  BP_TRY(bp_reserve(ctx, ctx->s_a, in_bytes, false));
  BP_TRY(bp_reserve(ctx, ctx->s_b, in_bytes, false));
  BP_TRY(bp_reserve(ctx, ctx->s_c, (size_t)count, false));
  BP_TRY(bp_reserve(ctx, ctx->s_d, (size_t)count * sizeof(int32_t), false));
  BP_TRY(bp_reserve(ctx, ctx->s_e, (size_t)count * sizeof(uint32_t), false));
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->s_a.p, indexed ? (const void*)fi : (const void*)from, in_bytes,
                               cudaMemcpyHostToDevice, ctx->stream));
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->s_b.p, indexed ? (const void*)ti : (const void*)to, in_bytes,
                               cudaMemcpyHostToDevice, ctx->stream));
  const double* df = indexed ? nullptr : ctx->s_a.as<double>();
  const double* dt = indexed ? nullptr : ctx->s_b.as<double>();
  const uint32_t* dfi = indexed ? ctx->s_a.as<uint32_t>() : nullptr;
  const uint32_t* dti = indexed ? ctx->s_b.as<uint32_t>() : nullptr;
  uint8_t* dv = ctx->s_c.as<uint8_t>();
  int32_t* dfirst = ctx->s_d.as<int32_t>();
  uint32_t* dn = ctx->s_e.as<uint32_t>();
  if (indexed) BP_TRY(launch_ids_validate(ctx, ctx->s_a.as<uint32_t>(), ctx->s_b.as<uint32_t>(), count, ctx->stream));
  BP_TRY(bp_launch_edge_check(ctx, df, dt, dfi, dti, count, dv, dfirst, dn, false));
  {
    const int rc = resolve_deferred(ctx, df, dt, dfi, dti, count, dv, dfirst, dn);
    BP_TRY(ids_flag_result(ctx));  // the flags were read by resolve_deferred
    BP_TRY(rc);
  }
  BP_CUDA(ctx, cudaMemcpyAsync(valid, dv, (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaMemcpyAsync(first, dfirst, (size_t)count * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (nstates)
    BP_CUDA(ctx, cudaMemcpyAsync(nstates, dn, (size_t)count * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BPOMP_OK;
}

extern "C" int bpomp_edges_check(bpomp_ctx* ctx, const double* from, const double* to, int64_t count,
                                 uint8_t* valid, int32_t* first_invalid_slot, uint32_t* nstates) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (count < 0 || (count > 0 && (!from || !to || !valid || !first_invalid_slot)))
    return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  return edges_check_host(ctx, from, to, nullptr, nullptr, count, valid, first_invalid_slot, nstates);
}

extern "C" int bpomp_edges_check_indexed(bpomp_ctx* ctx, const uint32_t* from_ids, const uint32_t* to_ids,
                                         int64_t count, uint8_t* valid, int32_t* first_invalid_slot,
                                         uint32_t* nstates) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (count < 0 || (count > 0 && (!from_ids || !to_ids || !valid || !first_invalid_slot)))
    return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  return edges_check_host(ctx, nullptr, nullptr, from_ids, to_ids, count, valid, first_invalid_slot, nstates);
}

// The states of one edge in slot order (debug / belief replay helper).
__global__ void edge_states_kernel(const double* f, const double* t, int D, const double* tf, uint32_t n,
                                   double* out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double tt = tf[i];
    for (int j = 0; j < D; ++j) out[(size_t)i * D + j] = __dadd_rn(f[j], __dmul_rn(__dsub_rn(t[j], f[j]), tt));
  }
}
// ---------------------------------------------------------------------------
// Status packing for the multi-GPU merge: 2 bits per edge, 16 edges per word.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) status_pack_kernel(const uint8_t* __restrict__ valid, int64_t count,
                                                          uint32_t* __restrict__ packed) {
  const int64_t words = (count + 15) / 16;
  for (int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; wi < words; wi += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e0 = wi * 16;
    uint32_t out = 0;
    if (e0 + 16 <= count && (reinterpret_cast<uintptr_t>(valid) & 15) == 0) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(valid + e0));
      const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int b = 0; b < 4; ++b) out |= ((w4[q] >> (8 * b)) & 3u) << (2 * (4 * q + b));
      }
    } else {
      for (int i = 0; i < 16 && e0 + i < count; ++i) out |= ((uint32_t)valid[e0 + i] & 3u) << (2 * i);
    }
    packed[wi] = out;
  }
}

__global__ void __launch_bounds__(256) status_unpack_kernel(const uint32_t* __restrict__ packed, int64_t count,
                                                            uint8_t* __restrict__ valid) {
  const int64_t words = (count + 15) / 16;
  for (int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; wi < words; wi += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e0 = wi * 16;
    const uint32_t in = __ldg(packed + wi);
    if (e0 + 16 <= count && (reinterpret_cast<uintptr_t>(valid) & 15) == 0) {
      uint32_t w4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        w4[q] = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) w4[q] |= ((in >> (2 * (4 * q + b))) & 3u) << (8 * b);
      }
      *reinterpret_cast<uint4*>(valid + e0) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    } else {
      for (int i = 0; i < 16 && e0 + i < count; ++i) valid[e0 + i] = (uint8_t)((in >> (2 * i)) & 3u);
    }
  }
}

static int status_grid(bpomp_ctx* ctx, int64_t count) {
  const int64_t words = (count + 15) / 16;
  return (int)std::max<int64_t>(1, std::min<int64_t>((words + 255) / 256, (int64_t)ctx->num_sms * 8));
}

extern "C" int bpomp_edge_status_pack_dev(bpomp_ctx* ctx, const uint8_t* d_valid, int64_t count, uint32_t* d_packed) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (count < 0 || (count > 0 && (!d_valid || !d_packed))) return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  if (count == 0) return BPOMP_OK;
  status_pack_kernel<<<status_grid(ctx, count), 256, 0, ctx->stream>>>(d_valid, count, d_packed);
  ctx->launches++;
  BP_CUDA(ctx, cudaGetLastError());
  return BPOMP_OK;
}

extern "C" int bpomp_edge_status_unpack_dev(bpomp_ctx* ctx, const uint32_t* d_packed, int64_t count,
                                            uint8_t* d_valid) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (count < 0 || (count > 0 && (!d_valid || !d_packed))) return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  if (count == 0) return BPOMP_OK;
  status_unpack_kernel<<<status_grid(ctx, count), 256, 0, ctx->stream>>>(d_packed, count, d_valid);
  ctx->launches++;
  BP_CUDA(ctx, cudaGetLastError());
  return BPOMP_OK;
}

__global__ void edge_nstates_kernel(const double* f, const double* t, int D, double L, uint32_t* n) {
  double s = 0.0;
  for (int j = 0; j < D; ++j) {
    double diff = __dsub_rn(f[j], t[j]);
    s = __dadd_rn(s, __dmul_rn(diff, diff));
  }
  *n = bp_nstates(__dsqrt_rn(s), L);
}

extern "C" int bpomp_edge_states(bpomp_ctx* ctx, const double* from, const double* to, double* out, uint32_t cap,
                                 uint32_t* nstates) {
  if (!ctx || !from || !to || !nstates) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  const int D = ctx->dim;
  BP_TRY(bp_reserve(ctx, ctx->s_a, 2 * D * sizeof(double) + 16, false));
  double* df = ctx->s_a.as<double>();
  double* dt = df + D;
  uint32_t* dn = reinterpret_cast<uint32_t*>(dt + D);
  BP_CUDA(ctx, cudaMemcpyAsync(df, from, D * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  BP_CUDA(ctx, cudaMemcpyAsync(dt, to, D * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  edge_nstates_kernel<<<1, 1, 0, ctx->stream>>>(df, dt, D, ctx->L, dn);
  ctx->launches++;
  uint32_t n = 0;
  BP_CUDA(ctx, cudaMemcpyAsync(&n, dn, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (n == 0) return bp_fail(ctx, BPOMP_ERR_RANGE, "edge has more than %u states", BPOMP_NSTATES_MAX);
  *nstates = n;
  if (!out || cap == 0) return BPOMP_OK;
  BP_TRY(bp_perm_ensure_rows(ctx, std::vector<uint32_t>{n}));
  uint32_t m = n < cap ? n : cap;
  BP_TRY(bp_reserve(ctx, ctx->s_b, (size_t)m * D * sizeof(double), false));
  edge_states_kernel<<<(m + 127) / 128, 128, 0, ctx->stream>>>(df, dt, D, ctx->d_tfrac.as<double>() + ctx->h_perm_off[n],
                                                               m, ctx->s_b.as<double>());
  ctx->launches++;
  BP_CUDA(ctx, cudaGetLastError());
  BP_CUDA(ctx, cudaMemcpyAsync(out, ctx->s_b.p, (size_t)m * D * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BPOMP_OK;
}
