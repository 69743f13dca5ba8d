// ctx.cu — context lifetime, device buffers, sample-order tables, vertex table.
// Reference citations are relative to /root/reference/.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <queue>

#include "common.cuh"
#include "fastpath.cuh"

static std::string g_create_error;

int bp_fail(bpomp_ctx* ctx, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  else g_create_error = buf;
  return code;
}

int bp_reserve(bpomp_ctx* ctx, DevBuf& b, size_t bytes, bool keep) {
  if (bytes <= b.cap) return BPOMP_OK;
  size_t ncap = b.cap ? b.cap : 256;
  while (ncap < bytes) ncap += ncap / 2 + 256;
  ncap = (ncap + 255) & ~(size_t)255;
  void* np = nullptr;
  cudaError_t e = cudaMalloc(&np, ncap);
  if (e != cudaSuccess) {
    // retry with the exact size before giving up
    ncap = (bytes + 255) & ~(size_t)255;
    e = cudaMalloc(&np, ncap);
    if (e != cudaSuccess)
      return bp_fail(ctx, BPOMP_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", ncap, cudaGetErrorString(e));
  }
  if (keep && b.p && b.cap) {
    e = cudaMemcpyAsync(np, b.p, b.cap, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      cudaFree(np);
      return bp_fail(ctx, BPOMP_ERR_CUDA, "device buffer grow copy failed: %s", cudaGetErrorString(e));
    }
  }
  if (b.p) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(b.p);
  }
  b.p = np;
  b.cap = ncap;
  return BPOMP_OK;
}

void bp_free(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

PermView bpomp_ctx::perm_view() const {
  PermView v;
  v.off = d_perm_off.as<uint32_t>();
  v.tfrac = d_tfrac.as<double>();
  v.step = d_step.as<uint32_t>();
  v.need = d_need.as<uint8_t>();
  v.flags = d_flags;
  return v;
}

// ---------------------------------------------------------------------------
// util::BisectPerm::get — include/batching_pomp/util/BisectPerm.hpp:45-89.
// The reference repeatedly picks the first index farthest from every already-picked index
// (virtual picks at -1 and n), O(n^2).  Equivalent gap formulation, O(n log n): a gap (a,b) of
// unpicked indices offers value floor((b-a)/2) first reached at a + floor((b-a)/2); always take
// the largest value, leftmost index on ties.
// ---------------------------------------------------------------------------
void bp_bisect_perm_host(uint32_t n, std::vector<int32_t>& first) {
  first.clear();
  first.reserve(n);
  struct Gap {
    int64_t val, idx, a, b;
    bool operator<(const Gap& o) const { return val != o.val ? val < o.val : idx > o.idx; }
  };
  std::priority_queue<Gap> pq;
  auto push = [&](int64_t a, int64_t b) {
    int64_t v = (b - a) / 2;
    if (v >= 1) pq.push(Gap{v, a + v, a, b});
  };
  push(-1, (int64_t)n);
  while (!pq.empty()) {
    Gap g = pq.top();
    pq.pop();
    first.push_back((int32_t)g.idx);
    push(g.a, g.idx);
    push(g.idx, g.b);
  }
}

extern "C" int bpomp_bisect_perm(uint32_t n, int32_t* out) {
  if (!out && n) return BPOMP_ERR_INVALID;
  std::vector<int32_t> f;
  bp_bisect_perm_host(n, f);
  for (uint32_t i = 0; i < n; ++i) out[i] = f[i];
  return BPOMP_OK;
}

// Uploads the rows for the given nStates values (those not yet resident).
int bp_perm_ensure_rows(bpomp_ctx* ctx, const std::vector<uint32_t>& ns) {
  std::vector<double> add;
  std::vector<std::pair<uint32_t, uint32_t>> newoff;
  std::vector<int32_t> perm;
  size_t base = ctx->tfrac_used;
  for (uint32_t n : ns) {
    if (n < 2 || n > BPOMP_NSTATES_MAX) continue;
    if (ctx->h_perm_off[n] != BP_ABSENT) continue;
    bp_bisect_perm_host(n, perm);
    size_t off = base + add.size();
    if (off + n >= (size_t)BP_ABSENT) return bp_fail(ctx, BPOMP_ERR_RANGE, "sample-order table overflow");
    ctx->h_perm_off[n] = (uint32_t)off;
    newoff.emplace_back(n, (uint32_t)off);
    // BP.cpp:459: 1.0*(1+order[i].first)/(nStates+1)   (int -> double, unsigned -> double)
    for (uint32_t i = 0; i < n; ++i) add.push_back(1.0 * (1 + perm[i]) / (n + 1));
  }
  if (add.empty()) return BPOMP_OK;
  BP_TRY(bp_reserve(ctx, ctx->d_tfrac, (base + add.size()) * sizeof(double), true));
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->d_tfrac.as<double>() + base, add.data(), add.size() * sizeof(double),
                               cudaMemcpyHostToDevice, ctx->stream));
  ctx->tfrac_used = base + add.size();
  // offsets: contiguous runs are uploaded as one copy
  size_t i = 0;
  while (i < newoff.size()) {
    size_t j = i + 1;
    while (j < newoff.size() && newoff[j].first == newoff[j - 1].first + 1) ++j;
    uint32_t n0 = newoff[i].first;
    BP_CUDA(ctx, cudaMemcpyAsync(ctx->d_perm_off.as<uint32_t>() + n0, &ctx->h_perm_off[n0],
                                 (j - i) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    i = j;
  }
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `add` and offsets are host temporaries
  return BPOMP_OK;
}

extern "C" int bpomp_perm_reserve(bpomp_ctx* ctx, uint32_t nmax) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (nmax > BPOMP_NSTATES_MAX) nmax = BPOMP_NSTATES_MAX;
  if (nmax <= ctx->perm_reserved) return BPOMP_OK;
  std::vector<uint32_t> ns;
  for (uint32_t n = 2; n <= nmax; ++n)
    if (ctx->h_perm_off[n] == BP_ABSENT) ns.push_back(n);
  BP_TRY(bp_perm_ensure_rows(ctx, ns));
  ctx->perm_reserved = nmax;
  return BPOMP_OK;
}

// ---------------------------------------------------------------------------
// create / destroy
// ---------------------------------------------------------------------------
extern "C" const char* bpomp_last_error(const bpomp_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

static int create_impl(bpomp_ctx* ctx) {
  BP_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaDeviceProp prop;
  BP_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
  if (prop.major < 10)
    return bp_fail(ctx, BPOMP_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only",
                   ctx->device, prop.major, prop.minor);
  ctx->num_sms = prop.multiProcessorCount;
  ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
  BP_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  BP_CUDA(ctx, cudaHostAlloc((void**)&ctx->h_flags, 4 * sizeof(int), cudaHostAllocDefault));
  memset(ctx->h_flags, 0, 4 * sizeof(int));
  BP_CUDA(ctx, cudaMalloc((void**)&ctx->d_flags, 4 * sizeof(int)));
  BP_CUDA(ctx, cudaMemsetAsync(ctx->d_flags, 0, 4 * sizeof(int), ctx->stream));

  const size_t nent = (size_t)BPOMP_NSTATES_MAX + 1;
  ctx->h_perm_off.assign(nent, BP_ABSENT);
  BP_TRY(bp_reserve(ctx, ctx->d_perm_off, nent * sizeof(uint32_t), false));
  BP_CUDA(ctx, cudaMemsetAsync(ctx->d_perm_off.p, 0xFF, nent * sizeof(uint32_t), ctx->stream));
  BP_TRY(bp_reserve(ctx, ctx->d_need, nent, false));
  BP_CUDA(ctx, cudaMemsetAsync(ctx->d_need.p, 0, nent, ctx->stream));
  // stepSize table of computeAndSetEdgeFreeProbability, BP.cpp:472, with the host's libm
  std::vector<uint32_t> step(nent, 1u);
  for (size_t n = 2; n < nent; ++n) step[n] = static_cast<unsigned int>(n / std::log(n));
  BP_TRY(bp_reserve(ctx, ctx->d_step, nent * sizeof(uint32_t), false));
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->d_step.p, step.data(), nent * sizeof(uint32_t), cudaMemcpyHostToDevice,
                               ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  BP_TRY(bpomp_perm_reserve(ctx, 256));
  // sample coordinates s = 1 + perm_n[slot] as floats for the fast kernel's range search (fastpath.cuh)
  {
    std::vector<float> ps(fp_perm_off(FP_NMAX + 1u), 0.0f);
    std::vector<int32_t> perm;
    for (uint32_t n = 2; n <= FP_NMAX; ++n) {
      bp_bisect_perm_host(n, perm);
      for (uint32_t i = 0; i < n; ++i) ps[fp_perm_off(n) + i] = (float)(1 + perm[i]);
    }
    BP_TRY(bp_reserve(ctx, ctx->d_perm_s, ps.size() * sizeof(float), false));
    BP_CUDA(ctx, cudaMemcpyAsync(ctx->d_perm_s.p, ps.data(), ps.size() * sizeof(float), cudaMemcpyHostToDevice,
                                 ctx->stream));
    BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->fast.perm_s = ctx->d_perm_s.as<float>();
    std::vector<float> p4(FP_PERM4_SMEM_BYTES / 4, -1.0f);
    for (uint32_t n = 3; n <= FP_PERM4_NMAX; ++n) {
      bp_bisect_perm_host(n, perm);
      for (uint32_t i = 0; i + 2 < n; ++i) p4[fp_perm4_off(n) + i] = (float)(1 + perm[1 + i]);
    }
    // thresholds of nStates on the squared distance (fastpath.cuh: fp_nstates_thr), by bisection over the bit
    // patterns of the positive doubles with the host's IEEE sqrt and division
    {
      double thr[FP_NTHR + 2u];
      const double L = ctx->L;
      auto nst = [L](double s) { return std::floor(std::sqrt(s) / L); };
      for (uint32_t k = 0; k < FP_NTHR + 2u; ++k) {
        thr[k] = 0.0;
        if (k < 3) continue;
        double hi = ((double)k + 2.0) * L;
        hi *= hi;
        uint64_t a = 0, b;
        memcpy(&b, &hi, 8);
        if (!(nst(hi) >= (double)k) || !std::isfinite(hi)) {  // degenerate L: never use the table beyond here
          thr[k] = std::numeric_limits<double>::infinity();
          continue;
        }
        while (b - a > 1) {  // invariant: nst(a) < k <= nst(b)
          const uint64_t m = a + (b - a) / 2;
          double sm;
          memcpy(&sm, &m, 8);
          if (nst(sm) >= (double)k) b = m;
          else a = m;
        }
        memcpy(&thr[k], &b, 8);
      }
      memcpy(reinterpret_cast<char*>(p4.data()) + FP_PERM4_BYTES, thr, FP_THR_BYTES);
    }
    BP_TRY(bp_reserve(ctx, ctx->d_perm4, p4.size() * sizeof(float), false));
    BP_CUDA(ctx, cudaMemcpyAsync(ctx->d_perm4.p, p4.data(), p4.size() * sizeof(float), cudaMemcpyHostToDevice,
                                 ctx->stream));
    BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->fast.perm4 = ctx->d_perm4.as<float>();
  }
  return BPOMP_OK;
}

extern "C" int bpomp_create(int device, int dim, double segment_length, bpomp_ctx** out) {
  if (!out) return bp_fail(nullptr, BPOMP_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (dim < 1 || dim > BPOMP_DIM_MAX)
    return bp_fail(nullptr, BPOMP_ERR_INVALID, "dim must be in 1..%d (got %d)", BPOMP_DIM_MAX, dim);
  if (!(segment_length > 0.0) || !std::isfinite(segment_length))
    return bp_fail(nullptr, BPOMP_ERR_INVALID,
                   "segment_length must be finite and > 0 (set up the state space before the planner, "
                   "src/BatchingPOMP.cpp:241)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return bp_fail(nullptr, BPOMP_ERR_CUDA, "no CUDA device available (%s); there is no CPU fallback",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= ndev)
    return bp_fail(nullptr, BPOMP_ERR_INVALID, "device %d out of range (have %d)", device, ndev);
  bpomp_ctx* ctx = new bpomp_ctx();
  ctx->device = device;
  ctx->dim = dim;
  ctx->L = segment_length;
  int r = create_impl(ctx);
  if (r != BPOMP_OK) {
    g_create_error = ctx->err;
    bpomp_destroy(ctx);
    return r;
  }
  *out = ctx;
  return BPOMP_OK;
}

extern "C" void bpomp_destroy(bpomp_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  bpomp_comm_destroy(ctx);
  DevBuf* bufs[] = {&ctx->d_perm_off, &ctx->d_tfrac, &ctx->d_step, &ctx->d_need, &ctx->d_lohi, &ctx->d_rmask,
                    &ctx->d_bits, &ctx->d_work, &ctx->d_tmask, &ctx->d_qbox, &ctx->d_perm_s, &ctx->d_perm4, &ctx->d_verts, &ctx->d_active, &ctx->d_nb_rowptr, &ctx->d_nb_col,
                    &ctx->d_nb_dist, &ctx->d_nb_query, &ctx->d_nb_tmp, &ctx->d_nb_tmp2, &ctx->d_grid_cell_start,
                    &ctx->d_grid_ids, &ctx->d_grid_pts, &ctx->d_grid_keys, &ctx->d_bel_pts, &ctx->d_bel_vals,
                    &ctx->d_bg_start, &ctx->d_bg_cnt, &ctx->d_bg_cell, &ctx->d_bg_pts, &ctx->d_bg_idx, &ctx->d_bg_geom,
                    &ctx->s_a, &ctx->s_b, &ctx->s_c, &ctx->s_d, &ctx->s_e, &ctx->s_f, &ctx->s_g, &ctx->s_h, &ctx->s_i,
                    &ctx->s_j, &ctx->s_scan};
  for (DevBuf* b : bufs) bp_free(*b);
  for (int l = 0; l < 4; ++l) bp_free(ctx->d_slow_list[l]);
  for (int l = 0; l < 3; ++l) {
    bp_free(ctx->lane_a[l]);
    bp_free(ctx->lane_b[l]);
    bp_free(ctx->lane_o[l]);
    if (ctx->lane_stream[l]) cudaStreamDestroy(ctx->lane_stream[l]);
  }
  if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
  if (ctx->d_flags) cudaFree(ctx->d_flags);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

// Waits for the stream and brings the kernels' status flags ([0] deferred edges, [1] range errors)
// to the host; non-zero device flags are cleared.
int bp_read_flags(bpomp_ctx* ctx) {
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->h_flags[0] || ctx->h_flags[1] || ctx->h_flags[2])
    BP_CUDA(ctx, cudaMemsetAsync(ctx->d_flags, 0, 4 * sizeof(int), ctx->stream));
  return BPOMP_OK;
}

extern "C" int bpomp_sync(bpomp_ctx* ctx) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  BP_TRY(bp_read_flags(ctx));
  if (ctx->h_flags[1]) {
    ctx->h_flags[1] = 0;
    return bp_fail(ctx, BPOMP_ERR_RANGE, "an edge has more than %u states (segment_length too small)",
                   BPOMP_NSTATES_MAX);
  }
  if (ctx->h_flags[0]) {
    int n = ctx->h_flags[0];
    ctx->h_flags[0] = 0;
    return bp_fail(ctx, BPOMP_ERR_DEFERRED,
                   "%d edges of a _dev call need sample orders beyond bpomp_perm_reserve(); they were "
                   "marked BPOMP_EDGE_DEFERRED", n);
  }
  return BPOMP_OK;
}

extern "C" int bpomp_set_stream(bpomp_ctx* ctx, void* cuda_stream) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return BPOMP_OK;
}

extern "C" int bpomp_set_option(bpomp_ctx* ctx, const char* name, int64_t value) {
  if (!ctx || !name) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (!strcmp(name, "fast_edge_kernel")) ctx->fast_mode = value ? 1 : 0;
  else if (!strcmp(name, "fast_edge_min_edges")) ctx->fast_min_edges = value < 0 ? 0 : value;
  else if (!strcmp(name, "fast_edge_warps")) ctx->fast_warps = (int)value;
  else if (!strcmp(name, "reserve_sms")) ctx->reserve_sms = (int)value;
  else if (!strcmp(name, "nccl_max_ctas")) ctx->nccl_max_ctas = (int)value;
  else if (!strcmp(name, "fused_pfree")) ctx->fused_pfree = value ? 1 : 0;
  else if (!strcmp(name, "fp32_prefilter")) ctx->prefilter_f32 = value ? 1 : 0;
  else if (!strcmp(name, "belief_grid")) { ctx->belief_grid = value ? 1 : 0; ctx->bg_count = 0; }
  else return bp_fail(ctx, BPOMP_ERR_INVALID, "unknown option '%s'", name);
  return BPOMP_OK;
}

extern "C" int bpomp_get_counter(bpomp_ctx* ctx, const char* name, int64_t* value) {
  if (!ctx || !name || !value) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (!strcmp(name, "slow_edges")) {  // edges the fast kernel left to the general kernel in the last edge call
    *value = 0;
    if (!ctx->last_slow_count) return BPOMP_OK;
    int v = 0;
    BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    BP_CUDA(ctx, cudaMemcpy(&v, ctx->last_slow_count, sizeof(int), cudaMemcpyDeviceToHost));
    *value = v;
    return BPOMP_OK;
  }
  if (!strcmp(name, "fast_world")) {
    *value = ctx->fast.enabled;
    return BPOMP_OK;
  }
  if (!strcmp(name, "belief_grid_builds")) {  // builds of the belief model's grid index so far
    *value = (int64_t)ctx->bg_builds;
    return BPOMP_OK;
  }
  if (!strcmp(name, "belief_grid_points")) {  // belief points covered by the index
    *value = (int64_t)ctx->bg_count;
    return BPOMP_OK;
  }
  return bp_fail(ctx, BPOMP_ERR_INVALID, "unknown counter '%s'", name);
}

extern "C" void* bpomp_get_stream(bpomp_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" uint64_t bpomp_launch_count(const bpomp_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------
// roadmap generation: the two sample sequences of this project, same arithmetic as
// oracle/bpomp_oracle.cpp: orc_halton and batching_pomp_b200/workloads.py: uniform
// ---------------------------------------------------------------------------
__constant__ int c_primes[16] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53};

__global__ void __launch_bounds__(256) roadmap_generate_kernel(int kind, unsigned long long first, uint32_t count,
                                                               int dim, double* __restrict__ out) {
  const uint64_t total = (uint64_t)count * dim;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = t / dim;
    const int j = (int)(t - k * dim);
    double r;
    if (kind == BPOMP_ROADMAP_HALTON) {
      unsigned long long idx = first + k;
      const unsigned long long p = (unsigned long long)c_primes[j];
      const double base = (double)c_primes[j];
      double f = 1.0;
      r = 0.0;
      while (idx > 0) {
        f = __ddiv_rn(f, base);
        r = __dadd_rn(r, __dmul_rn(f, (double)(idx % p)));
        idx /= p;
      }
    } else {
      unsigned long long z = first + (t + 1ull) * 0x9E3779B97F4A7C15ull;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z = z ^ (z >> 31);
      r = __dmul_rn((double)(z >> 11), 1.0 / 9007199254740992.0);  // 53 bits: exact
    }
    out[t] = r;
  }
}

extern "C" int bpomp_roadmap_generate(bpomp_ctx* ctx, int kind, uint64_t first, uint32_t count, double* host_out) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (kind != BPOMP_ROADMAP_HALTON && kind != BPOMP_ROADMAP_UNIFORM) return bp_fail(ctx, BPOMP_ERR_INVALID, "unknown roadmap kind %d", kind);
  if (count == 0) return BPOMP_OK;
  if (!host_out) return bp_fail(ctx, BPOMP_ERR_INVALID, "host_out is NULL");
  if (kind == BPOMP_ROADMAP_HALTON && ctx->dim > 16) return bp_fail(ctx, BPOMP_ERR_RANGE, "Halton roadmaps are defined for dim <= 16");
  const size_t bytes = (size_t)count * ctx->dim * sizeof(double);
  BP_TRY(bp_reserve(ctx, ctx->s_a, bytes, false));
  const int grid = (int)std::min<uint64_t>(((uint64_t)count * ctx->dim + 255) / 256, (uint64_t)ctx->num_sms * 16);
  roadmap_generate_kernel<<<grid, 256, 0, ctx->stream>>>(kind, first, count, ctx->dim, ctx->s_a.as<double>());
  ctx->launches++;
  BP_CUDA(ctx, cudaGetLastError());
  BP_CUDA(ctx, cudaMemcpyAsync(host_out, ctx->s_a.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BPOMP_OK;
}

// ---------------------------------------------------------------------------
// vertex table
// ---------------------------------------------------------------------------
extern "C" int bpomp_vertices_append(bpomp_ctx* ctx, const double* states, uint32_t count, uint32_t* first_id) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (first_id) *first_id = ctx->nverts;
  if (count == 0) return BPOMP_OK;
  if (!states) return bp_fail(ctx, BPOMP_ERR_INVALID, "states is NULL");
  if ((uint64_t)ctx->nverts + count >= 0xFFFFFFFFull) return bp_fail(ctx, BPOMP_ERR_RANGE, "too many vertices");
  size_t row = (size_t)ctx->dim * sizeof(double);
  size_t n0 = ctx->nverts, n1 = n0 + count;
  BP_TRY(bp_reserve(ctx, ctx->d_verts, n1 * row, true));
  BP_TRY(bp_reserve(ctx, ctx->d_active, n1, true));
  BP_CUDA(ctx, cudaMemcpyAsync((char*)ctx->d_verts.p + n0 * row, states, (size_t)count * row,
                               cudaMemcpyHostToDevice, ctx->stream));
  BP_CUDA(ctx, cudaMemsetAsync((char*)ctx->d_active.p + n0, 1, count, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->h_active.resize(n1, 1);
  ctx->nverts = (uint32_t)n1;
  ctx->nactive += count;
  ctx->grid_valid = false;
  return BPOMP_OK;
}

extern "C" int bpomp_vertices_set_active(bpomp_ctx* ctx, const uint32_t* ids, uint32_t count, int flag) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (count == 0) return BPOMP_OK;
  if (!ids) return bp_fail(ctx, BPOMP_ERR_INVALID, "ids is NULL");
  for (uint32_t i = 0; i < count; ++i)
    if (ids[i] >= ctx->nverts) return bp_fail(ctx, BPOMP_ERR_RANGE, "vertex id %u out of range", ids[i]);
  for (uint32_t i = 0; i < count; ++i) {
    const uint8_t nv = flag ? 1 : 0;
    if (ctx->h_active[ids[i]] != nv) {
      ctx->h_active[ids[i]] = nv;
      if (nv) ++ctx->nactive; else --ctx->nactive;
    }
  }
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->d_active.p, ctx->h_active.data(), ctx->nverts, cudaMemcpyHostToDevice,
                               ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->grid_valid = false;
  return BPOMP_OK;
}

extern "C" int bpomp_vertices_count(bpomp_ctx* ctx, uint32_t* count) {
  if (!ctx || !count) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  *count = ctx->nverts;
  return BPOMP_OK;
}

extern "C" int bpomp_vertices_clear(bpomp_ctx* ctx) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  ctx->nverts = 0;
  ctx->nactive = 0;
  ctx->h_active.clear();
  ctx->grid_valid = false;
  return BPOMP_OK;
}
