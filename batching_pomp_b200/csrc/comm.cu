// comm.cu — multi-GPU form of the batch operators: one bpomp_ctx per GPU, NCCL over
// NVLink / NVSwitch for the only exchange the path has, the merge of per-shard results.
//
// The collision world, the vertex table and the belief points are REPLICATED (every rank made the same
// bpomp_world_* / bpomp_vertices_* / bpomp_belief_* calls); the units of a batch (edges, query vertices,
// belief queries) are split into contiguous shards, rank g owning [count*g/G, count*(g+1)/G).  Every rank
// evaluates its shard with the single-GPU kernels and the results are all-gathered device to device, so each
// rank ends with the complete result in the unsharded order.  Ragged shards are gathered as a group of
// ncclBroadcast calls (one per rank, in place at the rank's offset), equal shards with ncclAllGather.
//
// NCCL is loaded at run time (dlopen of libnccl.so.2 — inside a PyTorch process that is the copy PyTorch
// already loaded), so the library has no link-time dependency on it and single-GPU users never touch it.
// The reference has no counterpart (single process, single thread): these entry points replace nothing, they
// are the "bpomp_comm_init / sharded variants" row of the boundary table in SURVEY.md §8(b).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;  // optional
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};
NcclApi g_nccl;

bool nccl_load() {
  if (g_nccl.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) {
    g_nccl.err = std::string("cannot load NCCL: ") + dlerror();
    return false;
  }
  auto sym = [&](const char* s) { return dlsym(g_nccl.lib, s); };
#define BP_NCCL_SYM(field, name)                                                  \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(sym(name));             \
  if (!g_nccl.field) {                                                            \
    g_nccl.err = std::string("NCCL symbol missing: ") + name;                     \
    g_nccl.lib = nullptr;                                                         \
    return false;                                                                 \
  }
  BP_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  BP_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  BP_NCCL_SYM(CommInitAll, "ncclCommInitAll")
  BP_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  BP_NCCL_SYM(AllGather, "ncclAllGather")
  BP_NCCL_SYM(Broadcast, "ncclBroadcast")
  BP_NCCL_SYM(GroupStart, "ncclGroupStart")
  BP_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  BP_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef BP_NCCL_SYM
  g_nccl.CommInitRankConfig = reinterpret_cast<decltype(g_nccl.CommInitRankConfig)>(sym("ncclCommInitRankConfig"));
  return true;
}
}  // namespace

struct bpomp_comm {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  cudaStream_t stream = nullptr;  // the merge of one edge batch runs here, under the next batch's kernel
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
  cudaEvent_t ev_done2[2] = {nullptr, nullptr};  // completion of the merge of edge call k, by parity of k
  const void* last_packed[2] = {nullptr, nullptr};
  uint64_t edge_calls = 0;
  bool pending = false;
  DevBuf tmp_a, tmp_b, tmp_c;
};

#define BP_NCCL(ctx, call)                                                                                      \
  do {                                                                                                          \
    ncclResult_t _r = (call);                                                                                   \
    if (_r != ncclSuccess)                                                                                      \
      return bp_fail((ctx), BPOMP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(_r), __FILE__, \
                     __LINE__);                                                                                 \
  } while (0)

extern "C" int bpomp_shard_range(int64_t count, int nranks, int rank, int64_t* begin, int64_t* end) {
  if (count < 0 || nranks < 1 || rank < 0 || rank >= nranks || !begin || !end) return BPOMP_ERR_INVALID;
  *begin = (int64_t)(((__int128)count * rank) / nranks);
  *end = (int64_t)(((__int128)count * (rank + 1)) / nranks);
  return BPOMP_OK;
}

extern "C" int bpomp_comm_unique_id(void* id, size_t bytes) {
  if (!id || bytes < sizeof(ncclUniqueId)) return BPOMP_ERR_INVALID;
  if (!nccl_load()) return bp_fail(nullptr, BPOMP_ERR_CUDA, "%s", g_nccl.err.c_str());
  ncclUniqueId u;
  if (g_nccl.GetUniqueId(&u) != ncclSuccess) return bp_fail(nullptr, BPOMP_ERR_CUDA, "ncclGetUniqueId failed");
  memcpy(id, &u, sizeof(u));
  return BPOMP_OK;
}

// The communicator's kernels run UNDER an edge kernel that fills all but a few SMs (common.cuh: bp_reserved_sms)
// with one CTA each, so NCCL is asked for no more CTAs than fit there: a collective whose CTAs are not all
// resident would only finish when the edge kernel's CTAs exit, and the merge would no longer be hidden
// (measured at 8 GPUs: 0.64 ms per step with NCCL's default CTA count and 4 free SMs, 0.46 with 8 and 8).
static ncclResult_t comm_init_one(ncclComm_t* comm, int nranks, const ncclUniqueId& u, int rank, int max_ctas) {
  if (g_nccl.CommInitRankConfig && max_ctas > 0) {
    ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
    cfg.minCTAs = 1;
    cfg.maxCTAs = max_ctas < 1 ? 1 : max_ctas;
    return g_nccl.CommInitRankConfig(comm, nranks, u, rank, &cfg);
  }
  return g_nccl.CommInitRank(comm, nranks, u, rank);
}
static int comm_ctas(const bpomp_ctx* ctx, int nranks) {
  if (ctx->nccl_max_ctas >= 0) return ctx->nccl_max_ctas;  // 0: no limit
  return bp_reserved_sms(ctx, nranks);
}

static int comm_attach(bpomp_ctx* ctx, ncclComm_t comm, int nranks, int rank) {
  bpomp_comm* c = new bpomp_comm();
  c->comm = comm;
  c->nranks = nranks;
  c->rank = rank;
  ctx->comm = c;
  ctx->comm_nranks = nranks;
  BP_CUDA(ctx, cudaSetDevice(ctx->device));
  BP_CUDA(ctx, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  BP_CUDA(ctx, cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
  BP_CUDA(ctx, cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
  for (int i = 0; i < 2; ++i) BP_CUDA(ctx, cudaEventCreateWithFlags(&c->ev_done2[i], cudaEventDisableTiming));
  return BPOMP_OK;
}

extern "C" int bpomp_comm_init_rank(bpomp_ctx* ctx, int nranks, int rank, const void* id) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (nranks < 1 || rank < 0 || rank >= nranks || !id) return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  if (ctx->comm) return bp_fail(ctx, BPOMP_ERR_INVALID, "this context already has a communicator");
  if (!nccl_load()) return bp_fail(ctx, BPOMP_ERR_CUDA, "%s", g_nccl.err.c_str());
  BP_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclComm_t comm;
  BP_NCCL(ctx, comm_init_one(&comm, nranks, u, rank, comm_ctas(ctx, nranks)));
  return comm_attach(ctx, comm, nranks, rank);
}

extern "C" int bpomp_comm_init(bpomp_ctx** ctxs, int ngpus) {
  if (!ctxs || ngpus < 1) return BPOMP_ERR_INVALID;
  for (int i = 0; i < ngpus; ++i)
    if (!ctxs[i] || ctxs[i]->comm) return BPOMP_ERR_INVALID;
  if (!nccl_load()) return bp_fail(ctxs[0], BPOMP_ERR_CUDA, "%s", g_nccl.err.c_str());
  std::vector<int> devs(ngpus);
  for (int i = 0; i < ngpus; ++i) devs[i] = ctxs[i]->device;
  std::vector<ncclComm_t> comms(ngpus);
  if (g_nccl.CommInitRankConfig) {  // ncclCommInitAll, spelled out so that the CTA limit applies
    ncclUniqueId u;
    BP_NCCL(ctxs[0], g_nccl.GetUniqueId(&u));
    BP_NCCL(ctxs[0], g_nccl.GroupStart());
    for (int i = 0; i < ngpus; ++i) {
      BP_CUDA(ctxs[0], cudaSetDevice(devs[i]));
      BP_NCCL(ctxs[0], comm_init_one(&comms[i], ngpus, u, i, comm_ctas(ctxs[i], ngpus)));
    }
    BP_NCCL(ctxs[0], g_nccl.GroupEnd());
  } else {
    BP_NCCL(ctxs[0], g_nccl.CommInitAll(comms.data(), ngpus, devs.data()));
  }
  for (int i = 0; i < ngpus; ++i) BP_TRY(comm_attach(ctxs[i], comms[i], ngpus, i));
  return BPOMP_OK;
}

extern "C" int bpomp_comm_destroy(bpomp_ctx* ctx) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  bpomp_comm* c = ctx->comm;
  if (!c) return BPOMP_OK;
  cudaSetDevice(ctx->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  if (c->ev_ready) cudaEventDestroy(c->ev_ready);
  if (c->ev_done) cudaEventDestroy(c->ev_done);
  for (int i = 0; i < 2; ++i)
    if (c->ev_done2[i]) cudaEventDestroy(c->ev_done2[i]);
  if (c->stream) cudaStreamDestroy(c->stream);
  bp_free(c->tmp_a);
  bp_free(c->tmp_b);
  bp_free(c->tmp_c);
  delete c;
  ctx->comm = nullptr;
  ctx->comm_nranks = 0;
  return BPOMP_OK;
}

extern "C" int bpomp_comm_rank(bpomp_ctx* ctx, int* rank, int* nranks) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (rank) *rank = ctx->comm ? ctx->comm->rank : 0;
  if (nranks) *nranks = ctx->comm ? ctx->comm->nranks : 1;
  return BPOMP_OK;
}

// In-place all-gather of ragged shards: buf holds `elem`-byte items, rank g's shard is [bounds[g], bounds[g+1]).
static int allgatherv_inplace(bpomp_ctx* ctx, void* buf, size_t elem, const std::vector<int64_t>& bounds,
                              cudaStream_t stream) {
  bpomp_comm* c = ctx->comm;
  BP_NCCL(ctx, g_nccl.GroupStart());
  for (int g = 0; g < c->nranks; ++g) {
    const size_t bytes = (size_t)(bounds[g + 1] - bounds[g]) * elem;
    if (bytes == 0) continue;
    char* p = static_cast<char*>(buf) + (size_t)bounds[g] * elem;
    ncclResult_t r = g_nccl.Broadcast(p, p, bytes, ncclChar, g, c->comm, stream);
    if (r != ncclSuccess) {
      g_nccl.GroupEnd();
      return bp_fail(ctx, BPOMP_ERR_CUDA, "ncclBroadcast failed: %s", g_nccl.GetErrorString(r));
    }
  }
  BP_NCCL(ctx, g_nccl.GroupEnd());
  return BPOMP_OK;
}

static std::vector<int64_t> shard_bounds(int64_t count, int nranks) {
  std::vector<int64_t> b(nranks + 1);
  for (int g = 0; g <= nranks; ++g) b[g] = (int64_t)(((__int128)count * g) / nranks);
  return b;
}

#define BP_NEED_COMM(ctx)                                                                                \
  if (!(ctx)) return BPOMP_ERR_INVALID;                                                                  \
  bp_enter(ctx);                                                                                         \
  if (!(ctx)->comm) return bp_fail((ctx), BPOMP_ERR_INVALID, "no communicator: call bpomp_comm_init / bpomp_comm_init_rank first")

// ---------------------------------------------------------------------------
// Edges
// ---------------------------------------------------------------------------
extern "C" int bpomp_edges_check_sharded(bpomp_ctx* ctx, const double* from, const double* to, int64_t count,
                                         uint8_t* valid, int32_t* first_invalid_slot, uint32_t* nstates) {
  BP_NEED_COMM(ctx);
  if (count < 0 || (count > 0 && (!from || !to || !valid || !first_invalid_slot)))
    return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  if (count == 0) return BPOMP_OK;
  bpomp_comm* c = ctx->comm;
  const std::vector<int64_t> b = shard_bounds(count, c->nranks);
  const int64_t b0 = b[c->rank], m = b[c->rank + 1] - b0;
  const int D = ctx->dim;
  // complete result buffers on the device; this rank fills its slice, the gathers fill the rest
  BP_TRY(bp_reserve(ctx, c->tmp_a, (size_t)count, false));
  BP_TRY(bp_reserve(ctx, c->tmp_b, (size_t)count * 4, false));
  BP_TRY(bp_reserve(ctx, c->tmp_c, (size_t)count * 4, false));
  uint8_t* dv = c->tmp_a.as<uint8_t>();
  int32_t* df = c->tmp_b.as<int32_t>();
  uint32_t* dn = c->tmp_c.as<uint32_t>();
  if (m > 0) {
    // the shard goes through the host-buffer entry point's machinery: upload, kernel(s), missing sample orders
    BP_TRY(bp_reserve(ctx, ctx->s_a, (size_t)m * D * 8, false));
    BP_TRY(bp_reserve(ctx, ctx->s_b, (size_t)m * D * 8, false));
    BP_CUDA(ctx, cudaMemcpyAsync(ctx->s_a.p, from + (size_t)b0 * D, (size_t)m * D * 8, cudaMemcpyHostToDevice, ctx->stream));
    BP_CUDA(ctx, cudaMemcpyAsync(ctx->s_b.p, to + (size_t)b0 * D, (size_t)m * D * 8, cudaMemcpyHostToDevice, ctx->stream));
    BP_TRY(bp_edges_check_resolved(ctx, ctx->s_a.as<double>(), ctx->s_b.as<double>(), nullptr, nullptr, m, dv + b0, df + b0,
                                   dn + b0));
  }
  BP_TRY(allgatherv_inplace(ctx, dv, 1, b, ctx->stream));
  BP_TRY(allgatherv_inplace(ctx, df, 4, b, ctx->stream));
  if (nstates) BP_TRY(allgatherv_inplace(ctx, dn, 4, b, ctx->stream));
  BP_CUDA(ctx, cudaMemcpyAsync(valid, dv, (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaMemcpyAsync(first_invalid_slot, df, (size_t)count * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (nstates) BP_CUDA(ctx, cudaMemcpyAsync(nstates, dn, (size_t)count * 4, cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BPOMP_OK;
}

// Device-resident shard in, merged status out.  Asynchronous: the kernel runs on the ctx stream, the merge on
// the communicator's own stream, so the merge of batch k overlaps the kernel of batch k+1 (use alternating
// result buffers).  bpomp_comm_wait makes the ctx stream wait for the last merge.
extern "C" int bpomp_edges_check_sharded_dev(bpomp_ctx* ctx, const double* d_from_shard, const double* d_to_shard,
                                             int64_t shard_count, uint8_t* d_valid_shard, int32_t* d_first_shard,
                                             uint32_t* d_nstates_shard, uint32_t* d_packed_shard,
                                             uint32_t* d_packed_all) {
  BP_NEED_COMM(ctx);
  if (shard_count < 0 || (shard_count > 0 && (!d_from_shard || !d_to_shard || !d_valid_shard || !d_first_shard)) ||
      !d_packed_shard || !d_packed_all)
    return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  bpomp_comm* c = ctx->comm;
  BP_TRY(bpomp_edges_check_dev(ctx, d_from_shard, d_to_shard, shard_count, d_valid_shard, d_first_shard, d_nstates_shard));
  // The merge that last read this set of packed buffers must be done before they are written again.  A caller
  // that alternates two buffer sets waits for the merge of call k-2, which ended long ago, while the merge of
  // call k-1 runs under this call's kernel; a caller that passes the same buffers again waits for the previous
  // merge (correct, no overlap).
  const uint64_t k = c->edge_calls;
  if (k >= 1 && (c->last_packed[(k - 1) & 1] == (const void*)d_packed_shard))
    BP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, c->ev_done2[(k - 1) & 1], 0));
  if (k >= 2) BP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, c->ev_done2[k & 1], 0));
  BP_TRY(bpomp_edge_status_pack_dev(ctx, d_valid_shard, shard_count, d_packed_shard));
  BP_CUDA(ctx, cudaEventRecord(c->ev_ready, ctx->stream));
  BP_CUDA(ctx, cudaStreamWaitEvent(c->stream, c->ev_ready, 0));
  const size_t words = (size_t)((shard_count + 15) / 16);
  BP_NCCL(ctx, g_nccl.AllGather(d_packed_shard, d_packed_all, words, ncclUint32, c->comm, c->stream));
  BP_CUDA(ctx, cudaEventRecord(c->ev_done2[k & 1], c->stream));
  BP_CUDA(ctx, cudaEventRecord(c->ev_done, c->stream));
  c->last_packed[k & 1] = d_packed_shard;
  c->edge_calls = k + 1;
  c->pending = true;
  return BPOMP_OK;
}

extern "C" int bpomp_comm_wait(bpomp_ctx* ctx) {
  BP_NEED_COMM(ctx);
  bpomp_comm* c = ctx->comm;
  if (c->pending) {
    BP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, c->ev_done, 0));
    c->pending = false;
  }
  return BPOMP_OK;
}

// ---------------------------------------------------------------------------
// Neighbours: query vertices sharded; the complete CSR ends up in the ctx's neighbour result on every rank.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) comm_row_lens_kernel(const uint64_t* __restrict__ rp, int64_t m,
                                                            uint32_t* __restrict__ lens) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
    lens[i] = (uint32_t)(rp[i + 1] - rp[i]);
}

extern "C" int bpomp_neighbors_radius_sharded(bpomp_ctx* ctx, const uint32_t* query_ids, uint32_t nq, double r,
                                              uint64_t* nnz) {
  BP_NEED_COMM(ctx);
  bpomp_comm* c = ctx->comm;
  if (nq > 0 && !query_ids) return bp_fail(ctx, BPOMP_ERR_INVALID, "query_ids is NULL");
  const std::vector<int64_t> b = shard_bounds(nq, c->nranks);
  const int64_t b0 = b[c->rank], m = b[c->rank + 1] - b0;
  // 1. this rank's rows (the single-GPU operator on the slice of the query list)
  uint64_t local_nnz = 0;
  BP_TRY(bpomp_neighbors_radius(ctx, query_ids + b0, (uint32_t)m, r, &local_nnz));
  if (nq == 0) {
    if (nnz) *nnz = 0;
    return BPOMP_OK;
  }
  // 2. row lengths of every rank -> complete row_ptr
  BP_TRY(bp_reserve(ctx, c->tmp_a, (size_t)nq * sizeof(uint32_t), false));
  uint32_t* lens = c->tmp_a.as<uint32_t>();
  if (m > 0) {
    comm_row_lens_kernel<<<(int)std::min<int64_t>((m + 255) / 256, 1024), 256, 0, ctx->stream>>>(
        ctx->d_nb_rowptr.as<uint64_t>(), m, lens + b0);
    ctx->launches++;
    BP_CUDA(ctx, cudaGetLastError());
  }
  BP_TRY(allgatherv_inplace(ctx, lens, sizeof(uint32_t), b, ctx->stream));
  DevBuf new_rp, new_col, new_dist;
  BP_TRY(bp_reserve(ctx, new_rp, ((size_t)nq + 2) * sizeof(uint64_t), false));
  BP_TRY(bp_exclusive_scan_u32(ctx, lens, nq, new_rp.as<uint64_t>(), ctx->s_scan));
  std::vector<uint64_t> off(c->nranks + 1);
  for (int g = 0; g <= c->nranks; ++g)
    BP_CUDA(ctx, cudaMemcpyAsync(&off[g], new_rp.as<uint64_t>() + b[g], sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const uint64_t total = off[c->nranks];
  if (off[c->rank + 1] - off[c->rank] != local_nnz)
    return bp_fail(ctx, BPOMP_ERR_CUDA, "internal: sharded neighbour offsets are inconsistent");
  // 3. columns and distances: own slice copied into place, the others gathered
  BP_TRY(bp_reserve(ctx, new_col, std::max<size_t>(total, 1) * sizeof(uint32_t), false));
  BP_TRY(bp_reserve(ctx, new_dist, std::max<size_t>(total, 1) * sizeof(double), false));
  if (local_nnz) {
    BP_CUDA(ctx, cudaMemcpyAsync(new_col.as<uint32_t>() + off[c->rank], ctx->d_nb_col.p, local_nnz * sizeof(uint32_t),
                                 cudaMemcpyDeviceToDevice, ctx->stream));
    BP_CUDA(ctx, cudaMemcpyAsync(new_dist.as<double>() + off[c->rank], ctx->d_nb_dist.p, local_nnz * sizeof(double),
                                 cudaMemcpyDeviceToDevice, ctx->stream));
  }
  std::vector<int64_t> ob(off.begin(), off.end());
  BP_TRY(allgatherv_inplace(ctx, new_col.p, sizeof(uint32_t), ob, ctx->stream));
  BP_TRY(allgatherv_inplace(ctx, new_dist.p, sizeof(double), ob, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  // 4. the merged CSR becomes the ctx's neighbour result (bpomp_neighbors_fetch / _device see it)
  bp_free(ctx->d_nb_rowptr);
  bp_free(ctx->d_nb_col);
  bp_free(ctx->d_nb_dist);
  ctx->d_nb_rowptr = new_rp;
  ctx->d_nb_col = new_col;
  ctx->d_nb_dist = new_dist;
  ctx->nb_nq = nq;
  ctx->nb_nnz = total;
  if (nnz) *nnz = total;
  return BPOMP_OK;
}

// ---------------------------------------------------------------------------
// Belief estimates: queries sharded, belief points replicated.
// ---------------------------------------------------------------------------
extern "C" int bpomp_belief_estimate_sharded(bpomp_ctx* ctx, const double* queries, int64_t nq, double* out) {
  BP_NEED_COMM(ctx);
  if (nq < 0 || (nq > 0 && (!queries || !out))) return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  if (nq == 0) return BPOMP_OK;
  bpomp_comm* c = ctx->comm;
  const std::vector<int64_t> b = shard_bounds(nq, c->nranks);
  const int64_t b0 = b[c->rank], m = b[c->rank + 1] - b0;
  BP_TRY(bp_reserve(ctx, c->tmp_a, (size_t)nq * sizeof(double), false));
  double* all = c->tmp_a.as<double>();
  if (m > 0) {
    BP_TRY(bp_reserve(ctx, c->tmp_b, (size_t)m * ctx->dim * sizeof(double), false));
    BP_CUDA(ctx, cudaMemcpyAsync(c->tmp_b.p, queries + (size_t)b0 * ctx->dim, (size_t)m * ctx->dim * sizeof(double),
                                 cudaMemcpyHostToDevice, ctx->stream));
    BP_TRY(bpomp_belief_estimate_dev(ctx, c->tmp_b.as<double>(), m, all + b0));
  }
  BP_TRY(allgatherv_inplace(ctx, all, sizeof(double), b, ctx->stream));
  BP_CUDA(ctx, cudaMemcpyAsync(out, all, (size_t)nq * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BPOMP_OK;
}
