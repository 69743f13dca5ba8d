/* bpomp_cuda.h — C ABI of the B200 (sm_100a) hot path of batching_pomp.
 *
 * This is the inner boundary of SURVEY.md §8(b): the entry points a maintainer of
 * personalrobotics/batching_pomp would bind to replace the CPU work done at
 *   - src/BatchingPOMP.cpp:150            (NearestNeighborsGNAT::nearestR in examine_vertex)
 *   - src/BatchingPOMP.cpp:435-461        (initializeEdgePoints: never materialised here)
 *   - src/BatchingPOMP.cpp:510-538        (the isValid loop of checkAndSetEdgeBlocked)
 *   - src/BatchingPOMP.cpp:464-493 + include/batching_pomp/cspacebelief/KNNModel.hpp:89-134
 *                                         (computeAndSetEdgeFreeProbability / KNNModel)
 *   - src/BatchingPOMP.cpp:636-656        (isVertexInadmissible / vertexPruneFunction)
 * The reference has no FFI layer of its own; INTEGRATION.md shows the binding code.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary;
 *   - one opaque bpomp_ctx per device; a ctx is NOT thread-safe (the reference planner is
 *     single-threaded, one planner instance per query);
 *   - every function returns 0 (BPOMP_OK) or a negative error code; the message is
 *     available from bpomp_last_error(ctx) (bpomp_create: bpomp_last_error(NULL));
 *   - functions without a _dev suffix take HOST buffers, copy in/out and return when the
 *     results are in the caller's buffers;
 *   - functions with a _dev suffix take DEVICE buffers (e.g. torch tensor data_ptr()),
 *     enqueue on the ctx stream and return immediately; call bpomp_sync to wait;
 *   - states are row-major double[count][dim] exactly as OMPL RealVectorStateSpace lays a
 *     state out (double values[dim]);
 *   - there is NO CPU fallback: without a CUDA device bpomp_create fails.
 *
 * All floating-point results are bit-identical to the reference's unfused x86-64 FP64
 * arithmetic (OMPL RealVectorStateSpace::distance / ::interpolate); kernels are compiled
 * with --fmad=false and use IEEE sqrt/div.
 */
#ifndef BPOMP_CUDA_H_
#define BPOMP_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BPOMP_API __attribute__((visibility("default")))
#else
#define BPOMP_API
#endif

typedef struct bpomp_ctx bpomp_ctx;

enum {
  BPOMP_OK = 0,
  BPOMP_ERR_INVALID = -1,  /* bad argument */
  BPOMP_ERR_CUDA = -2,     /* CUDA runtime error (message has the cudaError string) */
  BPOMP_ERR_NO_WORLD = -3, /* a collision query was made before a world was set */
  BPOMP_ERR_RANGE = -4,    /* nStates of an edge exceeds BPOMP_NSTATES_MAX, or an id is out of range */
  BPOMP_ERR_NOMEM = -5,
  BPOMP_ERR_DEFERRED = -6  /* a _dev edge call met an nStates whose sample order was not reserved */
};

#define BPOMP_NSTATES_MAX 65535u
#define BPOMP_DIM_MAX 8

/* Edge status values written to `valid` */
#define BPOMP_EDGE_BLOCKED 0u
#define BPOMP_EDGE_FREE 1u
#define BPOMP_EDGE_DEFERRED 2u /* _dev calls only; see bpomp_perm_reserve */

/* ---- context ---------------------------------------------------------------------------- */

/* Replaces the planner constructor's capture of the check resolution,
 * src/BatchingPOMP.cpp:241,278 (mCheckRadius = 0.5*getLongestValidSegmentLength()).
 * segment_length = getLongestValidSegmentLength() as computed by the host (must be > 0). */
BPOMP_API int bpomp_create(int device, int dim, double segment_length, bpomp_ctx** out);
BPOMP_API void bpomp_destroy(bpomp_ctx* ctx);
BPOMP_API const char* bpomp_last_error(const bpomp_ctx* ctx);
BPOMP_API int bpomp_sync(bpomp_ctx* ctx);
/* Use an external cudaStream_t (passed as void*) for all work of this ctx; NULL restores the
 * ctx's own stream. */
BPOMP_API int bpomp_set_stream(bpomp_ctx* ctx, void* cuda_stream);
BPOMP_API void* bpomp_get_stream(bpomp_ctx* ctx);
/* Tuning / testing switches; results never depend on them.  name: "fast_edge_kernel" (1 = use the fast
 * box-world edge kernel where it applies (default), 0 = general kernel only), "fast_edge_min_edges" (batches
 * below this go to the general kernel alone, default 4096), "fast_edge_warps" (cap on its warps per CTA, 0 = auto), "fused_pfree" (1 = small free-probability batches run
 * as one launch (default), 0 = always the tiled multi-launch path). */
BPOMP_API int bpomp_set_option(bpomp_ctx* ctx, const char* name, int64_t value);
/* Diagnostics: "slow_edges" = edges of the last edge call that the fast kernel left to the general one (waits
 * for the stream); "fast_world" = 1 if the current world has the fast kernel's tables. */
BPOMP_API int bpomp_get_counter(bpomp_ctx* ctx, const char* name, int64_t* value);
/* Number of kernel launches issued by this ctx so far (bench.py's gpu_launches). */
BPOMP_API uint64_t bpomp_launch_count(const bpomp_ctx* ctx);

/* Makes the per-nStates sample orders (util/BisectPerm.hpp:45-89 and the fractions of
 * src/BatchingPOMP.cpp:459) resident for every nStates <= nmax.  Host-buffer edge calls extend
 * the table on demand; _dev calls need it reserved up front.  Default at create: 256. */
BPOMP_API int bpomp_perm_reserve(bpomp_ctx* ctx, uint32_t nmax);
/* First components of BisectPerm::get(n) (host-side helper; out has n entries). */
BPOMP_API int bpomp_bisect_perm(uint32_t n, int32_t* out);

/* ---- collision world (replaces the user's ompl::base::StateValidityChecker) -------------- */

/* 2-D occupancy image on the unit square, (0,0) = upper-left, (1,1) = lower-right
 * (README.md:56): col = min(W-1, floor(x0*W)), row = min(H-1, floor(x1*H));
 * a state is valid iff pix[row*W+col] >= threshold and 0 <= x0,x1 <= 1.  dim must be 2. */
BPOMP_API int bpomp_world_set_image(bpomp_ctx* ctx, const uint8_t* pix, int width, int height, int threshold);
/* n-D axis-aligned boxes lo/hi[nboxes][dim]; a state is invalid iff it lies in the closed
 * box lo <= x <= hi of any box. */
BPOMP_API int bpomp_world_set_boxes(bpomp_ctx* ctx, const double* lo, const double* hi, int nboxes);

/* Optional: the bounds of the state space (RealVectorStateSpace::getBounds(), low/high[dim]).  A hint only —
 * results never depend on it: the broad-phase grid of the box world is laid over the part of the obstacles'
 * bounding region that lies inside the bounds, which is where the planner's edges are (finer cells, fewer
 * candidate boxes per edge).  NULL, NULL removes the hint. */
BPOMP_API int bpomp_world_set_bounds(bpomp_ctx* ctx, const double* lo, const double* hi);

/* isValid for a batch of states — src/BatchingPOMP.cpp:654, 777, 784.  out[i] = 1 valid / 0 not. */
BPOMP_API int bpomp_states_valid(bpomp_ctx* ctx, const double* states, int64_t count, uint8_t* out);
BPOMP_API int bpomp_states_valid_dev(bpomp_ctx* ctx, const double* d_states, int64_t count, uint8_t* d_out);

/* vertexPruneFunction / isVertexInadmissible — src/BatchingPOMP.cpp:636-656.
 * out[i] = 1 iff (best_cost < DBL_MAX && d(start,s_i)+d(goal,s_i) > best_cost)
 *               || (check_validity && !isValid(s_i)). */
BPOMP_API int bpomp_states_prune(bpomp_ctx* ctx, const double* states, int64_t count, const double* start,
                                 const double* goal, double best_cost, int check_validity, uint8_t* out);

/* The same predicate over the resident vertex table (BatchingManager::pruneVertices,
 * BatchingManager.hpp:144-165, walks every vertex of the current roadmap): out[v] for v in
 * 0..count-1, start/goal given as vertex ids. */
BPOMP_API int bpomp_vertices_prune(bpomp_ctx* ctx, uint32_t start_id, uint32_t goal_id, double best_cost,
                                   int check_validity, uint8_t* out);

/* ---- roadmap generation ------------------------------------------------------
 * The roadmap states a GraphML file would hold (util/RoadmapFromFile.hpp:136-147 reads them as text), generated on
 * the device and copied to host_out ([count][dim] doubles), for roadmaps that follow this project's two sample
 * sequences bit for bit:
 *   BPOMP_ROADMAP_HALTON  point i = first + k (1-based index `first`), coordinate j = radical inverse of i in
 *                         the j-th prime, evaluated as  f = f / base;  r = r + f * digit  in FP64;
 *   BPOMP_ROADMAP_UNIFORM coordinate j of point k = (splitmix64(seed = `first`, k*dim + j + 1) >> 11) * 2^-53.
 * Nothing is added to the vertex table (the planner admits roadmap vertices batch by batch). */
#define BPOMP_ROADMAP_HALTON 0
#define BPOMP_ROADMAP_UNIFORM 1
BPOMP_API int bpomp_roadmap_generate(bpomp_ctx* ctx, int kind, uint64_t first, uint32_t count, double* host_out);

/* ---- vertex table = the elements of mVertexNN ------------------------------------------- */

/* mVertexNN->add — src/BatchingPOMP.cpp:782,789; VertexBatching.hpp:133; EdgeBatching.hpp:146;
 * HybridBatching.hpp:184.  Ids are consecutive (the g vertex descriptors: start = 0, goal = 1);
 * appended vertices are active. */
BPOMP_API int bpomp_vertices_append(bpomp_ctx* ctx, const double* states, uint32_t count, uint32_t* first_id);
/* mVertexNN->remove — BatchingManager.hpp:162 (flag = 0); flag = 1 re-activates. */
BPOMP_API int bpomp_vertices_set_active(bpomp_ctx* ctx, const uint32_t* ids, uint32_t count, int flag);
BPOMP_API int bpomp_vertices_count(bpomp_ctx* ctx, uint32_t* count);
BPOMP_API int bpomp_vertices_clear(bpomp_ctx* ctx);

/* nearestR — src/BatchingPOMP.cpp:150.  For each query vertex id: every ACTIVE vertex v with
 * distance(q,v) <= r (inclusive), ascending by (distance, id); the query itself is included
 * when active (the caller skips it, :158).  Distances are bit-equal to
 * RealVectorStateSpace::distance.  The CSR result stays on the device until fetched.
 * r = DBL_MAX is allowed (vertex batching, VertexBatching.hpp:71). */
BPOMP_API int bpomp_neighbors_radius(bpomp_ctx* ctx, const uint32_t* query_ids, uint32_t nq, double r,
                                     uint64_t* nnz);
/* row_ptr[nq+1], col[nnz], dist[nnz] (dist may be NULL). */
BPOMP_API int bpomp_neighbors_fetch(bpomp_ctx* ctx, uint64_t* row_ptr, uint32_t* col, double* dist);
/* Device pointers of the last CSR result (valid until the next neighbour call). */
BPOMP_API int bpomp_neighbors_device(bpomp_ctx* ctx, const uint64_t** d_row_ptr, const uint32_t** d_col,
                                     const double** d_dist);
/* Whole-roadmap densification: rows for ALL vertices 0..count-1 (inactive ones get empty rows). */
BPOMP_API int bpomp_neighbors_all(bpomp_ctx* ctx, double r, uint64_t* nnz);

/* ---- edge evaluation --------------------------------------------------------------------- */

/* The isValid loop of checkAndSetEdgeBlocked — src/BatchingPOMP.cpp:496-544, on edges discretised
 * as initializeEdgePoints does (:435-461): n = max(2, floor(distance(from,to)/segment_length)),
 * slot i holds interpolate(from, to, (1+perm_n[i])/(n+1)); slots 1..n-2 are tested in order.
 *   valid[e]              1 = FREE, 0 = BLOCKED
 *   first_invalid_slot[e] slot (1..n-2) of the first invalid state, -1 if FREE.  The reference
 *                         makes `slot` isValid calls for a BLOCKED edge and max(0,n-2) for a
 *                         FREE one, and appends those states to the belief model (value 0.0,
 *                         the last one 1.0 if BLOCKED) — the host replays that from the slot.
 *   nstates[e]            n (may be NULL)
 * Orientation matters: from = source(e), to = target(e) of the edge as created (:162-165). */
BPOMP_API int bpomp_edges_check(bpomp_ctx* ctx, const double* from, const double* to, int64_t count,
                                uint8_t* valid, int32_t* first_invalid_slot, uint32_t* nstates);
/* Same, endpoints given as ids into the vertex table. */
BPOMP_API int bpomp_edges_check_indexed(bpomp_ctx* ctx, const uint32_t* from_ids, const uint32_t* to_ids,
                                        int64_t count, uint8_t* valid, int32_t* first_invalid_slot,
                                        uint32_t* nstates);
BPOMP_API int bpomp_edges_check_dev(bpomp_ctx* ctx, const double* d_from, const double* d_to, int64_t count,
                                    uint8_t* d_valid, int32_t* d_first_invalid_slot, uint32_t* d_nstates);
BPOMP_API int bpomp_edges_check_indexed_dev(bpomp_ctx* ctx, const uint32_t* d_from_ids, const uint32_t* d_to_ids,
                                            int64_t count, uint8_t* d_valid, int32_t* d_first_invalid_slot,
                                            uint32_t* d_nstates);
/* Multi-GPU merge helpers (no counterpart in the reference, which is single-process: BASELINE north_star,
 * "NCCL all-gather ... only for merging edge-status").  The status bytes of bpomp_edges_check*_dev packed
 * to 2 bits per edge, 16 edges per word (edge e -> word e/16, bits 2*(e%16)..+1, unused bits 0), so the
 * all-gather moves a quarter of the bytes; unpack restores the byte flags.  d_packed holds (count+15)/16
 * words.  Asynchronous on the ctx stream. */
BPOMP_API int bpomp_edge_status_pack_dev(bpomp_ctx* ctx, const uint8_t* d_valid, int64_t count, uint32_t* d_packed);
BPOMP_API int bpomp_edge_status_unpack_dev(bpomp_ctx* ctx, const uint32_t* d_packed, int64_t count,
                                           uint8_t* d_valid);
/* The states of one edge in slot order (debug / belief replay helper): out[n][dim]. */
BPOMP_API int bpomp_edge_states(bpomp_ctx* ctx, const double* from, const double* to, double* out,
                                uint32_t cap, uint32_t* nstates);

/* ---- C-space belief (cspacebelief::KNNModel) --------------------------------------------- */

/* KNNModel ctor — KNNModel.hpp:56-67; defaults 15, 0.5, 0.25 (src/BatchingPOMP.cpp:300). */
BPOMP_API int bpomp_belief_config(bpomp_ctx* ctx, int k, double prior, double prior_weight);
/* addPoint — KNNModel.hpp:89-93.  Append-only; the insertion index is the tie-break key. */
BPOMP_API int bpomp_belief_append(bpomp_ctx* ctx, const double* states, const double* values, int64_t count);
/* Replays the belief inserts of checkAndSetEdgeBlocked (:523-537) for already-checked edges
 * entirely on the device: for each edge, slots 1..(first_invalid_slot or n-2) are appended in
 * order with value 0.0, the first invalid one with 1.0.  Edges are processed in array order. */
BPOMP_API int bpomp_belief_append_checked(bpomp_ctx* ctx, const double* from, const double* to,
                                          const int32_t* first_invalid_slot, int64_t count);
/* The same for edges given as vertex ids, with the nStates bpomp_edges_check* reported for them (the caller has
 * both from the check it just made): no endpoint states travel back to the device. */
BPOMP_API int bpomp_belief_append_checked_indexed(bpomp_ctx* ctx, const uint32_t* from_ids, const uint32_t* to_ids,
                                                  const int32_t* first_invalid_slot, const uint32_t* nstates,
                                                  int64_t count);
BPOMP_API int bpomp_belief_size(bpomp_ctx* ctx, uint64_t* count);
BPOMP_API int bpomp_belief_clear(bpomp_ctx* ctx);
/* estimate — KNNModel.hpp:101-134 with beliefDistanceFunction (src/BatchingPOMP.cpp:213-219). */
BPOMP_API int bpomp_belief_estimate(bpomp_ctx* ctx, const double* queries, int64_t nq, double* out);
BPOMP_API int bpomp_belief_estimate_dev(bpomp_ctx* ctx, const double* d_queries, int64_t nq, double* d_out);
/* computeAndSetEdgeFreeProbability — src/BatchingPOMP.cpp:464-493: pfree[e] = prod over slots
 * i = 1, 1+step, ... < n-1 of (1 - estimate(state_i)), step = (unsigned)(n/log(n));
 * nlookups[e] = number of estimate calls (mNumLookups, :480).  nlookups may be NULL. */
BPOMP_API int bpomp_belief_edge_pfree(bpomp_ctx* ctx, const double* from, const double* to, int64_t count,
                                      double* pfree, uint32_t* nlookups);
BPOMP_API int bpomp_belief_edge_pfree_indexed(bpomp_ctx* ctx, const uint32_t* from_ids, const uint32_t* to_ids,
                                              int64_t count, double* pfree, uint32_t* nlookups);

/* ---- multi-GPU (no counterpart in the reference, which is one thread in one process) ------
 * One bpomp_ctx per GPU.  The collision world, the vertex table and the belief points are REPLICATED: every
 * rank makes the same bpomp_world_* / bpomp_vertices_* / bpomp_belief_* calls.  The units of a batch (edges,
 * query vertices, belief queries) are sharded, rank g owning [count*g/G, count*(g+1)/G) (bpomp_shard_range);
 * every rank evaluates its shard and the results are all-gathered device-to-device over NCCL (NVLink /
 * NVSwitch), so each rank returns the complete result in the unsharded order.  The *_sharded calls are
 * collective: every rank calls them with the same arguments.  NCCL is loaded at run time (libnccl.so.2). */
BPOMP_API int bpomp_shard_range(int64_t count, int nranks, int rank, int64_t* begin, int64_t* end);
/* One process (or thread) per GPU: rank 0 makes the id (128 bytes), every rank passes it to init_rank. */
BPOMP_API int bpomp_comm_unique_id(void* id, size_t bytes);
BPOMP_API int bpomp_comm_init_rank(bpomp_ctx* ctx, int nranks, int rank, const void* id);
/* One process driving ngpus contexts (ctxs[i] becomes rank i).  Collective calls on them must then be made
 * from one thread per context. */
BPOMP_API int bpomp_comm_init(bpomp_ctx** ctxs, int ngpus);
BPOMP_API int bpomp_comm_destroy(bpomp_ctx* ctx);
BPOMP_API int bpomp_comm_rank(bpomp_ctx* ctx, int* rank, int* nranks);
/* bpomp_edges_check over all ranks: the same host arrays on every rank, results complete on every rank. */
BPOMP_API int bpomp_edges_check_sharded(bpomp_ctx* ctx, const double* from, const double* to, int64_t count,
                                        uint8_t* valid, int32_t* first_invalid_slot, uint32_t* nstates);
/* Device-resident form for equal shards: this rank's shard in, its results out, and the status flags of ALL
 * ranks merged in packed form (bpomp_edge_status_pack_dev layout per shard, shards in rank order:
 * d_packed_all holds nranks * ((shard_count+15)/16) words).  First-invalid slots stay with the shard owner
 * (it replays the belief inserts).  Asynchronous: the merge runs on the communicator's own stream and overlaps
 * the next batch's kernel when consecutive calls alternate between two sets of d_packed_shard / d_packed_all
 * buffers (a call that reuses the previous call's buffers first waits for that merge); bpomp_comm_wait makes the
 * ctx stream wait for the last merge. */
BPOMP_API int bpomp_edges_check_sharded_dev(bpomp_ctx* ctx, const double* d_from_shard, const double* d_to_shard,
                                            int64_t shard_count, uint8_t* d_valid_shard, int32_t* d_first_shard,
                                            uint32_t* d_nstates_shard, uint32_t* d_packed_shard,
                                            uint32_t* d_packed_all);
BPOMP_API int bpomp_comm_wait(bpomp_ctx* ctx);
/* bpomp_neighbors_radius with the query vertices sharded; afterwards the complete CSR is this ctx's
 * neighbour result (bpomp_neighbors_fetch / bpomp_neighbors_device). */
BPOMP_API int bpomp_neighbors_radius_sharded(bpomp_ctx* ctx, const uint32_t* query_ids, uint32_t nq, double r,
                                             uint64_t* nnz);
/* bpomp_belief_estimate with the queries sharded. */
BPOMP_API int bpomp_belief_estimate_sharded(bpomp_ctx* ctx, const double* queries, int64_t nq, double* out);

#ifdef __cplusplus
}
#endif
#endif /* BPOMP_CUDA_H_ */
