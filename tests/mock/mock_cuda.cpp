// mock_cuda.cpp — TEST DOUBLE of the C ABI in include/bpomp_cuda.h.  TEST INFRASTRUCTURE ONLY.
//
// Purpose: run the C++ host planner (batching_pomp_b200/host) in the CPU test-suite, where there is no
// GPU, so that its own logic — batching managers, lazy A*, selectors, the in-order commit of batched
// results, counters — is checked against the literal planner restatement of the oracle on every
// `pytest -m "not gpu"` run.  It is NOT a CPU fallback of the product: nothing under batching_pomp_b200/
// references it, it is built by tests/test_host_mock_cpu.py into a temporary directory next to a COPY of
// libbpomp_host.so (whose RUNPATH is $ORIGIN), and it is loaded only in the subprocess that test starts.
// Every operator is answered by the CPU oracle (oracle/bpomp_oracle.cpp, compiled into this library),
// so a planner run through this file says nothing about the CUDA kernels — those are checked by the
// `-m gpu` parity tests through the real library.
//
// Only the entry points libbpomp_host.so imports are provided.
#include <cfloat>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "bpomp_cuda.h"

extern "C" {
// oracle/bpomp_oracle.cpp
void orc_states_valid_boxes(const double* states, int64_t count, int d, const double* lo, const double* hi, int nboxes,
                            uint8_t* out);
void orc_states_valid_image(const double* states, int64_t count, const uint8_t* pix, int W, int H, int threshold,
                            uint8_t* out);
void orc_edges_check_boxes(const double* from, const double* to, int64_t count, int d, double L, const double* lo,
                           const double* hi, int nboxes, uint8_t* valid, int32_t* first, uint32_t* nstates, int nthreads);
void orc_edges_check_image(const double* from, const double* to, int64_t count, double L, const uint8_t* pix, int W,
                           int H, int threshold, uint8_t* valid, int32_t* first, uint32_t* nstates, int nthreads);
void orc_neighbors_radius(const double* verts, const uint8_t* active, int64_t N, int d, const uint32_t* query_ids,
                          int64_t nq, double r, uint64_t* row_ptr, uint32_t* col, double* dist, int nthreads);
void orc_vertices_inadmissible(const double* states, int64_t count, int d, const double* start, const double* goal,
                               double best_cost, uint8_t* out);
void orc_belief_estimate(const double* pts, const double* vals, int64_t M, int d, const double* queries, int64_t nq,
                         int K, double prior, double prior_weight, double* out, int nthreads);
void orc_belief_edge_pfree(const double* from, const double* to, int64_t count, int d, double L, const double* pts,
                           const double* vals, int64_t M, int K, double prior, double prior_weight, double* pfree,
                           uint32_t* nlookups, int nthreads);
unsigned orc_edge_states(const double* from, const double* to, int d, double L, double* states, unsigned cap);
int orc_max_threads();
}

struct bpomp_ctx {
  int dim = 0;
  double L = 0.0;
  int world = 0;  // 0 none, 1 boxes, 2 image
  std::vector<double> lo, hi;
  std::vector<uint8_t> pix;
  int W = 0, H = 0, thr = 128;
  std::vector<double> verts;
  std::vector<uint8_t> active;
  std::vector<double> bel_pts, bel_vals;
  int knn = 15;
  double prior = 0.5, prior_weight = 0.25;
  std::vector<uint64_t> row_ptr;
  std::vector<uint32_t> col;
  std::vector<double> dist;
  uint64_t launches = 0;
  std::string err;
};

static std::string g_err;

static int fail(bpomp_ctx* ctx, int code, const char* fmt, ...) {
  char buf[256];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  (ctx ? ctx->err : g_err) = buf;
  return code;
}

static void valid_of(bpomp_ctx* c, const double* states, int64_t count, uint8_t* out) {
  if (c->world == 1)
    orc_states_valid_boxes(states, count, c->dim, c->lo.data(), c->hi.data(), (int)(c->lo.size() / c->dim), out);
  else
    orc_states_valid_image(states, count, c->pix.data(), c->W, c->H, c->thr, out);
}

extern "C" {

BPOMP_API int bpomp_create(int device, int dim, double segment_length, bpomp_ctx** out) {
  (void)device;
  if (!out) return fail(nullptr, BPOMP_ERR_INVALID, "out is NULL");
  if (dim < 1 || dim > BPOMP_DIM_MAX) return fail(nullptr, BPOMP_ERR_INVALID, "dim must be in 1..%d", BPOMP_DIM_MAX);
  if (!(segment_length > 0.0)) return fail(nullptr, BPOMP_ERR_INVALID, "segment_length must be > 0");
  bpomp_ctx* c = new bpomp_ctx();
  c->dim = dim;
  c->L = segment_length;
  *out = c;
  return BPOMP_OK;
}
BPOMP_API void bpomp_destroy(bpomp_ctx* ctx) { delete ctx; }
BPOMP_API const char* bpomp_last_error(const bpomp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }
BPOMP_API int bpomp_sync(bpomp_ctx* ctx) { return ctx ? BPOMP_OK : BPOMP_ERR_INVALID; }
BPOMP_API uint64_t bpomp_launch_count(const bpomp_ctx* ctx) { return ctx ? ctx->launches : 0; }

BPOMP_API int bpomp_world_set_image(bpomp_ctx* c, const uint8_t* pix, int width, int height, int threshold) {
  if (!c || !pix || width <= 0 || height <= 0 || c->dim != 2) return fail(c, BPOMP_ERR_INVALID, "bad image world");
  c->pix.assign(pix, pix + (size_t)width * height);
  c->W = width; c->H = height; c->thr = threshold;
  c->world = 2;
  return BPOMP_OK;
}
BPOMP_API int bpomp_world_set_boxes(bpomp_ctx* c, const double* lo, const double* hi, int nboxes) {
  if (!c || nboxes < 0 || (nboxes > 0 && (!lo || !hi))) return fail(c, BPOMP_ERR_INVALID, "bad box world");
  c->lo.assign(lo, lo + (size_t)nboxes * c->dim);
  c->hi.assign(hi, hi + (size_t)nboxes * c->dim);
  c->world = 1;
  return BPOMP_OK;
}

BPOMP_API int bpomp_states_valid(bpomp_ctx* c, const double* states, int64_t count, uint8_t* out) {
  if (!c) return BPOMP_ERR_INVALID;
  if (c->world == 0) return fail(c, BPOMP_ERR_NO_WORLD, "no collision world set");
  c->launches++;
  valid_of(c, states, count, out);
  return BPOMP_OK;
}

static void prune_of(bpomp_ctx* c, const double* states, int64_t count, const double* start, const double* goal,
                     double best, int check_validity, uint8_t* out) {
  orc_vertices_inadmissible(states, count, c->dim, start, goal, best, out);  // cost part, BP.cpp:636-647
  if (check_validity) {                                                     // || !isValid, BP.cpp:650-656
    std::vector<uint8_t> v((size_t)count);
    valid_of(c, states, count, v.data());
    for (int64_t i = 0; i < count; ++i) out[i] = (out[i] || !v[i]) ? 1 : 0;
  }
}

BPOMP_API int bpomp_states_prune(bpomp_ctx* c, const double* states, int64_t count, const double* start,
                                 const double* goal, double best_cost, int check_validity, uint8_t* out) {
  if (!c) return BPOMP_ERR_INVALID;
  if (check_validity && c->world == 0) return fail(c, BPOMP_ERR_NO_WORLD, "no collision world set");
  if (count == 0) return BPOMP_OK;
  c->launches++;
  prune_of(c, states, count, start, goal, best_cost, check_validity, out);
  return BPOMP_OK;
}
BPOMP_API int bpomp_vertices_prune(bpomp_ctx* c, uint32_t start_id, uint32_t goal_id, double best_cost,
                                   int check_validity, uint8_t* out) {
  if (!c) return BPOMP_ERR_INVALID;
  const size_t n = c->active.size();
  if (n == 0) return BPOMP_OK;
  if (start_id >= n || goal_id >= n) return fail(c, BPOMP_ERR_RANGE, "start/goal id out of range");
  c->launches++;
  prune_of(c, c->verts.data(), (int64_t)n, &c->verts[(size_t)start_id * c->dim], &c->verts[(size_t)goal_id * c->dim],
           best_cost, check_validity, out);
  return BPOMP_OK;
}

BPOMP_API int bpomp_vertices_append(bpomp_ctx* c, const double* states, uint32_t count, uint32_t* first_id) {
  if (!c) return BPOMP_ERR_INVALID;
  if (first_id) *first_id = (uint32_t)c->active.size();
  c->verts.insert(c->verts.end(), states, states + (size_t)count * c->dim);
  c->active.insert(c->active.end(), count, 1);
  return BPOMP_OK;
}
BPOMP_API int bpomp_vertices_set_active(bpomp_ctx* c, const uint32_t* ids, uint32_t count, int flag) {
  if (!c) return BPOMP_ERR_INVALID;
  for (uint32_t i = 0; i < count; ++i) {
    if (ids[i] >= c->active.size()) return fail(c, BPOMP_ERR_RANGE, "vertex id out of range");
    c->active[ids[i]] = flag ? 1 : 0;
  }
  return BPOMP_OK;
}
BPOMP_API int bpomp_vertices_clear(bpomp_ctx* c) {
  if (!c) return BPOMP_ERR_INVALID;
  c->verts.clear();
  c->active.clear();
  return BPOMP_OK;
}

static int neighbours(bpomp_ctx* c, const uint32_t* q, uint32_t nq, double r, uint64_t* nnz) {
  c->launches++;
  const int64_t N = (int64_t)c->active.size();
  c->row_ptr.assign((size_t)nq + 1, 0);
  orc_neighbors_radius(c->verts.data(), c->active.data(), N, c->dim, q, nq, r, c->row_ptr.data(), nullptr, nullptr,
                       orc_max_threads());
  c->col.assign((size_t)c->row_ptr[nq] + 1, 0);
  c->dist.assign((size_t)c->row_ptr[nq] + 1, 0.0);
  orc_neighbors_radius(c->verts.data(), c->active.data(), N, c->dim, q, nq, r, c->row_ptr.data(), c->col.data(),
                       c->dist.data(), orc_max_threads());
  if (nnz) *nnz = c->row_ptr[nq];
  return BPOMP_OK;
}
BPOMP_API int bpomp_neighbors_radius(bpomp_ctx* c, const uint32_t* query_ids, uint32_t nq, double r, uint64_t* nnz) {
  if (!c) return BPOMP_ERR_INVALID;
  for (uint32_t i = 0; i < nq; ++i)
    if (query_ids[i] >= c->active.size()) return fail(c, BPOMP_ERR_RANGE, "query id out of range");
  return neighbours(c, query_ids, nq, r, nnz);
}
BPOMP_API int bpomp_neighbors_all(bpomp_ctx* c, double r, uint64_t* nnz) {
  if (!c) return BPOMP_ERR_INVALID;
  const uint32_t N = (uint32_t)c->active.size();
  // rows for ALL vertices; inactive ones get empty rows: query only the active ones, then spread the rows
  std::vector<uint32_t> q;
  for (uint32_t v = 0; v < N; ++v)
    if (c->active[v]) q.push_back(v);
  uint64_t total = 0;
  int rc = neighbours(c, q.data(), (uint32_t)q.size(), r, &total);
  if (rc) return rc;
  std::vector<uint64_t> rp((size_t)N + 1, 0);
  size_t k = 0;
  for (uint32_t v = 0; v < N; ++v) {
    uint64_t len = 0;
    if (c->active[v]) {
      len = c->row_ptr[k + 1] - c->row_ptr[k];
      ++k;
    }
    rp[v + 1] = rp[v] + len;
  }
  c->row_ptr.swap(rp);
  if (nnz) *nnz = total;
  return BPOMP_OK;
}
BPOMP_API int bpomp_neighbors_fetch(bpomp_ctx* c, uint64_t* row_ptr, uint32_t* col, double* dist) {
  if (!c || c->row_ptr.empty()) return fail(c, BPOMP_ERR_INVALID, "no neighbour result");
  const uint64_t nnz = c->row_ptr.back();
  if (row_ptr) memcpy(row_ptr, c->row_ptr.data(), c->row_ptr.size() * sizeof(uint64_t));
  if (col) memcpy(col, c->col.data(), nnz * sizeof(uint32_t));
  if (dist) memcpy(dist, c->dist.data(), nnz * sizeof(double));
  return BPOMP_OK;
}

static int gather(bpomp_ctx* c, const uint32_t* ids, int64_t count, std::vector<double>& out) {
  out.resize((size_t)count * c->dim);
  for (int64_t e = 0; e < count; ++e) {
    if (ids[e] >= c->active.size()) return fail(c, BPOMP_ERR_RANGE, "vertex id out of range");
    memcpy(&out[(size_t)e * c->dim], &c->verts[(size_t)ids[e] * c->dim], c->dim * sizeof(double));
  }
  return BPOMP_OK;
}

BPOMP_API int bpomp_edges_check_indexed(bpomp_ctx* c, const uint32_t* from_ids, const uint32_t* to_ids, int64_t count,
                                        uint8_t* valid, int32_t* first, uint32_t* nstates) {
  if (!c) return BPOMP_ERR_INVALID;
  if (c->world == 0) return fail(c, BPOMP_ERR_NO_WORLD, "no collision world set");
  if (count == 0) return BPOMP_OK;
  std::vector<double> f, t;
  int rc = gather(c, from_ids, count, f);
  if (rc) return rc;
  rc = gather(c, to_ids, count, t);
  if (rc) return rc;
  std::vector<uint32_t> ns((size_t)count);
  c->launches++;
  if (c->world == 1)
    orc_edges_check_boxes(f.data(), t.data(), count, c->dim, c->L, c->lo.data(), c->hi.data(),
                          (int)(c->lo.size() / c->dim), valid, first, ns.data(), orc_max_threads());
  else
    orc_edges_check_image(f.data(), t.data(), count, c->L, c->pix.data(), c->W, c->H, c->thr, valid, first, ns.data(),
                          orc_max_threads());
  if (nstates) memcpy(nstates, ns.data(), ns.size() * sizeof(uint32_t));
  return BPOMP_OK;
}

BPOMP_API int bpomp_belief_config(bpomp_ctx* c, int k, double prior, double prior_weight) {
  if (!c) return BPOMP_ERR_INVALID;
  if (k < 1 || k > 32) return fail(c, BPOMP_ERR_INVALID, "k must be in 1..32");
  c->knn = k; c->prior = prior; c->prior_weight = prior_weight;
  return BPOMP_OK;
}
BPOMP_API int bpomp_belief_clear(bpomp_ctx* c) {
  if (!c) return BPOMP_ERR_INVALID;
  c->bel_pts.clear();
  c->bel_vals.clear();
  return BPOMP_OK;
}
extern "C" void orc_halton(int64_t first_index, int64_t count, int d, double* out);
BPOMP_API int bpomp_roadmap_generate(bpomp_ctx* c, int kind, uint64_t first, uint32_t count, double* host_out) {
  if (!c || (kind != 0 && kind != 1) || (count > 0 && !host_out)) return BPOMP_ERR_INVALID;
  if (kind == 0) {
    orc_halton((int64_t)first, (int64_t)count, c->dim, host_out);
  } else {
    for (uint64_t t = 0; t < (uint64_t)count * c->dim; ++t) {  // workloads.uniform: splitmix64(seed, t + 1)
      uint64_t z = first + (t + 1) * 0x9E3779B97F4A7C15ull;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z = z ^ (z >> 31);
      host_out[t] = (double)(z >> 11) * (1.0 / 9007199254740992.0);
    }
  }
  return BPOMP_OK;
}
BPOMP_API int bpomp_belief_append(bpomp_ctx* c, const double* states, const double* values, int64_t count) {
  if (!c || count < 0 || (count > 0 && (!states || !values))) return BPOMP_ERR_INVALID;
  c->bel_pts.insert(c->bel_pts.end(), states, states + (size_t)count * c->dim);
  c->bel_vals.insert(c->bel_vals.end(), values, values + count);
  return BPOMP_OK;
}
BPOMP_API int bpomp_belief_size(bpomp_ctx* c, uint64_t* count) {
  if (!c || !count) return BPOMP_ERR_INVALID;
  *count = c->bel_vals.size();
  return BPOMP_OK;
}
BPOMP_API int bpomp_belief_estimate(bpomp_ctx* c, const double* queries, int64_t nq, double* out) {
  if (!c || nq < 0 || (nq > 0 && (!queries || !out))) return BPOMP_ERR_INVALID;
  c->launches++;
  orc_belief_estimate(c->bel_pts.data(), c->bel_vals.data(), (int64_t)c->bel_vals.size(), c->dim, queries, nq, c->knn,
                      c->prior, c->prior_weight, out, 1);
  return BPOMP_OK;
}
// checkAndSetEdgeBlocked's addPoint calls (BP.cpp:523-537), replayed from the first-invalid slot.
BPOMP_API int bpomp_belief_append_checked(bpomp_ctx* c, const double* from, const double* to,
                                          const int32_t* first_invalid_slot, int64_t count) {
  if (!c) return BPOMP_ERR_INVALID;
  c->launches++;
  std::vector<double> states;
  for (int64_t e = 0; e < count; ++e) {
    const double* f = from + (size_t)e * c->dim;
    const double* t = to + (size_t)e * c->dim;
    unsigned n = orc_edge_states(f, t, c->dim, c->L, nullptr, 0);
    states.resize((size_t)n * c->dim);
    orc_edge_states(f, t, c->dim, c->L, states.data(), n);
    const int32_t bad = first_invalid_slot[e];
    const unsigned last = bad >= 1 ? (unsigned)bad : (n >= 2 ? n - 2 : 0);
    for (unsigned s = 1; s <= last && s + 1 < n; ++s) {
      c->bel_pts.insert(c->bel_pts.end(), &states[(size_t)s * c->dim], &states[(size_t)(s + 1) * c->dim]);
      c->bel_vals.push_back((bad >= 1 && s == (unsigned)bad) ? 1.0 : 0.0);
    }
  }
  return BPOMP_OK;
}
BPOMP_API int bpomp_belief_append_checked_indexed(bpomp_ctx* c, const uint32_t* from_ids, const uint32_t* to_ids,
                                                  const int32_t* first_invalid_slot, const uint32_t* nstates,
                                                  int64_t count) {
  if (!c) return BPOMP_ERR_INVALID;
  if (count == 0) return BPOMP_OK;
  std::vector<double> f, t;
  int rc = gather(c, from_ids, count, f);
  if (rc) return rc;
  rc = gather(c, to_ids, count, t);
  if (rc) return rc;
  for (int64_t e = 0; e < count; ++e)
    if (nstates[e] != orc_edge_states(&f[(size_t)e * c->dim], &t[(size_t)e * c->dim], c->dim, c->L, nullptr, 0))
      return BPOMP_ERR_RANGE;
  return bpomp_belief_append_checked(c, f.data(), t.data(), first_invalid_slot, count);
}
BPOMP_API int bpomp_belief_edge_pfree_indexed(bpomp_ctx* c, const uint32_t* from_ids, const uint32_t* to_ids,
                                              int64_t count, double* pfree, uint32_t* nlookups) {
  if (!c) return BPOMP_ERR_INVALID;
  if (count == 0) return BPOMP_OK;
  std::vector<double> f, t;
  int rc = gather(c, from_ids, count, f);
  if (rc) return rc;
  rc = gather(c, to_ids, count, t);
  if (rc) return rc;
  c->launches++;
  orc_belief_edge_pfree(f.data(), t.data(), count, c->dim, c->L, c->bel_pts.data(), c->bel_vals.data(),
                        (int64_t)c->bel_vals.size(), c->knn, c->prior, c->prior_weight, pfree, nlookups,
                        orc_max_threads());
  return BPOMP_OK;
}

}  // extern "C"
