// bpomp_oracle.cpp — CPU ORACLE for the batching_pomp hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under batching_pomp_b200/ (the product) may
// include, link or call this file.  It is used by tests/, __graft_entry__.smoke()
// and the cpu_baseline / --impl reference legs of bench.py, as the checker and as
// the timed CPU baseline, never as the thing shipped.
//
// PARITY UNPINNED: the reference (personalrobotics/batching_pomp) ships no tests,
// golden vectors or fixtures, and cannot be compiled here (OMPL, Boost and Eigen
// are absent).  This file is a dependency-free restatement of the reference's
// arithmetic, written by following the cited lines; third-party arithmetic
// (OMPL RealVectorStateSpace, GNAT result semantics) is restated from its
// published behaviour.  It is pinned only against the
// hand-derived known-answer values of SURVEY.md §4 (tests/test_oracle_kat.py).
// One exception: util/BisectPerm.hpp is header-only standard C++, so it IS
// compiled from where it lies (oracle/ref_shim.cpp -> oracle/_ref/libbpomp_ref.so)
// and orc_bisect_perm is pinned to it, live and through tests/golden/
// bisect_perm_ref.npz.  Everything else in this file stays unpinned.
//
// Compile: g++ -O2 -std=c++17 -fPIC -shared -fopenmp -ffp-contract=off   (no
// -march, no -ffast-math: the reference sets only -std=c++11, CMakeLists.txt:4,
// so x86-64 gcc emits unfused mulsd/addsd and a correctly rounded sqrtsd.)
//
// File:line citations are relative to /root/reference/.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <utility>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API extern "C" __attribute__((visibility("default")))

// ---------------------------------------------------------------------------
// OMPL RealVectorStateSpace arithmetic (third party, not in tree)
// ---------------------------------------------------------------------------

// RealVectorStateSpace::distance — called from src/BatchingPOMP.cpp:431 (vertexDistFun),
// :643-644 (isVertexInadmissible), util/RoadmapFromFile.hpp:158.
// dist = 0; for j: diff = a[j]-b[j]; dist += diff*diff; return sqrt(dist)
static inline double rv_distance(const double* a, const double* b, int d) {
  double dist = 0.0;
  for (int j = 0; j < d; ++j) {
    double diff = a[j] - b[j];
    dist += diff * diff;
  }
  return std::sqrt(dist);
}

// RealVectorStateSpace::interpolate — called from src/BatchingPOMP.cpp:458.
// out[j] = from[j] + (to[j]-from[j])*t
static inline void rv_interpolate(const double* from, const double* to, double t, int d, double* out) {
  for (int j = 0; j < d; ++j) out[j] = from[j] + (to[j] - from[j]) * t;
}

ORC_API double orc_distance(const double* a, const double* b, int d) { return rv_distance(a, b, d); }
ORC_API void orc_interpolate(const double* from, const double* to, double t, int d, double* out) {
  rv_interpolate(from, to, t, d, out);
}

// ---------------------------------------------------------------------------
// util::BisectPerm::get — include/batching_pomp/util/BisectPerm.hpp:45-89.
// Literal O(n^2) restatement (farthest-from-done greedy, first maximum wins).
// ---------------------------------------------------------------------------
static void bisect_perm(int n, std::vector<std::pair<int, int>>& perm) {
  perm.clear();
  if (n <= 0) return;
  std::vector<char> done(n, 0);
  std::vector<int> dist(n);
  for (;;) {
    int last_true = -1;
    for (int i = 0; i < n; i++) {  // BisectPerm.hpp:60-65
      if (done[i]) last_true = i;
      dist[i] = i - last_true;
    }
    last_true = n;
    for (int i = n - 1; i >= 0; i--) {  // :66-71
      if (done[i]) last_true = i;
      dist[i] = (last_true - i) < dist[i] ? (last_true - i) : dist[i];
    }
    int max_val = 0, max_i = 0;
    for (int i = 0; i < n; i++)  // :72-78 strict '<' => first maximum
      if (max_val < dist[i]) {
        max_val = dist[i];
        max_i = i;
      }
    if (!max_val) break;  // :79-80
    perm.push_back(std::make_pair(max_i, max_val));
    done[max_i] = 1;
  }
}

// Per-thread cache of permutations, as the reference caches them per n (BisectPerm.hpp:47-50).
struct PermCache {
  std::vector<std::vector<int>> rows;  // rows[n] = first components
  const std::vector<int>& get(int n) {
    if ((int)rows.size() <= n) rows.resize(n + 1);
    if (rows[n].empty() && n > 0) {
      std::vector<std::pair<int, int>> p;
      bisect_perm(n, p);
      rows[n].resize(n);
      for (int i = 0; i < n; ++i) rows[n][i] = p[i].first;
    }
    return rows[n];
  }
};

ORC_API int orc_bisect_perm(int n, int* out_first, int* out_second) {
  std::vector<std::pair<int, int>> p;
  bisect_perm(n, p);
  for (size_t i = 0; i < p.size(); ++i) {
    if (out_first) out_first[i] = p[i].first;
    if (out_second) out_second[i] = p[i].second;
  }
  return (int)p.size();
}

// ---------------------------------------------------------------------------
// initializeEdgePoints — src/BatchingPOMP.cpp:435-461
// nStates = (unsigned)floor(distance / (2.0*mCheckRadius)), at least 2; 2.0*mCheckRadius == L
// exactly (mCheckRadius = 0.5*L, :241,278).  Slot i holds interpolate(from,to,(1+perm[i])/(n+1)).
// ---------------------------------------------------------------------------
static inline unsigned nstates_of(double distance, double L) {
  unsigned nStates = static_cast<unsigned int>(std::floor(distance / L));  // :440
  if (nStates < 2u) nStates = 2u;                                          // :443-445
  return nStates;
}
ORC_API unsigned orc_nstates(double distance, double L) { return nstates_of(distance, L); }

static inline double slot_fraction(int perm_first, unsigned nStates) {
  return 1.0 * (1 + perm_first) / (nStates + 1);  // :459  (double / unsigned->double)
}

// Fills states[n][d] in slot order; returns n.
ORC_API unsigned orc_edge_states(const double* from, const double* to, int d, double L, double* states,
                                 unsigned cap) {
  static thread_local PermCache pc;
  double distance = rv_distance(from, to, d);  // BP.cpp:163 distance(source,target)
  unsigned n = nstates_of(distance, L);
  if (!states) return n;
  const std::vector<int>& order = pc.get((int)n);
  for (unsigned i = 0; i < n && i < cap; ++i)
    rv_interpolate(from, to, slot_fraction(order[i], n), d, states + (size_t)i * d);
  return n;
}

// ---------------------------------------------------------------------------
// Collision worlds.  The reference's StateValidityChecker is user code (README.md:54-67
// example not in tree); these two worlds are defined by this project.
// ---------------------------------------------------------------------------

// n-D box world: invalid iff exists box b with lo[b][j] <= x[j] <= hi[b][j] for all j (closed).
static inline bool valid_boxes(const double* x, int d, const double* lo, const double* hi, int nb) {
  for (int b = 0; b < nb; ++b) {
    const double* l = lo + (size_t)b * d;
    const double* h = hi + (size_t)b * d;
    bool inside = true;
    for (int j = 0; j < d; ++j)
      if (!(l[j] <= x[j] && x[j] <= h[j])) {
        inside = false;
        break;
      }
    if (inside) return false;
  }
  return true;
}

// 2-D occupancy image on the unit square, (0,0)=upper-left, (1,1)=lower-right (README.md:56):
// col = min(W-1,(int)floor(x0*W)), row = min(H-1,(int)floor(x1*H)); valid iff pix[row*W+col] >= thr;
// states outside [0,1]^2 (or NaN) are invalid.
static inline bool valid_image(const double* x, const uint8_t* pix, int W, int H, int thr) {
  double x0 = x[0], x1 = x[1];
  if (!(x0 >= 0.0 && x0 <= 1.0 && x1 >= 0.0 && x1 <= 1.0)) return false;
  int col = (int)std::floor(x0 * (double)W);
  int row = (int)std::floor(x1 * (double)H);
  if (col > W - 1) col = W - 1;
  if (row > H - 1) row = H - 1;
  return (int)pix[(size_t)row * W + col] >= thr;
}

ORC_API void orc_states_valid_boxes(const double* states, int64_t count, int d, const double* lo,
                                    const double* hi, int nb, uint8_t* out) {
  for (int64_t i = 0; i < count; ++i) out[i] = valid_boxes(states + (size_t)i * d, d, lo, hi, nb) ? 1 : 0;
}
ORC_API void orc_states_valid_image(const double* states, int64_t count, const uint8_t* pix, int W, int H,
                                    int thr, uint8_t* out) {
  for (int64_t i = 0; i < count; ++i) out[i] = valid_image(states + (size_t)i * 2, pix, W, H, thr) ? 1 : 0;
}

// ---------------------------------------------------------------------------
// checkAndSetEdgeBlocked — src/BatchingPOMP.cpp:496-544
// Slots i = 1 .. n-2 in perm order (:510); first invalid => BLOCKED (:523-533).
// Output per edge: valid (1 = FREE, 0 = BLOCKED), first_invalid_slot (slot index in 1..n-2 or -1),
// nstates.  The number of isValid calls the reference makes is (first_invalid_slot) if blocked,
// else max(0,n-2)  (mNumCollChecks, :516).
// ---------------------------------------------------------------------------
template <class ValidFn>
static inline void edge_check_one(const double* from, const double* to, int d, double L, PermCache& pc,
                                  double* scratch, ValidFn&& is_valid, uint8_t* valid, int32_t* first_invalid,
                                  uint32_t* nstates) {
  double distance = rv_distance(from, to, d);
  unsigned n = nstates_of(distance, L);
  const std::vector<int>& order = pc.get((int)n);
  int32_t bad = -1;
  for (unsigned i = 1; i + 1 < n; ++i) {  // :510  i < nStates-1
    rv_interpolate(from, to, slot_fraction(order[i], n), d, scratch);
    if (!is_valid(scratch)) {
      bad = (int32_t)i;
      break;
    }
  }
  *valid = bad < 0 ? 1 : 0;
  *first_invalid = bad;
  if (nstates) *nstates = n;
}

ORC_API void orc_edges_check_boxes(const double* from, const double* to, int64_t count, int d, double L,
                                   const double* lo, const double* hi, int nb, uint8_t* valid,
                                   int32_t* first_invalid, uint32_t* nstates, int nthreads) {
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    PermCache pc;
    std::vector<double> scratch(d);
#pragma omp for schedule(static, 4096)
    for (int64_t e = 0; e < count; ++e)
      edge_check_one(
          from + (size_t)e * d, to + (size_t)e * d, d, L, pc, scratch.data(),
          [&](const double* x) { return valid_boxes(x, d, lo, hi, nb); }, valid + e, first_invalid + e,
          nstates ? nstates + e : nullptr);
  }
}

ORC_API void orc_edges_check_image(const double* from, const double* to, int64_t count, double L,
                                   const uint8_t* pix, int W, int H, int thr, uint8_t* valid,
                                   int32_t* first_invalid, uint32_t* nstates, int nthreads) {
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    PermCache pc;
    double scratch[2];
#pragma omp for schedule(static, 4096)
    for (int64_t e = 0; e < count; ++e)
      edge_check_one(
          from + (size_t)e * 2, to + (size_t)e * 2, 2, L, pc, scratch,
          [&](const double* x) { return valid_image(x, pix, W, H, thr); }, valid + e, first_invalid + e,
          nstates ? nstates + e : nullptr);
  }
}

// "Reference-overhead" variant of the edge check: reproduces the allocation pattern of the
// reference for one edge — every slot state heap-allocated at edge creation (BP.cpp:449-452),
// a std::function-like indirect checker, and a scratch copy per sample (:512-513).  Reported
// beside the headline baseline; not the headline.
ORC_API void orc_edges_check_boxes_refstyle(const double* from, const double* to, int64_t count, int d,
                                            double L, const double* lo, const double* hi, int nb,
                                            uint8_t* valid, int32_t* first_invalid, uint32_t* nstates) {
  PermCache pc;
  for (int64_t e = 0; e < count; ++e) {
    const double* f = from + (size_t)e * d;
    const double* t = to + (size_t)e * d;
    double distance = rv_distance(f, t, d);
    unsigned n = nstates_of(distance, L);
    const std::vector<int>& order = pc.get((int)n);
    std::vector<double*> edgeStates(n);
    for (unsigned i = 0; i < n; ++i) edgeStates[i] = new double[d];
    for (unsigned i = 0; i < n; ++i) rv_interpolate(f, t, slot_fraction(order[i], n), d, edgeStates[i]);
    int32_t bad = -1;
    for (unsigned i = 1; i + 1 < n; ++i) {
      double* s = new double[d];
      std::memcpy(s, edgeStates[i], sizeof(double) * d);
      bool ok = valid_boxes(edgeStates[i], d, lo, hi, nb);
      delete[] s;
      if (!ok) {
        bad = (int32_t)i;
        break;
      }
    }
    for (unsigned i = 0; i < n; ++i) delete[] edgeStates[i];
    valid[e] = bad < 0;
    first_invalid[e] = bad;
    if (nstates) nstates[e] = n;
  }
}

// ---------------------------------------------------------------------------
// neighbours_visitor::examine_vertex -> NearestNeighborsGNAT::nearestR — BP.cpp:150
// All ACTIVE vertices v with distance(q,v) <= r (inclusive), ascending by (distance, id);
// the query itself is included when active (the caller skips it, BP.cpp:158).
// Canonical tie rule (dist, id) is this project's: GNAT orders ties by address.
// Two-call protocol: pass col==nullptr to get row_ptr only.
// ---------------------------------------------------------------------------
ORC_API void orc_neighbors_radius(const double* verts, const uint8_t* active, int64_t N, int d,
                                  const uint32_t* query_ids, int64_t nq, double r, uint64_t* row_ptr,
                                  uint32_t* col, double* dist, int nthreads) {
  std::vector<uint64_t> counts(nq, 0);
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(dynamic, 16)
  for (int64_t qi = 0; qi < nq; ++qi) {
    const double* q = verts + (size_t)query_ids[qi] * d;
    uint64_t c = 0;
    for (int64_t v = 0; v < N; ++v) {
      if (active && !active[v]) continue;
      if (rv_distance(q, verts + (size_t)v * d, d) <= r) ++c;
    }
    counts[qi] = c;
  }
  row_ptr[0] = 0;
  for (int64_t qi = 0; qi < nq; ++qi) row_ptr[qi + 1] = row_ptr[qi] + counts[qi];
  if (!col) return;
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(dynamic, 16)
  for (int64_t qi = 0; qi < nq; ++qi) {
    const double* q = verts + (size_t)query_ids[qi] * d;
    std::vector<std::pair<double, uint32_t>> row;
    row.reserve(counts[qi]);
    for (int64_t v = 0; v < N; ++v) {
      if (active && !active[v]) continue;
      double dv = rv_distance(q, verts + (size_t)v * d, d);
      if (dv <= r) row.emplace_back(dv, (uint32_t)v);
    }
    std::sort(row.begin(), row.end());
    uint64_t o = row_ptr[qi];
    for (size_t i = 0; i < row.size(); ++i) {
      col[o + i] = row[i].second;
      if (dist) dist[o + i] = row[i].first;
    }
  }
}

// ---------------------------------------------------------------------------
// vertexPruneFunction / isVertexInadmissible — BP.cpp:636-656
// prune iff (best < DBL_MAX && d(start,v)+d(goal,v) > best) || !isValid(v)
// ---------------------------------------------------------------------------
ORC_API void orc_vertices_inadmissible(const double* states, int64_t count, int d, const double* start,
                                       const double* goal, double best_cost, uint8_t* out) {
  for (int64_t i = 0; i < count; ++i) {
    if (best_cost >= std::numeric_limits<double>::max()) {  // :638-640
      out[i] = 0;
      continue;
    }
    const double* v = states + (size_t)i * d;
    double through = rv_distance(start, v, d) + rv_distance(goal, v, d);  // :642-644
    out[i] = through > best_cost ? 1 : 0;                                  // :646
  }
}

// ---------------------------------------------------------------------------
// KNNModel::estimate — include/batching_pomp/cspacebelief/KNNModel.hpp:101-134,
// beliefDistanceFunction — BP.cpp:213-219 ((a-b).norm(), restated with sequential summation).
// kNN = brute force, ascending (dist, insertion index); a later point never displaces an
// equidistant earlier one (GNAT replaces only if strictly closer).
// ---------------------------------------------------------------------------
static double knn_estimate(const double* pts, const double* vals, int64_t M, int d, const double* q, int K,
                           double prior, double prior_weight, std::vector<std::pair<double, int64_t>>& heap) {
  if (M == 0) return prior;  // KNNModel.hpp:104-106
  int64_t knn = std::min<int64_t>(K, M);  // :108
  heap.clear();
  for (int64_t i = 0; i < M; ++i) {
    double dist = rv_distance(q, pts + (size_t)i * d, d);
    if ((int64_t)heap.size() < knn) {
      heap.emplace_back(dist, i);
      std::push_heap(heap.begin(), heap.end());
    } else if (std::make_pair(dist, i) < heap.front()) {
      std::pop_heap(heap.begin(), heap.end());
      heap.back() = std::make_pair(dist, i);
      std::push_heap(heap.begin(), heap.end());
    }
  }
  std::sort(heap.begin(), heap.end());
  // :117-127 weights/values, dot / sum (Eigen; restated sequentially)
  double dot = 0.0, wsum = 0.0;
  for (int64_t i = 0; i < knn; ++i) {
    double distance = std::max(0.0001, heap[i].first);  // :121
    double w = 1.0 / distance;                          // :122
    double v = vals[heap[i].second];                    // :123
    dot += w * v;
    wsum += w;
  }
  double result = dot / wsum;                                    // :127
  result = (result + prior_weight * prior) / (1 + prior_weight);  // :130
  return result;
}

ORC_API void orc_belief_estimate(const double* pts, const double* vals, int64_t M, int d,
                                 const double* queries, int64_t nq, int K, double prior, double prior_weight,
                                 double* out, int nthreads) {
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    std::vector<std::pair<double, int64_t>> heap;
#pragma omp for schedule(dynamic, 8)
    for (int64_t i = 0; i < nq; ++i)
      out[i] = knn_estimate(pts, vals, M, d, queries + (size_t)i * d, K, prior, prior_weight, heap);
  }
}

// Also returns the neighbour indices (ascending (dist, idx)) for debugging / parity on ids.
ORC_API void orc_belief_knn_ids(const double* pts, int64_t M, int d, const double* q, int K, int64_t* ids,
                                double* dists) {
  std::vector<std::pair<double, int64_t>> all;
  for (int64_t i = 0; i < M; ++i) all.emplace_back(rv_distance(q, pts + (size_t)i * d, d), i);
  std::sort(all.begin(), all.end());
  int64_t knn = std::min<int64_t>(K, M);
  for (int64_t i = 0; i < knn; ++i) {
    ids[i] = all[i].second;
    if (dists) dists[i] = all[i].first;
  }
}

// ---------------------------------------------------------------------------
// computeAndSetEdgeFreeProbability — BP.cpp:464-493
// stepSize = (unsigned)(nStates / log(nStates)); for i=1; i<nStates-1; i+=stepSize:
//   result *= (1 - estimate(state_i))
// ---------------------------------------------------------------------------
ORC_API unsigned orc_belief_step(unsigned nStates) {
  return static_cast<unsigned int>(nStates / std::log(nStates));  // :472 (size_t / double)
}

ORC_API void orc_belief_edge_pfree(const double* from, const double* to, int64_t count, int d, double L,
                                   const double* pts, const double* vals, int64_t M, int K, double prior,
                                   double prior_weight, double* pfree, uint32_t* nlookups, int nthreads) {
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    PermCache pc;
    std::vector<double> scratch(d);
    std::vector<std::pair<double, int64_t>> heap;
#pragma omp for schedule(dynamic, 8)
    for (int64_t e = 0; e < count; ++e) {
      const double* f = from + (size_t)e * d;
      const double* t = to + (size_t)e * d;
      unsigned n = nstates_of(rv_distance(f, t, d), L);
      const std::vector<int>& order = pc.get((int)n);
      unsigned step = orc_belief_step(n);
      double result = 1.0;
      uint32_t looks = 0;
      for (unsigned i = 1; i < n - 1; i += step) {  // :475
        rv_interpolate(f, t, slot_fraction(order[i], n), d, scratch.data());
        ++looks;                                     // :480
        double collProb = knn_estimate(pts, vals, M, d, scratch.data(), K, prior, prior_weight, heap);
        result *= (1.0 - collProb);                  // :487
      }
      pfree[e] = result;
      if (nlookups) nlookups[e] = looks;
    }
  }
}

// ---------------------------------------------------------------------------
// haltonRadiusFun / rggRadiusFun — BP.cpp:548-574 ; ompl::unitNBallMeasure
// ---------------------------------------------------------------------------
ORC_API double orc_halton_radius(unsigned n, unsigned dim) {
  double dimDbl = static_cast<double>(dim);
  double cardDbl = static_cast<double>(n);
  return 2.5 * std::pow(1.0 / cardDbl, 1 / dimDbl);  // :554
}
ORC_API double orc_unit_nball_measure(unsigned N) {
  // ompl/util/GeometricEquations: pow(sqrt(pi), N) / tgamma(N/2 + 1)
  return std::pow(std::sqrt(M_PI), static_cast<double>(N)) / std::tgamma(static_cast<double>(N) / 2.0 + 1.0);
}
ORC_API double orc_rgg_radius(unsigned n, unsigned dim, double space_measure) {
  double dimDbl = static_cast<double>(dim);
  double minRggR =
      2.0 * std::pow((1.0 + 1.0 / dimDbl) * (space_measure / orc_unit_nball_measure(dim)), 1.0 / dimDbl);  // :568-569
  double cardDbl = static_cast<double>(n);
  return minRggR * std::pow(std::log(cardDbl) / cardDbl, 1 / dimDbl);  // :573
}

// ---------------------------------------------------------------------------
// Halton roadmap (defined by this project): point i (1-based), coordinate j =
// radical inverse of i in base p_j, p = (2,3,5,7,11,13,17,19,...).
// ---------------------------------------------------------------------------
static const int kPrimes[16] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53};
ORC_API void orc_halton(int64_t first_index, int64_t count, int d, double* out) {
  for (int64_t i = 0; i < count; ++i)
    for (int j = 0; j < d; ++j) {
      uint64_t idx = (uint64_t)(first_index + i);
      double base = (double)kPrimes[j], f = 1.0, r = 0.0;
      while (idx > 0) {
        f = f / base;
        r = r + f * (double)(idx % (uint64_t)kPrimes[j]);
        idx /= (uint64_t)kPrimes[j];
      }
      out[(size_t)i * d + j] = r;
    }
}

ORC_API int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
