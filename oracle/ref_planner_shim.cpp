// ref_planner_shim.cpp — TEST INFRASTRUCTURE ONLY.
//
// A C interface around the REFERENCE's own planner and models, compiled from where they lie under
// /root/reference against the API stand-ins of oracle/standins (OMPL, Boost.Graph and Eigen are not installed
// here; see oracle/standins/README.md for what the stand-ins are faithful to).  oracle/Makefile (target `ref`)
// builds  $(REF)/src/BatchingPOMP.cpp + this file + ref_shim.cpp  ->  oracle/_ref/libbpomp_ref.so.  Nothing of
// the reference is copied into this repository.  Used by tests/ and tests/golden/make_golden_ref.py to pin the
// CPU oracle (oracle/*.cpp) — and through it the CUDA path — to the reference's code:
//   refp_*            batching_pomp::BatchingPOMP itself: setup / setProblemDefinition / solve, its counters,
//                     radius / alpha schedule and solution paths (src/BatchingPOMP.cpp, batching/*.hpp,
//                     util/Selector.hpp, util/RoadmapFromFile.hpp, cspacebelief/KNNModel.hpp, ...)
//   ref_knn_estimate  cspacebelief::KNNModel::addPoint / estimate (KNNModel.hpp:89-134, BeliefPoint.hpp)
//   ref_selector      util::Selector<Graph>::selectEdges (Selector.hpp:84-146)
// The two collision worlds are this project's definitions, the same as in bpomp_oracle.cpp.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "batching_pomp/BatchingPOMP.hpp"
#include <ompl/geometric/PathGeometric.h>

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {
namespace ob = ompl::base;

class WorldChecker : public ob::StateValidityChecker {
public:
  WorldChecker(const ob::SpaceInformationPtr& si, int dim) : ob::StateValidityChecker(si), d(dim) {}
  bool isValid(const ob::State* s) const override {
    const double* x = s->as<ob::RealVectorStateSpace::StateType>()->values;
    if (kind == 1) {  // closed boxes
      for (int b = 0; b < nb; ++b) {
        bool inside = true;
        for (int j = 0; j < d; ++j)
          if (!(lo[(size_t)b * d + j] <= x[j] && x[j] <= hi[(size_t)b * d + j])) {
            inside = false;
            break;
          }
        if (inside) return false;
      }
      return true;
    }
    if (kind == 2) {  // occupancy image on the unit square (README.md:56)
      double x0 = x[0], x1 = x[1];
      if (!(x0 >= 0.0 && x0 <= 1.0 && x1 >= 0.0 && x1 <= 1.0)) return false;
      int col = (int)std::floor(x0 * (double)W), row = (int)std::floor(x1 * (double)H);
      if (col > W - 1) col = W - 1;
      if (row > H - 1) row = H - 1;
      return (int)pix[(size_t)row * W + col] >= thr;
    }
    return true;
  }
  int d, kind = 0, nb = 0, W = 0, H = 0, thr = 128;
  std::vector<double> lo, hi;
  std::vector<uint8_t> pix;
};

struct RefPlanner {
  int d = 0;
  std::shared_ptr<ob::RealVectorStateSpace> space;
  ob::SpaceInformationPtr si;
  std::shared_ptr<WorldChecker> checker;
  std::unique_ptr<batching_pomp::BatchingPOMP> planner;
  ob::ProblemDefinitionPtr pdef;
  std::string error;
  long budget = -1;
};
}  // namespace

// The state space is set up BEFORE the planner is constructed (the constructor captures the check radius,
// src/BatchingPOMP.cpp:241,278).
REF_API void* refp_create(int d, double lo, double hi, double fraction) {
  RefPlanner* r = new RefPlanner();
  r->d = d;
  r->space = std::make_shared<ob::RealVectorStateSpace>((unsigned)d);
  r->space->setBounds(lo, hi);
  r->si = std::make_shared<ob::SpaceInformation>(r->space);
  r->si->setStateValidityCheckingResolution(fraction);
  r->checker = std::make_shared<WorldChecker>(r->si, d);
  r->si->setStateValidityChecker(r->checker);
  r->si->setup();
  r->planner.reset(new batching_pomp::BatchingPOMP(r->si));
  return r;
}
REF_API void refp_destroy(void* h) { delete (RefPlanner*)h; }
REF_API const char* refp_error(void* h) { return ((RefPlanner*)h)->error.c_str(); }
REF_API double refp_segment_length(void* h) { return ((RefPlanner*)h)->space->getLongestValidSegmentLength(); }
REF_API void refp_world_boxes(void* h, const double* lo, const double* hi, int nb) {
  RefPlanner* r = (RefPlanner*)h;
  r->checker->kind = 1;
  r->checker->nb = nb;
  r->checker->lo.assign(lo, lo + (size_t)nb * r->d);
  r->checker->hi.assign(hi, hi + (size_t)nb * r->d);
}
REF_API void refp_world_image(void* h, const uint8_t* pix, int W, int H, int thr) {
  RefPlanner* r = (RefPlanner*)h;
  r->checker->kind = 2;
  r->checker->W = W;
  r->checker->H = H;
  r->checker->thr = thr;
  r->checker->pix.assign(pix, pix + (size_t)W * H);
}
REF_API int refp_set_param(void* h, const char* key, const char* value) {
  RefPlanner* r = (RefPlanner*)h;
  return r->planner->params().setParam(key, value) ? 0 : -1;
}
#define REF_TRY(r, body)                   \
  try {                                    \
    body;                                  \
  } catch (const std::invalid_argument& e) { \
    (r)->error = e.what();                 \
    return -2;                             \
  } catch (const std::exception& e) {      \
    (r)->error = e.what();                 \
    return -1;                             \
  }
REF_API int refp_setup(void* h) {
  RefPlanner* r = (RefPlanner*)h;
  REF_TRY(r, r->planner->setup());
  return 0;
}
REF_API int refp_set_problem(void* h, const double* start, const double* goal) {
  RefPlanner* r = (RefPlanner*)h;
  ob::ScopedState<ob::RealVectorStateSpace> s(r->space), g(r->space);
  for (int j = 0; j < r->d; ++j) {
    s[j] = start[j];
    g[j] = goal[j];
  }
  r->pdef = std::make_shared<ob::ProblemDefinition>(r->si);
  r->pdef->setStartAndGoalStates(s.get(), g.get());
  REF_TRY(r, r->planner->setProblemDefinition(r->pdef));
  return 0;
}
// ptc: true once it has been evaluated `budget` times (budget < 0: never) — the same rule as the oracle's.
REF_API void refp_set_ptc_budget(void* h, long n) { ((RefPlanner*)h)->budget = n; }
REF_API int refp_solve(void* h) {
  RefPlanner* r = (RefPlanner*)h;
  ob::PlannerTerminationCondition ptc([r]() {
    if (r->budget == 0) return true;
    if (r->budget > 0) --r->budget;
    return false;
  });
  int st = 0;
  REF_TRY(r, st = (int)(ob::PlannerStatus::StatusType)r->planner->solve(ptc));
  return st;  // ob::PlannerStatus::StatusType: TIMEOUT 4, APPROXIMATE_SOLUTION 5, EXACT_SOLUTION 6, ABORT 8
}
REF_API double refp_best_cost(void* h) { return ((RefPlanner*)h)->planner->getCurrentBestCost(); }
REF_API double refp_alpha(void* h) { return ((RefPlanner*)h)->planner->getCurrentAlpha(); }
REF_API double refp_radius(void* h) { return ((RefPlanner*)h)->planner->mBatchingPtr->getCurrentRadius(); }
REF_API int refp_exhausted(void* h) { return ((RefPlanner*)h)->planner->mBatchingPtr->isExhausted() ? 1 : 0; }
REF_API unsigned refp_num_batches(void* h) { return ((RefPlanner*)h)->planner->mBatchingPtr->getNumBatches(); }
REF_API uint64_t refp_num_vertices(void* h) { return num_vertices(((RefPlanner*)h)->planner->g); }
REF_API uint64_t refp_num_edges(void* h) { return num_edges(((RefPlanner*)h)->planner->g); }
REF_API void refp_counters(void* h, uint64_t* out) {
  batching_pomp::BatchingPOMP& p = *((RefPlanner*)h)->planner;
  out[0] = p.getNumEdgeChecks();
  out[1] = p.getNumCollChecks();
  out[2] = p.getNumSearches();
  out[3] = p.getNumLookups();
}
// States of the last solution path added to the problem definition (pdef_->addSolutionPath, :1006): out[n][d];
// returns n (0 if none).  cap in states.
REF_API int refp_path_states(void* h, double* out, int cap) {
  RefPlanner* r = (RefPlanner*)h;
  if (!r->pdef || r->pdef->getSolutionCount() == 0) return 0;
  auto* path = r->pdef->getSolutionPath()->as<ompl::geometric::PathGeometric>();
  const int n = (int)path->getStateCount();
  for (int i = 0; i < n && i < cap; ++i) {
    const double* v = path->getState((size_t)i)->as<ob::RealVectorStateSpace::StateType>()->values;
    for (int j = 0; j < r->d; ++j) out[(size_t)i * r->d + j] = v[j];
  }
  return n;
}
REF_API uint64_t refp_solution_count(void* h) {
  RefPlanner* r = (RefPlanner*)h;
  return r->pdef ? r->pdef->getSolutionCount() : 0;
}

// ---- BatchingPOMP::initializeEdgePoints + checkAndSetEdgeBlocked (src/BatchingPOMP.cpp:435-461, 496-544) ----
// The reference's own edge evaluation on arbitrary edges: every edge gets two vertices and an edge record in the
// planner's public graph `g`, its states are materialised by initializeEdgePoints and marched by
// checkAndSetEdgeBlocked (belief inserts included), exactly as the planner does for a candidate-path edge.
// first_invalid is recovered from the planner's own isValid counter (getNumCollChecks).  The reference is
// single-threaded; `nthreads` independent planner instances each take a contiguous slice of the edges (a fresh
// instance every 2048 edges, so the affected-edge set and the belief model stay as small as in a short query).
// Used to pin the oracle's edge operator to the reference and as bench.py's reference arm.
namespace {
void ref_edges_slice(int d, double fraction, const double* blo, const double* bhi, int nb, const double* from,
                     const double* to, int64_t begin, int64_t end, uint8_t* valid, int32_t* first, uint32_t* nstates) {
  typedef batching_pomp::BatchingPOMP BP;
  for (int64_t c0 = begin; c0 < end; c0 += 2048) {
    const int64_t c1 = c0 + 2048 < end ? c0 + 2048 : end;
    auto space = std::make_shared<ob::RealVectorStateSpace>((unsigned)d);
    space->setBounds(0.0, 1.0);
    ob::SpaceInformationPtr si = std::make_shared<ob::SpaceInformation>(space);
    si->setStateValidityCheckingResolution(fraction);
    auto checker = std::make_shared<WorldChecker>(si, d);
    checker->kind = 1;
    checker->nb = nb;
    checker->lo.assign(blo, blo + (size_t)nb * d);
    checker->hi.assign(bhi, bhi + (size_t)nb * d);
    si->setStateValidityChecker(checker);
    si->setup();
    BP planner(si);
    for (int64_t e = c0; e < c1; ++e) {
      BP::Vertex uv[2];
      for (int k = 0; k < 2; ++k) {
        uv[k] = boost::add_vertex(planner.g);
        batching_pomp::StateConPtr st = std::make_shared<batching_pomp::StateCon>(space);
        double* v = st->state->as<ob::RealVectorStateSpace::StateType>()->values;
        const double* src = (k == 0 ? from : to) + (size_t)e * d;
        for (int j = 0; j < d; ++j) v[j] = src[j];
        planner.g[uv[k]].v_state = st;
      }
      std::pair<BP::Edge, bool> ne = boost::add_edge(uv[0], uv[1], planner.g);
      planner.g[ne.first].distance = planner.vertexDistFun(uv[0], uv[1]);  // BP.cpp:162-166
      planner.g[ne.first].blockedStatus = BP::UNKNOWN;
      planner.initializeEdgePoints(ne.first);
      const unsigned int before = planner.getNumCollChecks();
      const bool blocked = planner.checkAndSetEdgeBlocked(ne.first);
      valid[e] = blocked ? 0 : 1;
      first[e] = blocked ? (int32_t)(planner.getNumCollChecks() - before) : -1;
      nstates[e] = (uint32_t)planner.g[ne.first].edgeStates.size();
    }
  }
}
}  // namespace
REF_API int ref_edges_check_boxes(int d, double fraction, const double* blo, const double* bhi, int nb, const double* from,
                                  const double* to, int64_t count, uint8_t* valid, int32_t* first, uint32_t* nstates,
                                  int nthreads) {
  try {
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> ts;
    for (int t = 0; t < nthreads; ++t) {
      const int64_t b = count * t / nthreads, e = count * (t + 1) / nthreads;
      if (e > b) ts.emplace_back(ref_edges_slice, d, fraction, blo, bhi, nb, from, to, b, e, valid, first, nstates);
    }
    for (auto& t : ts) t.join();
    return 0;
  } catch (...) {
    return -1;
  }
}


// ---- cspacebelief::KNNModel (KNNModel.hpp:56-134) with the planner's distance function
// (beliefDistanceFunction, src/BatchingPOMP.cpp:213-219: the Eigen norm of the difference) ------------------
REF_API int ref_knn_estimate(const double* pts, const double* vals, int64_t M, int d, const double* queries, int64_t nq,
                             int K, double prior, double prior_weight, double* out) {
  using batching_pomp::cspacebelief::BeliefPoint;
  auto space = std::make_shared<ob::RealVectorStateSpace>((unsigned)d);
  std::function<double(const BeliefPoint&, const BeliefPoint&)> dist = [](const BeliefPoint& a, const BeliefPoint& b) {
    Eigen::VectorXd diff{a.stateValues - b.stateValues};
    return diff.norm();
  };
  batching_pomp::cspacebelief::KNNModel model((size_t)K, prior, prior_weight, dist);
  ob::State* s = space->allocState();
  double* v = s->as<ob::RealVectorStateSpace::StateType>()->values;
  try {
    for (int64_t i = 0; i < M; ++i) {
      for (int j = 0; j < d; ++j) v[j] = pts[(size_t)i * d + j];
      model.addPoint(BeliefPoint(s, (unsigned)d, vals[i]));
    }
    for (int64_t i = 0; i < nq; ++i) {
      for (int j = 0; j < d; ++j) v[j] = queries[(size_t)i * d + j];
      out[i] = model.estimate(BeliefPoint(s, (unsigned)d, -1.0));
    }
  } catch (const std::exception&) {
    space->freeState(s);
    return -1;
  }
  space->freeState(s);
  return 0;
}

// ---- util::Selector (Selector.hpp:84-146) on a path of n edges with the given probFree values; order[i] =
// index (position on the path) of the i-th edge to evaluate -----------------------------------------------------
REF_API int ref_selector(const char* type, const double* prob_free, int n, int* order) {
  typedef batching_pomp::BatchingPOMP::Graph Graph;
  typedef batching_pomp::BatchingPOMP::Edge Edge;
  try {
    batching_pomp::util::Selector<Graph> sel(type);
    Graph g;
    for (int i = 0; i <= n; ++i) add_vertex(g);
    std::vector<Edge> path;
    for (int i = 0; i < n; ++i) {
      Edge e = add_edge((size_t)i, (size_t)i + 1, g).first;
      g[e].probFree = prob_free[i];
      g[e].distance = 1.0;
      g[e].blockedStatus = 0;
      path.push_back(e);
    }
    std::vector<Edge> out = sel.selectEdges(g, path);
    for (int i = 0; i < (int)out.size(); ++i) order[i] = (int)source(out[(size_t)i], g);
    return (int)out.size();
  } catch (const std::invalid_argument&) {
    return -2;
  }
}
