// planner_oracle.cpp — CPU ORACLE for the CALLER of the hot path: a literal, single-threaded
// restatement of batching_pomp::BatchingPOMP (planner shell, batching managers, selector,
// lazy A*) with every hot operator done on the CPU by brute force.
//
// TEST INFRASTRUCTURE ONLY (see bpomp_oracle.cpp header).  PARITY UNPINNED: the reference has
// no tests/fixtures and cannot be built here; Boost.Graph / OMPL semantics are restated from
// SURVEY.md Appendix A.  This file defines "bit-identical planned paths" for the project: the
// product's host planner (batching_pomp_b200/host/, flat arrays + CUDA through the C ABI) is
// an independent implementation and must return the same vertex sequences and counters.
//
// File:line citations are relative to /root/reference/ (BP.cpp = src/BatchingPOMP.cpp).

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>

#define ORC_API extern "C" __attribute__((visibility("default")))

extern "C" {
double orc_distance(const double* a, const double* b, int d);
void orc_interpolate(const double* from, const double* to, double t, int d, double* out);
unsigned orc_nstates(double distance, double L);
int orc_bisect_perm(int n, int* out_first, int* out_second);
void orc_states_valid_boxes(const double* states, int64_t count, int d, const double* lo, const double* hi,
                            int nb, uint8_t* out);
void orc_states_valid_image(const double* states, int64_t count, const uint8_t* pix, int W, int H, int thr,
                            uint8_t* out);
void orc_belief_estimate(const double* pts, const double* vals, int64_t M, int d, const double* queries,
                         int64_t nq, int K, double prior, double prior_weight, double* out, int nthreads);
double orc_halton_radius(unsigned n, unsigned dim);
double orc_rgg_radius(unsigned n, unsigned dim, double space_measure);
}

namespace {

const double kInf = std::numeric_limits<double>::max();

// ompl::base::PlannerStatus::StatusType values
enum Status { ST_UNKNOWN = 0, ST_INVALID_START = 1, ST_INVALID_GOAL = 2, ST_TIMEOUT = 4, ST_APPROXIMATE = 5,
              ST_EXACT = 6, ST_ABORT = 8 };

enum { FREE = 1, BLOCKED = -1, UNKNOWN = 0 };  // BatchingPOMP.hpp:74-80

struct EProps {  // BatchingPOMP.hpp:90-103 (edgeStates replaced by (nStates, orientation))
  uint32_t src, tgt;  // creation orientation: source(e), target(e) of the add_edge descriptor
  double distance = 0.0;
  double probFree = 0.0;
  int blockedStatus = 0;
  unsigned nStates = 0;  // 0 = edge points not initialised
  bool removed = false;
};

struct OutEdge { uint32_t target; uint32_t eid; };

// boost::adjacency_list<vecS,vecS,undirectedS> restated
struct Graph {
  std::vector<const double*> vstate;            // VProps::v_state
  std::vector<std::vector<OutEdge>> out;        // out-edge lists in insertion order
  std::vector<EProps> edges;                    // global edge store in creation order
  uint32_t add_vertex(const double* s) { vstate.push_back(s); out.emplace_back(); return (uint32_t)vstate.size() - 1; }
  uint32_t add_edge(uint32_t u, uint32_t v) {
    EProps e; e.src = u; e.tgt = v;
    edges.push_back(e);
    uint32_t id = (uint32_t)edges.size() - 1;
    out[u].push_back({v, id});
    out[v].push_back({u, id});
    return id;
  }
  // boost::edge(u,v,g): linear scan of u's out-edges
  int64_t edge(uint32_t u, uint32_t v) const {
    for (const OutEdge& oe : out[u]) if (oe.target == v) return oe.eid;
    return -1;
  }
  void clear_vertex(uint32_t v) {
    for (const OutEdge& oe : out[v]) {
      edges[oe.eid].removed = true;
      if (oe.target == v) continue;
      std::vector<OutEdge>& o = out[oe.target];
      o.erase(std::remove_if(o.begin(), o.end(), [&](const OutEdge& x) { return x.eid == oe.eid; }), o.end());
    }
    out[v].clear();
  }
};

// boost::d_ary_heap_indirect<Vertex,4,...,cost,std::less> restated literally
struct Heap4 {
  std::vector<uint32_t> data;
  std::vector<int64_t>& index_in_heap;
  const std::vector<double>& key;
  Heap4(std::vector<int64_t>& iih, const std::vector<double>& k) : index_in_heap(iih), key(k) {}
  bool empty() const { return data.empty(); }
  uint32_t top() const { return data[0]; }
  void push(uint32_t v) {
    size_t index = data.size();
    data.push_back(v);
    index_in_heap[v] = (int64_t)index;
    up(index);
  }
  void update(uint32_t v) { up((size_t)index_in_heap[v]); }
  void up(size_t index) {
    size_t orig_index = index, num_levels_moved = 0;
    if (index == 0) return;
    uint32_t moving = data[index];
    double moving_dist = key[moving];
    for (;;) {
      if (index == 0) break;
      size_t parent_index = (index - 1) / 4;
      if (moving_dist < key[data[parent_index]]) { ++num_levels_moved; index = parent_index; continue; }
      break;
    }
    index = orig_index;
    for (size_t i = 0; i < num_levels_moved; ++i) {
      size_t parent_index = (index - 1) / 4;
      uint32_t parent_value = data[parent_index];
      index_in_heap[parent_value] = (int64_t)index;
      data[index] = parent_value;
      index = parent_index;
    }
    data[index] = moving;
    index_in_heap[moving] = (int64_t)index;
  }
  void pop() {
    index_in_heap[data[0]] = -1;
    if (data.size() != 1) {
      data[0] = data.back();
      index_in_heap[data[0]] = 0;
      data.pop_back();
      down();
    } else {
      data.pop_back();
    }
  }
  void down() {
    if (data.empty()) return;
    size_t index = 0;
    double moving_dist = key[data[0]];
    size_t heap_size = data.size();
    for (;;) {
      size_t first_child = 4 * index + 1;
      if (first_child >= heap_size) break;
      size_t nchild = std::min<size_t>(4, heap_size - first_child);
      size_t smallest = 0;
      double smallest_dist = key[data[first_child]];
      for (size_t i = 1; i < nchild; ++i) {
        double dd = key[data[first_child + i]];
        if (dd < smallest_dist) { smallest = i; smallest_dist = dd; }
      }
      if (smallest_dist < moving_dist) {
        size_t c = first_child + smallest;
        std::swap(data[c], data[index]);
        index_in_heap[data[c]] = (int64_t)c;
        index_in_heap[data[index]] = (int64_t)index;
        index = c;
        continue;
      }
      break;
    }
  }
};

inline double closed_plus(double a, double b) {  // boost::closed_plus<double>(DBL_MAX)
  if (a == kInf) return kInf;
  if (b == kInf) return kInf;
  return a + b;
}

struct Planner {
  // ----- space / world -----
  int d = 0;
  double L = 0.0;            // getLongestValidSegmentLength(); mCheckRadius = 0.5*L (BP.cpp:241,278)
  double maxExtent = 0.0;    // getMaximumExtent()
  double spaceMeasure = 1.0; // si_->getSpaceMeasure()
  int worldKind = 0;         // 1 = boxes, 2 = image
  std::vector<double> boxLo, boxHi; int nb = 0;
  std::vector<uint8_t> pix; int W = 0, H = 0, thr = 128;

  // ----- params (BP.cpp:268-320) -----
  double mIncrement = 0.2, mStartGoalRadius = 0.0, mPruneThreshold = 1.05;
  std::string mGraphType, mBatchingType, mSelectorType;
  int K = 15; double prior = 0.5, priorWeight = 0.25;  // BP.cpp:300

  // ----- roadmap file contents -----
  std::vector<double> roadmap;  // N x d (graphml node order)
  std::vector<std::pair<uint32_t, uint32_t>> roadmapEdges;  // single batching only
  uint32_t N = 0;

  // ----- planner state -----
  Graph g;
  std::vector<uint8_t> nnActive;  // membership in mVertexNN, indexed by g vertex id
  std::vector<double> startState, goalState;
  uint32_t mStartVertex = 0, mGoalVertex = 0;
  double mCurrentAlpha = 0.0;
  bool mIsInitSearchBatch = true, mIsPathFound = false, mUsingSelector = false;
  double mBestPathCost = kInf;
  std::string selType = "normal";
  std::vector<uint32_t> mCurrBestPath;  // edge ids
  std::vector<uint32_t> mCurrBestPathVerts;
  std::set<uint32_t> mEdgesToUpdate;
  unsigned mNumEdgeChecks = 0, mNumCollChecks = 0, mNumSearches = 0, mNumLookups = 0;
  std::vector<double> beliefPts, beliefVals;  // KNNModel points (append-only)

  // ----- batching manager state (BM.hpp, VB.hpp, EB.hpp, HB.hpp, SB.hpp) -----
  unsigned bmNumBatches = 0; bool bmExhausted = false; double bmCurrRadius = 0.0;
  unsigned bmNumVerticesAdded = 0, bmNextVertexTarget = 0; double bmVertInfl = 2.0;
  double bmRadiusInfl = 1.0, bmInitRadius = 0.0, bmMaxRadius = 0.0;
  uint32_t bmCurrVertex = 0; bool bmEdgeMode = false;
  bool setupDone = false;

  std::string error;
  std::map<int, std::vector<int>> permCache;

  const std::vector<int>& perm(int n) {
    auto it = permCache.find(n);
    if (it != permCache.end()) return it->second;
    std::vector<int> f(n), s(n);
    int m = orc_bisect_perm(n, f.data(), s.data());
    f.resize(m);
    return permCache[n] = f;
  }

  bool isValid(const double* x) const {
    uint8_t o = 0;
    if (worldKind == 1) orc_states_valid_boxes(x, 1, d, boxLo.data(), boxHi.data(), nb, &o);
    else if (worldKind == 2) orc_states_valid_image(x, 1, pix.data(), W, H, thr, &o);
    else o = 1;
    return o != 0;
  }

  double vertexDistFun(uint32_t u, uint32_t v) const { return orc_distance(g.vstate[u], g.vstate[v], d); }  // BP.cpp:429-432

  double radiusFun(unsigned n) const {
    if (mGraphType == "halton") return orc_halton_radius(n, (unsigned)d);   // BP.cpp:548-556
    return orc_rgg_radius(n, (unsigned)d, spaceMeasure);                    // BP.cpp:558-574
  }

  // BP.cpp:435-461
  void initializeEdgePoints(uint32_t e) { g.edges[e].nStates = orc_nstates(g.edges[e].distance, L); }

  void edgeState(uint32_t e, unsigned slot, double* out) {
    const EProps& ep = g.edges[e];
    const std::vector<int>& order = perm((int)ep.nStates);
    double t = 1.0 * (1 + order[slot]) / (ep.nStates + 1);  // BP.cpp:459
    orc_interpolate(g.vstate[ep.src], g.vstate[ep.tgt], t, d, out);
  }

  double estimate(const double* q) {  // KNNModel::estimate
    double out;
    orc_belief_estimate(beliefPts.data(), beliefVals.data(), (int64_t)beliefVals.size(), d, q, 1, K, prior,
                        priorWeight, &out, 1);
    return out;
  }

  // BP.cpp:464-493
  double computeAndSetEdgeFreeProbability(uint32_t e) {
    size_t nStates = g.edges[e].nStates;
    unsigned stepSize = static_cast<unsigned int>(nStates / std::log(nStates));
    double result = 1.0;
    std::vector<double> q(d);
    for (unsigned i = 1; i < nStates - 1; i += stepSize) {
      edgeState(e, i, q.data());
      mNumLookups++;
      double collProb = estimate(q.data());
      result *= (1.0 - collProb);
    }
    g.edges[e].probFree = result;
    return result;
  }

  // BP.cpp:600-620
  void addAffectedEdges(uint32_t e) {
    uint32_t u = g.edges[e].src, v = g.edges[e].tgt;
    for (const OutEdge& oe : g.out[u]) mEdgesToUpdate.insert(oe.eid);
    for (const OutEdge& oe : g.out[v]) mEdgesToUpdate.insert(oe.eid);
  }

  // BP.cpp:496-544
  bool checkAndSetEdgeBlocked(uint32_t e) {
    mNumEdgeChecks++;
    addAffectedEdges(e);
    size_t nStates = g.edges[e].nStates;
    std::vector<double> s(d);
    for (unsigned i = 1; i < nStates - 1; i++) {
      edgeState(e, i, s.data());
      mNumCollChecks++;
      bool checkResult = isValid(s.data());
      beliefPts.insert(beliefPts.end(), s.begin(), s.end());
      if (checkResult == false) {
        beliefVals.push_back(1.0);  // :525-526
        g.edges[e].blockedStatus = BLOCKED;
        g.edges[e].distance = kInf;
        g.edges[e].probFree = 0.0;
        return true;
      } else {
        beliefVals.push_back(0.0);  // :535-536
      }
    }
    g.edges[e].blockedStatus = FREE;
    g.edges[e].probFree = 1.0;
    return false;
  }

  // util/Selector.hpp:95-146
  std::vector<uint32_t> selectEdges(const std::vector<uint32_t>& epath) const {
    if (selType == "normal") return epath;
    std::vector<uint32_t> r(epath);
    if (selType == "alternate") { std::reverse(r.begin(), r.end()); return r; }
    // failfast / maxinf: std::sort of pair<double, Edge>.  Boost orders edge descriptors by the address of
    // their property object, so equal keys come out in allocation-address order.  Canonical rule here: edge
    // creation order (the address order of an allocator that hands out increasing addresses; the stand-in the
    // reference is compiled against for the pinning tests uses the same rule) — exact ties are otherwise
    // "parity unpinned" against a real Boost build.
    std::vector<std::pair<std::pair<double, size_t>, uint32_t>> lst;
    for (size_t i = 0; i < epath.size(); ++i) {
      double p = g.edges[epath[i]].probFree;
      double key = selType == "failfast" ? p : std::fabs(p - 0.5);
      lst.push_back({{key, (size_t)epath[i]}, epath[i]});
    }
    std::sort(lst.begin(), lst.end());
    for (size_t i = 0; i < lst.size(); ++i) r[i] = lst[i].second;
    return r;
  }

  // BP.cpp:577-598
  bool checkAndUpdatePathBlocked(const std::vector<uint32_t>& ePath) {
    std::vector<uint32_t> sel = selectEdges(ePath);
    bool pathBlocked = false;
    for (uint32_t e : sel) {
      if (g.edges[e].blockedStatus == UNKNOWN) {
        if (checkAndSetEdgeBlocked(e)) {
          pathBlocked = true;
          if (mUsingSelector) return true;
        }
      }
    }
    return pathBlocked;
  }

  // BP.cpp:623-633  (std::set<Edge> iteration order is irrelevant to results: each update is independent)
  void updateAffectedEdgeWeights() {
    for (uint32_t e : mEdgesToUpdate)
      if (!g.edges[e].removed && g.edges[e].blockedStatus == UNKNOWN) computeAndSetEdgeFreeProbability(e);
    mEdgesToUpdate.clear();
  }

  // BP.cpp:636-647
  bool isVertexInadmissible(const double* v) const {
    if (mBestPathCost >= kInf) return false;
    double through = orc_distance(g.vstate[mStartVertex], v, d) + orc_distance(g.vstate[mGoalVertex], v, d);
    return through > mBestPathCost;
  }
  // BP.cpp:650-656
  bool vertexPruneFunction(const double* v) const { return isVertexInadmissible(v) == true || isValid(v) == false; }

  // BM.hpp:144-165
  void pruneVertices(bool fullPrune) {
    std::vector<uint32_t> rm;
    for (uint32_t v = 0; v < g.vstate.size(); ++v) {
      bool p = fullPrune ? vertexPruneFunction(g.vstate[v]) : isVertexInadmissible(g.vstate[v]);
      if (p) rm.push_back(v);
    }
    for (uint32_t v : rm) {
      nnActive[v] = 0;  // _vertexNN.remove
      // pending affected-edge entries of removed edges are dropped in updateAffectedEdgeWeights
      g.clear_vertex(v);
    }
  }

  uint32_t admit(uint32_t roadmapIdx) {
    uint32_t nv = g.add_vertex(&roadmap[(size_t)roadmapIdx * d]);
    nnActive.push_back(0);
    return nv;
  }

  // VB.hpp:94-138, EB.hpp:116-159, HB.hpp:146-202, SB.hpp:55-91
  void nextBatch() {
    if (bmExhausted) return;
    ++bmNumBatches;
    if (mBatchingType == "vertex" || (mBatchingType == "hybrid" && !bmEdgeMode)) {
      std::vector<uint32_t> added;
      while (bmNumVerticesAdded < bmNextVertexTarget) {
        if (!vertexPruneFunction(&roadmap[(size_t)bmCurrVertex * d])) added.push_back(admit(bmCurrVertex));
        ++bmCurrVertex;
        ++bmNumVerticesAdded;
        if (bmCurrVertex == N) {
          if (mBatchingType == "vertex") bmExhausted = true; else bmEdgeMode = true;
          break;
        }
      }
      for (uint32_t v : added) nnActive[v] = 1;
      if (mBatchingType == "hybrid") bmCurrRadius = radiusFun(std::min(bmNextVertexTarget, N));  // HB.hpp:187-189
      bmNextVertexTarget = static_cast<unsigned int>(bmNextVertexTarget * bmVertInfl);
    } else if (mBatchingType == "edge") {
      if (bmNumBatches == 1u) {
        std::vector<uint32_t> added;
        for (uint32_t i = 0; i < N; ++i)
          if (!vertexPruneFunction(&roadmap[(size_t)i * d])) added.push_back(admit(i));
        for (uint32_t v : added) nnActive[v] = 1;
        bmCurrRadius = bmInitRadius;
      } else {
        bmCurrRadius = bmCurrRadius * bmRadiusInfl;
      }
      if (bmCurrRadius > bmMaxRadius) { bmCurrRadius = bmMaxRadius; bmExhausted = true; }
    } else if (mBatchingType == "hybrid") {  // edge phase, HB.hpp:191-200
      bmCurrRadius = bmCurrRadius * bmRadiusInfl;
      if (bmCurrRadius > bmMaxRadius) { bmCurrRadius = bmMaxRadius; bmExhausted = true; }
    } else if (mBatchingType == "single") {
      pruneVertices(true);
      bmExhausted = true;
    }
  }
  void updateWithNewSolutionCost(double c) {
    if (mBatchingType == "edge" || mBatchingType == "hybrid") bmMaxRadius = std::min(bmMaxRadius, c);
  }

  // BP.cpp:668-757
  int setup() {
    if (mGraphType != "halton" && mGraphType != "rgg") {
      if (mBatchingType == "edge" || mBatchingType == "hybrid") {
        error = "Invalid graph type specified! Graph type needed for hybrid or edge batching!";
        return -1;
      }
    }
    if (N == 0) { error = "Roadmap name must be set before creating batching manager!"; return -1; }
    bmNumBatches = 0; bmExhausted = false; bmCurrRadius = 0.0;
    if (mBatchingType == "vertex") {
      bmNumVerticesAdded = 0; bmNextVertexTarget = 100u; bmVertInfl = 2.0;
      bmCurrRadius = kInf; bmCurrVertex = 0;  // VB.hpp:71
    } else if (mBatchingType == "edge") {
      bmRadiusInfl = std::pow(2.0, 1 / static_cast<double>(d));
      bmMaxRadius = maxExtent;
      bmInitRadius = radiusFun(N);  // EB.hpp:74
    } else if (mBatchingType == "hybrid") {
      bmNumVerticesAdded = 0; bmNextVertexTarget = 100u; bmVertInfl = 2.0;
      bmRadiusInfl = std::pow(2.0, 1 / static_cast<double>(d));
      bmMaxRadius = maxExtent;
      bmInitRadius = radiusFun(100u);  // HB.hpp:83
      bmCurrRadius = bmInitRadius; bmCurrVertex = 0; bmEdgeMode = false;
    } else if (mBatchingType == "single") {
      // SB.hpp:65: g = full_g (vertices 0..N-1 and the file's edges), then BP.cpp:739-743
      g = Graph(); nnActive.clear();
      for (uint32_t i = 0; i < N; ++i) admit(i);
      for (auto& pr : roadmapEdges) {
        uint32_t e = g.add_edge(pr.first, pr.second);
        g.edges[e].distance = vertexDistFun(pr.first, pr.second);  // RFF.hpp:149-160
      }
      for (uint32_t e = 0; e < g.edges.size(); ++e) initializeEdgePoints(e);
    } else {
      error = "Invalid batching type specified - " + mBatchingType + "!";
      return -1;
    }
    if (mSelectorType == "") { selType = "normal"; }
    else {
      if (mSelectorType != "normal" && mSelectorType != "alternate" && mSelectorType != "failfast" &&
          mSelectorType != "maxinf") { error = "Invalid selector type specified - " + mSelectorType + "!"; return -2; }
      mUsingSelector = true; selType = mSelectorType;
    }
    setupDone = true;
    return 0;
  }

  // BP.cpp:760-816
  int setProblemDefinition(const double* start, const double* goal) {
    startState.assign(start, start + d);
    goalState.assign(goal, goal + d);
    if (!isValid(startState.data())) { error = "Start configuration is in collision!"; return -3; }
    mStartVertex = g.add_vertex(startState.data()); nnActive.push_back(1);
    if (!isValid(goalState.data())) { error = "Goal configuration is in collision!"; return -4; }
    mGoalVertex = g.add_vertex(goalState.data()); nnActive.push_back(1);
    if (mBatchingType == "single") {
      for (uint32_t vi = 0; vi < g.vstate.size(); ++vi) {
        if (vi == mStartVertex || vi == mGoalVertex) continue;
        double startDist = vertexDistFun(mStartVertex, vi);
        if (startDist < mStartGoalRadius) {
          uint32_t e = g.add_edge(vi, mStartVertex);
          g.edges[e].blockedStatus = UNKNOWN; g.edges[e].distance = startDist; initializeEdgePoints(e);
        }
        double goalDist = vertexDistFun(mGoalVertex, vi);
        if (goalDist < mStartGoalRadius) {
          uint32_t e = g.add_edge(vi, mGoalVertex);
          g.edges[e].blockedStatus = UNKNOWN; g.edges[e].distance = goalDist; initializeEdgePoints(e);
        }
      }
    }
    return 0;
  }

  // BP.cpp:66-83
  double edgeWeight(uint32_t e) const {
    const EProps& ep = g.edges[e];
    if (ep.blockedStatus == BLOCKED) return kInf;
    double alpha = mCurrentAlpha;
    if (ep.blockedStatus == FREE) return alpha * ep.distance;
    double w_m = -std::log(ep.probFree);
    double w_l = ep.distance;
    double exp_cost = alpha * w_l + (1 - alpha) * w_m;
    return exp_cost;
  }

  // neighbours_visitor::examine_vertex, BP.cpp:143-169 (after the goal test)
  void expandNeighbours(uint32_t u, double radius) {
    std::vector<std::pair<double, uint32_t>> nbrs;
    for (uint32_t v = 0; v < g.vstate.size(); ++v) {
      if (!nnActive[v]) continue;
      double dv = vertexDistFun(u, v);
      if (dv <= radius) nbrs.emplace_back(dv, v);
    }
    std::sort(nbrs.begin(), nbrs.end());  // canonical (dist, id)
    for (auto& pr : nbrs) {
      uint32_t nbr = pr.second;
      if (nbr == u) continue;
      if (g.edge(u, nbr) < 0) {
        uint32_t e = g.add_edge(u, nbr);
        g.edges[e].distance = vertexDistFun(g.edges[e].src, g.edges[e].tgt);
        g.edges[e].blockedStatus = UNKNOWN;
        initializeEdgePoints(e);
        computeAndSetEdgeFreeProbability(e);
      }
    }
  }

  // boost::astar_search (initialising overload) + astar_bfs_visitor, SURVEY A.4.
  // Returns true if the goal was popped (throw_visitor_exception).
  bool astar(bool zeroHeuristic, bool single, std::vector<uint32_t>& pred, std::vector<double>& dist) {
    size_t nv = g.vstate.size();
    std::vector<double> cost(nv, kInf);
    std::vector<uint8_t> color(nv, 0);  // 0 white, 1 gray, 2 black
    std::vector<int64_t> iih(nv, -1);
    pred.resize(nv); dist.assign(nv, kInf);
    for (uint32_t v = 0; v < nv; ++v) pred[v] = v;
    // Zero in every search: BP.cpp:870-915 hands `*heuristicFn` (static type boost::astar_heuristic<Graph,double>)
    // to boost::astar_search, which takes the heuristic by value — the derived heuristics of BP.cpp:179-207 are
    // sliced away and the base operator() (non-virtual) returns 0.  Pinned against the reference's own source
    // compiled here (oracle/_ref; tests/test_oracle_vs_reference.py).
    (void)zeroHeuristic;
    auto h = [&](uint32_t) -> double { return 0.0; };
    double radius = bmCurrRadius;  // captured at visitor construction, BP.cpp:904
    uint32_t s = mStartVertex;
    dist[s] = 0.0;
    cost[s] = h(s);
    Heap4 Q(iih, cost);
    color[s] = 1;
    Q.push(s);
    while (!Q.empty()) {
      uint32_t u = Q.top();
      Q.pop();
      // examine_vertex
      if (u == mGoalVertex) return true;
      if (!single) expandNeighbours(u, radius);
      for (size_t k = 0; k < g.out[u].size(); ++k) {
        uint32_t eid = g.out[u][k].eid, v = g.out[u][k].target;
        // examine_edge: negative-weight test, then the user visitor
        double w0 = edgeWeight(eid);
        if (w0 < 0.0) { error = "negative edge weight"; return false; }
        if (single) computeAndSetEdgeFreeProbability(eid);  // throw_visitor::examine_edge, BP.cpp:108-111
        // relax_target
        double d_u = dist[u], d_v = dist[v], w_e = edgeWeight(eid);
        bool decreased = false;
        double cand = closed_plus(d_u, w_e);
        if (cand < d_v) {
          dist[v] = cand;
          if (dist[v] < d_v) { pred[v] = u; decreased = true; }
        }
        if (color[v] == 0) {         // tree_edge
          if (decreased) cost[v] = closed_plus(dist[v], h(v));
          color[v] = 1;
          Q.push(v);
        } else if (color[v] == 1) {  // gray_target
          if (decreased) { cost[v] = closed_plus(dist[v], h(v)); Q.update(v); }
        } else {                     // black_target
          if (decreased) { cost[v] = closed_plus(dist[v], h(v)); Q.push(v); color[v] = 1; }
        }
      }
      color[u] = 2;
    }
    return false;
  }

  // BP.cpp:819-1017.  ptcBudget: number of loop iterations allowed before ptc becomes true (<0 = never).
  long ptcBudget = -1;
  double ptcDeadline = 0.0;  // steady_clock seconds; 0 = none
  static double nowSeconds() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }
  int solve() {
    double currSolnCost = kInf;
    mIsPathFound = false;
    std::vector<uint32_t> ePath, vPath;
    while (currSolnCost >= mBestPathCost) {
      if (bmExhausted && mIsInitSearchBatch) {
        if (mBestPathCost < kInf) return ST_EXACT;
        return ST_TIMEOUT;
      }
      if (ptcDeadline > 0.0 && nowSeconds() >= ptcDeadline) return ST_TIMEOUT;
      if (ptcBudget == 0) return ST_TIMEOUT;
      if (ptcBudget > 0) --ptcBudget;
      if (mIsInitSearchBatch) {
        mCurrentAlpha = 0.0;
        nextBatch();
      }
      std::vector<uint32_t> startPreds;
      std::vector<double> startDist;
      mNumSearches++;
      bool zeroH = mIsInitSearchBatch;
      mIsInitSearchBatch = false;
      astar(zeroH, mBatchingType == "single", startPreds, startDist);
      if (!error.empty()) return ST_UNKNOWN;
      if (startDist[mGoalVertex] == kInf) {
        if (bmExhausted) {
          if (mBestPathCost < kInf) return ST_EXACT;
          return ST_TIMEOUT;
        } else {
          mIsInitSearchBatch = true;
          if (mBestPathCost < kInf) return ST_APPROXIMATE;
          return ST_ABORT;
        }
      }
      uint32_t vWalk = mGoalVertex;
      ePath.clear(); vPath.clear();
      while (vWalk != mStartVertex) {
        uint32_t vPred = startPreds[vWalk];
        int64_t e = g.edge(vPred, vWalk);
        if (e < 0) { error = "Error! Edge present during search no longer exists in graph!"; return ST_UNKNOWN; }
        ePath.push_back((uint32_t)e);
        vPath.push_back(vWalk);
        vWalk = vPred;
      }
      std::reverse(ePath.begin(), ePath.end());
      std::reverse(vPath.begin(), vPath.end());
      bool pathBlocked = checkAndUpdatePathBlocked(ePath);
      updateAffectedEdgeWeights();
      if (!pathBlocked) {
        mIsPathFound = true;
        currSolnCost = 0.0;
        for (uint32_t e : ePath) currSolnCost += g.edges[e].distance;  // BP.cpp:658-666
        if (mCurrentAlpha >= 1.0) mIsInitSearchBatch = true;
        else mCurrentAlpha = std::min(mCurrentAlpha + mIncrement, 1.0);
      }
    }
    double improvementFactor = mBestPathCost / currSolnCost;
    mBestPathCost = currSolnCost;
    mCurrBestPath = ePath;
    mCurrBestPathVerts.clear();
    mCurrBestPathVerts.push_back(mStartVertex);
    for (uint32_t v : vPath) mCurrBestPathVerts.push_back(v);  // target(e) for e on path, BP.cpp:1001-1004
    updateWithNewSolutionCost(mBestPathCost);
    if (improvementFactor > mPruneThreshold) pruneVertices(false);  // BP.cpp:1009-1013
    return ST_APPROXIMATE;
  }
};

}  // namespace

// ---------------------------------------------------------------------------
// C API (ctypes)
// ---------------------------------------------------------------------------
ORC_API void* orcp_create(int d, double L, double maxExtent, double spaceMeasure) {
  Planner* p = new Planner();
  p->d = d; p->L = L; p->maxExtent = maxExtent; p->spaceMeasure = spaceMeasure;
  return p;
}
ORC_API void orcp_destroy(void* h) { delete (Planner*)h; }
ORC_API void orcp_world_boxes(void* h, const double* lo, const double* hi, int nb) {
  Planner* p = (Planner*)h;
  p->worldKind = 1; p->nb = nb;
  p->boxLo.assign(lo, lo + (size_t)nb * p->d);
  p->boxHi.assign(hi, hi + (size_t)nb * p->d);
}
ORC_API void orcp_world_image(void* h, const uint8_t* pix, int W, int H, int thr) {
  Planner* p = (Planner*)h;
  p->worldKind = 2; p->W = W; p->H = H; p->thr = thr;
  p->pix.assign(pix, pix + (size_t)W * H);
}
ORC_API void orcp_roadmap(void* h, const double* states, uint32_t N, const uint32_t* edges, uint32_t nedges) {
  Planner* p = (Planner*)h;
  p->N = N;
  p->roadmap.assign(states, states + (size_t)N * p->d);
  p->roadmapEdges.clear();
  for (uint32_t i = 0; i < nedges; ++i) p->roadmapEdges.push_back({edges[2 * i], edges[2 * i + 1]});
}
ORC_API int orcp_set_param(void* h, const char* key, const char* val) {
  Planner* p = (Planner*)h;
  std::string k(key), v(val);
  if (k == "increment") p->mIncrement = std::stod(v);
  else if (k == "start_goal_radius") p->mStartGoalRadius = std::stod(v);
  else if (k == "prune_threshold") p->mPruneThreshold = std::stod(v);
  else if (k == "graph_type") p->mGraphType = v;
  else if (k == "batching_type") p->mBatchingType = v;
  else if (k == "selector_type") p->mSelectorType = v;
  else return -1;
  return 0;
}
ORC_API int orcp_setup(void* h) { return ((Planner*)h)->setup(); }
ORC_API int orcp_set_problem(void* h, const double* start, const double* goal) {
  return ((Planner*)h)->setProblemDefinition(start, goal);
}
ORC_API void orcp_set_ptc_budget(void* h, long n) { ((Planner*)h)->ptcBudget = n; }
ORC_API void orcp_set_time_limit(void* h, double seconds) {
  Planner* p = (Planner*)h;
  p->ptcDeadline = seconds > 0.0 ? Planner::nowSeconds() + seconds : 0.0;
}
ORC_API int orcp_solve(void* h) { return ((Planner*)h)->solve(); }
ORC_API const char* orcp_error(void* h) { return ((Planner*)h)->error.c_str(); }
ORC_API double orcp_best_cost(void* h) { return ((Planner*)h)->mBestPathCost; }
ORC_API double orcp_alpha(void* h) { return ((Planner*)h)->mCurrentAlpha; }
ORC_API double orcp_radius(void* h) { return ((Planner*)h)->bmCurrRadius; }
ORC_API int orcp_exhausted(void* h) { return ((Planner*)h)->bmExhausted ? 1 : 0; }
ORC_API unsigned orcp_num_batches(void* h) { return ((Planner*)h)->bmNumBatches; }
ORC_API int orcp_path(void* h, uint32_t* out, int cap) {
  Planner* p = (Planner*)h;
  int n = (int)p->mCurrBestPathVerts.size();
  for (int i = 0; i < n && i < cap; ++i) out[i] = p->mCurrBestPathVerts[i];
  return n;
}
// Maps a g vertex id to its state (so paths can be compared as state sequences).
ORC_API void orcp_vertex_state(void* h, uint32_t v, double* out) {
  Planner* p = (Planner*)h;
  std::memcpy(out, p->g.vstate[v], sizeof(double) * p->d);
}
ORC_API void orcp_counters(void* h, unsigned* out4) {
  Planner* p = (Planner*)h;
  out4[0] = p->mNumEdgeChecks; out4[1] = p->mNumCollChecks; out4[2] = p->mNumSearches; out4[3] = p->mNumLookups;
}
ORC_API uint64_t orcp_num_vertices(void* h) { return ((Planner*)h)->g.vstate.size(); }
// boost::num_edges(g): the edges clear_vertex removed (BatchingManager.hpp:163) do not count
ORC_API uint64_t orcp_num_edges(void* h) {
  uint64_t n = 0;
  for (const EProps& e : ((Planner*)h)->g.edges) n += e.removed ? 0 : 1;
  return n;
}
ORC_API uint64_t orcp_num_belief(void* h) { return ((Planner*)h)->beliefVals.size(); }
