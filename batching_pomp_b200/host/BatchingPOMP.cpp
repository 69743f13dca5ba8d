// BatchingPOMP.cpp — host planner of the B200 build.  Independent re-implementation of the
// control flow of /root/reference/src/BatchingPOMP.cpp (cited as BP.cpp) on flat arrays, with the
// hot operators batched into CUDA calls through include/bpomp_cuda.h:
//   nearestR (BP.cpp:150)                       -> bpomp_neighbors_all / bpomp_neighbors_radius
//   isValid loop (BP.cpp:510-538)               -> bpomp_edges_check_indexed (+ belief_append_checked)
//   computeAndSetEdgeFreeProbability (:464-493) -> bpomp_belief_edge_pfree_indexed
//   vertexPruneFunction (:650-656)              -> bpomp_states_prune / bpomp_vertices_prune
// Batching never changes observable behaviour: results are committed in the reference's order
// (edge creation order, path order, counters), see DESIGN.md "Host planner".
#include "BatchingPOMP.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>

#include <map>
#include <mutex>

#include "graphml.hpp"

namespace batching_pomp {
namespace core {

namespace {
const double kInf = std::numeric_limits<double>::max();
const uint32_t kNoEdge = 0xFFFFFFFFu;

inline double closedPlus(double a, double b) {  // boost::closed_plus<double>(DBL_MAX), BP.cpp:893
  if (a == kInf) return kInf;
  if (b == kInf) return kInf;
  return a + b;
}
inline uint64_t edgeKey(uint32_t u, uint32_t v) {
  uint32_t a = u < v ? u : v, b = u < v ? v : u;
  return ((uint64_t)a << 32) | b;
}
// RealVectorStateSpace::distance (host copy, used for the few scalar evaluations the host needs)
inline double rvDistance(const double* a, const double* b, unsigned d) {
  double dist = 0.0;
  for (unsigned j = 0; j < d; ++j) {
    double diff = a[j] - b[j];
    dist += diff * diff;
  }
  return std::sqrt(dist);
}
inline double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// Every call into the device library goes through this: its wall time (launch + wait for the results) is
// accumulated so that a run can be split into host and device shares (getDeviceTime()).
#define BP_DEV(expr)                 \
  do {                               \
    const double t0_ = now();        \
    const int rc_ = (expr);          \
    mDeviceTime += now() - t0_;      \
    check(rc_);                      \
  } while (0)
// Device contexts are recycled between planner instances of one process: the reference is
// one-planner-per-query, and creating a context (streams, tables, world upload) costs more than a
// short query.  A context is only reused for the same device, dimension, segment length and world.
struct CtxPool {
  std::mutex mu;
  struct Entry {
    std::vector<double> boxLo, boxHi;  // the world the contexts of this key were created with: compared on
    std::vector<uint8_t> image;        // every hit, so a hash collision can never hand out the wrong world
    std::vector<bpomp_ctx*> ctxs;
  };
  std::map<std::string, Entry> idle;
  // idle contexts are deliberately not destroyed at process exit: the CUDA runtime may already be
  // unloading when static destructors run, and the driver reclaims the memory with the process
};
CtxPool& pool() {
  static CtxPool p;
  return p;
}
std::string ctxKey(const SpaceInformation& si, double L) {
  uint64_t h = 1469598103934665603ull;  // FNV-1a over everything that defines the context
  auto mix = [&](const void* data, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < n; ++i) {
      h ^= b[i];
      h *= 1099511628211ull;
    }
  };
  int w = (int)si.world;
  mix(&si.device, sizeof(int));
  mix(&si.dim, sizeof(unsigned));
  mix(&L, sizeof(double));
  mix(&w, sizeof(int));
  if (!si.boxLo.empty()) mix(si.boxLo.data(), si.boxLo.size() * sizeof(double));
  if (!si.boxHi.empty()) mix(si.boxHi.data(), si.boxHi.size() * sizeof(double));
  if (!si.image.empty()) mix(si.image.data(), si.image.size());
  mix(&si.imageWidth, sizeof(int));
  mix(&si.imageHeight, sizeof(int));
  mix(&si.imageThreshold, sizeof(int));
  char buf[96];
  snprintf(buf, sizeof(buf), "%d/%u/%016llx/%zu/%zu", si.device, si.dim, (unsigned long long)h, si.boxLo.size(),
           si.image.size());
  return buf;
}
}  // namespace

// ---------------------------------------------------------------------------
// SpaceInformation
// ---------------------------------------------------------------------------
double SpaceInformation::getMaximumExtent() const {
  double e = 0.0;
  for (unsigned j = 0; j < dim; ++j) {
    double d = high[j] - low[j];
    e += d * d;
  }
  return std::sqrt(e);
}
double SpaceInformation::getLongestValidSegmentLength() const {
  return getMaximumExtent() * longestValidSegmentFraction;
}
double SpaceInformation::getSpaceMeasure() const {
  double m = 1.0;
  for (unsigned j = 0; j < dim; ++j) m *= high[j] - low[j];
  return m;
}

// ---------------------------------------------------------------------------
// construction
// ---------------------------------------------------------------------------
BatchingPOMP::BatchingPOMP(const SpaceInformation& si)
    : mDim(si.dim), mCtx(nullptr), mCurrentAlpha(0.0), mIncrement(0.2), mStartGoalRadius(0.0),
      mPruneThreshold(1.05), mIsInitSearchBatch(true), mIsPathFound(false), mUsingSelector(false),
      mBestPathCost(kInf), mSelType("normal"), mLaunchBase(0), mRoadmap(nullptr), mN(0), mFlushed(0), mStartVertex(0), mGoalVertex(0),
      mHaveProblem(false), mHashedUpTo(0), mEpoch(0), mSearchEdgesNeeded(0), mCsrValid(false), mCsrRadius(0.0), mCsrOlderEdges(false), mCsrBudget(30000000ull), bmNumBatches(0),
      bmExhausted(false), bmCurrRadius(0.0), bmNumVerticesAdded(0), bmNextVertexTarget(0), bmVertInfl(2.0),
      bmRadiusInfl(1.0), bmInitRadius(0.0), bmMaxRadius(0.0), bmCurrVertex(0), bmEdgeMode(false), mNumSolutions(0),
      mBeliefEmpty(true), mNumEdgeChecks(0), mNumCollChecks(0), mNumSearches(0), mNumLookups(0), mLookupTime(0.0),
      mCollCheckTime(0.0), mSearchTime(0.0), mDeviceTime(0.0), mSearchHostTime(0.0), mNumSpeculativeChecks(0), mBeliefK(15), mBeliefPrior(0.5),
      mBeliefPriorWeight(0.25) {
  if (mDim == 0 || si.low.size() != mDim || si.high.size() != mDim)
    throw Exception("SpaceInformation: bounds must have `dim` entries");
  // mCheckRadius = 0.5*getLongestValidSegmentLength() is captured at construction (BP.cpp:241,278);
  // 2.0*mCheckRadius == L exactly, which is what the device library takes.
  mL = si.getLongestValidSegmentLength();
  mMaxExtent = si.getMaximumExtent();
  mSpaceMeasure = si.getSpaceMeasure();
  if (si.world == SpaceInformation::World::NONE)
    throw Exception("the StateValidityChecker must be described as a device world (boxes or image); "
                    "there is no CPU fallback");
  mCtxKey = ctxKey(si, mL);
  // The device context is acquired at the first call that needs the GPU (setProblemDefinition, solve):
  // parameters, setup() and roadmap files are host work.  There is still no CPU fallback — that first
  // call throws when CUDA or the device library is unavailable.
  mSpace.reset(new SpaceInformation(si));
}

bpomp_ctx* BatchingPOMP::dev() {
  if (mCtx) return mCtx;
  if (!mSpace) throw Exception("device context was released");
  const SpaceInformation& si = *mSpace;
  {
    std::lock_guard<std::mutex> lk(pool().mu);
    auto it = pool().idle.find(mCtxKey);
    if (it != pool().idle.end() && !it->second.ctxs.empty() && it->second.boxLo == si.boxLo &&
        it->second.boxHi == si.boxHi && it->second.image == si.image) {
      mCtx = it->second.ctxs.back();
      it->second.ctxs.pop_back();
    }
  }
  if (mCtx) {  // recycled: same device / dim / L / world, emptied when it was returned
    mLaunchBase = bpomp_launch_count(mCtx);
    check(bpomp_belief_config(mCtx, mBeliefK, mBeliefPrior, mBeliefPriorWeight));
    return mCtx;
  }
  int rc = bpomp_create(si.device, (int)mDim, mL, &mCtx);
  if (rc != BPOMP_OK) {
    mCtx = nullptr;
    throw Exception(std::string("bpomp_create: ") + bpomp_last_error(nullptr));
  }
  try {
    if (si.world == SpaceInformation::World::BOXES) {
      int nb = (int)(si.boxLo.size() / mDim);
      check(bpomp_world_set_boxes(mCtx, si.boxLo.data(), si.boxHi.data(), nb));
    } else {
      check(bpomp_world_set_image(mCtx, si.image.data(), si.imageWidth, si.imageHeight, si.imageThreshold));
    }
    // belief model: KNNModel(15, 0.5, 0.25, ...) by default (BP.cpp:300), or the caller's (8-argument constructor)
    check(bpomp_belief_config(mCtx, mBeliefK, mBeliefPrior, mBeliefPriorWeight));
  } catch (...) {
    bpomp_destroy(mCtx);
    mCtx = nullptr;
    throw;
  }
  return mCtx;
}

void BatchingPOMP::setBeliefModel(int k, double prior, double priorWeight) {
  if (mCtx) throw Exception("setBeliefModel: the belief model is already resident on the device");
  if (k < 1 || k > 32) throw Exception("setBeliefModel: k must be in 1..32");
  mBeliefK = k;
  mBeliefPrior = prior;
  mBeliefPriorWeight = priorWeight;
}

BatchingPOMP::~BatchingPOMP() {
  if (!mCtx) return;
  // hand the emptied context back for the next planner instance (at most 4 idle per configuration)
  if (bpomp_vertices_clear(mCtx) == BPOMP_OK && bpomp_belief_clear(mCtx) == BPOMP_OK && bpomp_sync(mCtx) == BPOMP_OK) {
    std::lock_guard<std::mutex> lk(pool().mu);
    CtxPool::Entry& en = pool().idle[mCtxKey];
    const bool same = mSpace && en.boxLo == mSpace->boxLo && en.boxHi == mSpace->boxHi && en.image == mSpace->image;
    if (mSpace && en.ctxs.empty() && !same) {  // first context of this key (or a colliding key whose slots are free)
      en.boxLo = mSpace->boxLo;
      en.boxHi = mSpace->boxHi;
      en.image = mSpace->image;
    }
    if (mSpace && (same || en.ctxs.empty()) && en.ctxs.size() < 4) {
      en.ctxs.push_back(mCtx);
      return;
    }
  }
  bpomp_destroy(mCtx);
}

void BatchingPOMP::check(int rc) const {
  if (rc != BPOMP_OK) throw Exception(std::string("bpomp: ") + bpomp_last_error(mCtx));
}

void BatchingPOMP::setParam(const std::string& name, const std::string& value) {
  if (name == "increment") setIncrement(std::stod(value));
  else if (name == "start_goal_radius") setStartGoalRadius(std::stod(value));
  else if (name == "prune_threshold") setPruneThreshold(std::stod(value));
  else if (name == "graph_type") setGraphType(value);
  else if (name == "batching_type") setBatchingType(value);
  else if (name == "selector_type") setSelectorType(value);
  else if (name == "roadmap_filename") setRoadmapFileName(value);
  else throw Exception("unknown parameter " + name);
}

std::string BatchingPOMP::getParam(const std::string& name) const {
  if (name == "increment") return std::to_string(mIncrement);
  if (name == "start_goal_radius") return std::to_string(mStartGoalRadius);
  if (name == "prune_threshold") return std::to_string(mPruneThreshold);
  if (name == "graph_type") return mGraphType;
  if (name == "batching_type") return mBatchingType;
  if (name == "selector_type") return mSelectorType;
  if (name == "roadmap_filename") return mRoadmapName;
  throw Exception("unknown parameter " + name);
}

void BatchingPOMP::setRoadmap(const std::vector<double>& states,
                              const std::vector<std::pair<uint32_t, uint32_t>>& edges) {
  if (states.size() % mDim) throw Exception("roadmap states must be [n][dim]");
  mRoadmapOwned = states;
  mRoadmap = mRoadmapOwned.data();
  mRoadmapEdges = edges;
  mN = (uint32_t)(states.size() / mDim);
  if (mRoadmapName.empty()) mRoadmapName = "<memory>";
}

void BatchingPOMP::setRoadmapView(const double* states, uint32_t n) {
  mRoadmapOwned.clear();
  mRoadmap = states;
  mRoadmapEdges.clear();
  mN = n;
  if (mRoadmapName.empty()) mRoadmapName = "<memory>";
}

std::vector<double> BatchingPOMP::getSolutionStates() const {
  std::vector<double> out;
  for (Vertex v : mSolutionVertices) out.insert(out.end(), vertexState(v), vertexState(v) + mDim);
  return out;
}

// ---------------------------------------------------------------------------
// graph
// ---------------------------------------------------------------------------
BatchingPOMP::Vertex BatchingPOMP::addVertex(const double* state, bool inNN) {
  mVState.insert(mVState.end(), state, state + mDim);
  mOut.emplace_back();
  mNNActive.push_back(inNN ? 1 : 0);
  return (Vertex)mOut.size() - 1;
}

BatchingPOMP::Edge BatchingPOMP::addEdge(Vertex u, Vertex v, bool index) {
  EProps e;
  e.source = u;
  e.target = v;
  e.distance = 0.0;
  e.probFree = 0.0;
  e.negLogProbFree = -std::log(0.0);
  e.blockedStatus = UNKNOWN;
  e.nStates = 0;
  e.removed = false;
  mEdges.push_back(e);
  Edge id = (Edge)mEdges.size() - 1;
  mOut[u].push_back({v, id});
  mOut[v].push_back({u, id});
  // The pair index is only maintained while rows are fetched on demand; with a cached CSR the edge
  // ids live in the CSR slots (mCsrEdge) and the hash is brought up to date when it is needed again.
  if (index && mHashedUpTo == id) {
    mEdgeIndex.emplace(edgeKey(u, v), id);  // first edge wins, as boost::edge's scan returns the first
    mHashedUpTo = id + 1;
  }
  mEdgeMarked.push_back(0);
  return id;
}

void BatchingPOMP::completeEdgeIndex() {
  for (Edge e = mHashedUpTo; e < mEdges.size(); ++e)
    if (!mEdges[e].removed) mEdgeIndex.emplace(edgeKey(mEdges[e].source, mEdges[e].target), e);
  mHashedUpTo = (Edge)mEdges.size();
}

bool BatchingPOMP::hasEdge(Vertex u, Vertex v) {
  if (mHashedUpTo != mEdges.size()) completeEdgeIndex();
  return mEdgeIndex.find(edgeKey(u, v)) != nullptr;
}

// boost::edge(u, v, g): the first out-edge of u whose target is v (BP.cpp:960)
BatchingPOMP::Edge BatchingPOMP::findEdge(Vertex u, Vertex v) const {
  for (const OutEdge& oe : mOut[u])
    if (oe.target == v) return oe.eid;
  return (Edge)-1;
}

void BatchingPOMP::clearVertex(Vertex v) {  // boost::clear_vertex, BatchingManager.hpp:163
  for (const OutEdge& oe : mOut[v]) {
    if (!mEdges[oe.eid].removed) ++mNumRemovedEdges;
    mEdges[oe.eid].removed = true;
    if (oe.eid < mHashedUpTo) {
      const Edge* it = mEdgeIndex.find(edgeKey(v, oe.target));
      if (it && *it == oe.eid) mEdgeIndex.erase(edgeKey(v, oe.target));
    }
    if (oe.target == v) continue;
    std::vector<OutEdge>& o = mOut[oe.target];
    o.erase(std::remove_if(o.begin(), o.end(), [&](const OutEdge& x) { return x.eid == oe.eid; }), o.end());
  }
  mOut[v].clear();
}

// Appends not-yet-resident vertices to the device table (ids coincide with the g vertex ids).
void BatchingPOMP::flushVertices() {
  uint32_t nv = (uint32_t)mOut.size();
  if (mFlushed < nv) {
    uint32_t first = 0;
    BP_DEV(bpomp_vertices_append(dev(), &mVState[(size_t)mFlushed * mDim], nv - mFlushed, &first));
    if (first != mFlushed) throw Exception("internal: device vertex ids out of step with the roadmap");
    std::vector<uint32_t> off;
    for (uint32_t v = mFlushed; v < nv; ++v)
      if (!mNNActive[v]) off.push_back(v);
    if (!off.empty()) BP_DEV(bpomp_vertices_set_active(dev(), off.data(), (uint32_t)off.size(), 0));
    mFlushed = nv;
  }
  if (!mPendingDeactivate.empty()) {
    BP_DEV(bpomp_vertices_set_active(dev(), mPendingDeactivate.data(), (uint32_t)mPendingDeactivate.size(), 0));
    mPendingDeactivate.clear();
  }
}

// ---------------------------------------------------------------------------
// radius functions — BP.cpp:548-574 (host libm; results are passed to the device as doubles)
// ---------------------------------------------------------------------------
double BatchingPOMP::radiusFun(unsigned int n) const {
  auto dimDbl = static_cast<double>(mDim);
  auto cardDbl = static_cast<double>(n);
  if (mGraphType == "halton") return 2.5 * std::pow(1.0 / cardDbl, 1 / dimDbl);
  double approximationMeasure = mSpaceMeasure;
  if (!std::isfinite(approximationMeasure)) throw Exception("Measure of space is unbounded!");
  // ompl::unitNBallMeasure
  double unitBall = std::pow(std::sqrt(M_PI), dimDbl) / std::tgamma(dimDbl / 2.0 + 1.0);
  double minRggR = 2.0 * std::pow((1.0 + 1.0 / dimDbl) * (approximationMeasure / unitBall), 1.0 / dimDbl);
  return minRggR * std::pow(std::log(cardDbl) / cardDbl, 1 / dimDbl);
}

// ---------------------------------------------------------------------------
// setup — BP.cpp:668-757
// ---------------------------------------------------------------------------
void BatchingPOMP::setup() {
  if (mGraphType != "halton" && mGraphType != "rgg") {
    if (mBatchingType == "edge" || mBatchingType == "hybrid")
      throw Exception("Invalid graph type specified! Graph type needed for hybrid or edge batching!");
  }
  if (mRoadmapName == "") throw Exception("Roadmap name must be set before creating batching manager!");
  if (mN == 0 && (mRoadmapName.rfind("halton:", 0) == 0 || mRoadmapName.rfind("uniform:", 0) == 0)) {
    // B200 extension: "halton:<N>" / "uniform:<seed>:<N>" name a roadmap of this project's two
    // sample sequences; its states are generated on the device (bpomp_roadmap_generate), bit-equal to the GraphML
    // file of the same sequence, instead of being parsed from 200 MB of text.
    const bool halton = mRoadmapName[0] == 'h';
    unsigned long long first = 1, n = 0;
    char tail = 0;
    const int got = halton ? std::sscanf(mRoadmapName.c_str(), "halton:%llu%c", &n, &tail)
                           : std::sscanf(mRoadmapName.c_str(), "uniform:%llu:%llu%c", &first, &n, &tail);
    if (got != (halton ? 1 : 2) || n == 0 || n >= 0xFFFFFFF0ull)
      throw Exception("Roadmap name " + mRoadmapName + " is not halton:<N> or uniform:<seed>:<N>");
    mRoadmapOwned.assign((size_t)n * mDim, 0.0);
    BP_DEV(bpomp_roadmap_generate(dev(), halton ? BPOMP_ROADMAP_HALTON : BPOMP_ROADMAP_UNIFORM, first, (uint32_t)n,
                                  mRoadmapOwned.data()));
    mRoadmap = mRoadmapOwned.data();
    mRoadmapEdges.clear();
    mN = (uint32_t)n;
  }
  if (mN == 0 && mRoadmapName != "<memory>") {
    // RoadmapFromFile::generateVertices (util/RoadmapFromFile.hpp:136-147)
    std::vector<double> states;
    std::vector<std::pair<uint32_t, uint32_t>> edges;
    util::readGraphmlCached(mRoadmapName, mDim, states, edges);  // parsed once, then read from the binary cache
    mRoadmapOwned.swap(states);
    mRoadmap = mRoadmapOwned.data();
    mRoadmapEdges.swap(edges);
    mN = (uint32_t)(mRoadmapOwned.size() / mDim);
  }
  if (mN == 0) throw Exception("Roadmap " + mRoadmapName + " has no vertices");

  bmNumBatches = 0;
  bmExhausted = false;
  bmCurrRadius = 0.0;
  if (mBatchingType == "vertex") {
    bmNumVerticesAdded = 0;
    bmNextVertexTarget = 100u;  // initNumVertices, BP.cpp:692
    bmVertInfl = 2.0;           // vertInflFactor, BP.cpp:693
    bmCurrRadius = kInf;        // VertexBatching.hpp:71
    bmCurrVertex = 0;
  } else if (mBatchingType == "edge") {
    bmRadiusInfl = std::pow(2.0, 1 / static_cast<double>(mDim));  // BP.cpp:704
    bmMaxRadius = mMaxExtent;                                      // BP.cpp:705
    bmInitRadius = radiusFun(mN);                                  // EdgeBatching.hpp:74
  } else if (mBatchingType == "hybrid") {
    bmNumVerticesAdded = 0;
    bmNextVertexTarget = 100u;
    bmVertInfl = 2.0;
    bmRadiusInfl = std::pow(2.0, 1 / static_cast<double>(mDim));
    bmMaxRadius = mMaxExtent;
    bmInitRadius = radiusFun(100u);  // HybridBatching.hpp:83
    bmCurrRadius = bmInitRadius;
    bmCurrVertex = 0;
    bmEdgeMode = false;
  } else if (mBatchingType == "single") {
    // SingleBatching.hpp:65: g = full_g; then initializeEdgePoints for every edge (BP.cpp:739-743).
    if (mHaveProblem)
      throw Exception("single batching requires setup() before setProblemDefinition() (SingleBatching.hpp:65)");
    for (uint32_t i = 0; i < mN; ++i) addVertex(&mRoadmap[(size_t)i * mDim], false);
    for (auto& pr : mRoadmapEdges) {
      if (pr.first >= mN || pr.second >= mN) throw Exception("roadmap edge references an unknown node");
      Edge e = addEdge(pr.first, pr.second);
      mEdges[e].distance = rvDistance(vertexState(pr.first), vertexState(pr.second), mDim);  // RoadmapFromFile.hpp:158
      mEdges[e].nStates = 0;
    }
    for (EProps& e : mEdges) {  // initializeEdgePoints, BP.cpp:440-445
      unsigned int n = static_cast<unsigned int>(std::floor(e.distance / mL));
      e.nStates = n < 2u ? 2u : n;
    }
  } else {
    throw Exception("Invalid batching type specified - " + mBatchingType + "!");
  }

  if (mSelectorType == "") {
    mSelType = "normal";
  } else {
    if (mSelectorType != "normal" && mSelectorType != "alternate" && mSelectorType != "failfast" &&
        mSelectorType != "maxinf")
      throw std::invalid_argument("Invalid selector type specified - " + mSelectorType + "!");  // Selector.hpp:74
    mUsingSelector = true;
    mSelType = mSelectorType;
  }
}

// ---------------------------------------------------------------------------
// setProblemDefinition — BP.cpp:760-816
// ---------------------------------------------------------------------------
void BatchingPOMP::setProblemDefinition(const std::vector<double>& start, const std::vector<double>& goal) {
  if (start.size() != mDim || goal.size() != mDim) throw Exception("start/goal must have `dim` entries");
  std::vector<double> sg(start);
  sg.insert(sg.end(), goal.begin(), goal.end());
  uint8_t ok[2] = {0, 0};
  BP_DEV(bpomp_states_valid(dev(), sg.data(), 2, ok));
  if (!ok[0]) throw Exception("Start configuration is in collision!");
  mStartVertex = addVertex(start.data(), true);
  if (!ok[1]) throw Exception("Goal configuration is in collision!");
  mGoalVertex = addVertex(goal.data(), true);
  mHaveProblem = true;

  if (mBatchingType == "single") {
    for (Vertex vi = 0; vi < mOut.size(); ++vi) {
      if (vi == mStartVertex || vi == mGoalVertex) continue;
      double startDist = rvDistance(vertexState(mStartVertex), vertexState(vi), mDim);
      if (startDist < mStartGoalRadius) {
        Edge e = addEdge(vi, mStartVertex);
        mEdges[e].blockedStatus = UNKNOWN;
        mEdges[e].distance = startDist;
        unsigned int n = static_cast<unsigned int>(std::floor(startDist / mL));
        mEdges[e].nStates = n < 2u ? 2u : n;
      }
      double goalDist = rvDistance(vertexState(mGoalVertex), vertexState(vi), mDim);
      if (goalDist < mStartGoalRadius) {
        Edge e = addEdge(vi, mGoalVertex);
        mEdges[e].blockedStatus = UNKNOWN;
        mEdges[e].distance = goalDist;
        unsigned int n = static_cast<unsigned int>(std::floor(goalDist / mL));
        mEdges[e].nStates = n < 2u ? 2u : n;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// batching managers — VertexBatching.hpp:94-138, EdgeBatching.hpp:116-159,
// HybridBatching.hpp:146-202, SingleBatching.hpp:55-91.  Vertex admission
// (vertexPruneFunction per candidate) is one batched device call per batch.
// ---------------------------------------------------------------------------
void BatchingPOMP::nextBatch() {
  if (bmExhausted) return;
  ++bmNumBatches;
  auto admitRange = [&](uint32_t first, uint32_t count) {
    if (count == 0) return;
    std::vector<uint8_t> prune(count);
    BP_DEV(bpomp_states_prune(dev(), &mRoadmap[(size_t)first * mDim], count, vertexState(mStartVertex),
                             vertexState(mGoalVertex), mBestPathCost, 1, prune.data()));
    for (uint32_t i = 0; i < count; ++i)
      if (!prune[i]) addVertex(&mRoadmap[(size_t)(first + i) * mDim], true);
  };
  if (mBatchingType == "vertex" || (mBatchingType == "hybrid" && !bmEdgeMode)) {
    // the while loop adds candidates until the target is met or the file ends
    uint32_t want = bmNextVertexTarget > bmNumVerticesAdded ? bmNextVertexTarget - bmNumVerticesAdded : 0;
    uint32_t avail = mN - bmCurrVertex;
    uint32_t take = want < avail ? want : avail;
    admitRange(bmCurrVertex, take);
    bmCurrVertex += take;
    bmNumVerticesAdded += take;
    if (take > 0 && bmCurrVertex == mN) {
      if (mBatchingType == "vertex") bmExhausted = true;
      else bmEdgeMode = true;
    }
    if (mBatchingType == "hybrid") bmCurrRadius = radiusFun(std::min(bmNextVertexTarget, mN));  // HybridBatching.hpp:187-189
    bmNextVertexTarget = static_cast<unsigned int>(bmNextVertexTarget * bmVertInfl);
  } else if (mBatchingType == "edge") {
    if (bmNumBatches == 1u) {
      admitRange(0, mN);
      bmCurrRadius = bmInitRadius;
    } else {
      bmCurrRadius = bmCurrRadius * bmRadiusInfl;
    }
    if (bmCurrRadius > bmMaxRadius) {
      bmCurrRadius = bmMaxRadius;
      bmExhausted = true;
    }
  } else if (mBatchingType == "hybrid") {
    bmCurrRadius = bmCurrRadius * bmRadiusInfl;
    if (bmCurrRadius > bmMaxRadius) {
      bmCurrRadius = bmMaxRadius;
      bmExhausted = true;
    }
  } else if (mBatchingType == "single") {
    pruneVertices(true);
    bmExhausted = true;
  }
  mCsrValid = false;
}

void BatchingPOMP::updateWithNewSolutionCost(double c) {
  if (mBatchingType == "edge" || mBatchingType == "hybrid") bmMaxRadius = std::min(bmMaxRadius, c);
}

// BatchingManager::pruneVertices — BatchingManager.hpp:144-165 — over every vertex of g.
void BatchingPOMP::pruneVertices(bool withValidity) {
  flushVertices();
  uint32_t nv = (uint32_t)mOut.size();
  if (nv == 0) return;
  std::vector<uint8_t> prune(nv);
  BP_DEV(bpomp_vertices_prune(dev(), mStartVertex, mGoalVertex, mBestPathCost, withValidity ? 1 : 0, prune.data()));
  for (Vertex v = 0; v < nv; ++v) {
    if (!prune[v]) continue;
    if (mNNActive[v]) {
      mNNActive[v] = 0;
      mPendingDeactivate.push_back(v);
    }
    clearVertex(v);
  }
  flushVertices();
}

// ---------------------------------------------------------------------------
// neighbour rows
// ---------------------------------------------------------------------------
void BatchingPOMP::prepareNeighbourCache(double radius) {
  flushVertices();
  mCsrValid = false;
  uint32_t nv = (uint32_t)mOut.size();
  uint32_t nactive = 0;
  for (uint32_t v = 0; v < nv; ++v) nactive += mNNActive[v];
  // estimate the CSR size: complete graph for an unbounded radius, else from a sample of rows
  double est;
  if (radius >= kInf || nactive <= 64) {
    est = (double)nactive * (double)nactive;
  } else {
    std::vector<uint32_t> q;
    uint32_t stride = nv / 64 ? nv / 64 : 1;
    for (uint32_t v = 0; v < nv && q.size() < 64; v += stride) q.push_back(v);
    uint64_t nnz = 0;
    BP_DEV(bpomp_neighbors_radius(dev(), q.data(), (uint32_t)q.size(), radius, &nnz));
    est = (double)nnz / (double)q.size() * (double)nactive;
  }
  if (est > (double)mCsrBudget) return;  // rows are fetched on demand instead
  uint64_t nnz = 0;
  BP_DEV(bpomp_neighbors_all(dev(), radius, &nnz));
  mCsrRowPtr.resize((size_t)nv + 1);
  mCsrCol.resize(nnz ? nnz : 1);
  mCsrDist.resize(nnz ? nnz : 1);
  BP_DEV(bpomp_neighbors_fetch(dev(), mCsrRowPtr.data(), mCsrCol.data(), mCsrDist.data()));
  mCsrValid = true;
  mCsrRadius = radius;
  mCsrEdge.assign(nnz ? nnz : 1, kNoEdge);
  mCsrOlderEdges = !mEdges.empty();
  // the CSR bounds what this batch can create: size the host graph once instead of growing it edge by edge
  const size_t maxNew = (size_t)(nnz / 2 + 1);
  mEdges.reserve(mEdges.size() + maxNew);
  mEdgeMarked.reserve(mEdgeMarked.size() + maxNew);
  for (uint32_t v = 0; v < nv; ++v) {
    size_t len = (size_t)(mCsrRowPtr[v + 1] - mCsrRowPtr[v]);
    if (len > mOut[v].capacity()) mOut[v].reserve(len);
  }
}

const std::vector<std::pair<double, BatchingPOMP::Vertex>>& BatchingPOMP::neighbourRow(Vertex u, double radius) {
  mRowScratch.clear();
  // the cached CSR has rows for the vertices that were ACTIVE when it was built; rows of vertices
  // that have since been deactivated, or were never in the NN structure, are fetched on demand
  if (mCsrValid && radius == mCsrRadius && u + 1 < mCsrRowPtr.size() && mNNActive[u]) {
    for (uint64_t k = mCsrRowPtr[u]; k < mCsrRowPtr[u + 1]; ++k) {
      Vertex v = mCsrCol[k];
      if (mNNActive[v]) mRowScratch.emplace_back(mCsrDist[k], v);  // pruned since the CSR was built
    }
    return mRowScratch;
  }
  flushVertices();
  uint64_t nnz = 0;
  BP_DEV(bpomp_neighbors_radius(dev(), &u, 1, radius, &nnz));
  std::vector<uint64_t> rp(2);
  std::vector<uint32_t> col(nnz ? nnz : 1);
  std::vector<double> dist(nnz ? nnz : 1);
  BP_DEV(bpomp_neighbors_fetch(dev(), rp.data(), col.data(), dist.data()));
  for (uint64_t k = 0; k < nnz; ++k) mRowScratch.emplace_back(dist[k], col[k]);
  return mRowScratch;
}

// neighbours_visitor::examine_vertex after the goal test — BP.cpp:149-168
void BatchingPOMP::expandVertex(Vertex u, double radius) {
  // Within one batch (same radius, no vertices added) a second expansion of u finds every neighbour
  // already connected — by u's first expansion or by the neighbour's — so the row scan is skipped.
  // Vertices pruned in between only lose edges and leave the rows.
  if (mExpandedStamp.size() < mOut.size()) mExpandedStamp.resize(mOut.size(), 0);
  if (mExpandedStamp[u] == bmNumBatches) return;
  mExpandedStamp[u] = bmNumBatches;
  std::vector<Edge>& created = mCreatedScratch;
  created.clear();
  auto create = [&](Vertex nbr, double dist, bool index) -> Edge {
    Edge e = addEdge(u, nbr, index);
    mEdges[e].distance = dist;  // bit-equal to vertexDistFun(source, target)
    mEdges[e].blockedStatus = UNKNOWN;
    unsigned int n = static_cast<unsigned int>(std::floor(dist / mL));  // BP.cpp:440-445
    mEdges[e].nStates = n < 2u ? 2u : n;
    created.push_back(e);
    return e;
  };
  if (mCsrValid && radius == mCsrRadius && (size_t)u + 1 < mCsrRowPtr.size() && mNNActive[u]) {
    // Cached CSR: the edge id of every (u, neighbour) pair lives in the CSR slot itself, so the existence
    // test is a sequential read instead of a hash probe; the mirror slot (neighbour's row) is set too.
    const uint64_t rb = mCsrRowPtr[u], re = mCsrRowPtr[u + 1];
    if (mOut[u].capacity() == 0) mOut[u].reserve((size_t)(re - rb));  // one allocation per adjacency list
    for (uint64_t k = rb; k < re; ++k) {
      const Vertex nbr = mCsrCol[k];
      if (nbr == u || !mNNActive[nbr]) continue;  // pruned since the CSR was built
      if (mCsrEdge[k] != kNoEdge) continue;
      Edge e = mCsrOlderEdges ? findEdge(u, nbr) : kNoEdge;  // created before this CSR existed?
      if (e == kNoEdge) {
        if (mOut[nbr].capacity() == 0) mOut[nbr].reserve((size_t)(mCsrRowPtr[nbr + 1] - mCsrRowPtr[nbr]));
        e = create(nbr, mCsrDist[k], false);
      }
      mCsrEdge[k] = e;
      // mirror slot: rows are sorted by (distance, id) and distance(u, nbr) == distance(nbr, u) bit for bit,
      // so u sits in nbr's row at the first entry with this distance (then among the ties)
      const double* db = mCsrDist.data() + mCsrRowPtr[nbr];
      const double* de = mCsrDist.data() + mCsrRowPtr[nbr + 1];
      bool mirrored = false;
      for (uint64_t k2 = (uint64_t)(std::lower_bound(db, de, mCsrDist[k]) - mCsrDist.data());
           k2 < mCsrRowPtr[nbr + 1] && mCsrDist[k2] == mCsrDist[k]; ++k2)
        if (mCsrCol[k2] == u) {
          mCsrEdge[k2] = e;
          mirrored = true;
          break;
        }
      if (!mirrored) {  // not reached for rows in (distance, id) order; kept so that a slot is never missed
        for (uint64_t k2 = mCsrRowPtr[nbr]; k2 < mCsrRowPtr[nbr + 1]; ++k2)
          if (mCsrCol[k2] == u) {
            mCsrEdge[k2] = e;
            break;
          }
      }
    }
  } else {
    const auto& row = neighbourRow(u, radius);
    for (const auto& pr : row) {
      Vertex nbr = pr.second;
      if (nbr == u) continue;
      if (!hasEdge(u, nbr)) create(nbr, pr.first, true);
    }
  }
  computeCreatedEdgeProbabilities(u, created);  // BP.cpp:166
}

// Free probabilities of the edges one expansion created.  The belief model is read-only during a
// search, so probFree of an edge (source -> target) is the same whenever it is computed within that
// search.  Each device call therefore also evaluates, speculatively, the edges that the vertices at
// the head of the open list WOULD create when they are expanded (from the cached CSR rows), and keeps
// those values for this search only.  Later expansions then find their values in the cache and need
// no device call.  Values and the mNumLookups counter are taken only for edges that really get created,
// so results equal the reference's one-estimate-per-edge sequence.
void BatchingPOMP::computeCreatedEdgeProbabilities(Vertex u, const std::vector<Edge>& created) {
  if (created.empty()) return;
  if (mBeliefEmpty || !mCsrValid) {
    computeFreeProbabilities(created);
    return;
  }
  double t0 = now();
  std::vector<Edge> miss;
  for (Edge e : created) {
    const std::pair<double, uint32_t>* it = mSpecCache.find(((uint64_t)mEdges[e].source << 32) | mEdges[e].target);
    if (it) {
      setProbFree(mEdges[e], it->first);
      mNumLookups += it->second;
    } else {
      miss.push_back(e);
    }
  }
  if (!miss.empty()) {
    std::vector<uint32_t> from, to;
    for (Edge e : miss) {
      from.push_back(mEdges[e].source);
      to.push_back(mEdges[e].target);
    }
    const size_t nreal = from.size();
    // speculative part: potential edges of the first vertices of the open list (smallest f first-ish)
    // Speculation budget: half of what this search has really needed so far (geometric growth), so
    // the wasted work is bounded by that fraction when the goal is popped early (vertex batching pops
    // it after a few expansions of a complete graph) and the number of device calls stays logarithmic-ish
    // in the size of a search that sweeps a large roadmap.
    mSearchEdgesNeeded += nreal;
    const size_t kSpecVertices = mHeap.size();
    const size_t kSpecEdges = std::min<size_t>(262144, std::max<size_t>(512, mSearchEdgesNeeded / 2));
    size_t taken = 0;
    for (size_t h = 0; h < mHeap.size() && taken < kSpecVertices && from.size() < nreal + kSpecEdges; ++h) {
      Vertex v = mHeap[h].v;
      if (v == u || mSpecStamp[v] == mEpoch || !mNNActive[v] || v + 1 >= mCsrRowPtr.size()) continue;
      mSpecStamp[v] = mEpoch;
      ++taken;
      for (uint64_t k = mCsrRowPtr[v]; k < mCsrRowPtr[v + 1]; ++k) {
        Vertex w2 = mCsrCol[k];
        if (w2 == v || !mNNActive[w2] || mCsrEdge[k] != kNoEdge) continue;
        if (mCsrOlderEdges) {  // an edge from an earlier batch: remember it in the slot, nothing to evaluate
          Edge old = findEdge(v, w2);
          if (old != kNoEdge) {
            mCsrEdge[k] = old;
            continue;
          }
        }
        from.push_back(v);
        to.push_back(w2);
      }
    }
    std::vector<uint32_t> nl(from.size());
    std::vector<double> pf(from.size());
    flushVertices();
    BP_DEV(bpomp_belief_edge_pfree_indexed(dev(), from.data(), to.data(), (int64_t)from.size(), pf.data(), nl.data()));
    for (size_t i = 0; i < nreal; ++i) {
      setProbFree(mEdges[miss[i]], pf[i]);
      mNumLookups += nl[i];
    }
    for (size_t i = nreal; i < from.size(); ++i)
      mSpecCache.emplace(((uint64_t)from[i] << 32) | to[i], std::make_pair(pf[i], nl[i]));
  }
  mLookupTime += now() - t0;
}

// throw_visitor::examine_edge (single batching) recomputes probFree of every examined edge, BP.cpp:108-111
void BatchingPOMP::refreshOutEdgeProbabilities(Vertex u) {
  std::vector<Edge> es;
  es.reserve(mOut[u].size());
  for (const OutEdge& oe : mOut[u]) es.push_back(oe.eid);
  computeFreeProbabilities(es);
}

void BatchingPOMP::setProbFree(EProps& e, double p) {
  e.probFree = p;
  e.negLogProbFree = -std::log(p);
}

// computeAndSetEdgeFreeProbability for a set of edges — BP.cpp:464-493
void BatchingPOMP::computeFreeProbabilities(const std::vector<Edge>& edges) {
  if (edges.empty()) return;
  double t0 = now();
  if (mBeliefEmpty) {
    // KNNModel::estimate returns the prior while the model has no points (KNNModel.hpp:104-106)
    const double prior = mBeliefPrior;
    for (Edge e : edges) {
      size_t nStates = mEdges[e].nStates;
      unsigned int stepSize = static_cast<unsigned int>(nStates / std::log(nStates));
      double result = 1.0;
      for (unsigned int i = 1; i < nStates - 1; i += stepSize) {
        mNumLookups++;
        result *= (1.0 - prior);
      }
      setProbFree(mEdges[e], result);
    }
  } else {
    std::vector<uint32_t> from(edges.size()), to(edges.size()), nl(edges.size());
    std::vector<double> pf(edges.size());
    for (size_t i = 0; i < edges.size(); ++i) {
      from[i] = mEdges[edges[i]].source;
      to[i] = mEdges[edges[i]].target;
    }
    flushVertices();
    BP_DEV(bpomp_belief_edge_pfree_indexed(dev(), from.data(), to.data(), (int64_t)edges.size(), pf.data(), nl.data()));
    for (size_t i = 0; i < edges.size(); ++i) {
      setProbFree(mEdges[edges[i]], pf[i]);
      mNumLookups += nl[i];
    }
  }
  mLookupTime += now() - t0;
}

// Selector — util/Selector.hpp:95-146
std::vector<BatchingPOMP::Edge> BatchingPOMP::selectEdges(const std::vector<Edge>& ePath) const {
  if (mSelType == "normal") return ePath;
  std::vector<Edge> r(ePath);
  if (mSelType == "alternate") {
    std::reverse(r.begin(), r.end());
    return r;
  }
  // failfast: ascending probFree; maxinf: ascending |probFree-0.5|.  Exact ties: edge creation order —
  // std::sort over pair<double,Edge> orders equal keys by the edge descriptor, which Boost orders by the
  // address of the edge's property object, i.e. by allocation order on a fresh heap (Selector.hpp:113-146).
  std::vector<std::pair<std::pair<double, size_t>, Edge>> lst;
  for (size_t i = 0; i < ePath.size(); ++i) {
    double p = mEdges[ePath[i]].probFree;
    lst.push_back({{mSelType == "failfast" ? p : std::fabs(p - 0.5), (size_t)ePath[i]}, ePath[i]});
  }
  std::sort(lst.begin(), lst.end());
  for (size_t i = 0; i < lst.size(); ++i) r[i] = lst[i].second;
  return r;
}

// checkAndUpdatePathBlocked + checkAndSetEdgeBlocked — BP.cpp:577-598, 496-544.
// All UNKNOWN edges of the (re-ordered) path are evaluated in ONE device call; the results are then
// committed in order exactly as the reference's sequential loop would (stop at the first blocked
// edge when a selector type was set), so counters, statuses and belief inserts are identical.
bool BatchingPOMP::checkAndUpdatePathBlocked(const std::vector<Edge>& ePath) {
  std::vector<Edge> sel = selectEdges(ePath);
  std::vector<Edge> unknown;
  for (Edge e : sel)
    if (mEdges[e].blockedStatus == UNKNOWN) unknown.push_back(e);
  if (unknown.empty()) return false;
  double t0 = now();
  const size_t m = unknown.size();
  // Speculative edge evaluation.  Every still-unknown edge of the candidate path is evaluated in
  // ONE device call, although with a selector the reference stops at the first blocked edge (BP.cpp:591-593):
  // the collision world is static, so the result of an edge never changes, and the results of the edges behind
  // the stopping point are kept (mSpecStatus / mSpecFirst) until the reference's order gets to them — consecutive
  // candidate paths share most of their edges.  They are COMMITTED (status, counters, belief inserts) only then,
  // in the reference's order, so statuses, counters and the belief model are those of the sequential reference.
  if (mSpecStatus.size() < mEdges.size()) {
    mSpecStatus.resize(mEdges.size(), 0);
    mSpecFirst.resize(mEdges.size(), 0);
  }
  std::vector<uint32_t> from, to;
  std::vector<Edge> fresh;
  for (Edge e : unknown)
    if (!mSpecStatus[e]) {
      fresh.push_back(e);
      from.push_back(mEdges[e].source);
      to.push_back(mEdges[e].target);
    }
  if (!fresh.empty()) {
    const size_t k = fresh.size();
    std::vector<uint32_t> ns(k);
    std::vector<uint8_t> valid(k);
    std::vector<int32_t> first(k);
    flushVertices();
    BP_DEV(bpomp_edges_check_indexed(dev(), from.data(), to.data(), (int64_t)k, valid.data(), first.data(), ns.data()));
    for (size_t i = 0; i < k; ++i) {
      if (ns[i] != mEdges[fresh[i]].nStates) throw Exception("internal: device and host disagree on nStates");
      mSpecStatus[fresh[i]] = valid[i] ? 1 : 2;
      mSpecFirst[fresh[i]] = first[i];
    }
    mNumSpeculativeChecks += k;
  }
  bool pathBlocked = false;
  size_t committed = 0;
  for (size_t i = 0; i < m; ++i) {
    Edge e = unknown[i];
    EProps& ep = mEdges[e];
    mNumEdgeChecks++;  // BP.cpp:499
    // addAffectedEdges, BP.cpp:600-620
    for (Vertex x : {ep.source, ep.target})
      for (const OutEdge& oe : mOut[x])
        if (!mEdgeMarked[oe.eid]) {
          mEdgeMarked[oe.eid] = 1;
          mEdgesToUpdate.push_back(oe.eid);
        }
    committed = i + 1;
    if (mSpecStatus[e] == 2) {
      mNumCollChecks += (unsigned int)mSpecFirst[e];  // one isValid per slot up to the invalid one, BP.cpp:516
      ep.blockedStatus = BLOCKED;                      // BP.cpp:529-531
      ep.distance = kInf;
      setProbFree(ep, 0.0);
      pathBlocked = true;
      if (mUsingSelector) break;  // BP.cpp:591-593
    } else {
      mNumCollChecks += ep.nStates - 2u;
      ep.blockedStatus = FREE;  // BP.cpp:541-542
      setProbFree(ep, 1.0);
    }
  }
  // belief inserts of the committed edges, in check order (BP.cpp:523-537), replayed on the device from the
  // vertex ids, the first invalid slot and nStates
  uint64_t added = 0;
  std::vector<uint32_t> cf(committed), ct(committed), cn(committed);
  std::vector<int32_t> cfirst(committed);
  for (size_t i = 0; i < committed; ++i) {
    const EProps& ep = mEdges[unknown[i]];
    cf[i] = ep.source;
    ct[i] = ep.target;
    cn[i] = ep.nStates;
    cfirst[i] = mSpecStatus[unknown[i]] == 2 ? mSpecFirst[unknown[i]] : -1;
    added += cfirst[i] >= 1 ? (uint32_t)cfirst[i] : ep.nStates - 2u;
  }
  if (added > 0) {
    BP_DEV(bpomp_belief_append_checked_indexed(dev(), cf.data(), ct.data(), cfirst.data(), cn.data(), (int64_t)committed));
    mBeliefEmpty = false;
  }
  mCollCheckTime += now() - t0;
  return pathBlocked;
}

// updateAffectedEdgeWeights — BP.cpp:623-633
void BatchingPOMP::updateAffectedEdgeWeights() {
  std::vector<Edge> todo;
  for (Edge e : mEdgesToUpdate) {
    mEdgeMarked[e] = 0;
    if (!mEdges[e].removed && mEdges[e].blockedStatus == UNKNOWN) todo.push_back(e);
  }
  mEdgesToUpdate.clear();
  computeFreeProbabilities(todo);
}

// EdgeWeightMap get — BP.cpp:66-83
double BatchingPOMP::edgeWeight(Edge e) const {
  const EProps& ep = mEdges[e];
  if (ep.blockedStatus == BLOCKED) return kInf;
  double alpha = mCurrentAlpha;
  if (ep.blockedStatus == FREE) return alpha * ep.distance;
  double w_m = ep.negLogProbFree;  // == -std::log(ep.probFree), cached when probFree is set
  double w_l = ep.distance;
  double exp_cost = alpha * w_l + (1 - alpha) * w_m;
  return exp_cost;
}

// ---------------------------------------------------------------------------
// 4-ary indirect heap keyed on f (boost::d_ary_heap_indirect<Vertex,4,...>):
// sift-up while strictly smaller than the parent; pop moves the last element to the root and
// sifts down choosing the first strictly-smallest child.
// ---------------------------------------------------------------------------
void BatchingPOMP::touch(Vertex v) {
  SearchNode& n = mNode[v];
  if (n.stamp != mEpoch) {  // lazily applies astar_search's initialisation loop to v
    n.stamp = mEpoch;
    n.dist = kInf;
    n.cost = kInf;
    n.pred = v;
    n.color = 0;
    n.heapIndex = 0xffffffffu;
  }
}

void BatchingPOMP::heapUp(size_t index) {
  const HeapEntry moving{mNode[mHeap[index].v].cost, mHeap[index].v};  // refreshes the key after a decrease
  mHeap[index].key = moving.key;
  if (index == 0) return;
  size_t orig = index, moved = 0;
  for (;;) {
    if (index == 0) break;
    size_t parent = (index - 1) / 4;
    if (moving.key < mHeap[parent].key) {
      ++moved;
      index = parent;
      continue;
    }
    break;
  }
  index = orig;
  for (size_t i = 0; i < moved; ++i) {
    size_t parent = (index - 1) / 4;
    mHeap[index] = mHeap[parent];
    mNode[mHeap[index].v].heapIndex = (uint32_t)index;
    index = parent;
  }
  mHeap[index] = moving;
  mNode[moving.v].heapIndex = (uint32_t)index;
}

void BatchingPOMP::heapPush(Vertex v) {
  mHeap.push_back(HeapEntry{mNode[v].cost, v});
  mNode[v].heapIndex = (uint32_t)(mHeap.size() - 1);
  heapUp(mHeap.size() - 1);
}

void BatchingPOMP::heapDown() {
  if (mHeap.empty()) return;
  size_t index = 0, size = mHeap.size();
  const double key = mHeap[0].key;
  for (;;) {
    size_t firstChild = 4 * index + 1;
    if (firstChild >= size) break;
    size_t nchild = std::min<size_t>(4, size - firstChild);
    size_t best = 0;
    double bestKey = mHeap[firstChild].key;
    for (size_t i = 1; i < nchild; ++i) {
      double k = mHeap[firstChild + i].key;
      if (k < bestKey) {
        best = i;
        bestKey = k;
      }
    }
    if (bestKey < key) {
      size_t c = firstChild + best;
      std::swap(mHeap[c], mHeap[index]);
      mNode[mHeap[c].v].heapIndex = (uint32_t)c;
      mNode[mHeap[index].v].heapIndex = (uint32_t)index;
      index = c;
      continue;
    }
    break;
  }
}

void BatchingPOMP::heapPop() {
  mNode[mHeap[0].v].heapIndex = 0xffffffffu;
  if (mHeap.size() != 1) {
    mHeap[0] = mHeap.back();
    mNode[mHeap[0].v].heapIndex = 0;
    mHeap.pop_back();
    heapDown();
  } else {
    mHeap.pop_back();
  }
}

// boost::astar_search (initialising overload) with neighbours_visitor / throw_visitor —
// BP.cpp:860-924, 88-175; semantics in SURVEY.md A.4.  Returns true when the goal is popped.
bool BatchingPOMP::astar(bool zeroHeuristic, bool single) {
  size_t nv = mOut.size();
  if (mNode.size() < nv) mNode.resize(nv, SearchNode{0.0, 0.0, 0u, 0u, 0u, 0});
  if (++mEpoch == 0) {  // stamp wrap-around
    for (SearchNode& n : mNode) n.stamp = 0;
    std::fill(mSpecStamp.begin(), mSpecStamp.end(), 0);
    mEpoch = 1;
  }
  mHeap.clear();
  mSpecCache.clear();  // speculative probFree values are valid for one search (one belief-model state)
  mSearchEdgesNeeded = 0;
  if (mSpecStamp.size() < nv) mSpecStamp.resize(nv, 0);
  // The heuristic is ZERO in every search, as in the reference as compiled: BP.cpp:870-915 passes
  // `*heuristicFn` — static type boost::astar_heuristic<Graph,double> — to boost::astar_search, which takes
  // its heuristic BY VALUE; the derived zero_heuristic / exp_distance_heuristic objects (BP.cpp:179-207) are
  // sliced away and the base class's operator(), which is not virtual, returns 0.  Verified here by running
  // the reference's own source (oracle/_ref, tests/test_oracle_vs_reference.py): with alpha*distance as the
  // heuristic the later searches of a batch expand fewer vertices than the reference does (fewer edges
  // created, fewer belief lookups).
  (void)zeroHeuristic;
  auto h = [&](Vertex) -> double { return 0.0; };
  const double radius = bmCurrRadius;  // captured when the visitor is built, BP.cpp:904
  Vertex s = mStartVertex;
  touch(s);
  mNode[s].dist = 0.0;
  mNode[s].cost = h(s);
  mNode[s].color = 1;
  heapPush(s);
  while (!mHeap.empty()) {
    Vertex u = mHeap[0].v;
    heapPop();
    if (u == mGoalVertex) return true;  // examine_vertex throws, BP.cpp:104-106 / 145-147
    if (single) refreshOutEdgeProbabilities(u);
    else expandVertex(u, radius);
    const std::vector<OutEdge>& out = mOut[u];
    for (size_t k = 0; k < out.size(); ++k) {
      Edge eid = out[k].eid;
      Vertex v = out[k].target;
      double w_e = edgeWeight(eid);
      if (w_e < 0.0) throw Exception("negative edge weight");  // boost::negative_edge
      touch(v);
      SearchNode& nv_ = mNode[v];
      double d_u = mNode[u].dist, d_v = nv_.dist;
      bool decreased = false;
      double cand = closedPlus(d_u, w_e);
      if (cand < d_v) {  // relax_target
        nv_.dist = cand;
        nv_.pred = u;
        decreased = true;
      }
      if (nv_.color == 0) {  // tree_edge: pushed even if the relaxation failed
        if (decreased) nv_.cost = closedPlus(nv_.dist, h(v));
        nv_.color = 1;
        heapPush(v);
      } else if (nv_.color == 1) {  // gray_target
        if (decreased) {
          nv_.cost = closedPlus(nv_.dist, h(v));
          heapUp((size_t)nv_.heapIndex);
        }
      } else {  // black_target: re-open
        if (decreased) {
          nv_.cost = closedPlus(nv_.dist, h(v));
          heapPush(v);
          nv_.color = 1;
        }
      }
    }
    mNode[u].color = 2;
  }
  return false;
}

// ---------------------------------------------------------------------------
// solve — BP.cpp:819-1017
// ---------------------------------------------------------------------------
PlannerStatus BatchingPOMP::solve(const PlannerTerminationCondition& ptc) {
  if (!mHaveProblem) throw Exception("setProblemDefinition() has not been called");
  double currSolnCost = kInf;
  mIsPathFound = false;
  std::vector<Edge> ePath;
  std::vector<Vertex> vPath;

  while (currSolnCost >= mBestPathCost) {
    if (bmExhausted && mIsInitSearchBatch) {
      if (mBestPathCost < kInf) return PlannerStatus::EXACT_SOLUTION;
      return PlannerStatus::TIMEOUT;
    }
    if (ptc && ptc()) return PlannerStatus::TIMEOUT;

    if (mIsInitSearchBatch) {
      mCurrentAlpha = 0.0;
      nextBatch();
      if (mBatchingType != "single") prepareNeighbourCache(bmCurrRadius);
    }
    flushVertices();

    mNumSearches++;
    bool zeroH = mIsInitSearchBatch;
    mIsInitSearchBatch = false;
    double t0 = now();
    const double dev0 = mDeviceTime;
    astar(zeroH, mBatchingType == "single");
    const double dt = now() - t0;
    if (mBatchingType != "single") mSearchTime += dt;
    mSearchHostTime += dt - (mDeviceTime - dev0);

    touch(mGoalVertex);
    if (mNode[mGoalVertex].dist == kInf) {
      if (bmExhausted) {
        if (mBestPathCost < kInf) return PlannerStatus::EXACT_SOLUTION;
        return PlannerStatus::TIMEOUT;
      }
      mIsInitSearchBatch = true;
      if (mBestPathCost < kInf) return PlannerStatus::APPROXIMATE_SOLUTION;
      return PlannerStatus::ABORT;
    }

    Vertex vWalk = mGoalVertex;
    ePath.clear();
    vPath.clear();
    while (vWalk != mStartVertex) {
      Vertex vPred = mNode[vWalk].pred;
      Edge e = findEdge(vPred, vWalk);
      if (e == (Edge)-1) throw Exception("Error! Edge present during search no longer exists in graph!");
      ePath.push_back(e);
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
      for (Edge e : ePath) currSolnCost += mEdges[e].distance;  // getPathDistance, BP.cpp:658-666
      if (mCurrentAlpha >= 1.0) mIsInitSearchBatch = true;
      else mCurrentAlpha = std::min(mCurrentAlpha + mIncrement, 1.0);
    }
  }

  double improvementFactor = mBestPathCost / currSolnCost;
  mBestPathCost = currSolnCost;
  mCurrBestPath = ePath;
  updateWithNewSolutionCost(mBestPathCost);

  mSolutionVertices.clear();
  mSolutionVertices.push_back(mStartVertex);
  for (Vertex v : vPath) mSolutionVertices.push_back(v);  // target(e) for e on the path, BP.cpp:1001-1004
  ++mNumSolutions;

  if (improvementFactor > mPruneThreshold) pruneVertices(false);  // BP.cpp:1009-1013
  return PlannerStatus::APPROXIMATE_SOLUTION;
}

}  // namespace core
}  // namespace batching_pomp
