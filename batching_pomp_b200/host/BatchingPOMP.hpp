// BatchingPOMP.hpp — host side of the B200 build of batching_pomp: the planner shell, the roadmap
// graph, the batching managers, the selector and the lazy alpha-weighted A*, all on the CPU as in
// the reference, with every hot operator delegated to the CUDA library through the C ABI
// (include/bpomp_cuda.h).  It mirrors the public interface of
//   /root/reference/include/batching_pomp/BatchingPOMP.hpp:68-280
// (same method names, the same seven string/double parameters declared at
// /root/reference/src/BatchingPOMP.cpp:313-319, the same status values and counters) but does not
// depend on OMPL / Boost / Eigen, which are absent from this image.  ompl_adapter.hpp wraps it as
// an ompl::base::Planner where OMPL exists.
//
// There is no CPU fallback: constructing the planner without a CUDA device throws.
#pragma once

#include <cstdint>
#include <functional>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "bpomp_cuda.h"
#include "flat_map.hpp"

namespace batching_pomp {
// The planner core lives in batching_pomp::core; the class named batching_pomp::BatchingPOMP is the OMPL
// planner of include/batching_pomp/BatchingPOMP.hpp, which wraps this core.
namespace core {

/// ompl::base::PlannerStatus::StatusType values (OMPL is not linked; numeric values kept).
enum class PlannerStatus : int {
  UNKNOWN = 0,
  INVALID_START = 1,
  INVALID_GOAL = 2,
  UNRECOGNIZED_GOAL_TYPE = 3,
  TIMEOUT = 4,
  APPROXIMATE_SOLUTION = 5,
  EXACT_SOLUTION = 6,
  CRASH = 7,
  ABORT = 8
};

/// Stand-in for ompl::Exception (thrown at the places the reference throws it).
class Exception : public std::runtime_error {
public:
  explicit Exception(const std::string& what) : std::runtime_error(what) {}
};

/// What the planner needs from ompl::base::SpaceInformation for a RealVectorStateSpace, plus the
/// description of the StateValidityChecker as a device world (a checker that cannot
/// be described as one of these two worlds is an error — there is no CPU fallback).
struct SpaceInformation {
  unsigned int dim = 0;
  std::vector<double> low, high;                 // RealVectorBounds
  double longestValidSegmentFraction = 0.01;     // StateSpace default
  int device = 0;

  enum class World { NONE, BOXES, IMAGE } world = World::NONE;
  std::vector<double> boxLo, boxHi;              // [nboxes][dim]
  std::vector<uint8_t> image;                    // [height][width]
  int imageWidth = 0, imageHeight = 0, imageThreshold = 128;

  double getMaximumExtent() const;               // RealVectorStateSpace::getMaximumExtent
  double getLongestValidSegmentLength() const;   // maxExtent * fraction
  double getSpaceMeasure() const;                // prod(high - low)
};

typedef std::function<bool()> PlannerTerminationCondition;

class BatchingPOMP {
public:
  /// BatchingPOMP.hpp:74-80
  static const int FREE{1};
  static const int BLOCKED{-1};
  static const int UNKNOWN{0};

  typedef uint32_t Vertex;
  typedef uint32_t Edge;

  /// BatchingPOMP.hpp:90-103.  edgeStates are never materialised: (source, target, nStates) define them.
  struct EProps {
    Vertex source, target;  // creation orientation = interpolation direction (BP.cpp:162-165, 437-438)
    double distance;
    double probFree;
    double negLogProbFree;  // -log(probFree), kept beside it for the edge weight (BP.cpp:78)
    int blockedStatus;
    unsigned int nStates;
    bool removed;
  };

  explicit BatchingPOMP(const SpaceInformation& si);
  ~BatchingPOMP();
  BatchingPOMP(const BatchingPOMP&) = delete;
  BatchingPOMP& operator=(const BatchingPOMP&) = delete;

  // ---- OMPL ParamSet equivalents (BP.cpp:313-319) ----
  void setParam(const std::string& name, const std::string& value);
  std::string getParam(const std::string& name) const;

  // ---- setters / getters, BatchingPOMP.hpp:156-200 ----
  double getCurrentAlpha() const { return mCurrentAlpha; }
  double getCurrentBestCost() const { return mBestPathCost; }
  Vertex getStartVertex() const { return mStartVertex; }
  Vertex getGoalVertex() const { return mGoalVertex; }
  bool isInitSearchBatch() const { return mIsInitSearchBatch; }
  double getIncrement() const { return mIncrement; }
  void setIncrement(double v) { mIncrement = v; }
  double getStartGoalRadius() const { return mStartGoalRadius; }
  void setStartGoalRadius(double v) { mStartGoalRadius = v; }
  double getPruneThreshold() const { return mPruneThreshold; }
  void setPruneThreshold(double v) { mPruneThreshold = v; }
  std::string getGraphType() const { return mGraphType; }
  void setGraphType(const std::string& v) { mGraphType = v; }
  std::string getBatchingType() const { return mBatchingType; }
  void setBatchingType(const std::string& v) { mBatchingType = v; }
  std::string getSelectorType() const { return mSelectorType; }
  void setSelectorType(const std::string& v) { mSelectorType = v; }
  std::string getRoadmapFileName() const { return mRoadmapName; }
  void setRoadmapFileName(const std::string& v) { mRoadmapName = v; }

  unsigned int getNumEdgeChecks() const { return mNumEdgeChecks; }
  unsigned int getNumCollChecks() const { return mNumCollChecks; }
  unsigned int getNumSearches() const { return mNumSearches; }
  unsigned int getNumLookups() const { return mNumLookups; }
  double getLookupTime() const { return mLookupTime; }
  double getSearchTime() const { return mSearchTime; }
  double getCollCheckTime() const { return mCollCheckTime; }
  /// Wall time spent inside calls of the device library (not in the reference; the three times above
  /// include it where those calls are made from).
  double getDeviceTime() const { return mDeviceTime; }
  /// Wall time of the searches (astar, including edge creation) outside the device library.
  double getSearchHostTime() const { return mSearchHostTime; }
  /// Edges evaluated on the device so far, including those evaluated ahead of the reference's order whose
  /// results are waiting to be committed (getNumEdgeChecks counts committed evaluations, as the reference does).
  uint64_t getNumDeviceEdgeChecks() const { return mNumSpeculativeChecks; }

  // ---- OMPL required methods, BatchingPOMP.hpp:203-205 ----
  void setup();
  void setProblemDefinition(const std::vector<double>& start, const std::vector<double>& goal);
  PlannerStatus solve(const PlannerTerminationCondition& ptc);

  /// The roadmap may also be handed over in memory instead of through roadmap_filename
  /// (states [n][dim], optional edges for single batching).
  void setRoadmap(const std::vector<double>& states, const std::vector<std::pair<uint32_t, uint32_t>>& edges = {});
  /// Non-owning variant: `states` ([n][dim]) must outlive the planner.  Lets many planner instances
  /// (one per query, as in the reference) share one loaded roadmap.
  void setRoadmapView(const double* states, uint32_t n);

  /// Solutions appended by solve() (pdef_->addSolutionPath, BP.cpp:1006): vertex ids and states.
  const std::vector<Vertex>& getSolutionVertices() const { return mSolutionVertices; }
  std::vector<double> getSolutionStates() const;
  size_t getNumSolutions() const { return mNumSolutions; }

  // batching manager view (BatchingManager.hpp:86-104)
  unsigned int getNumBatches() const { return bmNumBatches; }
  bool isExhausted() const { return bmExhausted; }
  double getCurrentRadius() const { return bmCurrRadius; }

  /// The C-space belief model's parameters: KNNModel(k, prior, priorWeight, ...), BP.cpp:300 and the
  /// 8-argument constructor (BatchingPOMP.hpp:147-154).  Must be set before the first call that needs the GPU.
  void setBeliefModel(int k, double prior, double priorWeight);
  int getBeliefK() const { return mBeliefK; }
  double getBeliefPrior() const { return mBeliefPrior; }
  double getBeliefPriorWeight() const { return mBeliefPriorWeight; }

  // read-only views of the two roadmaps (the reference's public members g and full_g, BatchingPOMP.hpp:126-131)
  size_t roadmapSize() const { return mN; }                                  // num_vertices(full_g)
  const double* roadmapState(uint32_t i) const { return &mRoadmap[(size_t)i * mDim]; }
  size_t roadmapEdgeCount() const { return mRoadmapEdges.size(); }           // num_edges(full_g)
  size_t edgeSlots() const { return mEdges.size(); }                         // edge ids are 0 .. edgeSlots()-1
  const EProps& edgeProps(Edge e) const { return mEdges[e]; }                // g[e]; .removed: cleared by pruning
  size_t outDegree(Vertex v) const { return mOut[v].size(); }
  bool beliefEmpty() const { return mBeliefEmpty; }

  unsigned int getDimension() const { return mDim; }
  size_t numVertices() const { return mOut.size(); }
  // boost::num_edges(g): edges of vertices that pruneVertices cleared (BatchingManager.hpp:163) do not count
  size_t numEdges() const { return mEdges.size() - mNumRemovedEdges; }
  const double* vertexState(Vertex v) const { return &mVState[(size_t)v * mDim]; }
  /// The device context; null until the first call that needed the GPU (setProblemDefinition / solve).
  bpomp_ctx* context() const { return mCtx; }
  /// Kernel launches issued on behalf of this planner instance.
  uint64_t gpuLaunches() const { return mCtx ? bpomp_launch_count(mCtx) - mLaunchBase : 0; }

private:
  struct OutEdge { Vertex target; Edge eid; };

  // ---- graph (adjacency_list<vecS,vecS,undirectedS> semantics, flat storage) ----
  Vertex addVertex(const double* state, bool inNN);
  Edge addEdge(Vertex u, Vertex v, bool index = true);
  bool hasEdge(Vertex u, Vertex v);
  void completeEdgeIndex();
  Edge findEdge(Vertex u, Vertex v) const;
  void clearVertex(Vertex v);

  // ---- device helpers ----
  void check(int rc) const;
  void flushVertices();

  // ---- reference private helpers ----
  double radiusFun(unsigned int n) const;
  void nextBatch();
  void updateWithNewSolutionCost(double c);
  void pruneVertices(bool withValidity);
  void expandVertex(Vertex u, double radius);
  void refreshOutEdgeProbabilities(Vertex u);
  static void setProbFree(EProps& e, double p);
  void computeFreeProbabilities(const std::vector<Edge>& edges);
  void computeCreatedEdgeProbabilities(Vertex u, const std::vector<Edge>& created);
  bool checkAndUpdatePathBlocked(const std::vector<Edge>& ePath);
  void updateAffectedEdgeWeights();
  std::vector<Edge> selectEdges(const std::vector<Edge>& ePath) const;
  double edgeWeight(Edge e) const;
  bool astar(bool zeroHeuristic, bool single);
  const std::vector<std::pair<double, Vertex>>& neighbourRow(Vertex u, double radius);
  void prepareNeighbourCache(double radius);

  // ---- space ----
  unsigned int mDim;
  double mL, mMaxExtent, mSpaceMeasure;
  bpomp_ctx* mCtx;
  bpomp_ctx* dev();                         // acquires the context on first use; throws without CUDA
  std::unique_ptr<SpaceInformation> mSpace; // world description (upload, and the context pool compares it on reuse)
  std::string mCtxKey;  // device contexts are recycled between planner instances (BatchingPOMP.cpp, CtxPool)

  // ---- parameters ----
  std::string mRoadmapName, mGraphType, mBatchingType, mSelectorType;
  double mCurrentAlpha, mIncrement, mStartGoalRadius, mPruneThreshold;
  bool mIsInitSearchBatch, mIsPathFound, mUsingSelector;
  double mBestPathCost;
  std::string mSelType;
  uint64_t mLaunchBase;

  // ---- roadmap file contents ----
  std::vector<double> mRoadmapOwned;
  const double* mRoadmap;
  std::vector<std::pair<uint32_t, uint32_t>> mRoadmapEdges;
  uint32_t mN;

  // ---- current roadmap g ----
  std::vector<double> mVState;                 // [nv][dim]
  std::vector<std::vector<OutEdge>> mOut;
  std::vector<EProps> mEdges;
  size_t mNumRemovedEdges = 0;
  std::vector<uint8_t> mNNActive;              // membership in mVertexNN
  util::FlatMap64<Edge> mEdgeIndex;  // (min,max) -> edge id, replaces boost::edge's linear scan
  uint32_t mFlushed;                           // vertices already resident in the device table
  std::vector<uint32_t> mPendingDeactivate;
  Vertex mStartVertex, mGoalVertex;
  bool mHaveProblem;
  Edge mHashedUpTo;                            // edges [0, mHashedUpTo) are in mEdgeIndex

  // ---- search state: flat arrays + epoch stamps instead of four std::map (BP.cpp:860-863) ----
  // One 32-byte record per vertex (distance, f-cost, predecessor, colour, heap position, epoch stamp): a
  // relaxation touches one cache line of search state instead of six arrays.
  struct SearchNode {
    double dist, cost;
    uint32_t stamp;
    Vertex pred;
    uint32_t heapIndex;
    uint8_t color;  // 0 white, 1 gray, 2 black
  };
  std::vector<SearchNode> mNode;
  std::vector<Edge> mCreatedScratch;  // edges created by the current expansion
  uint32_t mEpoch;
  size_t mSearchEdgesNeeded;  // created edges of the current search whose probFree had to be computed
  // The heap stores the key beside the vertex (always equal to mNode[v].cost: keys only change through
  // heapUp after a decrease), so sifting compares within the heap array instead of chasing mNode[...].cost.
  struct HeapEntry {
    double key;
    Vertex v;
  };
  std::vector<HeapEntry> mHeap;
  void touch(Vertex v);
  void heapPush(Vertex v);
  void heapUp(size_t index);
  void heapDown();
  void heapPop();

  // ---- neighbour cache for the current batch (densification result) ----
  bool mCsrValid;
  double mCsrRadius;
  std::vector<uint64_t> mCsrRowPtr;
  std::vector<uint32_t> mCsrCol;
  std::vector<double> mCsrDist;
  std::vector<Edge> mCsrEdge;                  // edge id per CSR slot (kNoEdge: not created / not looked up yet)
  bool mCsrOlderEdges;                         // edges existed before this CSR was built
  std::vector<std::pair<double, Vertex>> mRowScratch;
  uint64_t mCsrBudget;
  // speculative probFree values for the current search: (source<<32 | target) -> (probFree, lookups)
  util::FlatMap64<std::pair<double, uint32_t>> mSpecCache;
  std::vector<uint32_t> mSpecStamp;
  std::vector<uint32_t> mExpandedStamp;  // batch number in which the vertex was last expanded

  // ---- batching manager state ----
  unsigned int bmNumBatches;
  bool bmExhausted;
  double bmCurrRadius;
  unsigned int bmNumVerticesAdded, bmNextVertexTarget;
  double bmVertInfl, bmRadiusInfl, bmInitRadius, bmMaxRadius;
  uint32_t bmCurrVertex;
  bool bmEdgeMode;

  // ---- results / evaluation ----
  std::vector<Edge> mCurrBestPath;
  std::vector<Vertex> mSolutionVertices;
  size_t mNumSolutions;
  std::vector<Edge> mEdgesToUpdate;
  std::vector<uint8_t> mEdgeMarked;
  bool mBeliefEmpty;
  unsigned int mNumEdgeChecks, mNumCollChecks, mNumSearches, mNumLookups;
  double mLookupTime, mCollCheckTime, mSearchTime;
  double mDeviceTime, mSearchHostTime;
  // speculative edge evaluation: device results of edges not yet committed (0 none, 1 free, 2 blocked)
  std::vector<uint8_t> mSpecStatus;
  std::vector<int32_t> mSpecFirst;
  uint64_t mNumSpeculativeChecks;
  int mBeliefK;
  double mBeliefPrior, mBeliefPriorWeight;
};

}  // namespace core
}  // namespace batching_pomp
