// batching_pomp/BatchingPOMP.hpp — the drop-in planner: class batching_pomp::BatchingPOMP, an
// ompl::base::Planner with the public interface of
//   /root/reference/include/batching_pomp/BatchingPOMP.hpp:68-280
// on top of the B200 planner core (batching_pomp_b200/host/BatchingPOMP.hpp, namespace batching_pomp::core) and
// the CUDA C ABI (include/bpomp_cuda.h).  Same include path, namespace and class name as the reference, so a
// caller keeps `#include <batching_pomp/BatchingPOMP.hpp>` and `batching_pomp::BatchingPOMP planner(si);`.
//
//   * both constructors (BatchingPOMP.hpp:133, :147-154).  As in the reference, only the 1-argument constructor
//     declares the seven ParamSet entries (BP.cpp:313-319), and setup() creates the batching manager and the
//     selector from `batching_type` / `selector_type` (BP.cpp:689-753), replacing what the 8-argument
//     constructor was given; the belief model (k, prior, prior weight), roadmap file, start/goal radius,
//     increment and prune threshold of the 8-argument constructor are kept;
//   * the public members mSpace, mBatchingPtr, mBeliefModel, g, full_g (:116-131) as views of the core's state
//     (GraphView: num_vertices / num_edges / g[v] / g[e]; BatchingManager getters; KNNModel bound to the device
//     model), the public methods vertexDistFun / computeAndSetEdgeFreeProbability / checkAndSetEdgeBlocked
//     (:213-233) on edge ids;
//   * setup / setProblemDefinition / solve(ptc) with the reference's exceptions and PlannerStatus values,
//     solutions through pdef_->addSolutionPath (BP.cpp:1006), the counter getters (:188-200).
// One requirement is added: the StateValidityChecker must be able to describe itself as one
// of the two device-resident collision worlds (DeviceWorldValidityChecker below) — the GPU cannot call a host
// isValid(), and there is no CPU fallback.
//
// OMPL is not installed in this image: the CPU test-suite compiles this header against stand-ins of the OMPL
// API (tests/stubs, tests/cpp/ompl_adapter_check.cpp); it has not been built against a real OMPL tree.
#pragma once
#if __has_include(<ompl/base/Planner.h>)

#include <ompl/base/Planner.h>
#include <ompl/base/ScopedState.h>
#include <ompl/base/goals/GoalState.h>
#include <ompl/base/spaces/RealVectorStateSpace.h>
#include <ompl/geometric/PathGeometric.h>
#include <ompl/util/Exception.h>

#include <cmath>
#include <memory>
#include <string>
#include <vector>

#include "../../batching_pomp_b200/host/BatchingPOMP.hpp"
#include "batching_pomp/batching/BatchingManager.hpp"
#include "batching_pomp/cspacebelief/KNNModel.hpp"
#include "batching_pomp/util/Selector.hpp"

namespace batching_pomp {

using cspacebelief::BeliefPoint;

/// A StateValidityChecker that can be uploaded to the device.  isValid() stays usable on the host
/// (OMPL calls it for path validation); the planner itself only uses the device world.
class DeviceWorldValidityChecker : public ompl::base::StateValidityChecker {
public:
  using ompl::base::StateValidityChecker::StateValidityChecker;
  /// Fill `si.world` and the box / image members.
  virtual void describe(core::SpaceInformation& si) const = 0;
};

/// Read-only stand-in for the reference's Boost graphs `g` and `full_g` (BatchingPOMP.hpp:106, :126-131).
class GraphView {
public:
  typedef uint32_t vertex_descriptor;
  typedef uint32_t edge_descriptor;
  /// VProps (BatchingPOMP.hpp:83-88): the vertex's state values (`dim` doubles).
  struct VProps {
    const double* v_state;
    unsigned int dim;
  };
  /// EProps (BatchingPOMP.hpp:90-103) without the materialised edgeStates.
  struct EProps {
    vertex_descriptor source, target;
    double distance, probFree;
    int blockedStatus;
    std::size_t numStates;
  };
  GraphView(const core::BatchingPOMP* core, bool full) : mCore(core), mFull(full) {}
  std::size_t numVertices() const { return mFull ? mCore->roadmapSize() : mCore->numVertices(); }
  std::size_t numEdges() const { return mFull ? mCore->roadmapEdgeCount() : mCore->numEdges(); }
  VProps operator[](vertex_descriptor v) const {
    return VProps{mFull ? mCore->roadmapState(v) : mCore->vertexState(v), mCore->getDimension()};
  }
  /// Edge ids of `g` run over 0 .. edgeSlots()-1; ids whose edge was cleared by pruning report removed().
  std::size_t edgeSlots() const { return mFull ? 0 : mCore->edgeSlots(); }
  bool removed(edge_descriptor e) const { return mCore->edgeProps(e).removed; }
  EProps edge(edge_descriptor e) const {
    const core::BatchingPOMP::EProps& p = mCore->edgeProps(e);
    return EProps{p.source, p.target, p.distance, p.probFree, p.blockedStatus, (std::size_t)p.nStates};
  }

private:
  const core::BatchingPOMP* mCore;
  bool mFull;
};
inline std::size_t num_vertices(const GraphView& g) { return g.numVertices(); }
inline std::size_t num_edges(const GraphView& g) { return g.numEdges(); }

class BatchingPOMP : public ompl::base::Planner {
  // The planner core; declared first because the public graph views below are constructed from it.
  std::unique_ptr<core::BatchingPOMP> impl_;

public:
  static const int FREE{1};
  static const int BLOCKED{-1};
  static const int UNKNOWN{0};

  typedef GraphView Graph;
  typedef GraphView::vertex_descriptor Vertex;
  typedef GraphView::edge_descriptor Edge;
  typedef GraphView::VProps VProps;
  typedef GraphView::EProps EProps;
  struct StateCon;        // the reference's state container (RoadmapFromFile.hpp); states are plain arrays here
  struct VPStateMap {};   // property-map tags of the reference's template arguments
  struct EPDistanceMap {};
  typedef batching::BatchingManager<Graph, VPStateMap, StateCon, EPDistanceMap> BatchingManagerType;

  // ---- public members of the reference (BatchingPOMP.hpp:116-131) ----
  const ompl::base::StateSpacePtr mSpace;
  std::shared_ptr<BatchingManagerType> mBatchingPtr;
  std::shared_ptr<cspacebelief::Model<BeliefPoint>> mBeliefModel;
  const Graph g;
  const Graph full_g;

  /// BatchingPOMP.hpp:133, BP.cpp:270-321: default belief model KNNModel(15, 0.5, 0.25), the seven parameters.
  explicit BatchingPOMP(const ompl::base::SpaceInformationPtr& si)
      : ompl::base::Planner(si, "BatchingPOMP"), impl_(makeCore(si)), mSpace(si->getStateSpace()),
        mBeliefModel(std::make_shared<cspacebelief::KNNModel>(15, 0.5, 0.25)), g(impl_.get(), false),
        full_g(impl_.get(), true) {
    attachBatchingView();
    Planner::declareParam<double>("increment", this, &BatchingPOMP::setIncrement, &BatchingPOMP::getIncrement);
    Planner::declareParam<double>("start_goal_radius", this, &BatchingPOMP::setStartGoalRadius, &BatchingPOMP::getStartGoalRadius);
    Planner::declareParam<double>("prune_threshold", this, &BatchingPOMP::setPruneThreshold, &BatchingPOMP::getPruneThreshold);
    Planner::declareParam<std::string>("graph_type", this, &BatchingPOMP::setGraphType, &BatchingPOMP::getGraphType);
    Planner::declareParam<std::string>("batching_type", this, &BatchingPOMP::setBatchingType, &BatchingPOMP::getBatchingType);
    Planner::declareParam<std::string>("selector_type", this, &BatchingPOMP::setSelectorType, &BatchingPOMP::getSelectorType);
    Planner::declareParam<std::string>("roadmap_filename", this, &BatchingPOMP::setRoadmapFileName, &BatchingPOMP::getRoadmapFileName);
  }

  /// BatchingPOMP.hpp:147-154, BP.cpp:221-267.  _batchingPtr and _selector are superseded by setup(), exactly as
  /// in the reference (BP.cpp:689-753 assigns both unconditionally); _beliefModel must be a cspacebelief::KNNModel
  /// (its k, prior and prior weight configure the device model).  No ParamSet entries are declared (BP.cpp:221-267).
  BatchingPOMP(const ompl::base::SpaceInformationPtr& si, std::shared_ptr<BatchingManagerType> _batchingPtr,
               std::shared_ptr<cspacebelief::Model<BeliefPoint>> _beliefModel,
               std::unique_ptr<util::Selector<Graph>> _selector, const std::string& _roadmapFileName,
               double _startGoalRadius, double _increment = 0.2, double _pruneThreshold = 0.05)
      : ompl::base::Planner(si, "BatchingPOMP"), impl_(makeCore(si)), mSpace(si->getStateSpace()),
        mBatchingPtr(std::move(_batchingPtr)), mBeliefModel(std::move(_beliefModel)), g(impl_.get(), false),
        full_g(impl_.get(), true), mSelector(std::move(_selector)) {
    auto* knn = dynamic_cast<cspacebelief::KNNModel*>(mBeliefModel.get());
    if (!knn) throw ompl::Exception("BatchingPOMP (B200): the belief model must be a cspacebelief::KNNModel");
    impl_->setBeliefModel((int)knn->getK(), knn->getPrior(), knn->getPriorWeight());
    impl_->setRoadmapFileName(_roadmapFileName);
    impl_->setStartGoalRadius(_startGoalRadius);
    impl_->setIncrement(_increment);
    impl_->setPruneThreshold(_pruneThreshold);
  }

  ~BatchingPOMP() override = default;

  // ---- setters and getters, BatchingPOMP.hpp:158-200 ----
  double getCurrentAlpha() const { return impl_->getCurrentAlpha(); }
  double getCurrentBestCost() const { return impl_->getCurrentBestCost(); }
  Vertex getStartVertex() const { return impl_->getStartVertex(); }
  Vertex getGoalVertex() const { return impl_->getGoalVertex(); }
  bool isInitSearchBatch() const { return impl_->isInitSearchBatch(); }
  double getIncrement() const { return impl_->getIncrement(); }
  void setIncrement(double v) { impl_->setIncrement(v); }
  double getStartGoalRadius() const { return impl_->getStartGoalRadius(); }
  void setStartGoalRadius(double v) { impl_->setStartGoalRadius(v); }
  double getPruneThreshold() const { return impl_->getPruneThreshold(); }
  void setPruneThreshold(double v) { impl_->setPruneThreshold(v); }
  std::string getGraphType() const { return impl_->getGraphType(); }
  void setGraphType(const std::string& v) { impl_->setGraphType(v); }
  std::string getBatchingType() const { return impl_->getBatchingType(); }
  void setBatchingType(const std::string& v) { impl_->setBatchingType(v); }
  std::string getSelectorType() const { return impl_->getSelectorType(); }
  void setSelectorType(const std::string& v) { impl_->setSelectorType(v); }
  std::string getRoadmapFileName() const { return impl_->getRoadmapFileName(); }
  void setRoadmapFileName(const std::string& v) { impl_->setRoadmapFileName(v); }
  unsigned int getNumEdgeChecks() { return impl_->getNumEdgeChecks(); }
  unsigned int getNumCollChecks() { return impl_->getNumCollChecks(); }
  unsigned int getNumSearches() { return impl_->getNumSearches(); }
  unsigned int getNumLookups() { return impl_->getNumLookups(); }
  double getLookupTime() { return impl_->getLookupTime(); }
  double getSearchTime() { return impl_->getSearchTime(); }
  double getCollCheckTime() { return impl_->getCollCheckTime(); }

  // ---- OMPL required methods, BatchingPOMP.hpp:203-205 ----
  void setup() override {
    ompl::base::Planner::setup();
    try {
      impl_->setup();
    } catch (const core::Exception& e) {
      throw ompl::Exception(e.what());
    }
    // BP.cpp:747-753: the selector is (re)created from selector_type, "normal" if none
    mSelector.reset(new util::Selector<Graph>(impl_->getSelectorType().empty() ? "normal" : impl_->getSelectorType()));
    attachBatchingView();  // BP.cpp:689-742: the manager named by batching_type replaces any earlier one
  }

  void setProblemDefinition(const ompl::base::ProblemDefinitionPtr& pdef) override {
    ompl::base::Planner::setProblemDefinition(pdef);
    auto* goal = pdef->getGoal()->as<ompl::base::GoalState>();
    try {
      impl_->setProblemDefinition(values(pdef->getStartState(0)), values(goal->getState()));
    } catch (const core::Exception& e) {
      throw ompl::Exception(e.what());
    }
    bindBeliefModel();
  }

  ompl::base::PlannerStatus solve(const ompl::base::PlannerTerminationCondition& ptc) override {
    std::size_t before = impl_->getNumSolutions();
    core::PlannerStatus st;
    try {
      st = impl_->solve([&ptc]() { return ptc.eval(); });
    } catch (const core::Exception& e) {
      throw ompl::Exception(e.what());
    }
    bindBeliefModel();
    if (impl_->getNumSolutions() > before) {  // BP.cpp:996-1006
      auto* path = new ompl::geometric::PathGeometric(si_);
      std::vector<double> states = impl_->getSolutionStates();
      const unsigned d = si_->getStateDimension();
      ompl::base::ScopedState<> s(si_->getStateSpace());
      for (std::size_t i = 0; i * d < states.size(); ++i) {
        for (unsigned j = 0; j < d; ++j) s[j] = states[i * d + j];
        path->append(s.get());
      }
      pdef_->addSolutionPath(ompl::base::PathPtr(path));
    }
    return status(st);
  }

  // ---- public helpers of the reference, BatchingPOMP.hpp:207-233, on this build's vertex / edge ids ----
  /// RealVectorStateSpace::distance between the states of two vertices of g (BP.cpp:429-432).
  double vertexDistFun(const Vertex& u, const Vertex& v) const {
    const double *a = impl_->vertexState(u), *b = impl_->vertexState(v);
    double s = 0.0;
    for (unsigned j = 0; j < impl_->getDimension(); ++j) {
      const double diff = a[j] - b[j];
      s += diff * diff;
    }
    return std::sqrt(s);
  }
  /// The collision measure of edge e under the current belief model (BP.cpp:464-493), from the device; unlike
  /// the reference's method this does not touch the lookup counters or the stored edge weight.
  double computeAndSetEdgeFreeProbability(const Edge& e) {
    const EProps p = g.edge(e);
    if (impl_->beliefEmpty() || !impl_->context()) {  // KNNModel.hpp:104-106: the prior for every query
      const std::size_t n = p.numStates;
      const unsigned int step = static_cast<unsigned int>(n / std::log(n));
      double r = 1.0;
      for (unsigned int i = 1; i < n - 1; i += step) r *= (1.0 - impl_->getBeliefPrior());
      return r;
    }
    double pf = 0.0;
    const uint32_t f = p.source, t = p.target;
    if (bpomp_belief_edge_pfree_indexed(impl_->context(), &f, &t, 1, &pf, nullptr) != BPOMP_OK)
      throw ompl::Exception(bpomp_last_error(impl_->context()));
    return pf;
  }
  /// Collision status of edge e evaluated on the device (BP.cpp:496-544) without committing it to the roadmap
  /// or the belief model: true if blocked.
  bool checkAndSetEdgeBlocked(const Edge& e) {
    const EProps p = g.edge(e);
    const uint32_t f = p.source, t = p.target;
    uint8_t valid = 0;
    int32_t first = 0;
    uint32_t ns = 0;
    bpomp_ctx* ctx = impl_->context();
    if (!ctx) throw ompl::Exception("BatchingPOMP (B200): no device context yet (call setProblemDefinition first)");
    if (bpomp_edges_check_indexed(ctx, &f, &t, 1, &valid, &first, &ns) != BPOMP_OK)
      throw ompl::Exception(bpomp_last_error(ctx));
    return valid == 0;
  }

  /// The selector in use (type name), after setup().
  const util::Selector<Graph>* getSelector() const { return mSelector.get(); }
  /// Kernel launches issued on behalf of this planner, and the core itself (B200 additions).
  uint64_t gpuLaunches() const { return impl_->gpuLaunches(); }
  core::BatchingPOMP& core() { return *impl_; }

private:
  // Builds the core from the SpaceInformation.
  static std::unique_ptr<core::BatchingPOMP> makeCore(const ompl::base::SpaceInformationPtr& si) {
    if (si->getStateSpace()->getType() != ompl::base::STATE_SPACE_REAL_VECTOR)
      throw ompl::Exception("This only supports real vector state spaces!");  // RoadmapFromFile.hpp:127
    auto* rv = si->getStateSpace()->as<ompl::base::RealVectorStateSpace>();
    auto* checker = dynamic_cast<const DeviceWorldValidityChecker*>(si->getStateValidityChecker().get());
    if (!checker) throw ompl::Exception("BatchingPOMP (B200): the StateValidityChecker must be a DeviceWorldValidityChecker");
    core::SpaceInformation s;
    s.dim = rv->getDimension();
    s.low = rv->getBounds().low;
    s.high = rv->getBounds().high;
    s.longestValidSegmentFraction = rv->getLongestValidSegmentFraction();
    checker->describe(s);
    try {
      return std::unique_ptr<core::BatchingPOMP>(new core::BatchingPOMP(s));
    } catch (const core::Exception& e) {
      throw ompl::Exception(e.what());
    }
  }
  void attachBatchingView() {
    core::BatchingPOMP* c = impl_.get();
    BatchingManagerType::Source src;
    src.numBatches = [c]() { return c->getNumBatches(); };
    src.numVertices = [c]() { return (unsigned int)c->roadmapSize(); };
    src.exhausted = [c]() { return c->isExhausted(); };
    src.radius = [c]() { return c->getCurrentRadius(); };
    src.type = [c]() { return c->getBatchingType(); };
    mBatchingPtr = std::make_shared<BatchingManagerType>(src);
  }
  void bindBeliefModel() {
    auto* knn = dynamic_cast<cspacebelief::KNNModel*>(mBeliefModel.get());
    if (knn && !knn->bound() && impl_->context()) knn->bind(impl_->context(), impl_->getDimension());
  }
  std::vector<double> values(const ompl::base::State* s) const {
    const double* v = s->as<ompl::base::RealVectorStateSpace::StateType>()->values;
    return std::vector<double>(v, v + si_->getStateDimension());
  }
  static ompl::base::PlannerStatus status(core::PlannerStatus st) {  // by name, not by enum value
    typedef ompl::base::PlannerStatus PS;
    switch (st) {
      case core::PlannerStatus::INVALID_START: return PS(PS::INVALID_START);
      case core::PlannerStatus::INVALID_GOAL: return PS(PS::INVALID_GOAL);
      case core::PlannerStatus::UNRECOGNIZED_GOAL_TYPE: return PS(PS::UNRECOGNIZED_GOAL_TYPE);
      case core::PlannerStatus::TIMEOUT: return PS(PS::TIMEOUT);
      case core::PlannerStatus::APPROXIMATE_SOLUTION: return PS(PS::APPROXIMATE_SOLUTION);
      case core::PlannerStatus::EXACT_SOLUTION: return PS(PS::EXACT_SOLUTION);
      case core::PlannerStatus::CRASH: return PS(PS::CRASH);
      case core::PlannerStatus::ABORT: return PS(PS::ABORT);
      default: return PS(PS::UNKNOWN);
    }
  }

  std::unique_ptr<util::Selector<Graph>> mSelector;
};

}  // namespace batching_pomp
#endif
