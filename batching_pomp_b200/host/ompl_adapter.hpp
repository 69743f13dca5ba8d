// ompl_adapter.hpp — drop-in ompl::base::Planner named batching_pomp::BatchingPOMP on top of the B200
// host planner.  Compiled only where OMPL is installed.  OMPL is NOT in this image: the CPU test-suite
// compiles this header against stand-ins of the OMPL API (tests/stubs) and drives it like an OMPL application
// (tests/cpp/ompl_adapter_check.cpp); it has not been built against a real OMPL tree.  It is the binding shown
// in INTEGRATION.md, kept next to the code it wraps.
//
// Keeps the outer contract of /root/reference/include/batching_pomp/BatchingPOMP.hpp:68-280:
//   BatchingPOMP(const ompl::base::SpaceInformationPtr&), the seven ParamSet entries
//   (/root/reference/src/BatchingPOMP.cpp:313-319), setup(), setProblemDefinition(), solve(ptc),
//   solutions through pdef_->addSolutionPath (BP.cpp:1006), the counter getters (BatchingPOMP.hpp:188-200).
// The StateValidityChecker must additionally describe itself as a device world.
#pragma once
#if __has_include(<ompl/base/Planner.h>)

#include <ompl/base/Planner.h>
#include <ompl/base/ScopedState.h>
#include <ompl/base/goals/GoalState.h>
#include <ompl/base/spaces/RealVectorStateSpace.h>
#include <ompl/geometric/PathGeometric.h>
#include <ompl/util/Exception.h>

#include "BatchingPOMP.hpp"

namespace batching_pomp {
namespace ompl_adapter {

/// A StateValidityChecker that can be uploaded to the device.  isValid() stays usable on the host
/// (OMPL calls it for path validation); the planner itself only uses the device world.
class DeviceWorldValidityChecker : public ompl::base::StateValidityChecker {
public:
  using ompl::base::StateValidityChecker::StateValidityChecker;
  /// Fill `si.world` and the box / image members.
  virtual void describe(batching_pomp::SpaceInformation& si) const = 0;
};

class BatchingPOMP : public ompl::base::Planner {
public:
  explicit BatchingPOMP(const ompl::base::SpaceInformationPtr& si) : ompl::base::Planner(si, "BatchingPOMP") {
    auto* rv = si->getStateSpace()->as<ompl::base::RealVectorStateSpace>();
    if (si->getStateSpace()->getType() != ompl::base::STATE_SPACE_REAL_VECTOR)
      throw ompl::Exception("This only supports real vector state spaces!");  // RoadmapFromFile.hpp:127
    auto* checker = dynamic_cast<const DeviceWorldValidityChecker*>(si->getStateValidityChecker().get());
    if (!checker) throw ompl::Exception("BatchingPOMP (B200): the StateValidityChecker must be a DeviceWorldValidityChecker");
    batching_pomp::SpaceInformation s;
    s.dim = rv->getDimension();
    s.low = rv->getBounds().low;
    s.high = rv->getBounds().high;
    s.longestValidSegmentFraction = rv->getLongestValidSegmentFraction();
    checker->describe(s);
    impl_.reset(new batching_pomp::BatchingPOMP(s));
    Planner::declareParam<double>("increment", this, &BatchingPOMP::setIncrement, &BatchingPOMP::getIncrement);
    Planner::declareParam<double>("start_goal_radius", this, &BatchingPOMP::setStartGoalRadius, &BatchingPOMP::getStartGoalRadius);
    Planner::declareParam<double>("prune_threshold", this, &BatchingPOMP::setPruneThreshold, &BatchingPOMP::getPruneThreshold);
    Planner::declareParam<std::string>("graph_type", this, &BatchingPOMP::setGraphType, &BatchingPOMP::getGraphType);
    Planner::declareParam<std::string>("batching_type", this, &BatchingPOMP::setBatchingType, &BatchingPOMP::getBatchingType);
    Planner::declareParam<std::string>("selector_type", this, &BatchingPOMP::setSelectorType, &BatchingPOMP::getSelectorType);
    Planner::declareParam<std::string>("roadmap_filename", this, &BatchingPOMP::setRoadmapFileName, &BatchingPOMP::getRoadmapFileName);
  }

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
  double getCurrentAlpha() const { return impl_->getCurrentAlpha(); }
  double getCurrentBestCost() const { return impl_->getCurrentBestCost(); }
  unsigned int getNumEdgeChecks() { return impl_->getNumEdgeChecks(); }
  unsigned int getNumCollChecks() { return impl_->getNumCollChecks(); }
  unsigned int getNumSearches() { return impl_->getNumSearches(); }
  unsigned int getNumLookups() { return impl_->getNumLookups(); }
  double getLookupTime() { return impl_->getLookupTime(); }
  double getSearchTime() { return impl_->getSearchTime(); }
  double getCollCheckTime() { return impl_->getCollCheckTime(); }

  void setup() override {
    ompl::base::Planner::setup();
    try { impl_->setup(); } catch (const batching_pomp::Exception& e) { throw ompl::Exception(e.what()); }
  }

  void setProblemDefinition(const ompl::base::ProblemDefinitionPtr& pdef) override {
    ompl::base::Planner::setProblemDefinition(pdef);
    const unsigned d = si_->getStateDimension();
    auto vals = [d](const ompl::base::State* s) {
      const double* v = s->as<ompl::base::RealVectorStateSpace::StateType>()->values;
      return std::vector<double>(v, v + d);
    };
    auto* goal = pdef->getGoal()->as<ompl::base::GoalState>();
    try { impl_->setProblemDefinition(vals(pdef->getStartState(0)), vals(goal->getState())); }
    catch (const batching_pomp::Exception& e) { throw ompl::Exception(e.what()); }
  }

  ompl::base::PlannerStatus solve(const ompl::base::PlannerTerminationCondition& ptc) override {
    size_t before = impl_->getNumSolutions();
    batching_pomp::PlannerStatus st = impl_->solve([&ptc]() { return ptc.eval(); });
    if (impl_->getNumSolutions() > before) {  // BP.cpp:996-1006
      auto* path = new ompl::geometric::PathGeometric(si_);
      std::vector<double> states = impl_->getSolutionStates();
      const unsigned d = si_->getStateDimension();
      ompl::base::ScopedState<> s(si_->getStateSpace());
      for (size_t i = 0; i * d < states.size(); ++i) {
        for (unsigned j = 0; j < d; ++j) s[j] = states[i * d + j];
        path->append(s.get());
      }
      pdef_->addSolutionPath(ompl::base::PathPtr(path));
    }
    return ompl::base::PlannerStatus(static_cast<ompl::base::PlannerStatus::StatusType>((int)st));
  }

private:
  std::unique_ptr<batching_pomp::BatchingPOMP> impl_;
};

}  // namespace ompl_adapter
}  // namespace batching_pomp
#endif
