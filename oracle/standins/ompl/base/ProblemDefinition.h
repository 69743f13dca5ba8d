// OMPL API stand-in (oracle/standins/README.md)
#pragma once
#include <vector>
#include "ompl/base/Path.h"
#include "ompl/base/goals/GoalState.h"
namespace ompl {
namespace base {
class ProblemDefinition {
public:
  explicit ProblemDefinition(const SpaceInformationPtr& si) : si_(si) {}
  ~ProblemDefinition() { for (auto* s : starts_) si_->freeState(s); }
  void addStartState(const State* s) {
    State* c = si_->allocState();
    si_->getStateSpace()->copyState(c, s);
    starts_.push_back(c);
  }
  void setStartAndGoalStates(const State* start, const State* goal) {
    addStartState(start);
    auto g = std::make_shared<GoalState>(si_);
    g->setState(goal);
    goal_ = g;
  }
  unsigned int getStartStateCount() const { return (unsigned int)starts_.size(); }
  const State* getStartState(unsigned int i) const { return starts_[i]; }
  const GoalPtr& getGoal() const { return goal_; }
  void setGoal(const GoalPtr& g) { goal_ = g; }
  void addSolutionPath(const PathPtr& p) { solutions_.push_back(p); }
  std::size_t getSolutionCount() const { return solutions_.size(); }
  PathPtr getSolutionPath() const { return solutions_.empty() ? PathPtr() : solutions_.back(); }
private:
  SpaceInformationPtr si_;
  std::vector<State*> starts_;
  GoalPtr goal_;
  std::vector<PathPtr> solutions_;
};
typedef std::shared_ptr<ProblemDefinition> ProblemDefinitionPtr;
}  // namespace base
}  // namespace ompl
