// OMPL API stand-in (oracle/standins/README.md)
#pragma once
#include <memory>
#include "ompl/base/StateSpace.h"
namespace ompl {
namespace base {
class SpaceInformation;
typedef std::shared_ptr<SpaceInformation> SpaceInformationPtr;
class StateValidityChecker {
public:
  explicit StateValidityChecker(SpaceInformation* si) : si_(si) {}
  explicit StateValidityChecker(const SpaceInformationPtr& si) : si_(si.get()) {}
  virtual ~StateValidityChecker() = default;
  virtual bool isValid(const State* state) const = 0;
protected:
  SpaceInformation* si_;
};
typedef std::shared_ptr<StateValidityChecker> StateValidityCheckerPtr;
class SpaceInformation {
public:
  explicit SpaceInformation(StateSpacePtr space) : space_(std::move(space)) {}
  const StateSpacePtr& getStateSpace() const { return space_; }
  unsigned int getStateDimension() const { return space_->getDimension(); }
  double getSpaceMeasure() const { return space_->getMeasure(); }
  void setStateValidityChecker(const StateValidityCheckerPtr& c) { checker_ = c; }
  const StateValidityCheckerPtr& getStateValidityChecker() const { return checker_; }
  void setStateValidityCheckingResolution(double f) { space_->setLongestValidSegmentFraction(f); }
  bool isValid(const State* s) const { return checker_->isValid(s); }
  State* allocState() const { return space_->allocState(); }
  void freeState(State* s) const { space_->freeState(s); }
  void setup() { space_->setup(); }
private:
  StateSpacePtr space_;
  StateValidityCheckerPtr checker_;
};
}  // namespace base
}  // namespace ompl
