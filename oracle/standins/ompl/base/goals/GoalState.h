// OMPL API stand-in (oracle/standins/README.md)
#pragma once
#include <memory>
#include "ompl/base/SpaceInformation.h"
namespace ompl {
namespace base {
class Goal {
public:
  explicit Goal(const SpaceInformationPtr& si) : si_(si) {}
  virtual ~Goal() = default;
  template <class T> T* as() { return static_cast<T*>(this); }
  template <class T> const T* as() const { return static_cast<const T*>(this); }
protected:
  SpaceInformationPtr si_;
};
typedef std::shared_ptr<Goal> GoalPtr;
class GoalState : public Goal {
public:
  explicit GoalState(const SpaceInformationPtr& si) : Goal(si), state_(nullptr) {}
  ~GoalState() override { if (state_) si_->freeState(state_); }
  void setState(const State* s) {
    if (!state_) state_ = si_->allocState();
    si_->getStateSpace()->copyState(state_, s);
  }
  const State* getState() const { return state_; }
  State* getState() { return state_; }
private:
  State* state_;
};
}  // namespace base
}  // namespace ompl
