// OMPL API stand-in (oracle/standins/README.md)
#pragma once
#include "ompl/base/StateSpace.h"
namespace ompl {
namespace base {
template <class T = StateSpace>
class ScopedState {
public:
  explicit ScopedState(const StateSpacePtr& space) : space_(space), state_(space->allocState()) {}
  ScopedState(const ScopedState& o) : space_(o.space_), state_(o.space_->allocState()) { space_->copyState(state_, o.state_); }
  ~ScopedState() { space_->freeState(state_); }
  ScopedState& operator=(const State* s) {
    space_->copyState(state_, s);
    return *this;
  }
  ScopedState& operator=(const ScopedState& o) {
    space_->copyState(state_, o.state_);
    return *this;
  }
  State* get() { return state_; }
  const State* get() const { return state_; }
  double& operator[](unsigned int i) { return static_cast<typename T::StateType*>(state_)->values[i]; }
private:
  StateSpacePtr space_;
  State* state_;
};
}  // namespace base
}  // namespace ompl
