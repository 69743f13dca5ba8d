// OMPL API stand-in (tests/stubs/README.md)
#pragma once
#include "ompl/base/SpaceInformation.h"
#include "ompl/base/spaces/RealVectorStateSpace.h"
namespace ompl {
namespace base {
template <class T = RealVectorStateSpace>
class ScopedState {
public:
  explicit ScopedState(const StateSpacePtr& space) : space_(space), state_(space->allocState()) {}
  explicit ScopedState(const SpaceInformationPtr& si) : ScopedState(si->getStateSpace()) {}
  ScopedState(const ScopedState&) = delete;
  ScopedState& operator=(const ScopedState&) = delete;
  ~ScopedState() { space_->freeState(state_); }
  double& operator[](unsigned int i) { return static_cast<typename T::StateType*>(state_)->values[i]; }
  double operator[](unsigned int i) const { return static_cast<const typename T::StateType*>(state_)->values[i]; }
  State* get() { return state_; }
  const State* get() const { return state_; }
private:
  StateSpacePtr space_;
  State* state_;
};
}  // namespace base
}  // namespace ompl
