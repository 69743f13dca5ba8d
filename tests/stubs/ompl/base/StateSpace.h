// OMPL API stand-in (tests/stubs/README.md)
#pragma once
#include <memory>
#include "ompl/base/State.h"
namespace ompl {
namespace base {
enum StateSpaceType { STATE_SPACE_UNKNOWN = 0, STATE_SPACE_REAL_VECTOR = 1 };
class StateSpace {
public:
  virtual ~StateSpace() = default;
  template <class T> T* as() { return static_cast<T*>(this); }
  template <class T> const T* as() const { return static_cast<const T*>(this); }
  int getType() const { return type_; }
  virtual unsigned int getDimension() const = 0;
  virtual State* allocState() const = 0;
  virtual void freeState(State* s) const = 0;
  virtual void copyState(State* dst, const State* src) const = 0;
  double getLongestValidSegmentFraction() const { return fraction_; }
  void setLongestValidSegmentFraction(double f) { fraction_ = f; }
protected:
  int type_ = STATE_SPACE_UNKNOWN;
  double fraction_ = 0.01;
};
typedef std::shared_ptr<StateSpace> StateSpacePtr;
}  // namespace base
}  // namespace ompl
