// OMPL API stand-in (oracle/standins/README.md)
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
  virtual double getMaximumExtent() const = 0;
  virtual double getMeasure() const = 0;
  virtual State* allocState() const = 0;
  virtual void freeState(State* s) const = 0;
  virtual void copyState(State* dst, const State* src) const = 0;
  virtual double distance(const State* a, const State* b) const = 0;
  virtual void interpolate(const State* from, const State* to, double t, State* out) const = 0;
  // StateSpace::setup(): maxExtent_ = getMaximumExtent(); longestValidSegment_ = maxExtent_ * fraction
  virtual void setup() {
    maxExtent_ = getMaximumExtent();
    longestValidSegment_ = maxExtent_ * fraction_;
  }
  double getLongestValidSegmentFraction() const { return fraction_; }
  void setLongestValidSegmentFraction(double f) { fraction_ = f; }
  double getLongestValidSegmentLength() const { return longestValidSegment_; }  // 0 before setup()
protected:
  int type_ = STATE_SPACE_UNKNOWN;
  double fraction_ = 0.01;
  double maxExtent_ = 0.0;
  double longestValidSegment_ = 0.0;
};
typedef std::shared_ptr<StateSpace> StateSpacePtr;
}  // namespace base
}  // namespace ompl
