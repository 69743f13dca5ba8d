// OMPL API stand-in (oracle/standins/README.md).  The arithmetic is the published RealVectorStateSpace
// arithmetic, in the same operation order.
#pragma once
#include <cmath>
#include <vector>
#include "ompl/base/StateSpace.h"
namespace ompl {
namespace base {
class RealVectorBounds {
public:
  explicit RealVectorBounds(unsigned int dim) : low(dim, 0.0), high(dim, 0.0) {}
  void setLow(double v) { low.assign(low.size(), v); }
  void setHigh(double v) { high.assign(high.size(), v); }
  std::vector<double> low, high;
};
class RealVectorStateSpace : public StateSpace {
public:
  class StateType : public State {
  public:
    double operator[](unsigned int i) const { return values[i]; }
    double& operator[](unsigned int i) { return values[i]; }
    double* values = nullptr;
  };
  explicit RealVectorStateSpace(unsigned int dim = 0) : dim_(dim), bounds_(dim) { type_ = STATE_SPACE_REAL_VECTOR; }
  unsigned int getDimension() const override { return dim_; }
  void setBounds(const RealVectorBounds& b) { bounds_ = b; }
  void setBounds(double lo, double hi) { bounds_.setLow(lo); bounds_.setHigh(hi); }
  const RealVectorBounds& getBounds() const { return bounds_; }
  double getMaximumExtent() const override {
    double e = 0.0;
    for (unsigned int i = 0; i < dim_; ++i) {
      double d = bounds_.high[i] - bounds_.low[i];
      e += d * d;
    }
    return std::sqrt(e);
  }
  double getMeasure() const override {
    double m = 1.0;
    for (unsigned int i = 0; i < dim_; ++i) m *= bounds_.high[i] - bounds_.low[i];
    return m;
  }
  State* allocState() const override {
    auto* s = new StateType();
    s->values = new double[dim_]();
    return s;
  }
  void freeState(State* s) const override {
    auto* r = static_cast<StateType*>(s);
    delete[] r->values;
    delete r;
  }
  void copyState(State* dst, const State* src) const override {
    for (unsigned int i = 0; i < dim_; ++i)
      static_cast<StateType*>(dst)->values[i] = static_cast<const StateType*>(src)->values[i];
  }
  double distance(const State* a, const State* b) const override {
    double dist = 0.0;
    const double* s1 = static_cast<const StateType*>(a)->values;
    const double* s2 = static_cast<const StateType*>(b)->values;
    for (unsigned int i = 0; i < dim_; ++i) {
      double diff = (*s1++) - (*s2++);
      dist += diff * diff;
    }
    return std::sqrt(dist);
  }
  void interpolate(const State* from, const State* to, double t, State* out) const override {
    const StateType* rfrom = static_cast<const StateType*>(from);
    const StateType* rto = static_cast<const StateType*>(to);
    StateType* rstate = static_cast<StateType*>(out);
    for (unsigned int i = 0; i < dim_; ++i) rstate->values[i] = rfrom->values[i] + (rto->values[i] - rfrom->values[i]) * t;
  }
private:
  unsigned int dim_;
  RealVectorBounds bounds_;
};
}  // namespace base
}  // namespace ompl
