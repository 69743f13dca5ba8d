// OMPL API stand-in (tests/stubs/README.md)
#pragma once
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
  const RealVectorBounds& getBounds() const { return bounds_; }
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
private:
  unsigned int dim_;
  RealVectorBounds bounds_;
};
}  // namespace base
}  // namespace ompl
