// batching_pomp/cspacebelief/BeliefPoint.hpp — a checked configuration and its collision result
// (/root/reference/include/batching_pomp/cspacebelief/BeliefPoint.hpp:43-71).  stateValues is an
// Eigen::VectorXd where Eigen is installed (as in the reference); otherwise a std::vector<double>, which
// offers the members this build needs (size(), operator[], data(), ==).
#pragma once

#include <ompl/base/State.h>
#include <ompl/base/spaces/RealVectorStateSpace.h>

#include <vector>
#if __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
#define BPOMP_HAVE_EIGEN 1
#endif

namespace batching_pomp {
namespace cspacebelief {

struct BeliefPoint {
#ifdef BPOMP_HAVE_EIGEN
  typedef Eigen::VectorXd StateVector;
#else
  typedef std::vector<double> StateVector;
#endif
  StateVector stateValues;
  double value;

  BeliefPoint() : value(0.0) {}
  BeliefPoint(const ompl::base::State* state, unsigned int dims, double val) : value(val) {
    const double* v = state->as<ompl::base::RealVectorStateSpace::StateType>()->values;
    stateValues = StateVector(dims);
    for (unsigned int j = 0; j < dims; ++j) stateValues[j] = v[j];
  }
  double getValue() { return value; }
  bool operator==(const BeliefPoint& o) const { return same(o); }
  bool operator!=(const BeliefPoint& o) const { return !same(o); }

private:
  bool same(const BeliefPoint& o) const {
    if (stateValues.size() != o.stateValues.size()) return false;
    for (long j = 0; j < (long)stateValues.size(); ++j)
      if (stateValues[j] != o.stateValues[j]) return false;
    return true;
  }
};

}  // namespace cspacebelief
}  // namespace batching_pomp
