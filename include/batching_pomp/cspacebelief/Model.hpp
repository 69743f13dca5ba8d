// batching_pomp/cspacebelief/Model.hpp — the abstract C-space belief model interface of
// /root/reference/include/batching_pomp/cspacebelief/Model.hpp:33-62 (addPoint / removePoint / estimate),
// kept so that code written against the reference's `mBeliefModel` member and 8-argument constructor
// (BatchingPOMP.hpp:122, 147-154) compiles against the B200 build.
#pragma once

namespace batching_pomp {
namespace cspacebelief {

template <class T>
class Model {
public:
  virtual ~Model() = default;
  /// Add the result of a collision check to the model.
  virtual void addPoint(const T& data) = 0;
  /// Remove a point (the reference calls this "rarely used"; the planner never does).
  virtual void removePoint(const T& data) = 0;
  /// Probability of collision of the query configuration, in (0, 1).
  virtual double estimate(const T& query) const = 0;
};

}  // namespace cspacebelief
}  // namespace batching_pomp
