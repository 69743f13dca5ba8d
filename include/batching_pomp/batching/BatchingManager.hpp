// batching_pomp/batching/BatchingManager.hpp — the batching manager seen through the planner's public member
// mBatchingPtr (/root/reference/include/batching_pomp/BatchingPOMP.hpp:119; interface
// batching/BatchingManager.hpp:114-132).  In this build the four managers (vertex / edge / hybrid / single,
// batching/*.hpp) are state machines inside the planner core, because their nextBatch() / pruneVertices()
// drive device calls (admission, densification); this class is the read-only view of that state with the
// reference's getter names.  As in the reference, setup() creates the manager named by the `batching_type`
// parameter (BP.cpp:689-742) and replaces whatever the 8-argument constructor was given.
#pragma once

#include <functional>
#include <string>

namespace batching_pomp {
namespace batching {

template <class Graph, class VStateMap, class StateCon, class EDistance>
class BatchingManager {
public:
  struct Source {  // filled in by the planner
    std::function<unsigned int()> numBatches, numVertices;
    std::function<bool()> exhausted;
    std::function<double()> radius;
    std::function<std::string()> type;
  };
  BatchingManager() = default;
  explicit BatchingManager(const Source& s) : mSource(s) {}
  virtual ~BatchingManager() = default;

  unsigned int getNumBatches() const { return mSource.numBatches ? mSource.numBatches() : 0u; }   // :114
  unsigned int getNumVertices() const { return mSource.numVertices ? mSource.numVertices() : 0u; } // :119
  bool isExhausted() const { return mSource.exhausted ? mSource.exhausted() : false; }             // :124
  double getCurrentRadius() const { return mSource.radius ? mSource.radius() : 0.0; }              // :129
  /// "vertex", "edge", "hybrid" or "single" once setup() has run.
  std::string getType() const { return mSource.type ? mSource.type() : std::string(); }

private:
  Source mSource;
};

}  // namespace batching
}  // namespace batching_pomp
