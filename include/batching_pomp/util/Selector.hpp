// batching_pomp/util/Selector.hpp — the edge-selector handle of
// /root/reference/include/batching_pomp/util/Selector.hpp:49-146: constructed from one of the four type names,
// std::invalid_argument for anything else (Selector.hpp:74).  The ordering itself (normal / alternate / failfast /
// maxinf, Selector.hpp:84-146) is applied inside the planner core (host/BatchingPOMP.cpp: selectEdges), which
// reads the type from this handle.
#pragma once

#include <stdexcept>
#include <string>

namespace batching_pomp {
namespace util {

template <class Graph>
class Selector {
public:
  explicit Selector(const std::string& type) : mType(type) {
    if (type != "normal" && type != "alternate" && type != "failfast" && type != "maxinf")
      throw std::invalid_argument("Invalid selector type specified - " + type + "!");
  }
  const std::string& getType() const { return mType; }

private:
  std::string mType;
};

}  // namespace util
}  // namespace batching_pomp
