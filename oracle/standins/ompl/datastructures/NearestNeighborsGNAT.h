// OMPL API stand-in (oracle/standins/README.md).  The real GNAT is an exact structure: its results are the
// linear scan's, only the order among equal distances differs (element address / traversal order).
#pragma once
#include "ompl/datastructures/NearestNeighborsLinear.h"
namespace ompl {
template <typename _T>
class NearestNeighborsGNAT : public NearestNeighborsLinear<_T> {
public:
  NearestNeighborsGNAT(unsigned int = 8, unsigned int = 4, unsigned int = 12, unsigned int = 50, unsigned int = 50, bool = false) {}
};
}  // namespace ompl
