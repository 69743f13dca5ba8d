// ref_shim.cpp — TEST INFRASTRUCTURE ONLY.
//
// The one piece of the reference that compiles without OMPL / Boost / Eigen is
// include/batching_pomp/util/BisectPerm.hpp (the sample order of every edge: BatchingPOMP.cpp:447-460
// takes perm[i].first from it).  This shim includes that header FROM WHERE IT LIES under
// /root/reference (oracle/Makefile passes -I$(REF)/include; nothing is copied into this repository)
// and exports its result through a C function, so that the oracle's restatement (orc_bisect_perm) and
// the product's O(n log n) construction (bpomp_bisect_perm) can be checked against the reference's own
// code.  Output: oracle/_ref/libbpomp_ref.so (git-ignored; travels to the GPU box with the snapshot).
#include <cstddef>
#include <map>
#include <utility>
#include <vector>

#include "batching_pomp/util/BisectPerm.hpp"

extern "C" __attribute__((visibility("default"))) int ref_bisect_perm(int n, int* out_first, int* out_second) {
  batching_pomp::util::BisectPerm bp;  // a fresh cache per call: no state between calls
  const std::vector<std::pair<int, int>>& p = bp.get(n);
  for (std::size_t i = 0; i < p.size(); ++i) {
    if (out_first) out_first[i] = p[i].first;
    if (out_second) out_second[i] = p[i].second;
  }
  return (int)p.size();
}
