// ompl_adapter.hpp — kept for callers of round 1: the OMPL planner now lives in
// include/batching_pomp/BatchingPOMP.hpp under the reference's own name, batching_pomp::BatchingPOMP
// (/root/reference/include/batching_pomp/BatchingPOMP.hpp:68).  This header only re-exports it under the old
// namespace batching_pomp::ompl_adapter.
#pragma once
#include "batching_pomp/BatchingPOMP.hpp"
#if __has_include(<ompl/base/Planner.h>)
namespace batching_pomp {
namespace ompl_adapter {
using ::batching_pomp::BatchingPOMP;
using ::batching_pomp::DeviceWorldValidityChecker;
}  // namespace ompl_adapter
}  // namespace batching_pomp
#endif
