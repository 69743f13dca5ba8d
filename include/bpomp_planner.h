/* bpomp_planner.h — C view of the host planner (batching_pomp_b200/host/BatchingPOMP.hpp) for
 * harnesses that cannot include C++ (the Python tests and benchmarks bind it with ctypes).
 * It mirrors the outer contract of the reference planner, SURVEY.md §8(b):
 *   batching_pomp::BatchingPOMP(si)            /root/reference/include/batching_pomp/BatchingPOMP.hpp:133
 *   params (OMPL ParamSet strings)             /root/reference/src/BatchingPOMP.cpp:313-319
 *   setup / setProblemDefinition / solve       BatchingPOMP.hpp:203-205
 * Status values are ompl::base::PlannerStatus::StatusType numbers (TIMEOUT 4, APPROXIMATE_SOLUTION 5,
 * EXACT_SOLUTION 6, ABORT 8).  Exceptions never cross: functions return 0 or a negative code
 * (-1 ompl::Exception-like, -2 std::invalid_argument, -3 other) with the text in
 * bpomp_planner_error().
 */
#ifndef BPOMP_PLANNER_H_
#define BPOMP_PLANNER_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BPOMP_PAPI __attribute__((visibility("default")))
#else
#define BPOMP_PAPI
#endif

typedef struct bpomp_planner bpomp_planner;

/* Space description (RealVectorStateSpace bounds low/high[dim], longestValidSegmentFraction) and
 * the collision world.  world_kind: 1 = boxes (lo, hi, count), 2 = image (pix, width, height,
 * threshold).  Returns NULL on failure (message from bpomp_planner_error(NULL)).  The device context
 * is acquired by the first call that needs the GPU (set_problem / solve); parameters, setup and
 * roadmap files are host work.  There is no CPU fallback: that first call fails without CUDA. */
BPOMP_PAPI bpomp_planner* bpomp_planner_create_boxes(int device, int dim, const double* low, const double* high,
                                                     double segment_fraction, const double* box_lo,
                                                     const double* box_hi, int nboxes);
BPOMP_PAPI bpomp_planner* bpomp_planner_create_image(int device, const double* low, const double* high,
                                                     double segment_fraction, const uint8_t* pix, int width,
                                                     int height, int threshold);
BPOMP_PAPI void bpomp_planner_destroy(bpomp_planner* p);
BPOMP_PAPI const char* bpomp_planner_error(const bpomp_planner* p);

BPOMP_PAPI int bpomp_planner_set_param(bpomp_planner* p, const char* name, const char* value);
BPOMP_PAPI int bpomp_planner_set_roadmap(bpomp_planner* p, const double* states, uint32_t n, const uint32_t* edges,
                                         uint32_t nedges);
/* Non-owning: `states` must outlive the planner (lets one loaded roadmap serve many planner instances). */
BPOMP_PAPI int bpomp_planner_set_roadmap_view(bpomp_planner* p, const double* states, uint32_t n);
BPOMP_PAPI int bpomp_planner_setup(bpomp_planner* p);
BPOMP_PAPI int bpomp_planner_set_problem(bpomp_planner* p, const double* start, const double* goal);
/* ptc: the planner termination condition becomes true after `max_iterations` evaluations
 * (< 0: never) or after `max_seconds` of wall time (<= 0: no time limit). */
BPOMP_PAPI int bpomp_planner_solve(bpomp_planner* p, long max_iterations, double max_seconds, int* status);

BPOMP_PAPI double bpomp_planner_best_cost(const bpomp_planner* p);
BPOMP_PAPI double bpomp_planner_alpha(const bpomp_planner* p);
BPOMP_PAPI double bpomp_planner_radius(const bpomp_planner* p);
BPOMP_PAPI int bpomp_planner_exhausted(const bpomp_planner* p);
BPOMP_PAPI uint32_t bpomp_planner_num_batches(const bpomp_planner* p);
BPOMP_PAPI int bpomp_planner_path(const bpomp_planner* p, uint32_t* vertices, int cap);
BPOMP_PAPI int bpomp_planner_vertex_state(const bpomp_planner* p, uint32_t v, double* out);
/* out[0..3] = numEdgeChecks, numCollChecks, numSearches, numLookups; times[0..2] = lookup, search,
 * collision-check seconds */
BPOMP_PAPI void bpomp_planner_counters(const bpomp_planner* p, uint32_t* out4, double* times3);
/* Wall seconds spent inside calls of the device library (launch + wait); the host share of a run is its
 * wall time minus this. */
BPOMP_PAPI double bpomp_planner_device_seconds(const bpomp_planner* p);
BPOMP_PAPI uint64_t bpomp_planner_num_vertices(const bpomp_planner* p);
BPOMP_PAPI uint64_t bpomp_planner_num_edges(const bpomp_planner* p);
BPOMP_PAPI uint64_t bpomp_planner_gpu_launches(const bpomp_planner* p);

/* GraphML helpers (the roadmap boundary format, util/RoadmapFromFile.hpp). */
BPOMP_PAPI int bpomp_graphml_write(const char* filename, int dim, const double* states, uint32_t n,
                                   const uint32_t* edges, uint32_t nedges);
/* Two-call protocol: states == NULL / edges == NULL returns the counts only.  With buffers, *n and *nedges hold
 * their capacities (in nodes / edges) on entry and the file's counts on return; a buffer that is too small is an
 * error and nothing is written to it. */
BPOMP_PAPI int bpomp_graphml_read(const char* filename, int dim, double* states, uint32_t* n, uint32_t* edges,
                                  uint32_t* nedges);
/* The same through the binary cache `<filename>.bpomp-cache` that the planner's setup() uses (host/graphml.hpp:
 * keyed by the source file's size, modification time and dimension; rewritten when stale).  *from_cache = 1 when
 * the roadmap was read from the cache instead of being parsed. */
BPOMP_PAPI int bpomp_graphml_read_cached(const char* filename, int dim, double* states, uint32_t* n, uint32_t* edges,
                                         uint32_t* nedges, int* from_cache);

#ifdef __cplusplus
}
#endif
#endif
