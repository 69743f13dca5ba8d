// planner_capi.cpp — C view of the host planner (include/bpomp_planner.h).
#include <chrono>
#include <cstring>
#include <string>

#include "BatchingPOMP.hpp"
#include "bpomp_planner.h"
#include "graphml.hpp"

using batching_pomp::core::BatchingPOMP;
using batching_pomp::core::SpaceInformation;

struct bpomp_planner {
  std::unique_ptr<BatchingPOMP> planner;
  std::string error;
};

static std::string g_error;

template <class F>
static int guarded(bpomp_planner* p, F&& f) {
  try {
    f();
    return 0;
  } catch (const batching_pomp::core::Exception& e) {
    (p ? p->error : g_error) = e.what();
    return -1;
  } catch (const std::invalid_argument& e) {
    (p ? p->error : g_error) = e.what();
    return -2;
  } catch (const std::exception& e) {
    (p ? p->error : g_error) = e.what();
    return -3;
  }
}

static bpomp_planner* create(const SpaceInformation& si) {
  bpomp_planner* p = new bpomp_planner();
  int rc = guarded(nullptr, [&] { p->planner.reset(new BatchingPOMP(si)); });
  if (rc != 0) {
    delete p;
    return nullptr;
  }
  return p;
}

extern "C" bpomp_planner* bpomp_planner_create_boxes(int device, int dim, const double* low, const double* high,
                                                      double segment_fraction, const double* box_lo,
                                                      const double* box_hi, int nboxes) {
  SpaceInformation si;
  si.dim = (unsigned)dim;
  si.device = device;
  si.low.assign(low, low + dim);
  si.high.assign(high, high + dim);
  si.longestValidSegmentFraction = segment_fraction;
  si.world = SpaceInformation::World::BOXES;
  si.boxLo.assign(box_lo, box_lo + (size_t)nboxes * dim);
  si.boxHi.assign(box_hi, box_hi + (size_t)nboxes * dim);
  return create(si);
}

extern "C" bpomp_planner* bpomp_planner_create_image(int device, const double* low, const double* high,
                                                      double segment_fraction, const uint8_t* pix, int width,
                                                      int height, int threshold) {
  SpaceInformation si;
  si.dim = 2;
  si.device = device;
  si.low.assign(low, low + 2);
  si.high.assign(high, high + 2);
  si.longestValidSegmentFraction = segment_fraction;
  si.world = SpaceInformation::World::IMAGE;
  si.image.assign(pix, pix + (size_t)width * height);
  si.imageWidth = width;
  si.imageHeight = height;
  si.imageThreshold = threshold;
  return create(si);
}

extern "C" void bpomp_planner_destroy(bpomp_planner* p) { delete p; }
extern "C" const char* bpomp_planner_error(const bpomp_planner* p) { return p ? p->error.c_str() : g_error.c_str(); }

extern "C" int bpomp_planner_set_param(bpomp_planner* p, const char* name, const char* value) {
  return guarded(p, [&] { p->planner->setParam(name, value); });
}

extern "C" int bpomp_planner_set_roadmap(bpomp_planner* p, const double* states, uint32_t n, const uint32_t* edges,
                                         uint32_t nedges) {
  return guarded(p, [&] {
    const unsigned d = p->planner->getDimension();
    std::vector<double> s(states, states + (size_t)n * d);
    std::vector<std::pair<uint32_t, uint32_t>> e;
    for (uint32_t i = 0; i < nedges; ++i) e.emplace_back(edges[2 * i], edges[2 * i + 1]);
    p->planner->setRoadmap(s, e);
  });
}

extern "C" int bpomp_planner_set_roadmap_view(bpomp_planner* p, const double* states, uint32_t n) {
  return guarded(p, [&] { p->planner->setRoadmapView(states, n); });
}

extern "C" int bpomp_planner_setup(bpomp_planner* p) { return guarded(p, [&] { p->planner->setup(); }); }

extern "C" int bpomp_planner_set_problem(bpomp_planner* p, const double* start, const double* goal) {
  return guarded(p, [&] {
    const unsigned d = p->planner->getDimension();
    p->planner->setProblemDefinition(std::vector<double>(start, start + d), std::vector<double>(goal, goal + d));
  });
}

extern "C" int bpomp_planner_solve(bpomp_planner* p, long max_iterations, double max_seconds, int* status) {
  return guarded(p, [&] {
    long budget = max_iterations;
    auto t0 = std::chrono::steady_clock::now();
    auto ptc = [&]() -> bool {
      if (max_seconds > 0.0 &&
          std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() >= max_seconds)
        return true;
      if (budget < 0) return false;
      if (budget == 0) return true;
      --budget;
      return false;
    };
    batching_pomp::core::PlannerStatus st = p->planner->solve(ptc);
    if (status) *status = (int)st;
  });
}

extern "C" double bpomp_planner_best_cost(const bpomp_planner* p) { return p->planner->getCurrentBestCost(); }
extern "C" double bpomp_planner_alpha(const bpomp_planner* p) { return p->planner->getCurrentAlpha(); }
extern "C" double bpomp_planner_radius(const bpomp_planner* p) { return p->planner->getCurrentRadius(); }
extern "C" int bpomp_planner_exhausted(const bpomp_planner* p) { return p->planner->isExhausted() ? 1 : 0; }
extern "C" uint32_t bpomp_planner_num_batches(const bpomp_planner* p) { return p->planner->getNumBatches(); }

extern "C" int bpomp_planner_path(const bpomp_planner* p, uint32_t* vertices, int cap) {
  const auto& v = p->planner->getSolutionVertices();
  for (int i = 0; i < (int)v.size() && i < cap; ++i) vertices[i] = v[i];
  return (int)v.size();
}

extern "C" int bpomp_planner_vertex_state(const bpomp_planner* p, uint32_t v, double* out) {
  if (v >= p->planner->numVertices()) return -1;
  std::memcpy(out, p->planner->vertexState(v), sizeof(double) * p->planner->getDimension());
  return 0;
}

extern "C" void bpomp_planner_counters(const bpomp_planner* p, uint32_t* out4, double* times3) {
  if (out4) {
    out4[0] = p->planner->getNumEdgeChecks();
    out4[1] = p->planner->getNumCollChecks();
    out4[2] = p->planner->getNumSearches();
    out4[3] = p->planner->getNumLookups();
  }
  if (times3) {
    times3[0] = p->planner->getLookupTime();
    times3[1] = p->planner->getSearchTime();
    times3[2] = p->planner->getCollCheckTime();
  }
}

extern "C" double bpomp_planner_device_seconds(const bpomp_planner* p) { return p->planner->getDeviceTime(); }

extern "C" uint64_t bpomp_planner_num_vertices(const bpomp_planner* p) { return p->planner->numVertices(); }
extern "C" uint64_t bpomp_planner_num_edges(const bpomp_planner* p) { return p->planner->numEdges(); }
extern "C" uint64_t bpomp_planner_gpu_launches(const bpomp_planner* p) {
  return p->planner->gpuLaunches();
}

extern "C" int bpomp_graphml_write(const char* filename, int dim, const double* states, uint32_t n,
                                   const uint32_t* edges, uint32_t nedges) {
  return guarded(nullptr, [&] {
    std::vector<double> s(states, states + (size_t)n * dim);
    std::vector<std::pair<uint32_t, uint32_t>> e;
    for (uint32_t i = 0; i < nedges; ++i) e.emplace_back(edges[2 * i], edges[2 * i + 1]);
    batching_pomp::util::writeGraphml(filename, (unsigned)dim, s, e);
  });
}

extern "C" int bpomp_graphml_read(const char* filename, int dim, double* states, uint32_t* n, uint32_t* edges,
                                  uint32_t* nedges) {
  return guarded(nullptr, [&] {
    std::vector<double> s;
    std::vector<std::pair<uint32_t, uint32_t>> e;
    batching_pomp::util::readGraphml(filename, (unsigned)dim, s, e);
    // in: capacities of the caller's buffers (when given), out: the counts in the file
    const uint32_t cap_n = n ? *n : 0, cap_e = nedges ? *nedges : 0;
    if (n) *n = (uint32_t)(s.size() / dim);
    if (nedges) *nedges = (uint32_t)e.size();
    if (states && (!n || cap_n < s.size() / dim))
      throw batching_pomp::core::Exception("bpomp_graphml_read: the states buffer is too small for the file's nodes");
    if (edges && (!nedges || cap_e < e.size()))
      throw batching_pomp::core::Exception("bpomp_graphml_read: the edges buffer is too small for the file's edges");
    if (states) std::memcpy(states, s.data(), s.size() * sizeof(double));
    if (edges)
      for (size_t i = 0; i < e.size(); ++i) {
        edges[2 * i] = e[i].first;
        edges[2 * i + 1] = e[i].second;
      }
  });
}

extern "C" int bpomp_graphml_read_cached(const char* filename, int dim, double* states, uint32_t* n,
                                         uint32_t* edges, uint32_t* nedges, int* from_cache) {
  return guarded(nullptr, [&] {
    std::vector<double> s;
    std::vector<std::pair<uint32_t, uint32_t>> e;
    const bool hit = batching_pomp::util::readGraphmlCached(filename, (unsigned)dim, s, e);
    if (from_cache) *from_cache = hit ? 1 : 0;
    // in: capacities of the caller's buffers (when given), out: the counts in the file
    const uint32_t cap_n = n ? *n : 0, cap_e = nedges ? *nedges : 0;
    if (n) *n = (uint32_t)(s.size() / dim);
    if (nedges) *nedges = (uint32_t)e.size();
    if (states && (!n || cap_n < s.size() / dim))
      throw batching_pomp::core::Exception("bpomp_graphml_read_cached: the states buffer is too small for the file's nodes");
    if (edges && (!nedges || cap_e < e.size()))
      throw batching_pomp::core::Exception("bpomp_graphml_read_cached: the edges buffer is too small for the file's edges");
    if (states) std::memcpy(states, s.data(), s.size() * sizeof(double));
    if (edges)
      for (size_t i = 0; i < e.size(); ++i) {
        edges[2 * i] = e[i].first;
        edges[2 * i + 1] = e[i].second;
      }
  });
}
