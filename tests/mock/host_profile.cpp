// host_profile.cpp — developer tool (TEST INFRASTRUCTURE): the C++ host planner linked statically with the
// test double of the CUDA C ABI (mock_cuda.cpp + the CPU oracle) into one executable, so that the HOST side
// (A*, batching, edge bookkeeping) can be profiled with gprof on a machine without a GPU.  The time spent
// inside the double is reported separately; it says nothing about the product's device time.
//   host_profile <roadmap.f64> <N> <image.u8> <size> <fraction> <batching> <selector|-> <max_searches>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "BatchingPOMP.hpp"

using batching_pomp::core::BatchingPOMP;
using batching_pomp::core::SpaceInformation;

template <typename T>
static std::vector<T> slurp(const char* path, size_t count) {
  std::vector<T> v(count);
  std::ifstream f(path, std::ios::binary);
  f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(count * sizeof(T)));
  if (!f) { fprintf(stderr, "short read: %s\n", path); exit(2); }
  return v;
}

int main(int argc, char** argv) {
  if (argc < 9) { fprintf(stderr, "usage: see header\n"); return 2; }
  const uint32_t N = (uint32_t)atol(argv[2]);
  const int size = atoi(argv[4]);
  SpaceInformation si;
  si.dim = 2;
  si.low = {0.0, 0.0};
  si.high = {1.0, 1.0};
  si.longestValidSegmentFraction = atof(argv[5]);
  si.world = SpaceInformation::World::IMAGE;
  si.image = slurp<uint8_t>(argv[3], (size_t)size * size);
  si.imageWidth = size; si.imageHeight = size; si.imageThreshold = 128;
  std::vector<double> roadmap = slurp<double>(argv[1], (size_t)N * 2);
  const long max_searches = atol(argv[8]);

  auto t0 = std::chrono::steady_clock::now();
  BatchingPOMP p(si);
  p.setRoadmapView(roadmap.data(), N);
  p.setParam("batching_type", argv[6]);
  p.setParam("graph_type", "halton");
  if (std::string(argv[7]) != "-") p.setParam("selector_type", argv[7]);
  p.setProblemDefinition({0.1, 0.1}, {0.9, 0.9});
  p.setup();
  long searches = 0;
  int solutions = 0;
  while (searches < max_searches) {
    long before = p.getNumSearches();
    long left = max_searches - searches;
    auto st = p.solve([&]() { return --left < 0; });
    searches += p.getNumSearches() - before;
    double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (st == batching_pomp::core::PlannerStatus::APPROXIMATE_SOLUTION && p.getNumSolutions() > (size_t)solutions) {
      solutions = (int)p.getNumSolutions();
      printf("t=%.2fs solution %d cost %.6f searches %u edges %zu\n", t, solutions, p.getCurrentBestCost(),
             p.getNumSearches(), p.numEdges());
    }
    if (st == batching_pomp::core::PlannerStatus::EXACT_SOLUTION || st == batching_pomp::core::PlannerStatus::TIMEOUT) break;
  }
  double total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("total %.2fs, searches %u, search_s %.2f lookup_s %.2f collcheck_s %.2f, vertices %zu edges %zu\n", total,
         p.getNumSearches(), p.getSearchTime(), p.getLookupTime(), p.getCollCheckTime(), p.numVertices(), p.numEdges());
  printf("device(double)_s %.3f  host_s %.3f  search_host_s %.3f = %.2f ms per search\n", p.getDeviceTime(),
         total - p.getDeviceTime(), p.getSearchHostTime(), 1e3 * p.getSearchHostTime() / p.getNumSearches());
  return 0;
}
