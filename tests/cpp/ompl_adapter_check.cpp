// Compiles batching_pomp_b200/host/ompl_adapter.hpp against the OMPL API stand-ins of tests/stubs and drives it
// the way an OMPL application drives the reference planner (README.md:54-67 of the reference): params through
// the ParamSet by name, a GraphML roadmap file, a ProblemDefinition with one start and a GoalState, solve(ptc),
// solution read back from the ProblemDefinition.  Linked with the test double of the CUDA C ABI
// (tests/mock/mock_cuda.cpp) — TEST INFRASTRUCTURE ONLY.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "ompl_adapter.hpp"
#include "graphml.hpp"

extern "C" void orc_halton(int64_t first_index, int64_t count, int d, double* out);

namespace ob = ompl::base;
using batching_pomp::ompl_adapter::DeviceWorldValidityChecker;

class WallWorld : public DeviceWorldValidityChecker {
public:
  explicit WallWorld(const ob::SpaceInformationPtr& si) : DeviceWorldValidityChecker(si), pix_(256 * 256, 255) {
    for (int r = 0; r < 256; ++r)          // a wall at x in [0.45, 0.55] with a door at y in [0.6, 0.75]
      for (int c = 115; c <= 140; ++c)
        if (r < 153 || r > 192) pix_[r * 256 + c] = 0;
  }
  bool isValid(const ob::State* s) const override {
    const double* v = s->as<ob::RealVectorStateSpace::StateType>()->values;
    if (!(v[0] >= 0.0 && v[0] <= 1.0 && v[1] >= 0.0 && v[1] <= 1.0)) return false;
    int col = std::min(255, (int)std::floor(v[0] * 256)), row = std::min(255, (int)std::floor(v[1] * 256));
    return pix_[row * 256 + col] >= 128;
  }
  void describe(batching_pomp::SpaceInformation& si) const override {
    si.world = batching_pomp::SpaceInformation::World::IMAGE;
    si.image = pix_;
    si.imageWidth = 256; si.imageHeight = 256; si.imageThreshold = 128;
  }
private:
  std::vector<uint8_t> pix_;
};

#define REQUIRE(c) do { if (!(c)) { fprintf(stderr, "FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
  const std::string roadmap_file = argc > 1 ? argv[1] : "/tmp/ompl_adapter_roadmap.graphml";
  const uint32_t N = 2000;
  std::vector<double> pts((size_t)N * 2);
  orc_halton(1, N, 2, pts.data());
  batching_pomp::util::writeGraphml(roadmap_file, 2, pts, {});

  auto space = std::make_shared<ob::RealVectorStateSpace>(2);
  ob::RealVectorBounds bounds(2);
  bounds.setLow(0.0);
  bounds.setHigh(1.0);
  space->setBounds(bounds);
  auto si = std::make_shared<ob::SpaceInformation>(space);
  si->setStateValidityChecker(std::make_shared<WallWorld>(si));
  si->setStateValidityCheckingResolution(2e-3);
  si->setup();

  batching_pomp::ompl_adapter::BatchingPOMP planner(si);
  REQUIRE(planner.getName() == "BatchingPOMP");
  REQUIRE(planner.params().size() == 7);  // BP.cpp:313-319
  REQUIRE(planner.params().setParam("roadmap_filename", roadmap_file));
  REQUIRE(planner.params().setParam("batching_type", "hybrid"));
  REQUIRE(planner.params().setParam("graph_type", "halton"));
  REQUIRE(planner.params().setParam("increment", "0.25"));
  REQUIRE(planner.params().setParam("prune_threshold", "1.05"));
  std::string v;
  REQUIRE(planner.params().getParam("increment", v) && std::stod(v) == 0.25);
  REQUIRE(planner.getBatchingType() == "hybrid");

  auto pdef = std::make_shared<ob::ProblemDefinition>(si);
  ob::ScopedState<> start(space), goal(space);
  start[0] = 0.1; start[1] = 0.1;
  goal[0] = 0.9; goal[1] = 0.9;
  pdef->setStartAndGoalStates(start, goal);
  planner.setProblemDefinition(pdef);
  planner.setup();

  long budget = 300;
  ob::PlannerTerminationCondition ptc([&budget]() { return --budget < 0; });
  int improved = 0;
  double last_cost = 1e300;
  for (int call = 0; call < 30; ++call) {
    std::size_t before = pdef->getSolutionCount();
    ob::PlannerStatus st = planner.solve(ptc);
    if (pdef->getSolutionCount() > before) {
      ++improved;
      REQUIRE(planner.getCurrentBestCost() < last_cost);  // anytime: every reported path is better
      last_cost = planner.getCurrentBestCost();
    }
    if (st == ob::PlannerStatus::EXACT_SOLUTION || st == ob::PlannerStatus::TIMEOUT) break;
    REQUIRE(st == ob::PlannerStatus::APPROXIMATE_SOLUTION);
  }
  REQUIRE(improved >= 1);
  auto* path = pdef->getSolutionPath()->as<ompl::geometric::PathGeometric>();
  REQUIRE(path->getStateCount() >= 3);  // the wall forces a detour through the door
  auto vals = [](ob::State* s) { return s->as<ob::RealVectorStateSpace::StateType>()->values; };
  REQUIRE(vals(path->getState(0))[0] == 0.1 && vals(path->getState(0))[1] == 0.1);
  unsigned last = (unsigned)path->getStateCount() - 1;
  REQUIRE(vals(path->getState(last))[0] == 0.9 && vals(path->getState(last))[1] == 0.9);
  double len = 0.0;
  for (unsigned i = 0; i + 1 <= last; ++i) {  // every segment collision-free under the host-side isValid
    double* a = vals(path->getState(i));
    double* b = vals(path->getState(i + 1));
    len += std::sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]));
    ob::ScopedState<> s(space);
    for (int k = 0; k <= 200; ++k) {
      s[0] = a[0] + (b[0] - a[0]) * k / 200.0;
      s[1] = a[1] + (b[1] - a[1]) * k / 200.0;
      REQUIRE(si->isValid(s.get()));
    }
  }
  REQUIRE(std::fabs(len - planner.getCurrentBestCost()) < 1e-9);
  REQUIRE(planner.getNumSearches() > 0 && planner.getNumEdgeChecks() > 0 && planner.getNumCollChecks() > 0);

  // error paths keep the reference's exception type
  bool threw = false;
  try {
    batching_pomp::ompl_adapter::BatchingPOMP p2(si);
    p2.params().setParam("batching_type", "bogus");
    p2.params().setParam("roadmap_filename", roadmap_file);
    p2.setup();
  } catch (const ompl::Exception& e) {
    threw = std::string(e.what()).find("Invalid batching type") != std::string::npos;
  }
  REQUIRE(threw);
  printf("ompl adapter ok: %d improving solutions, cost %.6f, %u searches\n", improved, last_cost,
         planner.getNumSearches());
  return 0;
}
