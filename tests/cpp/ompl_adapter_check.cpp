// Compiles include/batching_pomp/BatchingPOMP.hpp — the drop-in batching_pomp::BatchingPOMP, included and named
// exactly as the reference's (BatchingPOMP.hpp:52,68) — against the OMPL API stand-ins of tests/stubs and drives it
// the way an OMPL application drives the reference planner (README.md:43-67 of the reference): params through
// the ParamSet by name, a GraphML roadmap file, a ProblemDefinition with one start and a GoalState, solve(ptc),
// solution read back from the ProblemDefinition.  Linked with the test double of the CUDA C ABI
// (tests/mock/mock_cuda.cpp) — TEST INFRASTRUCTURE ONLY.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <string>
#include <vector>

#include <batching_pomp/BatchingPOMP.hpp>
#include "graphml.hpp"

extern "C" void orc_halton(int64_t first_index, int64_t count, int d, double* out);

namespace ob = ompl::base;
using batching_pomp::DeviceWorldValidityChecker;

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
  void describe(batching_pomp::core::SpaceInformation& si) const override {
    si.world = batching_pomp::core::SpaceInformation::World::IMAGE;
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

  batching_pomp::BatchingPOMP planner(si);
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
    batching_pomp::BatchingPOMP p2(si);
    p2.params().setParam("batching_type", "bogus");
    p2.params().setParam("roadmap_filename", roadmap_file);
    p2.setup();
  } catch (const ompl::Exception& e) {
    threw = std::string(e.what()).find("Invalid batching type") != std::string::npos;
  }
  REQUIRE(threw);
  // ---- the reference's public members (BatchingPOMP.hpp:116-131) as views ----
  REQUIRE(planner.mSpace.get() == space.get());
  REQUIRE(num_vertices(planner.full_g) == N);                    // the roadmap as loaded
  REQUIRE(num_vertices(planner.g) >= 2 && num_vertices(planner.g) <= N + 2);  // start, goal and the admitted batches
  REQUIRE(num_edges(planner.g) > 0 && num_edges(planner.g) <= planner.g.edgeSlots());
  REQUIRE(planner.g[planner.getStartVertex()].v_state[0] == 0.1 && planner.g[planner.getGoalVertex()].v_state[1] == 0.9);
  REQUIRE(planner.full_g[5].v_state[0] == pts[10] && planner.full_g[5].dim == 2);
  REQUIRE(planner.mBatchingPtr->getNumBatches() >= 1 && planner.mBatchingPtr->getType() == "hybrid");
  REQUIRE(planner.mBatchingPtr->getNumVertices() == N && planner.mBatchingPtr->getCurrentRadius() > 0.0);
  REQUIRE(planner.getSelector() && planner.getSelector()->getType() == "normal");
  {
    std::size_t blocked = 0, known = 0;
    for (uint32_t e = 0; e < planner.g.edgeSlots(); ++e) {
      if (planner.g.removed(e)) continue;
      auto ep = planner.g.edge(e);
      REQUIRE(ep.numStates >= 2 && ep.distance >= 0.0);
      if (ep.blockedStatus == batching_pomp::BatchingPOMP::BLOCKED)
        REQUIRE(ep.distance == std::numeric_limits<double>::max());  // BP.cpp:530
      else
        REQUIRE(planner.vertexDistFun(ep.source, ep.target) == ep.distance);
      if (ep.blockedStatus != batching_pomp::BatchingPOMP::UNKNOWN) {
        ++known;
        const bool isBlocked = ep.blockedStatus == batching_pomp::BatchingPOMP::BLOCKED;
        blocked += isBlocked;
        if (known <= 40) REQUIRE(planner.checkAndSetEdgeBlocked(e) == isBlocked);  // re-evaluation agrees
      }
    }
    REQUIRE(known == planner.getNumEdgeChecks() && blocked > 0);
  }
  // mBeliefModel is the device-resident model: the planner's checks are in it, estimates come from the device
  auto knn = std::dynamic_pointer_cast<batching_pomp::cspacebelief::KNNModel>(planner.mBeliefModel);
  REQUIRE(knn && knn->bound() && knn->getK() == 15 && knn->size() == planner.getNumCollChecks());
  {
    ob::ScopedState<> q(space);
    q[0] = 0.5; q[1] = 0.3;  // inside the wall
    batching_pomp::BeliefPoint bq(q.get(), 2, 0.0);
    const double inWall = planner.mBeliefModel->estimate(bq);
    q[0] = 0.1; q[1] = 0.1;
    const double atStart = planner.mBeliefModel->estimate(batching_pomp::BeliefPoint(q.get(), 2, 0.0));
    REQUIRE(inWall > 0.0 && inWall < 1.0 && atStart >= 0.0 && atStart < inWall);
    const std::size_t before = knn->size();
    planner.mBeliefModel->addPoint(batching_pomp::BeliefPoint(q.get(), 2, 1.0));
    REQUIRE(knn->size() == before + 1);
    REQUIRE(planner.mBeliefModel->estimate(batching_pomp::BeliefPoint(q.get(), 2, 0.0)) > atStart);
    const double pf = planner.computeAndSetEdgeFreeProbability(0);
    REQUIRE(pf > 0.0 && pf <= 1.0);
  }

  // ---- the 8-argument constructor (BatchingPOMP.hpp:147-154) ----
  {
    typedef batching_pomp::BatchingPOMP BP;
    auto model = std::make_shared<batching_pomp::cspacebelief::KNNModel>(7, 0.4, 0.5);
    std::unique_ptr<batching_pomp::util::Selector<BP::Graph>> sel(new batching_pomp::util::Selector<BP::Graph>("failfast"));
    BP p8(si, std::shared_ptr<BP::BatchingManagerType>(), model, std::move(sel), roadmap_file, 0.3, 0.5, 1.1);
    REQUIRE(p8.params().size() == 0);  // BP.cpp:221-267 declares no parameters
    REQUIRE(p8.getRoadmapFileName() == roadmap_file && p8.getStartGoalRadius() == 0.3 && p8.getIncrement() == 0.5 &&
            p8.getPruneThreshold() == 1.1);
    REQUIRE(p8.core().getBeliefK() == 7 && p8.core().getBeliefPrior() == 0.4 && p8.core().getBeliefPriorWeight() == 0.5);
    bool threw8 = false;
    try {
      p8.setup();  // batching_type is still "" -> the reference throws here too (BP.cpp:740-742)
    } catch (const ompl::Exception& e) {
      threw8 = std::string(e.what()).find("Invalid batching type") != std::string::npos;
    }
    REQUIRE(threw8);
    p8.setBatchingType("vertex");
    p8.setSelectorType("failfast");
    auto pdef8 = std::make_shared<ob::ProblemDefinition>(si);
    pdef8->setStartAndGoalStates(start, goal);
    p8.setProblemDefinition(pdef8);
    p8.setup();
    REQUIRE(p8.getSelector()->getType() == "failfast" && p8.mBatchingPtr->getType() == "vertex");
    long b8 = 200;
    ob::PlannerTerminationCondition ptc8([&b8]() { return --b8 < 0; });
    bool solved = false;
    for (int call = 0; call < 20 && !solved; ++call) {
      p8.solve(ptc8);
      solved = pdef8->getSolutionCount() > 0;
    }
    REQUIRE(solved && model->bound() && model->size() == p8.getNumCollChecks());
    bool bad = false;
    try {
      batching_pomp::util::Selector<BP::Graph> s2("bogus");
    } catch (const std::invalid_argument&) {
      bad = true;  // Selector.hpp:74
    }
    REQUIRE(bad);
  }
  // ---- generated roadmap (B200 extension): "halton:<N>" names the same roadmap as the GraphML file written above ----
  {
    batching_pomp::BatchingPOMP pg(si);
    REQUIRE(pg.params().setParam("roadmap_filename", "halton:" + std::to_string(N)));
    REQUIRE(pg.params().setParam("batching_type", "hybrid"));
    REQUIRE(pg.params().setParam("graph_type", "halton"));
    REQUIRE(pg.params().setParam("increment", "0.25"));
    REQUIRE(pg.params().setParam("prune_threshold", "1.05"));
    auto pdefg = std::make_shared<ob::ProblemDefinition>(si);
    pdefg->setStartAndGoalStates(start, goal);
    pg.setProblemDefinition(pdefg);
    pg.setup();
    REQUIRE(num_vertices(pg.full_g) == N);
    for (uint32_t i = 0; i < N; i += 97)
      REQUIRE(pg.full_g[i].v_state[0] == pts[2 * i] && pg.full_g[i].v_state[1] == pts[2 * i + 1]);
    long bg = 300;
    ob::PlannerTerminationCondition ptcg([&bg]() { return --bg < 0; });
    double cost_g = 1e300;
    for (int call = 0; call < 30; ++call) {
      ob::PlannerStatus st = pg.solve(ptcg);
      if (pdefg->getSolutionCount() > 0) cost_g = pg.getCurrentBestCost();
      if (st == ob::PlannerStatus::EXACT_SOLUTION || st == ob::PlannerStatus::TIMEOUT) break;
    }
    REQUIRE(cost_g == last_cost && pg.getNumSearches() == planner.getNumSearches() &&
            pg.getNumEdgeChecks() == planner.getNumEdgeChecks());  // the same run as from the file
    bool badname = false;
    try {
      batching_pomp::BatchingPOMP pb(si);
      pb.params().setParam("roadmap_filename", "halton:abc");
      pb.params().setParam("batching_type", "vertex");
      pb.setup();
    } catch (const ompl::Exception&) {
      badname = true;
    }
    REQUIRE(badname);
  }
  printf("ompl adapter ok: %d improving solutions, cost %.6f, %u searches\n", improved, last_cost,
         planner.getNumSearches());
  return 0;
}
