"""Host-planner logic on CPU: libbpomp_host.so (the C++ planner: batching managers, lazy A* on flat arrays,
selectors, in-order commit of batched results, counters) driven through a TEST DOUBLE of the CUDA C ABI
(tests/mock/mock_cuda.cpp, answered by the CPU oracle) and compared, solve() call by solve() call, with
the oracle's literal restatement of batching_pomp::BatchingPOMP.  This checks the host code on every CPU
run; it says nothing about the CUDA kernels (tests/test_gpu_planner.py does the same comparison through the
real library on a B200).  The double is built into a temp directory next to a copy of libbpomp_host.so and
only the subprocess started here loads it."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mock_dir(tmp_path_factory):
    host = os.path.join(ROOT, "batching_pomp_b200", "libbpomp_host.so")
    if not os.path.exists(host):
        import __graft_entry__ as g
        g.build()
    d = str(tmp_path_factory.mktemp("mockcuda"))
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-shared",
                           "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "mock", "mock_cuda.cpp"),
                           os.path.join(ROOT, "oracle", "bpomp_oracle.cpp"),
                           "-o", os.path.join(d, "libbpomp_cuda.so")])
    shutil.copy(host, os.path.join(d, "libbpomp_host.so"))  # RUNPATH $ORIGIN -> resolves the double above
    return d


@pytest.mark.parametrize("case", ["hybrid_image", "edge_selectors", "vertex_7d", "rgg_edge_7d", "single_graphml"])
def test_host_planner_matches_oracle_planner_on_cpu(mock_dir, case):
    env = dict(os.environ)
    env["BPOMP_HOST_LIB"] = os.path.join(mock_dir, "libbpomp_host.so")
    env.pop("BPOMP_CUDA_LIB", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "mock", "planner_parity_worker.py"), case],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "mock parity ok" in r.stdout


def test_host_planner_fuzz_against_oracle_planner(mock_dir):
    """40 random configurations (dimension, roadmap size, world, batching / graph / selector types, increment,
    prune threshold, search budget) — tests/mock/planner_fuzz_worker.py; 200 seeds pass when run by hand."""
    env = dict(os.environ)
    env["BPOMP_HOST_LIB"] = os.path.join(mock_dir, "libbpomp_host.so")
    env.pop("BPOMP_CUDA_LIB", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "mock", "planner_fuzz_worker.py"), "1", "40"],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "mock fuzz ok" in r.stdout


def test_ompl_adapter_compiles_and_plans(tmp_path):
    """include/batching_pomp/BatchingPOMP.hpp (the drop-in batching_pomp::BatchingPOMP: both constructors, the public
    members g / full_g / mBatchingPtr / mBeliefModel as views) compiled against the OMPL API
    stand-ins of tests/stubs and driven like an OMPL application: ParamSet by name, GraphML roadmap file,
    ProblemDefinition + GoalState, solve(ptc), anytime solutions through addSolutionPath, ompl::Exception on a
    bad parameter (tests/cpp/ompl_adapter_check.cpp).  Device side = the test double."""
    exe = str(tmp_path / "ompl_adapter_check")
    host = os.path.join(ROOT, "batching_pomp_b200", "host")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off",
                           "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "include"), "-I", host,
                           os.path.join(ROOT, "tests", "cpp", "ompl_adapter_check.cpp"),
                           os.path.join(host, "BatchingPOMP.cpp"), os.path.join(host, "graphml.cpp"),
                           os.path.join(ROOT, "tests", "mock", "mock_cuda.cpp"),
                           os.path.join(ROOT, "oracle", "bpomp_oracle.cpp"), "-o", exe])
    out = subprocess.run([exe, str(tmp_path / "roadmap.graphml")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "ompl adapter ok" in out.stdout
