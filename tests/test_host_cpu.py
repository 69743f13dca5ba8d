"""CPU-side checks of the host planner library: exports, GraphML boundary format, Halton roadmap."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def planner_mod():
    import __graft_entry__ as g
    from batching_pomp_b200 import planner as p
    if not os.path.exists(p.LIB_PATH):
        g.build()
    p.load()
    return p


def test_planner_header_symbols_exported(planner_mod):
    src = open(os.path.join(ROOT, "include", "bpomp_planner.h")).read()
    names = sorted(set(re.findall(r"BPOMP_PAPI\s+[\w\s\*]+?\b(bpomp_\w+)\s*\(", src)))
    assert sorted(planner_mod.SIGNATURES) == names
    lib = ctypes.CDLL(planner_mod.LIB_PATH)
    for n in names:
        assert hasattr(lib, n)


def test_graphml_roundtrip_exact(planner_mod, tmp_path, orc):
    from batching_pomp_b200 import workloads
    st = workloads.halton(500, 7)
    np.testing.assert_array_equal(st, orc.halton(1, 500, 7))  # numpy Halton == oracle Halton, bit for bit
    edges = np.array([[0, 1], [1, 2], [499, 0]], dtype=np.uint32)
    f = str(tmp_path / "r.graphml")
    planner_mod.write_graphml(f, st, edges)
    s2, e2 = planner_mod.read_graphml(f, 7)
    np.testing.assert_array_equal(s2, st)  # %.17g round-trips through `ss >> double`
    np.testing.assert_array_equal(e2, edges)


def test_graphml_reads_networkx_files(planner_mod, tmp_path):
    nx = pytest.importorskip("networkx")
    rng = np.random.default_rng(0)
    st = rng.random((50, 3))
    g = nx.Graph()
    for i in range(50):
        g.add_node(i, state=" ".join(repr(float(x)) for x in st[i]))
    g.add_edge(3, 7)
    g.add_edge(7, 9)
    f = str(tmp_path / "nx.graphml")
    nx.write_graphml(g, f)
    s2, e2 = planner_mod.read_graphml(f, 3)
    np.testing.assert_array_equal(s2, st)
    assert e2.tolist() == [[3, 7], [7, 9]]


def test_graphml_errors(planner_mod, tmp_path):
    with pytest.raises(planner_mod.PlannerError):
        planner_mod.read_graphml(str(tmp_path / "missing.graphml"), 2)
    bad = tmp_path / "bad.graphml"
    bad.write_text('<graphml><key id="d0" for="node" attr.name="state" attr.type="string"/>'
                   '<key id="d1" for="node" attr.name="color" attr.type="string"/><graph>'
                   '<node id="a"><data key="d0">0.1 0.2</data><data key="d1">red</data></node></graph></graphml>')
    with pytest.raises(planner_mod.PlannerError) as ei:
        planner_mod.read_graphml(str(bad), 2)
    assert "property not found" in str(ei.value)
    bad2 = tmp_path / "bad2.graphml"
    bad2.write_text('<graphml><key id="d0" for="node" attr.name="state" attr.type="string"/><graph>'
                    '<node id="a"><data key="d0">0.1 0.2</data></node><edge source="a" target="zz"/></graph></graphml>')
    with pytest.raises(planner_mod.PlannerError):
        planner_mod.read_graphml(str(bad2), 2)


def test_planner_needs_gpu(planner_mod):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(planner_mod.PlannerError) as ei:
        planner_mod.Planner(2, boxes=(np.zeros((1, 2)), np.ones((1, 2))))
    assert "CUDA" in str(ei.value)


def test_flat_map_against_std_unordered_map(tmp_path):
    """The planner's open-addressing edge index (host/flat_map.hpp) fuzzed against std::unordered_map."""
    import subprocess
    exe = str(tmp_path / "fm")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "batching_pomp_b200", "host"),
                           os.path.join(ROOT, "tests", "cpp", "flat_map_fuzz.cpp"), "-o", exe])
    out = subprocess.check_output([exe], timeout=120).decode()
    assert "flat map ok" in out
