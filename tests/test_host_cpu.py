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


def test_graphml_binary_cache(planner_mod, tmp_path, monkeypatch):
    """f3: the binary roadmap cache beside the GraphML file (RoadmapFromFile.hpp:136-147 is a text parse every
    time): first read parses and writes the cache, later reads come from it bit for bit; a changed source file,
    a different dimension or a corrupted cache fall back to parsing; a read-only directory is not an error."""
    import os
    from batching_pomp_b200 import workloads
    st = workloads.halton(3000, 7)
    edges = np.array([[0, 1], [5, 2999]], dtype=np.uint32)
    f = str(tmp_path / "r.graphml")
    planner_mod.write_graphml(f, st, edges)
    s1, e1, hit1 = planner_mod.read_graphml_cached(f, 7)   # two calls inside: counts (parse + write), then data (cache)
    assert os.path.exists(f + ".bpomp-cache")
    s2, e2, hit2 = planner_mod.read_graphml_cached(f, 7)
    assert hit2
    for s, e in ((s1, e1), (s2, e2)):
        np.testing.assert_array_equal(s, st)
        np.testing.assert_array_equal(e, edges)
    # the planner's setup() takes the same path
    p = planner_mod.Planner(7, boxes=(np.zeros((1, 7)), np.full((1, 7), 0.1)))
    p.set_param("roadmap_filename", f)
    p.set_param("batching_type", "vertex")
    p.setup()
    # stale: the source changes (other contents, other size / mtime) -> parsed again, cache rewritten
    st2 = workloads.halton(2000, 7) * 0.5
    planner_mod.write_graphml(f, st2)
    os.utime(f, ns=(1, 1))
    s3, e3, _ = planner_mod.read_graphml_cached(f, 7)
    np.testing.assert_array_equal(s3, st2)
    assert e3.shape[0] == 0
    assert planner_mod.read_graphml_cached(f, 7)[2]
    # corrupted payload: checksum mismatch -> parse
    with open(f + ".bpomp-cache", "r+b") as c:
        c.seek(200)
        c.write(b"\xff\xff\xff\xff")
    lib = planner_mod.load()
    import ctypes as C
    n, ne, hit = C.c_uint32(0), C.c_uint32(0), C.c_int(1)
    assert lib.bpomp_graphml_read_cached(f.encode(), 7, None, C.byref(n), None, C.byref(ne), C.byref(hit)) == 0
    assert hit.value == 0 and n.value == 2000
    # disabled by the environment
    monkeypatch.setenv("BPOMP_ROADMAP_CACHE", "0")
    os.remove(f + ".bpomp-cache")
    s4, _, hit4 = planner_mod.read_graphml_cached(f, 7)
    assert not hit4 and not os.path.exists(f + ".bpomp-cache")
    np.testing.assert_array_equal(s4, st2)
    # too-small buffers are refused (capacities travel in *n / *nedges)
    monkeypatch.delenv("BPOMP_ROADMAP_CACHE")
    small = np.zeros((10, 7))
    n = C.c_uint32(10)
    assert lib.bpomp_graphml_read(f.encode(), 7, small.ctypes.data_as(C.c_void_p), C.byref(n), None, None) != 0
    assert n.value == 2000 and not small.any()


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
    """No CPU fallback: parameters, setup() and roadmap files are host work, the first call that needs the
    device (setProblemDefinition validates start/goal there, BP.cpp:772-780) fails loudly without CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = planner_mod.Planner(2, boxes=(np.zeros((1, 2)), np.full((1, 2), 0.1)))
    assert p.gpu_launches == 0
    with pytest.raises(planner_mod.PlannerError) as ei:
        p.set_problem([0.5, 0.5], [0.9, 0.9])
    assert "CUDA" in str(ei.value)
    with pytest.raises(planner_mod.PlannerError):
        p.solve(max_iterations=1)


def test_planner_setup_errors_mirror_reference(planner_mod, tmp_path):
    """setup() parameter validation in the reference's order and wording (BP.cpp:676-750, Selector.hpp:74)."""
    from batching_pomp_b200 import workloads as wl
    P, E = planner_mod.Planner, planner_mod.PlannerError
    lo, hi = wl.box_world(2, 10, seed=3)
    p = P(2, boxes=(lo, hi))
    p.set_roadmap(wl.halton(100, 2))
    p.set_param("batching_type", "edge")
    with pytest.raises(E) as ei:
        p.setup()  # BP.cpp:681-683
    assert "Graph type needed" in str(ei.value)
    p.set_param("graph_type", "halton")
    p.set_param("batching_type", "bogus")
    with pytest.raises(E) as ei:
        p.setup()  # BP.cpp:745-747
    assert "Invalid batching type" in str(ei.value)
    p.set_param("batching_type", "vertex")
    p.set_param("selector_type", "bogus")
    with pytest.raises(E) as ei:
        p.setup()  # Selector.hpp:74
    assert ei.value.code == -2
    r = P(2, boxes=(lo, hi))
    r.set_param("batching_type", "vertex")
    with pytest.raises(E) as ei:
        r.setup()  # BP.cpp:686-688
    assert "Roadmap name must be set" in str(ei.value)
    r.set_param("roadmap_filename", str(tmp_path / "missing.graphml"))
    with pytest.raises(E):
        r.setup()


@pytest.mark.parametrize("batching", ["vertex", "edge", "hybrid", "single"])
def test_planner_setup_from_graphml_on_cpu(planner_mod, tmp_path, batching):
    """setup() with roadmap_filename reads the GraphML roadmap and builds the batching manager on the host."""
    from batching_pomp_b200 import workloads as wl
    d, N = 3, 400
    roadmap = wl.halton(N, d)
    edges = np.stack([np.arange(N - 1), np.arange(1, N)], 1).astype(np.uint32)
    f = str(tmp_path / "roadmap.graphml")
    planner_mod.write_graphml(f, roadmap, edges)
    lo, hi = wl.box_world(d, 10, seed=3)
    p = planner_mod.Planner(d, boxes=(lo, hi))
    p.set_param("roadmap_filename", f)
    p.set_param("batching_type", batching)
    p.set_param("graph_type", "halton")
    p.set_param("start_goal_radius", 0.3)
    p.setup()
    assert p.num_batches == 0 and not p.exhausted
    assert p.best_cost == np.finfo(np.float64).max  # std::numeric_limits<double>::max() (BP.cpp:242,279)
    assert p.gpu_launches == 0


def test_flat_map_against_std_unordered_map(tmp_path):
    """The planner's open-addressing edge index (host/flat_map.hpp) fuzzed against std::unordered_map."""
    import subprocess
    exe = str(tmp_path / "fm")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "batching_pomp_b200", "host"),
                           os.path.join(ROOT, "tests", "cpp", "flat_map_fuzz.cpp"), "-o", exe])
    out = subprocess.check_output([exe], timeout=120).decode()
    assert "flat map ok" in out
