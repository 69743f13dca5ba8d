#!/usr/bin/env python
"""bench.py — headline benchmark: batched edge evaluation (BASELINE.json config 4).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--edges E]

Workload: 10^7 random 7-DoF edges (explicit endpoints, lengths ~U[0, haltonRadius(10^6,7)]) against
500 axis-aligned box obstacles in the unit hypercube, default OMPL resolution (1 % of the extent).
One "step" = one pass of the hot path (checkAndSetEdgeBlocked's isValid loop,
/root/reference/src/BatchingPOMP.cpp:496-544) over the whole batch.

  value        edge checks / s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e          the same through the host-buffer C-ABI call bpomp_edges_check (pinned host inputs,
               H2D + kernel + D2H inside the timed region)
  roofline     117 algorithmic bytes per edge (2*7*8 in + 1 + 4 out) / kernel time, the DRAM
               traffic and the issue-slot ceiling from the committed ncu capture of the same kernel build
  cpu_baseline the reference's own edge code (oracle/_ref, compiled from /root/reference) on all host cores,
               bounded sample; the oracle port where that build is absent.  The GPU results are compared with
               both in the run (exit code 3 on a mismatch, before any number is printed)
  extra        N = 1: indexed box edges (config 3 rows), image edges (config 2), time to first / best path of
               the planner for configs 1-3 beside the CPU oracle planner; N > 1: host-link roof with all ranks
               copying, indexed end-to-end leg

N > 1 (torchrun, one rank per GPU): every rank evaluates its own 10^7-edge shard (weak scaling, world
replicated; --scaling strong splits 10^7 edges) through the library's communicator (bpomp_comm_init_rank): the
edge status flags are packed to 2 bits and merged with an NCCL all-gather that overlaps the next step's kernel.
--impl reference times the CPU arm only (rank 0; the other ranks exit at once).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# (On this pool NCCL_DEBUG=VERSION is exported, so under torchrun NCCL prints its one-line version banner to stdout
# before the JSON line; neither NCCL_DEBUG=WARN nor NCCL_DEBUG_FILE moved it, so it is left alone: the JSON line
# is the last line of stdout.)

from batching_pomp_b200 import workloads  # noqa: E402

DIM = 7
NBOXES = 500
ALGO_BYTES_PER_EDGE = 2 * DIM * 8 + 1 + 4
METRIC = "edge_checks_per_s"
UNIT = "edges/s"
KERNEL_TAG = "r2-fast-v7"  # must match profiles/edge_check_traffic.json (re-captured whenever the kernel changes)


def make_workload(edges, seed=4):
    L = workloads.segment_length(DIM)
    lo, hi = workloads.box_world(DIM, NBOXES, seed=3)
    frm, to = workloads.random_edges(edges, DIM, seed=seed, max_len=workloads.halton_radius(10 ** 6, DIM))
    return L, lo, hi, frm, to


def config_dict(edges, n_gpus, extra=None):
    cfg = {
        "workload": "config4: batched edge evaluation, %d random 7-DoF edges per GPU (explicit endpoints), "
                    "%d box obstacles, resolution 1%% (L=%.7f)" % (edges, NBOXES, workloads.segment_length(DIM)),
        "edges_per_gpu": edges, "dim": DIM, "boxes": NBOXES,
        "l2": "inputs (%.0f MB per step) larger than the 126 MB L2" % (edges * 2 * DIM * 8 / 1e6),
        "sharding": "edges sharded across %d rank(s), world replicated" % n_gpus,
    }
    if extra:
        cfg.update(extra)
    return cfg


MERGE_TEXT = ("bpomp_edges_check_sharded_dev: the status flags packed to 2 bits/edge and merged with one "
              "ncclAllGather per step on the library's communicator stream, overlapped with the next step's "
              "kernel (double-buffered results); first-invalid slots stay with the shard owner, which "
              "replays the belief inserts")


def full_config(edges_per_gpu, world, strong):
    """The `config` object of both arms (the driver compares them key by key)."""
    return config_dict(edges_per_gpu, world, {
        "scaling_mode": ("strong: %d edges in total" % (edges_per_gpu * world)) if strong else "weak: %d edges per GPU" % edges_per_gpu,
        "merge": MERGE_TEXT if world > 1 else "none (single GPU)"})


# --------------------------------------------------------------------------------------------
# CPU baseline (the oracle is test infrastructure; this and --impl reference are the only
# places bench.py executes it)
# --------------------------------------------------------------------------------------------
def cpu_baseline_run(L, lo, hi, frm, to, target_s=12.0, steps=1, warmup=0, prefer_reference=True, parity_edges=0):
    """Times the CPU arm on all host cores on a bounded sample of the batch.  kind "reference": the reference's own
    initializeEdgePoints + checkAndSetEdgeBlocked (src/BatchingPOMP.cpp:435-461, 496-544) compiled from
    /root/reference into oracle/_ref (one planner instance per thread: the reference is single-threaded); kind
    "port" (no oracle/_ref on this machine): the oracle's restatement with OpenMP over edges.  `parity_edges` > 0
    additionally returns the oracle port's outputs for that many edges (the in-run parity gate; not timed)."""
    from oracle import oracle
    # all host cores this process may use; torchrun exports OMP_NUM_THREADS=1, which must not throttle the
    # CPU arm, so the thread count is passed explicitly instead of read from it
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    use_ref = prefer_reference and oracle.ref_edges_available()
    frac = L / workloads.max_extent(DIM)

    def run(f, t, threads):
        if use_ref:
            return oracle.ref_edges_check_boxes(f, t, frac, lo, hi, nthreads=threads)
        return oracle.edges_check_boxes(f, t, L, lo, hi, nthreads=threads)

    probe = min(20000, frm.shape[0])
    t0 = time.perf_counter()
    run(frm[:probe], to[:probe], cores)
    dt = max(time.perf_counter() - t0, 1e-6)
    per_step = target_s / max(steps + warmup, 1)
    sample = int(min(frm.shape[0], max(probe, probe * per_step / dt)))
    times = []
    outputs = None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        outputs = run(frm[:sample], to[:sample], cores)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    # single-thread figure on a smaller sample (the reference itself is single-threaded)
    s1 = max(1000, sample // (4 * cores))
    t0 = time.perf_counter()
    run(frm[:s1], to[:s1], 1)
    v1 = s1 / max(time.perf_counter() - t0, 1e-9)
    res = {"value": sample * len(times) / total, "unit": UNIT, "cores": cores, "kind": "reference" if use_ref else "port",
           "sample": "first %d edges of the same batch, %d timed pass(es), %s; single-thread: %.4g edges/s"
                     % (sample, len(times),
                        "the reference's own edge code (oracle/_ref), one planner instance per thread" if use_ref
                        else "oracle port, OpenMP over edges", v1),
           "single_thread_value": v1, "ms_per_step": 1e3 * total / len(times),
           "outputs": outputs, "sample_edges": sample}
    if parity_edges > 0:  # the oracle port on a larger share of the batch: checker only, not timed
        m = int(min(frm.shape[0], parity_edges))
        t0 = time.perf_counter()
        res["port_outputs"] = oracle.edges_check_boxes(frm[:m], to[:m], L, lo, hi, nthreads=cores)
        res["port_value"] = m / max(time.perf_counter() - t0, 1e-9)
    return res


def parity_check(tag, got, want):
    """In-run parity gate: the GPU's (valid, first_invalid_slot, nstates) of the first len(want[0]) edges must be
    bit-equal to the oracle's.  A mismatch aborts the run (exit code 3) before any number is printed."""
    names = ("valid", "first_invalid_slot", "nstates")
    m = len(want[0])
    for name, g, w in zip(names, got, want):
        g = np.asarray(g[:m]).astype(np.int64)
        w = np.asarray(w).astype(np.int64)
        if not np.array_equal(g, w):
            bad = int(np.nonzero(g != w)[0][0])
            sys.stderr.write("PARITY FAILURE (%s): %s differs from the CPU oracle at edge %d: gpu %d, oracle %d "
                             "(%d of %d differ)\n" % (tag, name, bad, g[bad], w[bad], int((g != w).sum()), m))
            sys.stderr.flush()
            os._exit(3)
    return m


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    strong = args.scaling == "strong"
    E = args.edges if not strong else (args.edges // world) // 16 * 16
    L, lo, hi, frm, to = make_workload(min(E, 2_000_000))
    cb = cpu_baseline_run(L, lo, hi, frm, to, target_s=60.0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": full_config(E, world, strong),
        "note": ("the reference's own edge evaluation (src/BatchingPOMP.cpp:435-461, 496-544, compiled from /root/reference "
                 "against API stand-ins into oracle/_ref), one planner instance per host thread; each step is a bounded "
                 "sample of the batch" if cb["kind"] == "reference" else
                 "CPU oracle port of the reference loop (oracle/_ref is not present on this machine); each step is a "
                 "bounded sample of the batch"),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False
        self.active = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag:
            if self.active:
                try:
                    self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                    r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.004)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def device_time(torch, stream, ctx, fn, reps, warm=20):
    """Average device time of fn() (launches on `stream`) in ms: CUDA events on that stream."""
    for _ in range(warm):
        fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    ctx.sync()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def peak_hbm():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def extra_indexed_leg(torch, capi, oracle, stream, dev, L, lo, hi, peak, cores, rank=0, world=1, dist=None):
    """The planner's own edge path (BASELINE config 3): r-disk rows of the 10^6-vertex 7-D Halton roadmap as indexed
    edges (source = the expanded vertex) against the 500 boxes; vertex table resident, 8 B of ids per edge in."""
    N, nq = 10 ** 6, 4096
    verts = workloads.halton(N, DIM)
    r = workloads.halton_radius(N, DIM)
    ctx = capi.Context(DIM, L, device=dev.index)
    ctx.set_bounds(0.0, 1.0)
    ctx.set_world_boxes(lo, hi)
    ctx.vertices_append(verts)
    # N > 1 (weak scaling): every rank takes the rows of different vertices; the vertex table is replicated
    q = ((np.arange(nq, dtype=np.uint64) * (N // nq) + 17 * rank) % N).astype(np.uint32)
    t0 = time.perf_counter()
    rp, col, _ = ctx.neighbors_radius(q, r)
    nb_s = time.perf_counter() - t0
    fi = np.repeat(q, np.diff(rp.astype(np.int64)))
    keep = fi != col
    fi, ti = np.ascontiguousarray(fi[keep]), np.ascontiguousarray(col[keep])
    E = int(fi.size)
    ctx.set_stream(stream.cuda_stream)
    h_fi, h_ti = torch.from_numpy(fi.view(np.int32)).pin_memory(), torch.from_numpy(ti.view(np.int32)).pin_memory()
    d_fi, d_ti = h_fi.to(dev), h_ti.to(dev)
    tv = torch.empty(E, dtype=torch.uint8, device=dev)
    tf = torch.empty(E, dtype=torch.int32, device=dev)
    tn = torch.empty(E, dtype=torch.int32, device=dev)
    ms = device_time(torch, stream, ctx, lambda: ctx.edges_check_indexed_dev(d_fi.data_ptr(), d_ti.data_ptr(), E, tv.data_ptr(),
                                                                             tf.data_ptr(), tn.data_ptr()), reps=10)
    # end to end from pinned host id arrays
    hv = torch.empty(E, dtype=torch.uint8).pin_memory()
    hf = torch.empty(E, dtype=torch.int32).pin_memory()
    fi_p, ti_p = h_fi.numpy().view(np.uint32), h_ti.numpy().view(np.uint32)

    def e2e():
        ctx._ck(ctx.lib.bpomp_edges_check_indexed(ctx.h, capi._ptr(fi_p), capi._ptr(ti_p), E, capi._ptr(hv.numpy()),
                                                  capi._ptr(hf.numpy()), None))
    e2e()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    for _ in range(5):
        e2e()
    e2e_s = (time.perf_counter() - w0) / 5
    E_all = E
    if world > 1:  # max time over ranks, edges of all ranks
        t = torch.tensor([ms, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
        n = torch.tensor([E], dtype=torch.int64, device=dev)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        E_all = int(n[0])
    # parity + CPU column on a sample
    m = min(E, 1_000_000)
    t0 = time.perf_counter()
    ov, of, on = oracle.edges_check_boxes(verts[fi[:m]], verts[ti[:m]], L, lo, hi, nthreads=max(1, cores // world))
    cpu = m / (time.perf_counter() - t0)
    parity_check("indexed leg", (tv.cpu().numpy(), tf.cpu().numpy(), tn.cpu().numpy().view(np.uint32)), (ov, of, on))
    parity_check("indexed leg, end to end", (hv.numpy(), hf.numpy(), tn.cpu().numpy().view(np.uint32)), (ov, of, on))
    algo = 13 * E + N * DIM * 8
    ctx.close()
    return {"workload": "config3 roadmap edges: %d r-disk rows (r = haltonRadius) of the 10^6-vertex 7-D Halton roadmap, "
                        "indexed endpoints, 500 boxes" % nq,
            "edges": E_all, "n_gpus": world, "value": E_all / ms * 1e3, "unit": UNIT, "ms": ms,
            "roofline": {"bound": "hbm", "algorithmic_bytes": algo, "achieved": algo / ms * 1e-6, "peak": peak, "unit": "GB/s",
                         "frac": algo / ms * 1e-6 / peak,
                         "note": "per GPU: 13 B per edge + the 56 MB vertex table once; instruction-issue bound like the headline"},
            "e2e": {"value": E_all / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * E, "d2h_bytes_per_step": 5 * E,
                    "ms_per_step": 1e3 * e2e_s,
                    "call": "bpomp_edges_check_indexed (pinned host id arrays; vertex table resident; max over ranks)"},
            "cpu_port_value": cpu, "cpu_cores": cores, "parity_checked_edges": m,
            "neighbor_rows_seconds": nb_s, "blocked_fraction": float((tv == 0).float().mean())}


def h2d_roof(torch, dev, world, dist, mbytes=512, reps=5):
    """What the host link moves: every rank copies a pinned buffer to its GPU at the same time (barrier on both
    sides, max over ranks); the figure the end-to-end numbers are to be read against."""
    h = torch.empty(mbytes << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mbytes << 20, dtype=torch.uint8, device=dev)
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t[0])
    per_rank = reps * (mbytes << 20) / dt / 1e9
    return {"GBps_per_rank": per_rank, "GBps_aggregate": per_rank * world, "n_gpus": world,
            "how": "%d MiB pinned -> device, %d copies per rank, all ranks at once, max time over ranks" % (mbytes, reps)}


def extra_image_leg(torch, capi, oracle, stream, dev, peak, cores):
    """BASELINE config 2's edge evaluation: every roadmap edge (u < v) of the 10^6-vertex 2-D Halton roadmap at
    r = haltonRadius against the 4096^2 occupancy image, indexed endpoints."""
    d, N, size, frac = 2, 10 ** 6, 4096, 1e-4
    verts = workloads.halton(N, d)
    r = workloads.halton_radius(N, d)
    L2 = workloads.segment_length(d, frac)
    pix = workloads.image_world(size, seed=2)
    ctx = capi.Context(d, L2, device=dev.index)
    ctx.set_world_image(pix)
    ctx.vertices_append(verts)
    t0 = time.perf_counter()
    rp, col, _ = ctx.neighbors_all(r)
    nb_s = time.perf_counter() - t0
    rows = np.repeat(np.arange(N, dtype=np.uint32), np.diff(rp.astype(np.int64)))
    keep = rows < col
    fi, ti = np.ascontiguousarray(rows[keep]), np.ascontiguousarray(col[keep])
    E = int(fi.size)
    ctx.set_stream(stream.cuda_stream)
    d_fi, d_ti = torch.from_numpy(fi.view(np.int32)).to(dev), torch.from_numpy(ti.view(np.int32)).to(dev)
    tv = torch.empty(E, dtype=torch.uint8, device=dev)
    tf = torch.empty(E, dtype=torch.int32, device=dev)
    tn = torch.empty(E, dtype=torch.int32, device=dev)
    ms = device_time(torch, stream, ctx, lambda: ctx.edges_check_indexed_dev(d_fi.data_ptr(), d_ti.data_ptr(), E, tv.data_ptr(),
                                                                             tf.data_ptr(), tn.data_ptr()), reps=10)
    m = min(E, 1_000_000)
    t0 = time.perf_counter()
    ov, of, on = oracle.edges_check_image(verts[fi[:m]], verts[ti[:m]], L2, pix, nthreads=cores)
    cpu = m / (time.perf_counter() - t0)
    parity_check("image leg", (tv.cpu().numpy(), tf.cpu().numpy(), tn.cpu().numpy().view(np.uint32)), (ov, of, on))
    algo = 13 * E + N * d * 8
    ctx.close()
    return {"workload": "config2 roadmap edges: all %d edges (u < v) of the 10^6-vertex 2-D Halton roadmap at r = haltonRadius, "
                        "4096^2 occupancy image, resolution 1e-4, indexed endpoints" % E,
            "edges": E, "value": E / ms * 1e3, "unit": UNIT, "ms": ms,
            "roofline": {"bound": "hbm", "algorithmic_bytes": algo, "achieved": algo / ms * 1e-6, "peak": peak, "unit": "GB/s",
                         "frac": algo / ms * 1e-6 / peak},
            "cpu_port_value": cpu, "cpu_cores": cores, "parity_checked_edges": m,
            "densification_seconds": nb_s, "csr_neighbours": int(col.size),
            "blocked_fraction": float((tv == 0).float().mean()), "mean_nstates": float(tn.float().mean())}


def _anytime(p, start, goal, single, max_solutions, time_cap):
    t0 = time.perf_counter()
    if single:
        p.setup(); p.set_problem(start, goal)
    else:
        p.set_problem(start, goal); p.setup()
    first, costs, paths, st = None, [], [], "UNKNOWN"
    for _ in range(400):
        left = time_cap - (time.perf_counter() - t0)
        if left <= 0:
            st = "BENCH_TIME_CAP"
            break
        st = p.solve(max_seconds=left)
        if len(p.path()) and (not costs or p.best_cost < costs[-1][1]):
            costs.append((time.perf_counter() - t0, p.best_cost))
            paths.append(np.asarray(p.path_states()).copy())
            if first is None:
                first = time.perf_counter() - t0
        if st in ("EXACT_SOLUTION", "TIMEOUT") or len(costs) >= max_solutions:
            break
    return {"time_to_first_path_s": first, "time_to_best_path_s": costs[-1][0] if costs else None,
            "best_cost": costs[-1][1] if costs else None, "solutions": len(costs), "final_status": st,
            "counters": p.counters(), "paths": paths, "costs": [c for _, c in costs]}


def extra_planner_leg(oracle):
    """The second half of BASELINE's metric: wall time to the first and to the best path of the whole planner loop
    (src/BatchingPOMP.cpp:819-1017), GPU host planner beside the single-threaded CPU oracle planner (the reference is
    single-threaded), same roadmaps, paths compared solution by solution."""
    from batching_pomp_b200.planner import Planner
    out = {}
    # config 1: 2-D image, N = 10^4, hybrid batching
    d, N, frac = 2, 10 ** 4, 1e-3
    pix = workloads.image_world(1024, seed=1)
    roadmap = workloads.halton(N, d)
    params = {"batching_type": "hybrid", "graph_type": "halton", "increment": 0.2, "prune_threshold": 1.05}
    # config 3: 7-DoF boxes, N = 10^6, vertex batching
    d3, N3 = 7, 10 ** 6
    lo3, hi3 = workloads.box_world(d3, 500, seed=3)
    roadmap3 = workloads.halton(N3, d3)
    for name, mk in (("config1", lambda cls: (cls, d, frac, ("image", pix), roadmap, params, 12, 60.0)),
                     ("config3", lambda cls: (cls, d3, 0.01, ("boxes", lo3, hi3), roadmap3, {"batching_type": "vertex"}, 6, 40.0))):
        res = {}
        # The GPU planner runs twice: the first instance creates its device context inside the timed region (CUDA
        # module load, tables, world and vertex-table upload: "cold"); the second recycles that context from the
        # process-wide pool, which is how a process that plans more than one query runs (the reference is
        # one-planner-instance-per-query).  The headline figures are the warm run's; the cold ones are reported too.
        for side in ("gpu_cold", "gpu", "cpu_oracle"):
            _, dd, fr, world, rm, prm, nsol, cap = mk(side)
            if side.startswith("gpu"):
                kw = {"image": world[1]} if world[0] == "image" else {"boxes": (world[1], world[2])}
                p = Planner(dd, segment_fraction=fr, **kw)
                if hasattr(p, "set_roadmap_view"):
                    p.set_roadmap_view(rm)
                else:
                    p.set_roadmap(rm)
            else:
                p = oracle.OraclePlanner(dd, workloads.segment_length(dd, fr), workloads.max_extent(dd), 1.0)
                if world[0] == "image":
                    p.set_world_image(world[1])
                else:
                    p.set_world_boxes(world[1], world[2])
                p.set_roadmap(rm)
            for k, v in prm.items():
                p.set_param(k, v)
            res[side] = _anytime(p, np.full(dd, 0.1), np.full(dd, 0.9), False, nsol, cap)
            if side.startswith("gpu"):
                res[side]["gpu_launches"] = p.gpu_launches
                res[side]["phase_seconds"] = p.times()
                p.close()  # hands the device context back to the pool
        cold = res.pop("gpu_cold")
        same_cold = cold["costs"] == res["gpu"]["costs"] and all(np.array_equal(a, b) for a, b in zip(cold["paths"], res["gpu"]["paths"]))
        if not same_cold:
            sys.stderr.write("PARITY FAILURE (planner %s): two runs of the GPU planner differ\n" % name)
            os._exit(3)
        res["gpu"]["cold_start"] = {"time_to_first_path_s": cold["time_to_first_path_s"],
                                    "time_to_best_path_s": cold["time_to_best_path_s"],
                                    "note": "first planner instance of the process: device context created inside the timed region"}
        k = min(res["gpu"]["solutions"], res["cpu_oracle"]["solutions"])
        same = k > 0 and all(np.array_equal(a, b) for a, b in zip(res["gpu"]["paths"][:k], res["cpu_oracle"]["paths"][:k])) \
            and res["gpu"]["costs"][:k] == res["cpu_oracle"]["costs"][:k]
        if not same:
            sys.stderr.write("PARITY FAILURE (planner %s): the GPU planner's solutions differ from the CPU oracle planner's\n" % name)
            os._exit(3)
        for side in res:
            res[side].pop("paths")
            res[side].pop("costs")
        out[name] = {"gpu": res["gpu"], "cpu_oracle": res["cpu_oracle"], "paths_identical": True, "solutions_compared": k}
    # config 2: 2-D image, N = 10^6, edge batching with the first-unevaluated-edge selector.  GPU planner only, first two
    # solutions: the CPU oracle's nearest-neighbour search is a linear scan and cannot expand the vertices of a
    # 10^6-vertex roadmap in reasonable time (its parity with this planner is tested on smaller roadmaps).
    pix2 = workloads.image_world(4096, seed=2)
    roadmap2 = workloads.halton(10 ** 6, 2)
    p = Planner(2, segment_fraction=1e-4, image=pix2)
    if hasattr(p, "set_roadmap_view"):
        p.set_roadmap_view(roadmap2)
    else:
        p.set_roadmap(roadmap2)
    for k, v in (("batching_type", "edge"), ("graph_type", "halton"), ("selector_type", "normal")):
        p.set_param(k, v)
    r2 = _anytime(p, np.full(2, 0.1), np.full(2, 0.9), False, 2, 30.0)
    r2["gpu_launches"] = p.gpu_launches
    r2["phase_seconds"] = p.times()
    r2.pop("paths")
    r2.pop("costs")
    p.close()
    out["config2"] = {"gpu": r2, "cpu_oracle": None,
                      "note": "first two solutions on the full 10^6-vertex roadmap (9.3e6 lazily created edges); no CPU column: "
                              "the oracle planner's linear-scan neighbour search is O(N) per expansion"}
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from batching_pomp_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    strong = args.scaling == "strong"
    # weak scaling: every rank owns args.edges edges; strong scaling: args.edges in total, split evenly
    E = args.edges if not strong else (args.edges // world) // 16 * 16
    L, lo, hi, frm, to = make_workload(E, seed=4 + 100 * rank)  # every rank owns a different shard
    ctx = capi.Context(DIM, L, device=local_rank)
    ctx.set_bounds(0.0, 1.0)  # RealVectorStateSpace bounds of the unit hypercube (a hint: results do not depend on it)
    ctx.set_world_boxes(lo, hi)
    # a dedicated (non-default) stream: the ctx launches on it and the CUDA events are recorded on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    if world > 1:  # the library's own communicator (bpomp_comm_init_rank): the id travels through the process group
        ids = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        if args.reserve_sms >= 0:  # before the communicator exists: it also caps NCCL's CTAs
            ctx.set_option("reserve_sms", args.reserve_sms)
        for kv in args.ctx_option:
            ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
        ctx.comm_init_rank(world, rank, ids[0])

    # pinned host copies (e2e) and device-resident copies (value)
    h_from = torch.from_numpy(frm).pin_memory()
    h_to = torch.from_numpy(to).pin_memory()
    h_valid = torch.empty(E, dtype=torch.uint8).pin_memory()
    h_first = torch.empty(E, dtype=torch.int32).pin_memory()
    d_from = h_from.to(dev, non_blocking=True)
    d_to = h_to.to(dev, non_blocking=True)
    d_ns = torch.empty(E, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    # N > 1: the merge of step k's status flags (packed to 2 bits per edge, one ncclAllGather on the communicator's
    # stream inside bpomp_edges_check_sharded_dev) overlaps step k+1's kernel; results alternate between two buffer
    # sets so that a kernel never overwrites data that is still being gathered.
    nbuf = 2 if world > 1 else 1
    d_valid = [torch.empty(E, dtype=torch.uint8, device=dev) for _ in range(nbuf)]
    d_first = [torch.empty(E, dtype=torch.int32, device=dev) for _ in range(nbuf)]
    words = (E + 15) // 16
    packed = [torch.empty(words, dtype=torch.int32, device=dev) for _ in range(nbuf)] if world > 1 else None
    packed_all = [torch.empty(world * words, dtype=torch.int32, device=dev) for _ in range(nbuf)] if world > 1 else None
    kernel_ms = []
    step_no = [0]

    def step(timed):
        k = step_no[0] % nbuf
        step_no[0] += 1
        if timed:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        if world > 1:
            # (the wait for the merge that last used this buffer set is inside: same communicator stream order)
            ctx.edges_check_sharded_dev(d_from.data_ptr(), d_to.data_ptr(), E, d_valid[k].data_ptr(), d_first[k].data_ptr(),
                                        d_ns.data_ptr(), packed[k].data_ptr(), packed_all[k].data_ptr())
        else:
            ctx.edges_check_dev(d_from.data_ptr(), d_to.data_ptr(), E, d_valid[k].data_ptr(), d_first[k].data_ptr(),
                                d_ns.data_ptr())
        if timed:
            e1.record(stream)
            kernel_ms.append((e0, e1, E))

    def drain():
        if world > 1:
            ctx.comm_wait()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- device-resident throughput ("value") ------------------------------------------------
    for _ in range(args.warmup):
        step(False)
    drain()
    barrier()
    launches0 = ctx.launches
    t_beg = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    sampler.active = True
    t_beg.record(stream)
    for _ in range(args.steps):
        step(True)
    drain()
    t_end.record(stream)
    barrier()
    sampler.active = False
    launches = ctx.launches - launches0
    ctx.sync()  # also surfaces deferred / range flags
    elapsed_ms = t_beg.elapsed_time(t_end)
    per_launch = [(e0.elapsed_time(e1), n) for e0, e1, n in kernel_ms]
    kern_ms_per_step = sum(ms for ms, _ in per_launch) / args.steps
    avg_launch_ms = sum(ms for ms, _ in per_launch) / len(per_launch)
    edges_per_launch = sum(n for _, n in per_launch) / len(per_launch)
    last = (step_no[0] - 1) % nbuf  # buffer set of the last step

    # ---- end to end through the host-buffer C ABI ("e2e") -------------------------------------
    f_np, t_np = h_from.numpy(), h_to.numpy()
    v_np, fi_np = h_valid.numpy(), h_first.numpy()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        ctx.edges_check_into(f_np, t_np, v_np, fi_np, None)
    barrier()
    sampler.active = True
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.edges_check_into(f_np, t_np, v_np, fi_np, None)  # valid + first-invalid slot; nstates is optional
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0
    sampler.active = False
    sampler.stop_flag = True

    merged_ok = True
    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_s = float(t[0]), float(t[1])
        # merged result sanity: this rank's shard must sit at its slot of the gathered packed flags
        mine = torch.empty(E, dtype=torch.uint8, device=dev)
        ctx.edge_status_unpack_dev(packed_all[last][rank * words:].data_ptr(), E, mine.data_ptr())
        ctx.sync()
        merged_ok = bool(torch.equal(mine, d_valid[last]))
        assert merged_ok, "merged status flags differ from the shard's own"

    # parity gate on every rank: the oracle outputs of the cpu_baseline sample (rank 0, N = 1) or of a short oracle
    # pass against what the device-resident step and the end-to-end call produced in this very run
    cores = len(os.sched_getaffinity(0))
    cb = cpu_baseline_run(L, lo, hi, frm, to, target_s=12.0, parity_edges=7_000_000) if (not args.no_cpu and world == 1) else None
    if cb:
        want = cb["port_outputs"]
    else:
        from oracle import oracle
        m = min(E, 200000)
        want = oracle.edges_check_boxes(frm[:m], to[:m], L, lo, hi, nthreads=max(1, cores // world))
    ns_dev = d_ns.cpu().numpy().view(np.uint32)
    checked = parity_check("device-resident step", (d_valid[last].cpu().numpy(), d_first[last].cpu().numpy(), ns_dev), want)
    parity_check("end-to-end call", (v_np, fi_np, ns_dev), want)
    ref_checked = 0
    if cb and cb["kind"] == "reference":  # and against the reference's own code on its timed sample
        ref_checked = parity_check("device-resident step vs the reference's own code",
                                   (d_valid[last].cpu().numpy(), d_first[last].cpu().numpy(), ns_dev), cb["outputs"])

    # legs every rank takes part in (N > 1: the indexed end-to-end figure and the host link's roof)
    multi_extras = None
    if world > 1 and not args.no_extras:
        from oracle import oracle as _orc
        multi_extras = {"h2d_roof": h2d_roof(torch, dev, world, dist),
                        "indexed_edges": extra_indexed_leg(torch, capi, _orc, stream, dev, L, lo, hi, peak_hbm()[0], cores,
                                                           rank=rank, world=world, dist=dist)}
    if rank == 0:
        peak, peak_src = peak_hbm()
        achieved = ALGO_BYTES_PER_EDGE * edges_per_launch / (avg_launch_ms * 1e-3) / 1e9
        # DRAM bytes per launch of the timed kernel come from the committed `ncu --set full` capture of this
        # same kernel (a profiler cannot run inside a timed bench); the record names the kernel and capture it
        # belongs to, and is ignored (traffic = null) if it was taken for a different kernel build.
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "edge_check_traffic.json")
        if os.path.exists(tpath):
            try:
                rec = json.load(open(tpath))
                if rec.get("kernel_tag") == KERNEL_TAG and int(rec.get("edges_per_launch", 0)) == int(edges_per_launch):
                    traffic = rec.get("dram_bytes_per_launch")
                    traffic_src = "profiles/%s (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)" % rec.get("source", "?")
                else:
                    traffic_src = "no ncu capture of kernel build '%s' at %d edges/launch committed" % (KERNEL_TAG, int(edges_per_launch))
            except Exception:
                traffic = None
        # second ceiling (SURVEY.md 8d asks for the ALU one beside HBM): the kernel is instruction-issue bound, so the
        # floor is its warp instructions (same ncu capture as `traffic`) at one instruction per scheduler per cycle
        issue = None
        try:
            if traffic is not None and rec.get("warp_instructions_per_launch"):
                import torch as _t
                prop = _t.cuda.get_device_properties(dev)
                sm_hz = (sampler.summary().get("sm_mhz") or sampler.max_mhz or 1965) * 1e6
                floor_ms = rec["warp_instructions_per_launch"] / (prop.multi_processor_count * 4 * sm_hz) * 1e3
                issue = {"warp_instructions_per_edge": rec["warp_instructions_per_launch"] / edges_per_launch,
                         "floor_ms_per_launch": floor_ms, "frac": floor_ms / avg_launch_ms,
                         "how": "smsp__inst_executed.sum of the committed capture / (SMs x 4 schedulers x SM clock)"}
        except Exception:
            issue = None
        value = world * E * args.steps / (elapsed_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": full_config(E, world, strong),
            "kernel_ms_per_step": kern_ms_per_step,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_edge": ALGO_BYTES_PER_EDGE,
                         "kernel": "edge_check_fast_kernel<7,false> (+ the general kernel's launch for left-over edges: "
                                   "none on this workload, its CTAs exit at once)",
                         "avg_launch_ms": avg_launch_ms, "issue_ceiling": issue,
                         "note": "instruction-issue bound (DESIGN.md 4.1): ~32 warp instructions per edge at ~70% of the "
                                 "issue slots, shared-memory wavefronts ~64% of peak; DRAM at ~45% of the measured copy peak"},
            "e2e": {"value": world * E * e2e_steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(2 * E * DIM * 8), "d2h_bytes_per_step": int(E * 5),
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "h2d_GBps_per_rank": 2 * E * DIM * 8 / (e2e_s / e2e_steps) / 1e9,
                    "call": "bpomp_edges_check (pinned host buffers; chunked over 3 streams: H2D, kernel, D2H overlap): "
                            "bound by the host link (PCIe), 112 B per edge in"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "blocked_fraction": float((fi_np >= 1).mean()),
            "parity_checked_edges": int(checked),
            "parity": "valid, first_invalid_slot and nstates of the first %d edges bit-equal to the CPU oracle, "
                      "device-resident step and end-to-end call%s, checked in this run before printing"
                      % (checked, (", and of the first %d edges bit-equal to the reference's own edge code (oracle/_ref)"
                                   % ref_checked) if ref_checked else ""),
            "parity_checked_edges_reference": int(ref_checked),
        }
        if cb:
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"]["oracle_port_value"] = cb.get("port_value")
        if world == 1 and not args.no_extras:
            from oracle import oracle
            extras = {"h2d_roof": h2d_roof(torch, dev, 1, None)}
            extras["indexed_edges"] = extra_indexed_leg(torch, capi, oracle, stream, dev, L, lo, hi, peak, cores)
            extras["image_edges"] = extra_image_leg(torch, capi, oracle, stream, dev, peak, cores)
            extras["planner"] = extra_planner_leg(oracle)
            line["extra"] = extras
        if multi_extras:
            line["extra"] = multi_extras
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--edges", type=int, default=10 ** 7)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --edges per GPU (default); strong: --edges in total")
    ap.add_argument("--reserve-sms", type=int, default=-1,
                    help="SMs the edge kernel leaves free for NCCL when N > 1 (-1: library default: 4 on 2 GPUs, 8 beyond)")
    ap.add_argument("--ctx-option", action="append", default=[], metavar="NAME=VALUE",
                    help="library tuning option set before the communicator is created (bpomp_set_option)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra legs (indexed / image edges, planner times)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
