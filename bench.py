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
  roofline     117 algorithmic bytes per edge (2*7*8 in + 1 + 4 out) / kernel time
  cpu_baseline the CPU oracle (port of the reference loop) on all host cores, bounded sample

N > 1 (torchrun, one rank per GPU): every rank evaluates its own 10^7-edge shard (weak scaling, world
replicated) and the edge status flags are merged with an NCCL all-gather that overlaps the next
step's kernel.  --impl reference times the CPU oracle only (rank 0).
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


# --------------------------------------------------------------------------------------------
# CPU baseline (the oracle is test infrastructure; this and --impl reference are the only
# places bench.py executes it)
# --------------------------------------------------------------------------------------------
def cpu_baseline_run(L, lo, hi, frm, to, target_s=12.0, steps=1, warmup=0):
    from oracle import oracle
    # all host cores this process may use; torchrun exports OMP_NUM_THREADS=1, which must not throttle the
    # CPU arm, so the thread count is passed explicitly (OpenMP num_threads clause) instead of read from it
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    probe = min(20000, frm.shape[0])
    t0 = time.perf_counter()
    oracle.edges_check_boxes(frm[:probe], to[:probe], L, lo, hi, nthreads=cores)
    dt = max(time.perf_counter() - t0, 1e-6)
    per_step = target_s / max(steps + warmup, 1)
    sample = int(min(frm.shape[0], max(probe, probe * per_step / dt)))
    times = []
    outputs = None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        outputs = oracle.edges_check_boxes(frm[:sample], to[:sample], L, lo, hi, nthreads=cores)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    # single-thread figure on a smaller sample (the reference itself is single-threaded)
    s1 = max(1000, sample // (4 * cores))
    t0 = time.perf_counter()
    oracle.edges_check_boxes(frm[:s1], to[:s1], L, lo, hi, nthreads=1)
    v1 = s1 / max(time.perf_counter() - t0, 1e-9)
    return {"value": sample * len(times) / total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "first %d edges of the same batch, %d timed pass(es), OpenMP over edges; "
                      "single-thread: %.4g edges/s" % (sample, len(times), v1),
            "single_thread_value": v1, "ms_per_step": 1e3 * total / len(times),
            "outputs": outputs, "sample_edges": sample}


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
    L, lo, hi, frm, to = make_workload(min(args.edges, 2_000_000))
    cb = cpu_baseline_run(L, lo, hi, frm, to, target_s=60.0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.edges, args.gpus, {"note": "CPU oracle port of the reference loop (the reference "
                                                      "needs OMPL/Boost/Eigen and cannot be built here); each step is "
                                                      "a bounded sample of the batch"}),
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
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from batching_pomp_b200 import capi, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    E = args.edges
    L, lo, hi, frm, to = make_workload(E, seed=4 + 100 * rank)  # every rank owns a different shard
    ctx = capi.Context(DIM, L, device=local_rank)
    ctx.set_bounds(0.0, 1.0)  # RealVectorStateSpace bounds of the unit hypercube (a hint: results do not depend on it)
    ctx.set_world_boxes(lo, hi)
    # a dedicated (non-default) stream: the ctx launches on it and the CUDA events are recorded on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    # pinned host copies (e2e) and device-resident copies (value)
    h_from = torch.from_numpy(frm).pin_memory()
    h_to = torch.from_numpy(to).pin_memory()
    h_valid = torch.empty(E, dtype=torch.uint8).pin_memory()
    h_first = torch.empty(E, dtype=torch.int32).pin_memory()
    h_ns = torch.empty(E, dtype=torch.int32).pin_memory()
    d_from = h_from.to(dev, non_blocking=True)
    d_to = h_to.to(dev, non_blocking=True)
    d_valid = torch.empty(E, dtype=torch.uint8, device=dev)
    d_first = torch.empty(E, dtype=torch.int32, device=dev)
    d_ns = torch.empty(E, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    # N > 1: the all-gather that merges step k's first-invalid slots runs on its own stream and overlaps
    # step k+1's kernel; results alternate between two buffer sets so the kernel never overwrites data
    # that is still being gathered.
    nbuf = 2 if world > 1 else 1
    d_valid = [d_valid] + [torch.empty_like(d_valid) for _ in range(nbuf - 1)]
    d_first = [d_first] + [torch.empty_like(d_first) for _ in range(nbuf - 1)]
    gathered = [torch.empty(world * E, dtype=torch.uint8, device=dev) for _ in range(nbuf)] if world > 1 else None
    words = (E + 15) // 16  # status flags travel packed, 2 bits per edge (bpomp_edge_status_pack_dev)
    packed = [torch.empty(words, dtype=torch.int32, device=dev) for _ in range(nbuf)] if world > 1 else None
    packed_all = [torch.empty(world * words, dtype=torch.int32, device=dev) for _ in range(nbuf)] if world > 1 else None
    comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    kernel_ms = []
    pending = [None] * nbuf
    step_no = [0]

    def step(timed):
        k = step_no[0] % nbuf
        step_no[0] += 1
        if pending[k] is not None:  # the gather that last read this buffer set must be done
            finish(k)
        if timed:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        ctx.edges_check_dev(d_from.data_ptr(), d_to.data_ptr(), E, d_valid[k].data_ptr(), d_first[k].data_ptr(),
                            d_ns.data_ptr())
        if timed:
            e1.record(stream)
            kernel_ms.append((e0, e1, E))
        if world > 1:
            ctx.edge_status_pack_dev(d_valid[k].data_ptr(), E, packed[k].data_ptr())
            ev = torch.cuda.Event()
            ev.record(stream)
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(ev)
                pending[k] = sharding.allgather_equal(packed[k], packed_all[k], async_op=True)[1]

    def finish(k):
        """Buffer set k: wait for its gather and expand every rank's packed flags to the merged byte array."""
        pending[k].wait()  # the kernel stream waits for the collective
        pending[k] = None
        if E % 16 == 0:
            ctx.edge_status_unpack_dev(packed_all[k].data_ptr(), world * E, gathered[k].data_ptr())
        else:
            for g in range(world):
                ctx.edge_status_unpack_dev(packed_all[k][g * words:].data_ptr(), E, gathered[k][g * E:].data_ptr())

    def drain():
        for k in range(nbuf):
            if pending[k] is not None:
                finish(k)
        if world > 1:
            stream.wait_stream(comm_stream)

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

    # ---- end to end through the host-buffer C ABI ("e2e") -------------------------------------
    f_np, t_np = h_from.numpy(), h_to.numpy()
    v_np, fi_np, ns_np = h_valid.numpy(), h_first.numpy(), h_ns.numpy().view(np.uint32)
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
    # the device-resident and the end-to-end paths must agree
    assert np.array_equal(fi_np, d_first[0].cpu().numpy()), "e2e and device-resident results differ"

    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_s = float(t[0]), float(t[1])
        # merged result sanity: this rank's shard must sit at its slot of the gathered buffer
        assert torch.equal(gathered[0][rank * E:(rank + 1) * E], d_valid[0])

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
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
        # the CPU baseline is timed on rank 0 of the single-GPU run only (multi-GPU runs stay short)
        cb = cpu_baseline_run(L, lo, hi, frm, to, target_s=12.0) if (not args.no_cpu and world == 1) else None
        # parity gate: the oracle outputs of the cpu_baseline sample (or, without that leg, of a short oracle
        # pass) against what the device-resident step and the end-to-end call produced in this very run
        if cb:
            want = cb["outputs"]
        else:
            from oracle import oracle
            m = min(E, 200000)
            want = oracle.edges_check_boxes(frm[:m], to[:m], L, lo, hi, nthreads=len(os.sched_getaffinity(0)))
        ns_dev = d_ns.cpu().numpy().view(np.uint32)
        checked = parity_check("device-resident step", (d_valid[0].cpu().numpy(), d_first[0].cpu().numpy(), ns_dev), want)
        parity_check("end-to-end call", (v_np, fi_np, ns_dev), want)
        value = world * E * args.steps / (elapsed_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(E, world, {
                "merge": ("nccl all_gather of the edge status flags (packed to 2 bits/edge, unpacked to 1 B/edge on "
                          "every rank) per step, overlapped with the next "
                          "step's kernel (double-buffered results); first-invalid slots stay with the shard "
                          "owner, which replays the belief inserts") if world > 1 else "none (single GPU)"}),
            "kernel_ms_per_step": kern_ms_per_step,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_edge": ALGO_BYTES_PER_EDGE,
                         "kernel": "edge_check_fast_kernel<7,false> (+ the general kernel's launch for left-over edges: "
                                   "none on this workload, its CTAs exit at once)",
                         "avg_launch_ms": avg_launch_ms,
                         "note": "instruction-issue bound (DESIGN.md 4.1): ~32 warp instructions per edge at ~70% of the "
                                 "issue slots, shared-memory wavefronts ~64% of peak; DRAM at ~40% of the measured copy peak"},
            "e2e": {"value": world * E * e2e_steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(2 * E * DIM * 8), "d2h_bytes_per_step": int(E * 5),
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "call": "bpomp_edges_check (pinned host buffers; chunked over 3 streams: H2D, kernel, D2H overlap)"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "blocked_fraction": float((fi_np >= 1).mean()),
            "parity_checked_edges": int(checked),
            "parity": "valid, first_invalid_slot and nstates of the first %d edges bit-equal to the CPU oracle, "
                      "device-resident step and end-to-end call, checked in this run before printing" % checked,
        }
        if cb:
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
