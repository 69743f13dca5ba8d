#!/usr/bin/env python
"""Edge-kernel tuning run (config 4): general kernel vs the fast kernel with / without the state-space
bounds hint and with different warp counts; every variant is first compared with the CPU oracle.

  python benchmarks/edge_micro.py [--edges E] [--check M] [--out FILE]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from batching_pomp_b200 import capi, workloads  # noqa: E402


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--edges", type=int, default=10 ** 7)
    ap.add_argument("--check", type=int, default=10 ** 6)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--out", default=None)
    ap.add_argument("--image", type=int, default=0, help="N > 0: time the image-world kernel on all edges (u < v) of an "
                    "N-vertex 2-D Halton roadmap against the 4096^2 image (config 2), indexed")
    ap.add_argument("--indexed", type=int, default=0, help="N > 0: also time the indexed form on r-disk rows of an "
                    "N-vertex 7-D Halton roadmap (the planner's edge path, config 3)")
    ap.add_argument("--variants", default="general,fast,fast+bounds,fast+bounds/8,fast+bounds/12,fast+bounds/16,fast+bounds/20")
    args = ap.parse_args()
    from oracle import oracle
    d = 7
    E = args.edges
    L = workloads.segment_length(d)
    lo, hi = workloads.box_world(d, 500, seed=3)
    frm, to = workloads.random_edges(E, d, seed=4, max_len=workloads.halton_radius(10 ** 6, d))
    m = min(args.check, E)
    t0 = time.perf_counter()
    ov, of, on = oracle.edges_check_boxes(frm[:m], to[:m], L, lo, hi, nthreads=oracle.max_threads())
    print("oracle: %d edges in %.2f s" % (m, time.perf_counter() - t0), flush=True)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    tf, tt = torch.from_numpy(frm).to(dev), torch.from_numpy(to).to(dev)
    tv = torch.empty(E, dtype=torch.uint8, device=dev)
    tfi = torch.empty(E, dtype=torch.int32, device=dev)
    tn = torch.empty(E, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    results = []
    for var in args.variants.split(","):
        ctx = capi.Context(d, L)
        name, _, warps = var.partition("/")
        if "bounds" in name:
            ctx.set_bounds(0.0, 1.0)
        ctx.set_world_boxes(lo, hi)
        ctx.set_option("fast_edge_kernel", 0 if name == "general" else 1)
        if warps:
            ctx.set_option("fast_edge_warps", int(warps))
        ctx.set_stream(stream.cuda_stream)

        def run():
            ctx.edges_check_dev(tf.data_ptr(), tt.data_ptr(), E, tv.data_ptr(), tfi.data_ptr(), tn.data_ptr())
        tv.fill_(9); tfi.fill_(-9); tn.fill_(0)
        run()
        ctx.sync()
        slow = ctx.counter("slow_edges")
        ok = (np.array_equal(tv[:m].cpu().numpy(), ov) and np.array_equal(tfi[:m].cpu().numpy(), of)
              and np.array_equal(tn[:m].cpu().numpy().view(np.uint32), on))
        if not ok:
            gv, gf = tv[:m].cpu().numpy(), tfi[:m].cpu().numpy()
            bad = np.nonzero((gv != ov) | (gf != of))[0]
            print("PARITY FAILURE in variant %s: %d edges differ, first %s: gpu (%d,%d) oracle (%d,%d) n=%d" % (
                var, bad.size, bad[:5], gv[bad[0]], gf[bad[0]], ov[bad[0]], of[bad[0]], on[bad[0]]), flush=True)
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.reps):
            run()
        e1.record(stream)
        ctx.sync()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        rec = {"variant": var, "edges": E, "ms": ms, "edges_per_s": E / ms * 1e3, "frac_hbm": 117 * E / ms * 1e-6 / 6550.4,
               "slow_edges": slow, "parity_ok": bool(ok), "checked": m, "fast_world": ctx.counter("fast_world")}
        print(json.dumps(rec), flush=True)
        results.append(rec)
        ctx.close()
    if args.image:
        N, d2, size, frac = args.image, 2, 4096, 1e-4
        v2 = workloads.halton(N, d2)
        r2 = workloads.halton_radius(N, d2)
        L2 = workloads.segment_length(d2, frac)
        pix = workloads.image_world(size, seed=2)
        ctx = capi.Context(d2, L2)
        ctx.set_world_image(pix)
        ctx.vertices_append(v2)
        rp, col, dist = ctx.neighbors_all(r2)
        rows = np.repeat(np.arange(N, dtype=np.uint32), np.diff(rp.astype(np.int64)))
        keep = rows < col
        fi, ti = np.ascontiguousarray(rows[keep]), np.ascontiguousarray(col[keep])
        Ei = fi.size
        tfi_, tti_ = torch.from_numpy(fi.astype(np.int32)).to(dev), torch.from_numpy(ti.astype(np.int32)).to(dev)
        iv = torch.empty(Ei, dtype=torch.uint8, device=dev)
        ifi = torch.empty(Ei, dtype=torch.int32, device=dev)
        inn = torch.empty(Ei, dtype=torch.int32, device=dev)
        mm = min(m, Ei)
        ov_, of_, on_ = oracle.edges_check_image(v2[fi[:mm]], v2[ti[:mm]], L2, pix, nthreads=oracle.max_threads())
        ctx.set_stream(stream.cuda_stream)

        def runim():
            ctx.edges_check_indexed_dev(tfi_.data_ptr(), tti_.data_ptr(), Ei, iv.data_ptr(), ifi.data_ptr(), inn.data_ptr())
        runim()
        ctx.sync()
        ok = (np.array_equal(iv[:mm].cpu().numpy(), ov_) and np.array_equal(ifi[:mm].cpu().numpy(), of_)
              and np.array_equal(inn[:mm].cpu().numpy().view(np.uint32), on_))
        for _ in range(20):
            runim()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.reps):
            runim()
        e1.record(stream)
        ctx.sync()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        rec = {"variant": "image:indexed", "edges": int(Ei), "ms": ms, "edges_per_s": Ei / ms * 1e3,
               "frac_hbm_13B": (13 * Ei + N * 16) / ms * 1e-6 / 6550.4, "parity_ok": bool(ok),
               "blocked": float((iv == 0).float().mean()), "mean_n": float(inn.float().mean())}
        print(json.dumps(rec), flush=True)
        results.append(rec)
        ctx.close()
    if args.indexed:
        N = args.indexed
        verts = workloads.halton(N, d)
        r = workloads.halton_radius(N, d)
        for var in ("general", "fast+bounds"):
            ctx = capi.Context(d, L)
            if "bounds" in var:
                ctx.set_bounds(0.0, 1.0)
            ctx.set_world_boxes(lo, hi)
            ctx.set_option("fast_edge_kernel", 0 if var == "general" else 1)
            ctx.vertices_append(verts)
            nq = 4096
            q = (np.arange(nq, dtype=np.uint64) * (N // nq)).astype(np.uint32)
            rp, col, dist = ctx.neighbors_radius(q, r)
            fi = np.repeat(q, np.diff(rp.astype(np.int64)))
            keep = fi != col
            fi, ti = np.ascontiguousarray(fi[keep]), np.ascontiguousarray(col[keep])
            Ei = fi.size
            tfi_, tti_ = torch.from_numpy(fi.astype(np.int32)).to(dev), torch.from_numpy(ti.astype(np.int32)).to(dev)
            ov_, of_, on_ = oracle.edges_check_boxes(verts[fi[:m]], verts[ti[:m]], L, lo, hi, nthreads=oracle.max_threads())
            ctx.set_stream(stream.cuda_stream)

            def runi():
                ctx.edges_check_indexed_dev(tfi_.data_ptr(), tti_.data_ptr(), Ei, tv.data_ptr(), tfi.data_ptr(), tn.data_ptr())
            runi()
            ctx.sync()
            slow = ctx.counter("slow_edges")
            mm = min(m, Ei)
            ok = (np.array_equal(tv[:mm].cpu().numpy(), ov_[:mm]) and np.array_equal(tfi[:mm].cpu().numpy(), of_[:mm])
                  and np.array_equal(tn[:mm].cpu().numpy().view(np.uint32), on_[:mm]))
            for _ in range(20):
                runi()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.reps):
                runi()
            e1.record(stream)
            ctx.sync()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            rec = {"variant": "indexed:" + var, "edges": int(Ei), "ms": ms, "edges_per_s": Ei / ms * 1e3,
                   "frac_hbm_13B": 13 * Ei / ms * 1e-6 / 6550.4, "slow_edges": slow, "parity_ok": bool(ok),
                   "blocked": float((tv[:Ei] == 0).float().mean())}
            print(json.dumps(rec), flush=True)
            results.append(rec)
            ctx.close()
    if args.out:
        with open(args.out, "w") as f:
            for r in results:
                f.write(json.dumps(r) + "\n")
    return 0 if all(r["parity_ok"] for r in results) else 1


if __name__ == "__main__":
    sys.exit(main())
