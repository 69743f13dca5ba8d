#!/usr/bin/env python
"""CPU model of the shared-memory wavefronts of the broad-phase mask-row lookups of
edge_check_boxes_kernel on the config-4 workload.  A 16-byte shared-memory load is served
per quarter-warp; lanes of a quarter that read different addresses in the same bank group (row mod 8)
serialise.  The model reproduces the ncu counters of the round-1 kernels (span-minor rows 11.3 wavefronts
per LDS.128, c0-minor 8.4) and is used to rank layouts before spending GPU time.

    python benchmarks/sim_mask_row_banks.py [tiles]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from batching_pomp_b200 import workloads as wl  # noqa: E402


def main():
    tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    d = 7
    lo, hi = wl.box_world(d, 500, seed=3)
    L = wl.segment_length(d)
    E = 32 * tiles
    f, t = wl.random_edges(E, d, seed=4, max_len=wl.halton_radius(10 ** 6, d))
    dist = np.sqrt(((f - t) ** 2).sum(1))
    need = np.maximum(2, np.floor(dist / L)) > 2
    g0 = lo.min(0)
    gs = 32 / (hi.max(0) - g0)

    def cell(x):
        return np.clip(np.floor((x - g0) * gs), 0, 31).astype(int)

    cf, ct = cell(f), cell(t)
    c0 = np.minimum(cf, ct)
    span = np.maximum(cf, ct) - c0

    layouts = {
        "span-minor  (j*32+c0)*8+span": lambda j, c, s: (j * 32 + c) * 8 + s,
        "c0-minor    (j*8+span)*32+c0": lambda j, c, s: (j * 8 + s) * 32 + c,
        "rotated     (j*8+span)*32+((c0+3span)&31)": lambda j, c, s: (j * 8 + s) * 32 + ((c + 3 * s) & 31),
    }
    print("lane = edge, one LDS.128 per (dimension, 128-box group): wavefronts per instruction")
    per_tile = {}
    for name, row in layouts.items():
        tot = cnt = 0
        for w in range(0, E, 32):
            act = need[w:w + 32]
            for j in range(d):
                rows = row(j, c0[w:w + 32, j], span[w:w + 32, j])
                for q in range(4):
                    r = np.unique(rows[q * 8:(q + 1) * 8][act[q * 8:(q + 1) * 8]])  # same address: broadcast
                    if r.size:
                        tot += np.bincount(r % 8, minlength=8).max()
                cnt += 1
        per_tile[name] = 4 * d * tot / cnt
        print("  %-46s %.2f   (%.0f wavefronts per 32 edges)" % (name, tot / cnt, per_tile[name]))

    # team of 4 lanes per edge, table row-major [rows][4 groups] (64 B per row): a quarter-warp holds two
    # edges; their rows collide when they are different rows of the same parity
    row = layouts["rotated     (j*8+span)*32+((c0+3span)&31)"]
    idx = np.nonzero(need)[0]
    tot = cnt = 0
    for w in range(0, len(idx) - 8, 8):
        e = idx[w:w + 8]
        for j in range(d):
            rows = row(j, c0[e, j], span[e, j])
            for q in range(4):
                r = np.unique(rows[q * 2:q * 2 + 2])
                tot += 1 if (r.size == 1 or (r[0] % 2) != (r[1] % 2)) else 2
            cnt += 1
    instr = need.mean() * 32 / 8
    print("team of 4 lanes per edge, row-major rows: %.2f wavefronts per instruction, %.1f instructions per 32 "
          "edges and dimension -> %.0f wavefronts per 32 edges (ideal %.0f)"
          % (tot / cnt, instr, d * instr * tot / cnt, need.mean() * 32 * d * 64 / 128))


if __name__ == "__main__":
    main()
