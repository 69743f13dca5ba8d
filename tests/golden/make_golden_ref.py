#!/usr/bin/env python
"""Golden vectors taken from the REFERENCE ITSELF: BisectPerm::get(n) of
/root/reference/include/batching_pomp/util/BisectPerm.hpp, compiled by oracle/Makefile (target `ref`,
oracle/ref_shim.cpp) into oracle/_ref/libbpomp_ref.so.  Run in the build container (the reference is not
on the GPU box); the .npz travels.

    python tests/golden/make_golden_ref.py        # writes tests/golden/bisect_perm_ref.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

NS = list(range(0, 257)) + [511, 512, 513, 1000, 4097, 10007]


def main():
    if not oracle.ref_available():
        raise SystemExit("oracle/_ref/libbpomp_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    first, second, offs = [], [], [0]
    for n in NS:
        f, s = oracle.ref_bisect_perm(n)
        first.append(f)
        second.append(s)
        offs.append(offs[-1] + f.size)
    out = os.path.join(ROOT, "tests", "golden", "bisect_perm_ref.npz")
    np.savez_compressed(out, n=np.asarray(NS, dtype=np.int32), offsets=np.asarray(offs, dtype=np.int64),
                        first=np.concatenate(first).astype(np.int32), second=np.concatenate(second).astype(np.int32))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
