"""Synthetic inputs for the BASELINE.json configs, numpy only.

PRNG is splitmix64 so a C++ caller can regenerate the same streams.  Everything lives on the
unit hypercube [0,1]^d.
"""
import math

import numpy as np

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)
PRIMES = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53)


def splitmix64(seed, count):
    """count 64-bit outputs of splitmix64 seeded with `seed` (vectorised)."""
    with np.errstate(over="ignore"):
        i = np.arange(1, count + 1, dtype=np.uint64)
        z = np.uint64(seed) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform(seed, shape):
    n = int(np.prod(shape))
    return ((splitmix64(seed, n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)).reshape(shape)


def normal(seed, shape):
    n = int(np.prod(shape))
    u = uniform(seed, (2, n))
    r = np.sqrt(-2.0 * np.log(1.0 - u[0]))
    return (r * np.cos(2.0 * math.pi * u[1])).reshape(shape)


def halton(count, d, first_index=1):
    """Halton points i = first_index .. first_index+count-1, coordinate j = radical inverse in base PRIMES[j].

    Same arithmetic as oracle orc_halton (f = f/base; r = r + f*digit), so results are bit-equal.
    """
    out = np.zeros((count, d), dtype=np.float64)
    idx0 = np.arange(first_index, first_index + count, dtype=np.int64)
    for j in range(d):
        base = PRIMES[j]
        idx = idx0.copy()
        f = np.ones(count, dtype=np.float64)
        r = np.zeros(count, dtype=np.float64)
        while np.any(idx > 0):
            live = idx > 0
            f = np.where(live, f / float(base), f)
            r = np.where(live, r + f * (idx % base).astype(np.float64), r)
            idx = idx // base
        out[:, j] = r
    return out


def segment_length(d, fraction=0.01):
    """OMPL getLongestValidSegmentLength() on the unit hypercube: maxExtent * fraction."""
    return max_extent(d) * fraction


def max_extent(d):
    e = 0.0
    for _ in range(d):
        delta = 1.0 - 0.0
        e += delta * delta
    return math.sqrt(e)


def image_world(size, seed, nrect=40):
    """size x size u8 occupancy image, white (255) = free, with `nrect` black axis-aligned rectangles.

    Rectangle sides ~U[0.04, 0.16] of the image (≈30 % black after overlaps); a margin around
    (0.1,0.1) and (0.9,0.9) is kept free so the README start/goal are valid.
    """
    u = uniform(seed, (nrect, 4))
    img = np.full((size, size), 255, dtype=np.uint8)
    for k in range(nrect):
        w = 0.04 + 0.12 * u[k, 2]
        h = 0.04 + 0.12 * u[k, 3]
        x0 = u[k, 0] * (1.0 - w)
        y0 = u[k, 1] * (1.0 - h)
        c0, c1 = int(x0 * size), int((x0 + w) * size)
        r0, r1 = int(y0 * size), int((y0 + h) * size)
        img[r0:r1 + 1, c0:c1 + 1] = 0
    for cx, cy in ((0.1, 0.1), (0.9, 0.9)):
        m = max(2, size // 32)
        c, r = int(cx * size), int(cy * size)
        img[max(0, r - m):r + m + 1, max(0, c - m):c + m + 1] = 255
    return img


def box_world(d, nboxes, seed, keep_free=((0.1,), (0.9,))):
    """nboxes axis-aligned boxes: centre ~U[0,1]^d, half-extent per dim ~U[0.10,0.25]; boxes that
    contain any `keep_free` point (broadcast to d dims) are dropped."""
    u = uniform(seed, (nboxes, 2, d))
    c = u[:, 0, :]
    h = 0.10 + 0.15 * u[:, 1, :]
    lo, hi = c - h, c + h
    keep = np.ones(nboxes, dtype=bool)
    for p in keep_free:
        pt = np.broadcast_to(np.asarray(p, dtype=np.float64), (d,))
        keep &= ~np.all((lo <= pt) & (pt <= hi), axis=1)
    return np.ascontiguousarray(lo[keep]), np.ascontiguousarray(hi[keep])


def random_edges(count, d, seed, max_len):
    """Config-4 edges: from ~U[0,1]^d; to = clamp(from + l*u), u uniform on the sphere, l ~U[0,max_len]."""
    frm = uniform(seed, (count, d))
    g = normal(seed + 1, (count, d))
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    ell = uniform(seed + 2, (count, 1)) * max_len
    to = np.clip(frm + ell * g, 0.0, 1.0)
    return np.ascontiguousarray(frm), np.ascontiguousarray(to)


def halton_radius(n, d):
    """haltonRadiusFun, src/BatchingPOMP.cpp:548-556."""
    return 2.5 * math.pow(1.0 / float(n), 1 / float(d))
