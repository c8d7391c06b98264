"""How much of the Chamfer pair block could a spatial ordering skip?  (CPU study for the next round, numpy only.)

Points of each cloud are sorted by Morton code and cut into tiles of 128 (the tensor kernel's tile).  A (query tile,
target tile) pair must be evaluated only if the squared distance between the two tiles' bounding boxes does not exceed the
largest nearest-neighbour distance of the tile's queries (optimistic: the FINAL nearest-neighbour distances are used; a real
kernel only has a running bound, so it will skip somewhat less).  Prints the fraction of tile pairs that remain.

    python tools/prune_estimate.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def morton_order(p):
    lo, hi = p.min(0), p.max(0)
    q = np.clip(((p - lo) / np.maximum(hi - lo, 1e-12) * 1023).astype(np.uint64), 0, 1023)
    def spread(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    code = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
    return np.argsort(code, kind="stable")


def remaining_fraction(a, b, tile=128):
    a = a[morton_order(a)]; b = b[morton_order(b)]
    # exact NN distances of the queries (chunked brute force)
    nn = np.empty(len(a))
    for i in range(0, len(a), 1024):
        d = ((a[i:i + 1024, None, :] - b[None, :, :]) ** 2).sum(-1)
        nn[i:i + 1024] = d.min(1)
    qa = [a[i:i + tile] for i in range(0, len(a), tile)]
    tb = [b[i:i + tile] for i in range(0, len(b), tile)]
    qlo = np.array([t.min(0) for t in qa]); qhi = np.array([t.max(0) for t in qa])
    tlo = np.array([t.min(0) for t in tb]); thi = np.array([t.max(0) for t in tb])
    bound = np.array([nn[i:i + tile].max() for i in range(0, len(a), tile)])
    gap = np.maximum(0.0, np.maximum(qlo[:, None, :] - thi[None, :, :], tlo[None, :, :] - qhi[:, None, :]))
    need = (gap ** 2).sum(-1) <= bound[:, None]
    return need.mean()


def main():
    rng = np.random.default_rng(0)
    for n in (2048, 8192, 16384):
        a = rng.random((n, 3)) - 0.5; b = rng.random((n, 3)) - 0.5
        print("uniform cube      %6d <-> %6d : %5.1f %% of the tile pairs remain" % (n, n, 100 * remaining_fraction(a, b)))
    for n in (2048, 8192, 16384):
        def sphere(k):
            v = rng.normal(size=(k, 3)); return 0.5 * v / np.linalg.norm(v, axis=1, keepdims=True) + rng.normal(0, 1e-3, (k, 3))
        print("noisy sphere      %6d <-> %6d : %5.1f %% of the tile pairs remain" % (n, n, 100 * remaining_fraction(sphere(n), sphere(n))))
    g = np.load(os.path.join(ROOT, "tests", "golden", "chamfer_real.npz"))
    for s in range(g["a"].shape[0]):
        print("reference scan %d   %6d <-> %6d : %5.1f %% (scan -> gt), %5.1f %% (gt -> scan)"
              % (s, g["a"].shape[1], g["b"].shape[1], 100 * remaining_fraction(g["a"][s].astype(np.float64), g["b"][s].astype(np.float64)),
                 100 * remaining_fraction(g["b"][s].astype(np.float64), g["a"][s].astype(np.float64))))


if __name__ == "__main__":
    main()
