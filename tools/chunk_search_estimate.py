"""CPU study (numpy) for the hierarchical Chamfer search: how many 16-target chunks / exact distance evaluations
does a query need when the targets are cell-sorted (counting sort on a 2^g-per-axis Morton grid), cut into chunks of
CH consecutive points, and a chunk is evaluated only if the distance from the query to the chunk's bounding box
does not exceed the best distance found so far?  Procedure simulated per query (what the kernel does):
  1. best = distance to target 0 (the reference's initialisation);
  2. evaluate the chunk whose CENTRE is nearest (the tensor-core block gives |p - c_j|^2 for every chunk);
  3. sphere filter with the cloud-wide maximum radius (the packed 16-bit compare), then, in chunk order, the
     per-chunk box test against the running best; evaluate the survivors.

    python tools/chunk_search_estimate.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cell_order(p, g):
    lo, hi = p.min(0), p.max(0)
    G = 1 << g
    q = np.clip(((p - lo) / np.maximum(hi - lo, 1e-12) * G).astype(np.int64), 0, G - 1)
    code = np.zeros(len(p), np.int64)
    for b in range(g):
        for a in range(3):
            code |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return np.argsort(code, kind="stable")


def study(P, Q, g, CH, name):
    """queries P against targets Q"""
    Qs = Q[cell_order(Q, g)]
    Ps = P[cell_order(P, g)]
    m = len(Qs)
    nc = (m + CH - 1) // CH
    pad = nc * CH - m
    Qp = np.concatenate([Qs, np.repeat(Qs[-1:], pad, 0)]) if pad else Qs
    ch = Qp.reshape(nc, CH, 3)
    lo, hi = ch.min(1), ch.max(1)
    c = 0.5 * (lo + hi)
    r = np.sqrt(((ch - c[:, None]) ** 2).sum(2).max(1))
    rmax = r.max()
    n_sphere = n_box = n_eval = 0
    worst_warp = 0
    per_q = []
    for i0 in range(0, len(Ps), 512):
        p = Ps[i0:i0 + 512]
        dc = np.sqrt(((p[:, None] - c[None]) ** 2).sum(2))                   # (q, nc)
        best = ((p - Q[0]) ** 2).sum(1)
        j0 = dc.argmin(1)
        dj0 = ((p[:, None] - ch[j0]) ** 2).sum(2).min(1)
        best = np.minimum(best, dj0)
        s = np.sqrt(best)
        sph = dc <= (s + rmax)[:, None]
        dbox = np.maximum(np.maximum(lo[None] - p[:, None], p[:, None] - hi[None]), 0)
        dbox2 = (dbox ** 2).sum(2)
        cnt = np.ones(len(p), np.int64)
        nb = np.zeros(len(p), np.int64)
        for qi in range(len(p)):
            b = best[qi]
            for j in np.nonzero(sph[qi])[0]:
                if j == j0[qi]:
                    continue
                nb[qi] += 1
                if dbox2[qi, j] <= b:
                    cnt[qi] += 1
                    b = min(b, ((p[qi] - ch[j]) ** 2).sum(1).min())
        n_sphere += sph.sum(); n_box += nb.sum(); n_eval += cnt.sum()
        per_q.append(cnt)
    per_q = np.concatenate(per_q)
    nq = len(Ps)
    w = per_q[: nq // 32 * 32].reshape(-1, 32)
    print("%-34s g=%d CH=%2d chunks=%5d rmax=%.3f: sphere-pass %.1f, box tests %.1f, chunks evaluated %.2f (= %.0f distances) per query; "
          "warp max/mean %.2f" % (name, g, CH, nc, rmax, n_sphere / nq, n_box / nq, n_eval / nq, n_eval / nq * CH, w.max(1).mean() / per_q.mean()))


def main():
    rng = np.random.default_rng(0)
    for n in (2048, 8192):
        a = rng.random((n, 3), dtype=np.float32) - 0.5
        b = rng.random((n, 3), dtype=np.float32) - 0.5
        for g, CH in ((4, 16), (4, 8), (5, 16)):
            study(a, b, g, CH, "uniform cube %d<->%d" % (n, n))
    u = rng.standard_normal((8192, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    v = rng.standard_normal((8192, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    study((u * 0.5 + rng.normal(0, 0.005, u.shape)).astype(np.float32), (v * 0.5).astype(np.float32), 4, 16, "noisy sphere 8192<->8192")
    study((u * 0.5 + rng.normal(0, 0.005, u.shape)).astype(np.float32), (v * 0.5).astype(np.float32), 5, 16, "noisy sphere 8192<->8192")
    f = os.path.join(ROOT, "tests", "golden", "chamfer_real.npz")
    if os.path.exists(f):
        z = np.load(f)
        for g in (4, 5):
            study(z["a"][0], z["b"][0], g, 16, "reference scan 2048->16384")
            study(z["b"][0], z["a"][0], g, 16, "reference scan 16384->2048")


if __name__ == "__main__":
    sys.exit(main())
