"""Generate tests/golden/chamfer_real.npz: realistic geometry for the Chamfer kernels.

Inputs are the reference's own result artefacts (ASCII PLY written by val.py:365-440 into /root/reference/pcds/):
partial input scans (2048 points, `pcds/input/<cat>/<id>.ply`) against their ground-truth shapes (16384 points,
`pcds/gt/<cat>/<id>.ply`) -- the validation shape of val.py:302 -- and two network outputs (`pcds/all_clouds1`).
Surfaces, not uniform noise: clustered nearest neighbours, exact duplicates from resample_pcd (dataset.py:129-135).
Expected outputs come from the C restatement of the reference kernel (oracle/chamfer_oracle.c), which is itself
pinned against the reference's chamfer.cu compiled unmodified (tests/golden/chamfer_ref_*.npz).  /root/reference
does not exist on the GPU box, hence the committed fixture.

    python tests/golden/make_chamfer_real.py
"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import chamfer_oracle as co   # noqa: E402

REF = "/root/reference/pcds"


def read_ply(path):
    with open(path) as f:
        n = 0
        for line in f:
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            if line.startswith("end_header"):
                break
        pts = np.loadtxt(f, usecols=(0, 1, 2), max_rows=n, dtype=np.float64)
    return pts.astype(np.float32)


def main():
    ids = sorted(glob.glob(os.path.join(REF, "input", "*", "*.ply")))[:12:4]          # three categories' first shapes
    a = np.stack([read_ply(p) for p in ids])                                           # (3, 2048, 3)
    b = np.stack([read_ply(p.replace("/input/", "/gt/")) for p in ids])                # (3, 16384, 3)
    outs = sorted(glob.glob(os.path.join(REF, "all_clouds1", "*", "*-1.ply")))[:3]
    c = np.stack([read_ply(p) for p in outs])                                          # network outputs vs the same gt rows
    d1, d2, i1, i2 = co.forward(a, b)
    e1, e2, j1, j2 = co.forward(c, b[:, :4096].copy())
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "chamfer_real.npz"),
                        a=a, b=b, c=c, d1=d1, d2=d2, i1=i1, i2=i2, e1=e1, e2=e2, j1=j1, j2=j2,
                        files=np.array([os.path.relpath(p, REF) for p in ids + outs]))
    print("a", a.shape, "b", b.shape, "c", c.shape, "dup points in a:", [len(x) - len(np.unique(x, axis=0)) for x in a])


if __name__ == "__main__":
    main()
