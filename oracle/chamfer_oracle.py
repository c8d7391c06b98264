"""ctypes front end of oracle/chamfer_oracle.c (CPU restatement of the reference Chamfer kernels,
distance/chamfer/chamfer.cu:12-134,155-174).  TEST INFRASTRUCTURE ONLY -- see the C file's header
for what it restates and how it is pinned."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libchamfer_oracle.so")
_h = None


def build():
    src = os.path.join(_HERE, "chamfer_oracle.c")
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB


def _lib():
    global _h
    if _h is None:
        _h = ctypes.CDLL(build())
    return _h


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def default_threads():
    return max(1, len(os.sched_getaffinity(0)))


def forward(xyz1, xyz2, nthreads=None):
    """(B,n,3), (B,m,3) float32 -> dist1 (B,n) f32, dist2 (B,m) f32, idx1 (B,n) i32, idx2 (B,m) i32."""
    xyz1 = np.ascontiguousarray(xyz1, np.float32)
    xyz2 = np.ascontiguousarray(xyz2, np.float32)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist1 = np.zeros((B, n), np.float32); dist2 = np.zeros((B, m), np.float32)
    idx1 = np.zeros((B, n), np.int32); idx2 = np.zeros((B, m), np.int32)
    _lib().chamfer_oracle_forward(B, n, m, _p(xyz1), _p(xyz2), _p(dist1), _p(dist2), _p(idx1), _p(idx2),
                                  int(nthreads or default_threads()))
    return dist1, dist2, idx1, idx2


def backward(xyz1, xyz2, g1, g2, idx1, idx2, nthreads=None):
    """-> grad_xyz1 (B,n,3), grad_xyz2 (B,m,3)."""
    xyz1 = np.ascontiguousarray(xyz1, np.float32); xyz2 = np.ascontiguousarray(xyz2, np.float32)
    g1 = np.ascontiguousarray(g1, np.float32); g2 = np.ascontiguousarray(g2, np.float32)
    idx1 = np.ascontiguousarray(idx1, np.int32); idx2 = np.ascontiguousarray(idx2, np.int32)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    gx1 = np.empty((B, n, 3), np.float32); gx2 = np.empty((B, m, 3), np.float32)
    _lib().chamfer_oracle_backward(B, n, m, _p(xyz1), _p(xyz2), _p(g1), _p(g2), _p(idx1), _p(idx2), _p(gx1), _p(gx2),
                                   int(nthreads or default_threads()))
    return gx1, gx2
