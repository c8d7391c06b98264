"""Device time of chamfer_fwd_f32 for one shape: python tools/time_chamfer.py B n m  (CUDA-graph replay, 3 rotating sets)"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from softpool_b200 import _lib

B, n, m = (int(v) for v in sys.argv[1:4])
dev = torch.device("cuda:0")
L, p = _lib.lib(), _lib.ptr
g = torch.Generator().manual_seed(0)
sets = []
for i in range(3):
    a = (torch.rand(B, n, 3, generator=g) - 0.5).to(dev); b = (torch.rand(B, m, 3, generator=g) - 0.5).to(dev)
    sets.append((a, b, torch.empty(B, n, device=dev), torch.empty(B, m, device=dev),
                 torch.empty(B, n, dtype=torch.int32, device=dev), torch.empty(B, m, dtype=torch.int32, device=dev)))
wsb = int(L.chamfer_fwd_workspace_bytes(B, n, m))
ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
stream = torch.cuda.Stream()


def call(s):
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(L.chamfer_fwd_f32(p(s[0]), p(s[1]), B, n, m, p(s[2]), p(s[3]), p(s[4]), p(s[5]), p(ws), wsb, st), "chamfer_fwd_f32")


with torch.cuda.stream(stream):
    for s in sets:
        call(s)
    stream.synchronize()
    gr = torch.cuda.CUDAGraph()
    reps = 30
    with torch.cuda.graph(gr, stream=stream):
        for i in range(reps):
            call(sets[i % 3])
    gr.replay(); stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5):
        gr.replay()
    e1.record(stream); stream.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (5 * reps)
print("chamfer_fwd_f32 B=%d n=%d m=%d: %.2f us  (%.1f Gpairs/s)" % (B, n, m, us, B * n * m / us / 1e3))
