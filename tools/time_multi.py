"""P Chamfer predictions against one ground truth: one chamfer_fwd_multi_f32 call vs P chamfer_fwd_loss_f32 calls
(device time, CUDA-graph replay): python tools/time_multi.py P B n m"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from softpool_b200 import _lib

P, B, n, m = (int(v) for v in sys.argv[1:5])
dev = torch.device("cuda:0")
L, p = _lib.lib(), _lib.ptr
g = torch.Generator().manual_seed(0)
preds = (torch.rand(P * B, n, 3, generator=g) - 0.5).to(dev); gt = (torch.rand(B, m, 3, generator=g) - 0.5).to(dev)
d1 = torch.empty(P * B, n, device=dev); d2 = torch.empty(P * B, m, device=dev)
i1 = torch.empty(P * B, n, dtype=torch.int32, device=dev); i2 = torch.empty(P * B, m, dtype=torch.int32, device=dev)
loss = torch.empty(P * B, device=dev)
wsb = int(L.chamfer_fwd_workspace_bytes(P * B, n, m))
ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
stream = torch.cuda.Stream()


def multi():
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(L.chamfer_fwd_multi_f32(p(preds), p(gt), P, B, n, m, p(d1), p(d2), p(i1), p(i2), p(loss), p(ws), wsb, st), "multi")


def separate():
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for q in range(P):
        s = slice(q * B, (q + 1) * B)
        _lib.check(L.chamfer_fwd_loss_f32(p(preds[s]), p(gt), B, n, m, p(d1[s]), p(d2[s]), p(i1[s]), p(i2[s]), p(loss[s]), p(ws), wsb, st), "single")


for name, fn in (("one multi call", multi), ("%d separate calls" % P, separate)):
    with torch.cuda.stream(stream):
        fn(); stream.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=stream):
            for _ in range(10):
                fn()
        gr.replay(); stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            gr.replay()
        e1.record(stream); stream.synchronize()
    print("P=%d B=%d n=%d m=%d %-18s %.2f us" % (P, B, n, m, name + ":", e0.elapsed_time(e1) * 1e3 / 50))
