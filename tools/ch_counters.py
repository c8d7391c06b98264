"""Filter statistics of the Chamfer search kernel (needs a build with SPK_NVCC_EXTRA=-DSPK_TIMING):
python tools/ch_counters.py B n m  -> per query and pass: mask bits, box tests, chunk evaluations."""
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
a = (torch.rand(B, n, 3, generator=g) - 0.5).to(dev); b = (torch.rand(B, m, 3, generator=g) - 0.5).to(dev)
d1 = torch.empty(B, n, device=dev); d2 = torch.empty(B, m, device=dev)
i1 = torch.empty(B, n, dtype=torch.int32, device=dev); i2 = torch.empty(B, m, dtype=torch.int32, device=dev)
wsb = int(L.chamfer_fwd_workspace_bytes(B, n, m))
ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
out = (ctypes.c_ulonglong * 8)()
L.spk_debug_tc_counters(out, 1)
st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
_lib.check(L.chamfer_fwd_f32(p(a), p(b), B, n, m, p(d1), p(d2), p(i1), p(i2), p(ws), wsb, st), "chamfer_fwd_f32")
L.spk_debug_tc_counters(out, 0)
q = max(out[0], 1)
print("B=%d n=%d m=%d: query-passes %d; per query-pass: step-1 mask bits %.2f, step-2 mask bits %.2f, box tests %.2f, chunk evaluations %.2f"
      % (B, n, m, out[0], out[1] / q, out[2] / q, out[3] / q, out[4] / q))
