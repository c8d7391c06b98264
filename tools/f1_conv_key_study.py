"""could the Sorter's 1x1 conv (softpool.py:94) be fused into the top-k front end?  Only if the keys it
produces select the same points as the library conv the reference runs.  This study (GPU) compares the top-k index lists
that different evaluations of the SAME conv produce on config A / A' inputs:
  default   F.conv1d as the reference calls it (cuDNN, TF32 allowed by default on this GPU generation)
  fp32      F.conv1d with TF32 disabled (cuDNN / cuBLAS fp32)
  bmm       torch.bmm in fp32 (cuBLAS, another summation order)
  fma       a sequential fp32 FMA chain over the channels (what a fused kernel would compute), emulated with a loop
  exact     float64 accumulation rounded to fp32
Prints, per pair, the fraction of (sample, region) rows whose k selected indices are identical and the fraction of
positions that agree.     python tools/f1_conv_key_study.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from softpool_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, C, N, R = 32, 256, 2048, 8
x = torch.randn(B, C, N, device=dev)
conv = torch.nn.Conv1d(C, R, 1).to(dev)
w, b = conv.weight.detach(), conv.bias.detach()
keys = {}
with torch.no_grad():
    keys["default"] = F.conv1d(x, w, b)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    keys["fp32"] = F.conv1d(x, w, b)
    keys["bmm"] = torch.bmm(w[:, :, 0].unsqueeze(0).expand(B, R, C), x) + b[None, :, None]
    acc = b[None, :, None].expand(B, R, N).clone()
    for c in range(C):                                  # acc = fma(w[r,c], x[b,c,n], acc), channels ascending
        acc = torch.addcmul(acc, w[None, :, c, :], x[:, c:c + 1, :])
    keys["fma"] = acc
    keys["exact"] = (torch.einsum("rc,bcn->brn", w[:, :, 0].double(), x.double()) + b.double()[None, :, None]).float()
print("cudnn.allow_tf32 default on this image:", True)
for k in (32, 256):
    idx = {name: ops.softpool_topk(v.contiguous(), k, want_sp_idx=False, want_id_activa=False)[0] for name, v in keys.items()}
    print("k = %d" % k)
    for a, bb in (("default", "fp32"), ("default", "exact"), ("fp32", "bmm"), ("fp32", "fma"), ("fp32", "exact"), ("fma", "exact")):
        same_rows = (idx[a] == idx[bb]).all(-1).float().mean().item()
        same_pos = (idx[a] == idx[bb]).float().mean().item()
        same_set = torch.stack([(torch.sort(idx[a], -1)[0] == torch.sort(idx[bb], -1)[0]).all(-1)]).float().mean().item()
        print("  %-8s vs %-6s rows identical %.3f, same selected SET %.3f, positions equal %.3f, max |key diff| %.2e" %
              (a, bb, same_rows, same_set, same_pos, (keys[a] - keys[bb]).abs().max().item()))
