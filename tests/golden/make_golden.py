"""Generate tests/golden/softpool_*.npz by running the UNMODIFIED reference `softpool.py`
(/root/reference/softpool.py, imported in the build container; it cannot travel to the GPU box).

The reference hard-codes `.cuda()` (softpool.py:24,30,41-42,48,76,91,109-123,135,137); this
container has no GPU, so `.cuda()` is neutralised to the identity BEFORE the import -- no line of
the reference is changed.  Each fixture stores the inputs, the keys the reference's own Sorter
conv produced (`val_activa`), the four forward outputs (softpool.py:171) and autograd's grad_x for
fixed upstream gradients.  For the tie / NaN / +-0 cases the Sorter's conv is swapped for a stub
that returns preset keys, so that the region loop under test (softpool.py:139-151) sees exactly
those keys; `Sorter.forward`, `SoftPool.forward` and `train2cabins` stay the reference's code.

Tie order.  The reference calls `torch.sort(..., descending=True)` with the default stable=False
(softpool.py:140), so the order of EQUAL keys is formally unspecified; on this image's CPU build
(AVX-512 x86-simd-sort path) it is deterministic but not stable.  The library's contract is the
stable order (ties keep ascending index), i.e. the reference run with stable=True.  Every case is
therefore run twice: unmodified (arrays `sp_idx`, `sp_cube`, ...) and with `torch.sort` forced to
stable=True (arrays `st_*`).  `tie_free` = 1 when both runs agree bit for bit (then the unmodified
reference pins everything); otherwise the unmodified arrays pin the selected key VALUES and, given
its indices, the gather / window-max / backward, and the `st_*` arrays pin the tie order.

    python tests/golden/make_golden.py      # rewrites the .npz files next to this script
"""
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference/softpool.py"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    spec = importlib.util.spec_from_file_location("ref_softpool", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class PresetKeys(nn.Module):
    def __init__(self, keys):
        super().__init__()
        self.keys = keys

    def forward(self, x):
        return self.keys


_orig_sort = torch.sort


def _stable_sort(*a, **k):
    k["stable"] = True
    return _orig_sort(*a, **k)


def run_once(ref, B, C, N, R, sp_ratio, cab, seed, keys_fn, x_fn):
    torch.manual_seed(seed)
    m = ref.SoftPool(regions=R, cabins=cab, sp_ratio=sp_ratio, size_feat=C)
    x = torch.randn(B, C, N)
    if x_fn is not None:
        x = x_fn(x)
    x.requires_grad_(True)
    if keys_fn is not None:
        keys = keys_fn(torch.randn(B, R, N))
        m.sorter.conv1d = PresetKeys(keys)
    sp_cube, sp_idx, cabins, id_activa = m(x)
    with torch.no_grad():
        val_activa = m.sorter.conv1d(x).detach().clone()
    g_cube = torch.randn_like(sp_cube)
    g_cabins = torch.randn_like(cabins)
    ((sp_cube * g_cube).sum() + (cabins * g_cabins).sum()).backward()
    return dict(x=x.detach().numpy(), keys=val_activa.numpy(),
                sp_cube=sp_cube.detach().numpy(), sp_idx=sp_idx.detach().numpy(),
                cabins=cabins.detach().numpy(), id_activa=id_activa.numpy(),
                g_cube=g_cube.numpy(), g_cabins=g_cabins.numpy(), grad_x=x.grad.numpy())


def run_case(ref, name, B, C, N, R, sp_ratio, cab, seed, keys_fn=None, x_fn=None):
    torch.sort = _orig_sort
    a = run_once(ref, B, C, N, R, sp_ratio, cab, seed, keys_fn, x_fn)
    torch.sort = _stable_sort
    s = run_once(ref, B, C, N, R, sp_ratio, cab, seed, keys_fn, x_fn)
    torch.sort = _orig_sort
    for f in ("x", "keys", "g_cube", "g_cabins", "id_activa"):
        assert np.array_equal(a[f].view(np.uint8), s[f].view(np.uint8)), f
    tie_free = all(np.array_equal(a[f].view(np.uint8), s[f].view(np.uint8)) for f in a)
    k = N // sp_ratio
    out = dict(a)
    if not tie_free:
        out.update({"st_" + f: s[f] for f in ("sp_cube", "sp_idx", "cabins", "grad_x")})
    np.savez_compressed(
        os.path.join(OUT, "softpool_%s.npz" % name),
        meta=np.array([B, C, N, R, sp_ratio, cab, k, int(tie_free)], dtype=np.int64), **out)
    print("%-10s B=%d C=%d N=%d R=%d k=%d cab=%d  tie_free=%d" % (name, B, C, N, R, k, cab, tie_free))


def special_keys(keys):
    keys = torch.round(keys * 2) / 2                      # heavy ties
    keys[0, 0, :] = 1.25                                  # an all-equal row
    keys[0, 1, 5] = float("nan"); keys[0, 1, 77] = float("nan")
    keys[0, 1, 9] = float("inf"); keys[0, 1, 10] = float("-inf"); keys[0, 1, 11] = float("inf")
    keys[0, 2, ::2] = 0.0; keys[0, 2, 1::2] = -0.0        # +0 / -0 must tie
    keys[1, 0, 3] = float("nan")                          # NaN meets argmax over regions
    keys[1, 3, 3] = float("nan")
    return keys


def special_x(x):
    x = torch.round(x * 2) / 2                            # ties inside the max windows
    x[0, 0, :7] = float("nan")                            # NaN in a window: torch.max returns it
    x[0, 1, ::3] = -0.0
    return x


def run_feat(ref, name, B, N, R, sp_ratio, seed):
    """Reference SoftPoolFeat (softpool.py:174-241): PointNet MLP + BatchNorm (train mode, as the reference trains) + SoftPool
    + the index bookkeeping; stores the live weights (conv1-3, bn1-3, sorter.conv1d), the input, the MLP output, the keys and
    the three returned tensors."""
    torch.sort = _orig_sort
    torch.manual_seed(seed)
    m = ref.SoftPoolFeat(num_points=N, regions=R, sp_points=N, sp_ratio=sp_ratio)
    x = torch.rand(B, 3, N) - 0.5
    with torch.no_grad():
        feat = m.mlp(x)
        keys = m.softpool.sorter.conv1d(feat)
        sp_cube, cabins, sp_idx = m(x)
    sd = {k: v.numpy() for k, v in m.state_dict().items() if not k.startswith("softpool.conv2d_") and "num_batches" not in k
          and "running_" not in k}
    np.savez_compressed(os.path.join(OUT, "softpool_%s.npz" % name), meta=np.array([B, N, R, sp_ratio], dtype=np.int64),
                        x=x.numpy(), feat=feat.numpy(), keys=keys.numpy(), sp_cube=sp_cube.numpy(), cabins=cabins.numpy(),
                        sp_idx=sp_idx.numpy(), **{"w_" + k: v for k, v in sd.items()})
    print("%-10s SoftPoolFeat B=%d N=%d R=%d sp_ratio=%d -> sp_cube %s cabins %s sp_idx %s" %
          (name, B, N, R, sp_ratio, tuple(sp_cube.shape), tuple(cabins.shape), tuple(sp_idx.shape)))


def main():
    ref = load_reference()
    run_feat(ref, "feat", 2, 128, 8, 8, seed=6)
    # BASELINE.json config 1: (B=4, N=512, C=32), R=8, sp_ratio=8 -> k=64
    run_case(ref, "c1", 4, 32, 512, 8, 8, 8, seed=0)
    # ties, NaN, +-inf, +-0 in the keys and in the features
    run_case(ref, "ties", 2, 8, 256, 4, 4, 8, seed=1, keys_fn=special_keys, x_fn=special_x)
    # N not a power of two / not a multiple of 4, k % cab != 0 (trailing slots ignored), odd C
    run_case(ref, "ragged", 3, 5, 301, 3, 5, 8, seed=2)
    # k == N (sp_ratio 1): every point selected by every region (R-fold accumulation in backward)
    run_case(ref, "full", 2, 4, 64, 2, 1, 8, seed=3, keys_fn=lambda k: torch.round(k * 4) / 4)
    # k == cab: windows of one slot; R == 1
    run_case(ref, "kcab", 2, 6, 64, 1, 8, 8, seed=4)
    # the reference operating point shape in miniature: R*k == N, R=8, cab=8 (window 4.. here 2)
    run_case(ref, "oppoint", 2, 16, 128, 8, 8, 8, seed=5)


if __name__ == "__main__":
    sys.exit(main())
