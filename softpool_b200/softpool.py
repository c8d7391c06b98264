"""Drop-in for the reference's `softpool.py` (import as `import softpool as sp`, model.py:12).

Same class names, constructor arguments, parameter names/shapes (so reference checkpoints load,
SURVEY.md section 5) and return tuples as reference softpool.py:71-241; the sort/top-k, gather,
index-cube, argmax, window-max and their backward run as hand-written sm_100a kernels
(softpool_b200/csrc) instead of the reference's Python loop of R x torch.sort/gather/slice-assign.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


def _to_default_device(m):
    # the reference hard-codes `.cuda()` on every layer it builds (softpool.py:91,107-123)
    return m.cuda() if torch.cuda.is_available() else m


class Periodics(nn.Module):
    """SIREN-style sine layer (reference softpool.py:10-67; unused by the root model)."""

    def __init__(self, dim_input=2, dim_output=512, is_first=True):
        super().__init__()
        self.dim_input, self.dim_output, self.is_first = dim_input, dim_output, is_first
        self.with_frequency = True
        self.with_phase = True
        self.omega_0 = 30
        self.Li = _to_default_device(nn.Conv1d(dim_input, dim_output, 1, bias=True))
        bound = 1.0 / dim_input if is_first else float(np.sqrt(6.0 / dim_input)) / self.omega_0
        with torch.no_grad():
            self.Li.weight.uniform_(-bound, bound)

    def filter(self):
        return torch.ones(1, self.dim_output // 32 * 32, 1, device=self.Li.weight.device)

    def forward(self, x):
        return torch.sin(self.Li(x) * self.omega_0)


def train2cabins(windows, num_cabin=8):
    """(B,C,R,k) -> (B,C,R,num_cabin): max over num_cabin windows of k // num_cabin consecutive
    slots, trailing k % num_cabin slots ignored (reference softpool.py:71-85)."""
    return ops.cabins_max(windows, num_cabin)


class Sorter(nn.Module):
    """1x1 conv C -> R and the arg-max region of every point (reference softpool.py:88-96)."""

    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.conv1d = _to_default_device(nn.Conv1d(dim_in, dim_out, 1))

    def forward(self, x):
        val_activa = self.conv1d(x)
        return val_activa, ops.softpool_argmax(val_activa)


class SoftPool(nn.Module):
    """Reference softpool.py:99-171.

    forward(x (B,C,N)) -> sp_cube (B,C,R,k) f32, sp_idx (B,R+3,R,k) f32, cabins (B,C,R,cab) f32,
    id_activa (B,N) i64, with k = N // sp_ratio.  The conv2d_* layers are kept because the
    reference owns them (checkpoint keys) although its forward discards their output
    (`scope == 'local'`, softpool.py:154-169); they are not evaluated here.
    """

    def __init__(self, regions=16, cabins=8, sp_ratio=4, size_feat=256):
        super().__init__()
        if cabins < 5:
            raise ValueError("cabins must be >= 5 (conv2d_3 kernel width is cabins - 4)")
        self.regions, self.num_cabin, self.sp_ratio, self.size_feat = regions, cabins, sp_ratio, size_feat
        mk = lambda ks: _to_default_device(nn.Conv2d(size_feat, size_feat, kernel_size=ks, stride=(1, 1)))
        self.conv2d_1 = mk((1, 3))
        self.conv2d_2 = mk((1, 3))
        self.conv2d_3 = mk((1, cabins - 2 * (3 - 1)))
        self.conv2d_5 = mk((regions, 1))
        self.sorter = Sorter(size_feat, regions)

    def forward(self, x):
        self.size_bth, self.size_feat, n_points = list(x.shape)
        self.pnt_per_sort = n_points // self.sp_ratio
        if self.pnt_per_sort < self.num_cabin:
            raise RuntimeError("SoftPool: N // sp_ratio = %d points per region < %d cabins"
                               % (self.pnt_per_sort, self.num_cabin))
        val_activa = self.sorter.conv1d(x)                    # library 1x1 conv; boundary of the native path
        idx, sp_idx, id_activa = ops.softpool_topk(val_activa.detach(), self.pnt_per_sort)
        sp_cube, cabins = ops.softpool_gather(x, idx, self.num_cabin)
        self.last_idx = idx                                   # (B,R,k) i32: the integer index list behind the float sp_idx cube
        return sp_cube, sp_idx, cabins, id_activa


class SoftPoolFeat(nn.Module):
    """PointNet MLP 3->64->128->256 + SoftPool + index bookkeeping (reference softpool.py:174-241)."""

    def __init__(self, num_points=8192, regions=16, sp_points=2048, sp_ratio=8):
        super().__init__()
        self.conv1 = nn.Conv1d(3, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, 256, 1)
        self.bn1 = nn.BatchNorm1d(64)
        self.bn2 = nn.BatchNorm1d(128)
        self.bn3 = nn.BatchNorm1d(256)
        self.num_points = num_points
        self.regions = regions
        self.sp_points = sp_points // sp_ratio
        self.softpool = SoftPool(regions, cabins=8, sp_ratio=sp_ratio)

    def mlp(self, inputs):
        x = F.relu(self.bn1(self.conv1(inputs)))
        x = F.relu(self.bn2(self.conv2(x)))
        return self.bn3(self.conv3(x))

    def select_points(self, t):
        """t (B,F,N) gathered by the index list of the last forward -> (B,F,R*k): the caller's
        `torch.gather(part, dim=2, index=sp_idx[:, :3, 0, :].long())` (model.py:283-285) and the reference's own
        `point_wi_seg` gather (softpool.py:218-231) without the one_hot / cat / repeat / float-index detour."""
        return ops.softpool_select_points(t, self.softpool.last_idx)

    def point_wi_seg(self, x, x_seg=None):
        """The reference's `point_wi_seg` tensor (softpool.py:218-231): region one-hot of every point (or the given
        segmentation) stacked on xyz, (B, R+3, N), picked by the index list -> (B, R+3, 1, R*k).  Dead in the reference
        (it only feeds the commented-out `feature` return) but part of its forward; emitted index-driven here."""
        idx = self.softpool.last_idx
        B, R, k = idx.shape
        if x_seg is None:
            seg = F.one_hot(self.softpool.last_id_activa.to(torch.int64), self.regions).transpose(1, 2).float()
        else:
            seg = x_seg.float()
        return ops.softpool_select_points(torch.cat((seg, x), 1).contiguous(), idx).view(B, seg.shape[1] + x.shape[1], 1, R * k)

    def forward(self, x, x_seg=None):
        part = x
        sp_cube, sp_idx, cabins, id_activa = self.softpool(self.mlp(x))
        self.softpool.last_id_activa = id_activa
        # region one-hot (or the given segmentation) stacked on xyz, gathered by the same indices
        # (reference softpool.py:218-231); the result only feeds the commented-out `feature` return: see point_wi_seg()
        B = sp_cube.shape[0]
        flat = self.regions * self.sp_points
        sp_cube = sp_cube.view(B, sp_cube.shape[1], 1, flat)
        sp_idx = sp_idx.view(B, sp_idx.shape[1], 1, flat)
        return sp_cube, cabins, sp_idx
