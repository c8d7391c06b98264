"""Functional host layer over the C ABI (include/softpool_b200.h).

Every function here launches hand-written sm_100a kernels through ctypes on the tensor's
current CUDA stream.  There is no CPU / PyTorch fallback: non-CUDA tensors raise.
"""
import torch

from . import _lib
from ._lib import check, ptr, require_cuda, stream_of


# ------------------------------------------------------------------------------------------
# SoftPool
# ------------------------------------------------------------------------------------------
def softpool_topk(keys, k, want_sp_idx=True, want_id_activa=True):
    """keys (B,R,N) f32 -> idx (B,R,k) i32, sp_idx (B,R+3,R,k) f32 | None, id_activa (B,N) i64 | None.

    Reference: region loop sort + `[:, :k]` (softpool.py:139-142), the float index cube
    (softpool.py:136-137,146-147), `torch.argmax(val_activa, 1)` (softpool.py:95).
    """
    require_cuda(keys, "keys", torch.float32)
    if keys.dim() != 3:
        raise RuntimeError("keys must be (B,R,N)")
    keys = keys.contiguous()
    B, R, N = keys.shape
    k = int(k)
    if not 1 <= k <= N:
        raise RuntimeError("need 1 <= k <= N (k=%d, N=%d)" % (k, N))
    dev = keys.device
    idx = torch.empty((B, R, k), dtype=torch.int32, device=dev)
    sp_idx = torch.empty((B, R + 3, R, k), dtype=torch.float32, device=dev) if want_sp_idx else None
    id_activa = torch.empty((B, N), dtype=torch.int64, device=dev) if want_id_activa else None
    with torch.cuda.device(dev):
        check(_lib.lib().sp_topk_f32(ptr(keys), B, R, N, k, ptr(idx), ptr(sp_idx), ptr(id_activa),
                                     stream_of(keys)), "sp_topk_f32")
    return idx, sp_idx, id_activa


def softpool_argmax(keys):
    """keys (B,R,N) f32 -> (B,N) i64; `torch.argmax(val_activa, dim=1)` of softpool.py:95."""
    require_cuda(keys, "keys", torch.float32)
    keys = keys.contiguous()
    B, R, N = keys.shape
    out = torch.empty((B, N), dtype=torch.int64, device=keys.device)
    with torch.cuda.device(keys.device):
        check(_lib.lib().sp_argmax_i64(ptr(keys), B, R, N, ptr(out), stream_of(keys)), "sp_argmax_i64")
    return out


class _SoftPoolGather(torch.autograd.Function):
    """x (B,C,N), idx (B,R,k) -> sp_cube (B,C,R,k), cabins (B,C,R,cab).

    Forward = softpool.py:142-145 + train2cabins (softpool.py:71-85); backward = what autograd
    derives for them in the reference (scatter_add + CopySlices + MaxBackward), done by one
    deterministic kernel.
    """

    @staticmethod
    def forward(ctx, x, idx, cab):
        require_cuda(x, "x", torch.float32)
        require_cuda(idx, "idx", torch.int32)
        x = x.contiguous()
        idx = idx.contiguous()
        B, C, N = x.shape
        _, R, k = idx.shape
        dev = x.device
        sp_cube = torch.empty((B, C, R, k), dtype=torch.float32, device=dev)
        cabins = torch.empty((B, C, R, cab), dtype=torch.float32, device=dev)
        cab_arg = torch.empty((B, C, R, cab), dtype=torch.uint16, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().sp_gather_fwd_f32(ptr(x), ptr(idx), B, C, N, R, k, cab, ptr(sp_cube),
                                               ptr(cabins), ptr(cab_arg), stream_of(x)),
                  "sp_gather_fwd_f32")
        ctx.save_for_backward(idx, cab_arg)
        ctx.dims = (B, C, N, R, k, cab)
        return sp_cube, cabins

    @staticmethod
    def backward(ctx, g_cube, g_cabins):
        idx, cab_arg = ctx.saved_tensors
        B, C, N, R, k, cab = ctx.dims
        dev = idx.device
        if g_cube is None:
            g_cube = torch.zeros((B, C, R, k), dtype=torch.float32, device=dev)
        g_cube = g_cube.contiguous()
        if g_cabins is not None:
            g_cabins = g_cabins.contiguous()
        grad_x = torch.empty((B, C, N), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().sp_gather_bwd_f32(ptr(g_cube), ptr(g_cabins), ptr(idx), ptr(cab_arg),
                                               B, C, N, R, k, cab, ptr(grad_x), stream_of(g_cube)),
                  "sp_gather_bwd_f32")
        return grad_x, None, None


def softpool_gather(x, idx, cab):
    return _SoftPoolGather.apply(x, idx, int(cab))


class _SelectPoints(torch.autograd.Function):
    """Index-driven glue: any per-point tensor t (B,F,N) picked by the SoftPool index list idx (B,R,k) ->
    (B,F,R*k), with the same gather kernel (no window max) instead of the reference's
    `one_hot -> cat -> unsqueeze(2).repeat(1,1,R,1) -> torch.gather(index=sp_idx.long())` (softpool.py:218-231) and
    `torch.gather(part, dim=2, index=sp_idx[:, :3, 0, :].long())` (model.py:283-285).  Backward = the gather backward."""

    @staticmethod
    def forward(ctx, t, idx):
        require_cuda(t, "t", torch.float32)
        require_cuda(idx, "idx", torch.int32)
        t, idx = t.contiguous(), idx.contiguous()
        B, Fd, N = t.shape
        _, R, k = idx.shape
        out = torch.empty((B, Fd, R, k), dtype=torch.float32, device=t.device)
        with torch.cuda.device(t.device):
            check(_lib.lib().sp_gather_fwd_f32(ptr(t), ptr(idx), B, Fd, N, R, k, 1, ptr(out), None, None, stream_of(t)),
                  "sp_gather_fwd_f32")
        ctx.save_for_backward(idx)
        ctx.dims = (B, Fd, N, R, k)
        return out.view(B, Fd, R * k)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        B, Fd, N, R, k = ctx.dims
        g = g.contiguous().view(B, Fd, R, k)
        grad = torch.empty((B, Fd, N), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            check(_lib.lib().sp_gather_bwd_f32(ptr(g), None, ptr(idx), None, B, Fd, N, R, k, 1, ptr(grad), stream_of(g)),
                  "sp_gather_bwd_f32")
        return grad, None


def softpool_select_points(t, idx):
    """t (B,F,N) f32, idx (B,R,k) i32 -> (B,F,R*k): t[b,f,idx[b,r,j]] at column r*k+j (see _SelectPoints)."""
    return _SelectPoints.apply(t, idx)


class _Cabins(torch.autograd.Function):
    """Standalone train2cabins (softpool.py:71-85) with its MaxBackward."""

    @staticmethod
    def forward(ctx, windows, cab):
        require_cuda(windows, "windows", torch.float32)
        windows = windows.contiguous()
        k = windows.shape[-1]
        rows = windows.numel() // k if k else 0
        out_shape = tuple(windows.shape[:-1]) + (cab,)
        cabins = torch.empty(out_shape, dtype=torch.float32, device=windows.device)
        cab_arg = torch.empty(out_shape, dtype=torch.uint16, device=windows.device)
        with torch.cuda.device(windows.device):
            check(_lib.lib().sp_cabins_fwd_f32(ptr(windows), rows, k, cab, ptr(cabins), ptr(cab_arg),
                                               stream_of(windows)), "sp_cabins_fwd_f32")
        ctx.save_for_backward(cab_arg)
        ctx.meta = (tuple(windows.shape), rows, k, cab)
        return cabins

    @staticmethod
    def backward(ctx, g_cabins):
        (cab_arg,) = ctx.saved_tensors
        shape, rows, k, cab = ctx.meta
        g_cabins = g_cabins.contiguous()
        g_windows = torch.empty(shape, dtype=torch.float32, device=g_cabins.device)
        with torch.cuda.device(g_cabins.device):
            check(_lib.lib().sp_cabins_bwd_f32(ptr(g_cabins), ptr(cab_arg), rows, k, cab, ptr(g_windows),
                                               stream_of(g_cabins)), "sp_cabins_bwd_f32")
        return g_windows, None


def cabins_max(windows, cab):
    return _Cabins.apply(windows, int(cab))


def gather_operation(features, idx):
    """Drop-in for the MSN decoder's `gather_operation` (reference MSN/MDS/MDS_module.py:40-84, kernels
    MSN/MDS/MDS_cuda.cu:29-75): features (B,C,N) f32, idx (B,npoint) int32 -> (B,C,npoint) = features[b,c,idx[b,j]].
    Forward and backward run on the SoftPool gather kernels (one region of npoint slots).  The backward sums
    deterministically; the reference's `grad_points[..] += grad_out[..]` (MDS_cuda.cu:66-67, its atomicAdd commented out)
    races when an index repeats.  The indices of one sample must be DISTINCT, which is what its producer, minimum density
    sampling, returns (the inverse table of the backward holds one slot per point and region)."""
    require_cuda(idx, "idx", torch.int32)
    if idx.dim() != 2:
        raise RuntimeError("gather_operation: idx must be (B, npoint)")
    return softpool_select_points(features, idx.unsqueeze(1))


# ------------------------------------------------------------------------------------------
# Chamfer
# ------------------------------------------------------------------------------------------
def _chamfer_inputs(xyz1, xyz2):
    require_cuda(xyz1, "xyz1", torch.float32)
    require_cuda(xyz2, "xyz2", torch.float32)
    if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.shape[2] != 3 or xyz2.shape[2] != 3:
        raise RuntimeError("chamfer: inputs must be (B,n,3) and (B,m,3)")
    if xyz1.shape[0] != xyz2.shape[0]:
        raise RuntimeError("chamfer: batch sizes differ")
    if xyz1.device != xyz2.device:
        raise RuntimeError("chamfer: inputs on different devices")
    return xyz1.contiguous(), xyz2.contiguous()


def chamfer_forward(xyz1, xyz2, want_loss=False):
    """-> dist1 (B,n) f32, dist2 (B,m) f32, idx1 (B,n) i32, idx2 (B,m) i32 (chamfer.cu:136-152)
    [, loss (B,) f32 = mean(dist1,1) + mean(dist2,1) from the same launch when want_loss]."""
    xyz1, xyz2 = _chamfer_inputs(xyz1, xyz2)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dev = xyz1.device
    dist1 = torch.empty((B, n), dtype=torch.float32, device=dev)
    dist2 = torch.empty((B, m), dtype=torch.float32, device=dev)
    idx1 = torch.empty((B, n), dtype=torch.int32, device=dev)
    idx2 = torch.empty((B, m), dtype=torch.int32, device=dev)
    L = _lib.lib()
    ws_bytes = int(L.chamfer_fwd_workspace_bytes(B, n, m))
    ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        if want_loss:
            loss = torch.empty((B,), dtype=torch.float32, device=dev)
            check(L.chamfer_fwd_loss_f32(ptr(xyz1), ptr(xyz2), B, n, m, ptr(dist1), ptr(dist2), ptr(idx1),
                                         ptr(idx2), ptr(loss), ptr(ws), ws_bytes, stream_of(xyz1)),
                  "chamfer_fwd_loss_f32")
            return dist1, dist2, idx1, idx2, loss
        check(L.chamfer_fwd_f32(ptr(xyz1), ptr(xyz2), B, n, m, ptr(dist1), ptr(dist2), ptr(idx1),
                                ptr(idx2), ptr(ws), ws_bytes, stream_of(xyz1)), "chamfer_fwd_f32")
    return dist1, dist2, idx1, idx2


class _ChamferMeanLoss(torch.autograd.Function):
    """xyz1 (B,n,3), xyz2 (B,m,3) -> loss (B,) = mean(dist1,1) + mean(dist2,1): the expression at every
    Chamfer call site of the reference (train.py:68-69,82-86), as ONE forward launch pair and one backward
    launch (the 1/n, 1/m scaling of the upstream gradient is folded into the backward's inputs)."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        dist1, dist2, idx1, idx2, loss = chamfer_forward(xyz1, xyz2, want_loss=True)
        ctx.save_for_backward(xyz1.contiguous(), xyz2.contiguous(), idx1, idx2)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        B, n = idx1.shape
        m = idx2.shape[1]
        g = g_loss.contiguous().to(torch.float32)
        g1 = (g / n)[:, None].expand(B, n).contiguous()
        g2 = (g / m)[:, None].expand(B, m).contiguous()
        return chamfer_backward(xyz1, xyz2, g1, g2, idx1, idx2)


def chamfer_mean_loss(xyz1, xyz2):
    """Differentiable per-sample Chamfer loss (B,), see _ChamferMeanLoss."""
    return _ChamferMeanLoss.apply(xyz1, xyz2)


class _ChamferMultiMeanLoss(torch.autograd.Function):
    """preds (P, B, n, 3) stacked predictions, gt (B, m, 3) -> loss (P, B): mean(dist1,1) + mean(dist2,1) of every
    prediction against the SAME ground truth -- the training step's `self.CD(output1[i], gt)` calls (train.py:68-86) as one
    forward call (chamfer_fwd_multi_f32: one prep + one tensor launch, the ground truth formatted once per sample) and one
    backward call; the ground truth's gradient is the sum over the predictions."""

    @staticmethod
    def forward(ctx, preds, gt):
        require_cuda(preds, "preds", torch.float32)
        require_cuda(gt, "gt", torch.float32)
        if preds.dim() != 4 or preds.shape[3] != 3 or gt.dim() != 3 or gt.shape[2] != 3 or preds.shape[1] != gt.shape[0]:
            raise RuntimeError("chamfer_multi: preds must be (P,B,n,3) and gt (B,m,3)")
        preds, gt = preds.contiguous(), gt.contiguous()
        P, B, n, _ = preds.shape
        m = gt.shape[1]
        dev = preds.device
        dist1 = torch.empty((P * B, n), dtype=torch.float32, device=dev)
        dist2 = torch.empty((P * B, m), dtype=torch.float32, device=dev)
        idx1 = torch.empty((P * B, n), dtype=torch.int32, device=dev)
        idx2 = torch.empty((P * B, m), dtype=torch.int32, device=dev)
        loss = torch.empty((P * B,), dtype=torch.float32, device=dev)
        L = _lib.lib()
        ws_bytes = int(L.chamfer_fwd_workspace_bytes(P * B, n, m))
        ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(L.chamfer_fwd_multi_f32(ptr(preds), ptr(gt), P, B, n, m, ptr(dist1), ptr(dist2), ptr(idx1), ptr(idx2),
                                          ptr(loss), ptr(ws), ws_bytes, stream_of(preds)), "chamfer_fwd_multi_f32")
        ctx.save_for_backward(preds, gt, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return loss.view(P, B), dist1.view(P, B, n), dist2.view(P, B, m), idx1.view(P, B, n), idx2.view(P, B, m)

    @staticmethod
    def backward(ctx, g_loss, g_d1, g_d2, _i1, _i2):
        preds, gt, idx1, idx2 = ctx.saved_tensors
        P, B, n, _ = preds.shape
        m = gt.shape[1]
        g1 = torch.zeros((P * B, n), dtype=torch.float32, device=preds.device)
        g2 = torch.zeros((P * B, m), dtype=torch.float32, device=preds.device)
        if g_loss is not None:
            g = g_loss.contiguous().to(torch.float32).view(P * B)
            g1 = g1 + (g / n)[:, None]
            g2 = g2 + (g / m)[:, None]
        if g_d1 is not None:
            g1 = g1 + g_d1.reshape(P * B, n)
        if g_d2 is not None:
            g2 = g2 + g_d2.reshape(P * B, m)
        gt_rep = gt.unsqueeze(0).expand(P, B, m, 3).reshape(P * B, m, 3)           # (the backward kernel pairs rows one to one)
        grad1, grad2 = chamfer_backward(preds.view(P * B, n, 3), gt_rep, g1, g2, idx1, idx2)
        return grad1.view(P, B, n, 3), grad2.view(P, B, m, 3).sum(0)


def chamfer_multi(preds, gt):
    """Several predictions against one ground truth in one call.  preds: (P,B,n,3) tensor or a list of P (B,n,3) tensors of the
    same n; gt (B,m,3).  -> loss (P,B), dist1 (P,B,n), dist2 (P,B,m), idx1 (P,B,n) i32, idx2 (P,B,m) i32; differentiable in
    preds and gt through loss / dist1 / dist2.  Bit-identical to P separate chamferDist calls."""
    if isinstance(preds, (list, tuple)):
        preds = torch.stack(list(preds), 0)
    return _ChamferMultiMeanLoss.apply(preds, gt)


def chamfer_backward(xyz1, xyz2, g1, g2, idx1, idx2):
    """-> grad_xyz1 (B,n,3), grad_xyz2 (B,m,3) (chamfer.cu:155-196)."""
    xyz1, xyz2 = _chamfer_inputs(xyz1, xyz2)       # (contiguous: an expanded / strided view must not reach the kernel as it is)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dev = xyz1.device
    g1 = g1.contiguous()          # dist_chamfer.py:37-38 does the same
    g2 = g2.contiguous()
    idx1, idx2 = idx1.contiguous(), idx2.contiguous()
    grad1 = torch.empty((B, n, 3), dtype=torch.float32, device=dev)
    grad2 = torch.empty((B, m, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().chamfer_bwd_f32(ptr(xyz1), ptr(xyz2), ptr(g1), ptr(g2), ptr(idx1), ptr(idx2),
                                         B, n, m, ptr(grad1), ptr(grad2), stream_of(xyz1)),
              "chamfer_bwd_f32")
    return grad1, grad2


def chamfer_loss(dist1, dist2):
    """loss (B,) = mean(dist1,1) + mean(dist2,1) -- the reduction at train.py:68-69,82-86."""
    require_cuda(dist1, "dist1", torch.float32)
    require_cuda(dist2, "dist2", torch.float32)
    dist1, dist2 = dist1.contiguous(), dist2.contiguous()
    B, n = dist1.shape
    m = dist2.shape[1]
    loss = torch.empty((B,), dtype=torch.float32, device=dist1.device)
    with torch.cuda.device(dist1.device):
        check(_lib.lib().chamfer_loss_f32(ptr(dist1), ptr(dist2), B, n, m, ptr(loss), stream_of(dist1)),
              "chamfer_loss_f32")
    return loss
