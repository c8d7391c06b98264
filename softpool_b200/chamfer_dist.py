"""Drop-in for the reference's GRNet flavour, `GRNet/extensions/chamfer_dist/__init__.py`
(`from extensions.chamfer_dist import ChamferFunction, ChamferDistance`)."""
import torch

from . import ops


class ChamferFunction(torch.autograd.Function):
    """forward -> (dist1, dist2) only (reference __init__.py:13-19); backward :21-25."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        dist1, dist2, idx1, idx2 = ops.chamfer_forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1.contiguous(), xyz2.contiguous(), idx1, idx2)
        return dist1, dist2

    @staticmethod
    def backward(ctx, grad_dist1, grad_dist2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        if grad_dist1 is None:
            grad_dist1 = torch.zeros(idx1.shape, dtype=torch.float32, device=xyz1.device)
        if grad_dist2 is None:
            grad_dist2 = torch.zeros(idx2.shape, dtype=torch.float32, device=xyz1.device)
        return ops.chamfer_backward(xyz1, xyz2, grad_dist1, grad_dist2, idx1, idx2)


class ChamferDistance(torch.nn.Module):
    """scalar mean(dist1) + mean(dist2); with ignore_zeros and batch 1, points whose coordinates
    sum to zero are dropped first (reference __init__.py:28-42)."""

    def __init__(self, ignore_zeros=False):
        super().__init__()
        self.ignore_zeros = ignore_zeros

    def forward(self, xyz1, xyz2):
        if xyz1.size(0) == 1 and self.ignore_zeros:
            xyz1 = xyz1[torch.sum(xyz1, dim=2).ne(0)].unsqueeze(dim=0)
            xyz2 = xyz2[torch.sum(xyz2, dim=2).ne(0)].unsqueeze(dim=0)
        dist1, dist2 = ChamferFunction.apply(xyz1, xyz2)
        return torch.mean(dist1) + torch.mean(dist2)
