"""CPU port of the reference SoftPool forward/backward with the reference's own torch calls --
the *timed* CPU baseline (bench.py `cpu_baseline` / `--impl reference`).  TEST/BENCH
INFRASTRUCTURE ONLY; never imported by softpool_b200/.

/root/reference does not exist on the GPU box, so the reference module cannot be imported there;
this restates its hot loop op for op (reference softpool.py:134-151 and train2cabins :71-85):
zero-filled outputs, a Python loop over regions of `torch.sort(descending=True)` on the strided
row slice, the (B,C,k) int64 index `repeat`, `torch.gather`, two slice assignments, then `cab`
`torch.max` calls; backward is torch autograd, exactly as in the reference.  The dead conv tail
(softpool.py:154-164) is NOT included, which only flatters the baseline.
Checked against tests/golden (tests/test_oracle_cpu.py::test_torch_port_matches_reference).
"""
import torch


def forward(x, keys, k, cab=8):
    """x (B,C,N) f32 (may require grad), keys (B,R,N) f32 -> sp_cube, sp_idx, cabins, id_activa."""
    B, C, N = x.shape
    R = keys.shape[1]
    id_activa = torch.argmax(keys, dim=1)
    sp_cube = torch.zeros(B, C, R, k, device=x.device)
    sp_idx = torch.zeros(B, R + 3, R, k, device=x.device)
    for region in range(R):
        _, order = torch.sort(keys[:, region, :], dim=1, descending=True)
        top = order[:, :k]
        sp_cube[:, :, region, :] = torch.gather(x, dim=2, index=top.unsqueeze(1).repeat(1, C, 1))
        sp_idx[:, :, region, :] = top.unsqueeze(1).repeat(1, R + 3, 1)
    per = k // cab
    cabins = torch.zeros(B, C, R, cab, device=x.device)
    for w in range(cab):
        cabins[:, :, :, w] = torch.max(sp_cube[:, :, :, w * per:(w + 1) * per], dim=3, keepdim=False)[0]
    return sp_cube, sp_idx, cabins, id_activa


def forward_backward(x, keys, k, cab, g_cube, g_cabins):
    """One fwd+bwd step as the reference trains it; returns grad_x."""
    x = x.detach().requires_grad_(True)
    sp_cube, sp_idx, cabins, id_activa = forward(x, keys, k, cab)
    torch.autograd.backward([sp_cube, cabins], [g_cube, g_cabins])
    return x.grad
