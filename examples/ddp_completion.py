#!/usr/bin/env python
"""BASELINE config 4 in miniature: encoder -> SoftPool -> decoder -> Chamfer on synthetic 2048-point clouds,
one process per GPU under DistributedDataParallel (NCCL all-reduce on the gradients only; the operators
themselves exchange nothing -- every sort row, gather row and Chamfer sample lives inside one batch element).

    python examples/ddp_completion.py                      # one GPU
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/ddp_completion.py --steps 50

The encoder is the drop-in `SoftPoolFeat` (reference softpool.py:174-241: PointNet MLP + SoftPool); the decoder is a
small 1x1-conv stack standing in for the reference's `model.py` decoder (out of scope), the loss
is the reference's `mean(dist1,1) + mean(dist2,1)` (train.py:68-69).  SoftPool's own parameters never receive a
gradient (the sort is a hard selection and the conv2d_* tail is dead code in the reference), so they
are frozen -- otherwise DDP would wait for gradients that never come.
"""
import argparse
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import softpool_b200 as spb                      # noqa: E402
from softpool_b200 import dist as spd            # noqa: E402
from softpool_b200 import ops                    # noqa: E402


class Completion(nn.Module):
    def __init__(self, regions=8, n_points=2048):
        super().__init__()
        self.enc = spb.SoftPoolFeat(num_points=n_points, regions=regions, sp_points=n_points, sp_ratio=regions)
        self.dec = nn.Sequential(nn.Conv2d(256, 128, 1), nn.ReLU(), nn.Conv2d(128, 64, 1), nn.ReLU(), nn.Conv2d(64, 3, 1), nn.Tanh())
        for name, p in self.enc.softpool.named_parameters():       # sorter.conv1d, conv2d_{1,2,3,5}: no gradient ever
            p.requires_grad_(False)

    def forward(self, part, gt):
        sp_cube, _, _ = self.enc(part)                              # (B,256,1,R*k)
        pred = 0.5 * self.dec(sp_cube)[:, :, 0, :].transpose(1, 2).contiguous()     # (B,R*k,3)
        return ops.chamfer_mean_loss(pred, gt).mean()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch", type=int, default=32, help="per GPU")
    args = ap.parse_args()
    rank, local_rank, world = spd.init()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)                                            # same initial weights on every rank
    model = Completion().to(dev)
    if world > 1:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[local_rank])
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-3)
    g = torch.Generator(device="cpu").manual_seed(1000 + rank)      # every rank its own shard of the synthetic set
    losses = []
    t0 = None
    for it in range(args.steps + 3):
        if it == 3:
            torch.cuda.synchronize(dev); spd.barrier(); t0 = time.perf_counter()
        gt = (torch.rand(args.batch, 2048, 3, generator=g) - 0.5).to(dev)
        part = gt.transpose(1, 2).contiguous() + 0.01 * torch.randn(args.batch, 3, 2048, generator=g).to(dev)
        loss = model(part, gt)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    torch.cuda.synchronize(dev); spd.barrier()
    dt = time.perf_counter() - t0
    first, last = spd.sum_over_ranks(losses[0], dev) / world, spd.sum_over_ranks(losses[-1], dev) / world
    if rank == 0:
        print("world %d: %d steps of %d clouds per GPU: %.1f clouds/s, loss %.5f -> %.5f"
              % (world, args.steps, args.batch, world * args.batch * args.steps / dt, first, last), flush=True)
    assert last < first, "the loss did not go down"
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
