"""Run the bench step eagerly a few times (for ncu): python tools/profile_step.py [workload] [reps]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "A"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
step = bench.Step(w, dev)
sets = [bench.BufferSet(w, dev, 1 + i) for i in range(3)]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
for i in range(reps):
    flush.zero_()
    step.run(sets[i % 3])
torch.cuda.synchronize()
print("ok")
