"""Debug counters of chamfer_tc_kernel: build with SPK_NVCC_EXTRA=-DSPK_TIMING first.
python tools/ch_phases.py [workload]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "A"]
dev = torch.device("cuda:0")
step = bench.Step(w, dev)
s = bench.BufferSet(w, dev, 1)
for i in range(2):
    step.calls(s)[3][1]()
    torch.cuda.synchronize()
    print("--")
