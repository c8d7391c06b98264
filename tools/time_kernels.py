"""Per-call device time of every C-ABI call of the bench step, replayed from CUDA graphs so that host
launch overhead is out of the picture: python tools/time_kernels.py [workload]
Rotates over 3 buffer sets (cold-ish L2) and also reports the single-set (warm) figure."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "A"]
dev = torch.device("cuda:0")
step = bench.Step(w, dev)
sets = [bench.BufferSet(w, dev, 1 + i) for i in range(3)]
calls = [step.calls(s) for s in sets]
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    for c in calls:
        for _, f in c:
            f()
    stream.synchronize()
    for j, (name, _) in enumerate(calls[0]):
        res = []
        for rot in (1, 3):
            g = torch.cuda.CUDAGraph()
            reps = 30
            with torch.cuda.graph(g, stream=stream):
                for i in range(reps):
                    calls[i % rot][j][1]()
            g.replay()
            stream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(5):
                g.replay()
            e1.record(stream)
            stream.synchronize()
            res.append(e0.elapsed_time(e1) * 1e3 / (5 * reps))
        print("%-20s warm %.2f us   rotating %.2f us" % (name, res[0], res[1]))
