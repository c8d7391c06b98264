"""Device time of a chain of C-ABI calls of one bench step (CUDA-graph replay over rotating buffer sets):
python tools/time_chain.py WORKLOAD first last   e.g.  A 0 2  = sp_topk_f32 -> sp_gather_fwd_f32"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

w = bench.WORKLOADS[sys.argv[1]]
lo, hi = int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda:0")
step = bench.Step(w, dev)
sets = [bench.BufferSet(w, dev, 1 + i) for i in range(max(3, int(-(-400e6 // bench.BufferSet(w, dev, 0).footprint()))))]
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    calls = [step.calls(s) for s in sets]
    for c in calls:
        for _, f in c:
            f()
    stream.synchronize()
    reps = 10 * len(sets)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        for i in range(reps):
            for _, f in calls[i % len(sets)][lo:hi]:
                f()
    g.replay(); stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5):
        g.replay()
    e1.record(stream); stream.synchronize()
print("%s calls %s: %.2f us" % (sys.argv[1], [n for n, _ in calls[0][lo:hi]], e0.elapsed_time(e1) * 1e3 / (5 * reps)))
