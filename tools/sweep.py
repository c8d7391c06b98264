"""N sweep (BASELINE config 5: N in {1k, 2k, 4k, 8k, 16k}) and C sweep (north_star: C = 32 -> 512 at N = 2048) of the bench step.

    python tools/sweep.py [--gpus N] [--out profiles/r2_sweep.json] [--workloads A,N8192,...]

Runs bench.py once per workload (under torchrun when --gpus > 1) and keeps, per workload: us per step, Mpoints/s, the per-call
device times, both roofline fractions, the kernels to beat on the same GPU (ref_gpu) and the end-to-end number."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_sweep.json"))
ap.add_argument("--workloads", default="N1024,A1,N4096,N8192,N16384,C32,C64,C128,A,C512")
ap.add_argument("--steps", type=int, default=60)
ap.add_argument("--no-ref-gpu", action="store_true")
args = ap.parse_args()
rows = {}
if os.path.exists(args.out):
    rows = json.load(open(args.out))
for wl in args.workloads.split(","):
    cmd = [sys.executable]
    if args.gpus > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1", "--master-port", "29511"]
    cmd += [os.path.join(ROOT, "bench.py"), "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", "5", "--workload", wl, "--no-cpu-baseline"]
    if args.no_ref_gpu or args.gpus > 1:
        cmd.append("--no-ref-gpu")
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(wl, "FAILED", r.stderr[-400:]); continue
    d = json.loads(line[-1])
    row = {"n_gpus": d["n_gpus"], "us_per_step": d["ms_per_step"] * 1e3, "mpoints_per_s": d["value"], "kernel_us": d["kernel_us"], "chain_us": d.get("chain_us"),
           "roofline_softpool_frac": d["roofline_softpool"]["frac"], "roofline_chamfer_frac": d["roofline_chamfer"]["frac"],
           "chamfer_path": d["roofline_chamfer"].get("path"), "e2e_mpoints_per_s": d["e2e"]["value"], "ref_gpu": d.get("ref_gpu"),
           "ranks_ms_per_step": d.get("measurement", {}).get("ranks_ms_per_step"), "clocks": d["clocks"], "workload": d["config"]["workload"]}
    rows["%s@%d" % (wl, d["n_gpus"])] = row
    print("%-7s x%d: %8.1f us/step %8.1f Mpoints/s  softpool %.3f  chamfer %.3f (%s)" % (wl, d["n_gpus"], row["us_per_step"], row["mpoints_per_s"],
          row["roofline_softpool_frac"], row["roofline_chamfer_frac"], row["chamfer_path"]), flush=True)
    json.dump(rows, open(args.out, "w"), indent=1, sort_keys=True)
