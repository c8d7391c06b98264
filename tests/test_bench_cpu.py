"""bench.py pieces that need no GPU: the algorithmic-byte counts of SURVEY.md section 8(d) and the JSON contract of
the reference arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_match_the_survey():
    fwd, bwd = bench.algorithmic_bytes(bench.WORKLOADS["A"])          # 21.9 MB + 77.6 MB = 99.5 MB
    assert round(fwd / 1e6, 1) == 21.9 and round(bwd / 1e6, 1) == 77.6
    fwd, bwd = bench.algorithmic_bytes(bench.WORKLOADS["A1"])         # 141.8 MB + 136.6 MB = 278.4 MB
    assert round(fwd / 1e6, 1) == 141.8 and round(bwd / 1e6, 1) == 136.6


def test_launch_count_matches_the_step():
    # chamfer_fwd_loss_f32 = prep + tensor kernel, every other call one kernel
    assert bench.LAUNCHES_PER_STEP == len(bench.KERNELS) + 1
    assert bench.KERNELS[:bench.N_SOFTPOOL_CALLS] == ["sp_topk_f32", "sp_gather_fwd_f32", "sp_gather_bwd_f32"]


def test_traffic_lookup_by_kernel_prefix():
    t = {"sp_topk_kernel": 1.0, "sp_gather_fwd_kernel": 2.0, "sp_gather_bwd_pull_kernel": 3.0, "chamfer_tc_kernel": 5.0}
    assert bench.traffic_of(t, "sp_topk", "sp_gather_fwd", "sp_gather_bwd") == 6.0
    assert bench.traffic_of(t, "chamfer_prep", "chamfer_tc") is None           # a missing kernel -> no number, not a wrong one
    wl = bench.load_traffic("A")
    assert bench.traffic_of(wl, "sp_topk", "sp_gather_fwd", "sp_gather_bwd") > 5e7     # the committed ncu capture


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU port of the reference path, bounded sample) on this host's cores."""
    w = dict(bench.WORKLOADS["A"], B=2, C=8, N=256, k=8, name="tiny")           # keep the CPU suite short
    class A: steps, warmup = 1, 1
    out = bench.run_reference(A, w, 0, 1)
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in out, key
    assert out["impl"] == "reference" and out["value"] > 0 and out["e2e"]["h2d_bytes_per_step"] == 0
    assert out["cpu_baseline"]["kind"] == "port" and out["cpu_baseline"]["cores"] >= 1
    json.dumps(out)
