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


def test_both_arms_print_the_same_config():
    """The driver compares the `config` dicts of the two arms (`same_config`): one function builds both."""
    for name, w in bench.WORKLOADS.items():
        c = bench.config_of(w)
        assert c == bench.config_of(dict(w)) and c["workload"] == w["name"] and c["per_gpu_batch"] == w["B"]
        json.dumps(c)
    assert {"A", "A1", "N1024", "N4096", "N8192", "N16384", "C32", "C64", "C128", "C512"} <= set(bench.WORKLOADS)


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU port of the reference path, bounded sample) on this host's cores."""
    w = dict(bench.WORKLOADS["A"], B=2, C=8, N=256, k=8, name="tiny")           # keep the CPU suite short
    class A: steps, warmup = 1, 1
    out = bench.run_reference(A, w, 0, 1)
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in out, key
    assert out["impl"] == "reference" and out["value"] > 0 and out["e2e"]["h2d_bytes_per_step"] == 0
    assert out["cpu_baseline"]["kind"] in ("port", "reference") and out["cpu_baseline"]["cores"] >= 1
    assert out["config"] == bench.config_of(w)
    json.dumps(out)


def test_reference_module_and_port_agree():
    """The reference arm runs the staged reference softpool.py when oracle/_ref holds it (kind "reference"), else the port;
    where both exist they must produce the same gradient for the same inputs."""
    import numpy as np
    import pytest
    from oracle import build_ref
    if build_ref.load_ref_softpool(cpu=True) is None:
        pytest.skip("reference softpool.py not staged (oracle/_ref) and /root/reference absent")
    w = dict(bench.WORKLOADS["A"], B=2, C=8, N=256, k=32, name="tiny")
    r = bench.CpuRef(w, 2)
    assert r.kind == "reference"
    xx = r.x.detach().requires_grad_(True)
    sp_cube, sp_idx, cabins, id_activa = r.sp(xx)
    import torch
    torch.autograd.backward([sp_cube, cabins], [r.gc, r.gb])
    g_port = r.port.forward_backward(r.x, r.keys, r.k, r.cab, r.gc, r.gb)
    assert np.allclose(xx.grad.numpy(), g_port.numpy(), rtol=1e-5, atol=1e-6)
