"""GPU parity: the sm_100a Chamfer kernels (through the C ABI) against the C restatement of the
reference kernels (oracle/chamfer_oracle.c) and, when oracle/_ref was built, against the
reference's own chamfer.cu compiled unmodified for sm_100."""
import os

import numpy as np
import pytest
import torch

from oracle import chamfer_oracle as co

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True, params=["auto", "dense", "sorted"])
def chamfer_path(request, monkeypatch):
    """Every test of this file runs three times: the library's own choice of forward kernel, the dense tensor-core
    kernel (chamfer_dense.cu) forced, and the sorted search (chamfer_tc.cu) forced -- all must be bit-identical."""
    if request.param != "auto":
        monkeypatch.setenv("SPK_CHAMFER_PATH", request.param)
    else:
        monkeypatch.delenv("SPK_CHAMFER_PATH", raising=False)
    return request.param


def dev():
    return torch.device("cuda:0")


def clouds(B, n, m, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    a = ((rng.random((B, n, 3), dtype=np.float32) - 0.5) * scale).astype(np.float32)
    b = ((rng.random((B, m, 3), dtype=np.float32) - 0.5) * scale).astype(np.float32)
    return a, b


def cuda_forward(a, b):
    from softpool_b200 import ops
    d1, d2, i1, i2 = ops.chamfer_forward(torch.from_numpy(a).to(dev()), torch.from_numpy(b).to(dev()))
    return d1.cpu().numpy(), d2.cpu().numpy(), i1.cpu().numpy(), i2.cpu().numpy()


SHAPES = [(4, 64, 128), (2, 700, 1100), (3, 1, 5), (3, 5, 1), (2, 513, 511), (1, 1, 1),
          (32, 2048, 2048),          # BASELINE config 3
          (4, 4096, 2048),           # training shape (model.py:296-303 vs gt)
          (1, 2048, 16384),          # validation shape (val.py:270,302)
          (2, 8192, 8192)]


@pytest.mark.parametrize("B,n,m", SHAPES)
def test_forward_bit_exact_vs_oracle(B, n, m):
    a, b = clouds(B, n, m, seed=n * 31 + m)
    r = co.forward(a, b)
    o = cuda_forward(a, b)
    assert o[2].dtype == np.int32 and o[3].dtype == np.int32
    assert np.array_equal(o[2], r[2]) and np.array_equal(o[3], r[3])                    # indices
    assert np.array_equal(o[0].view(np.uint32), r[0].view(np.uint32))                   # distances, bit for bit
    assert np.array_equal(o[1].view(np.uint32), r[1].view(np.uint32))


def test_first_minimum_on_duplicates_and_scales():
    a, b = clouds(2, 300, 400, seed=5)
    b[:, 100:200] = b[:, 0:100]            # exact duplicates: the first one must win
    a[:, 50:60] = b[:, 120:130]            # zero distances
    for scale in (1.0, 1e-3, 1e3):
        r = co.forward(a * scale, b * scale)
        o = cuda_forward((a * scale).astype(np.float32), (b * scale).astype(np.float32))
        for x, y in zip(o, r):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
    assert (r[2][:, 50:60] == np.arange(20, 30)).all()


@pytest.mark.parametrize("B,n,m", [(4, 64, 128), (2, 700, 1100), (8, 2048, 2048), (2, 4096, 2048)])
def test_backward_vs_oracle(B, n, m):
    from softpool_b200 import ops
    a, b = clouds(B, n, m, seed=n + m)
    rng = np.random.default_rng(1)
    g1 = rng.random((B, n), dtype=np.float32)
    g2 = rng.random((B, m), dtype=np.float32)
    d1, d2, i1, i2 = co.forward(a, b)
    rg1, rg2 = co.backward(a, b, g1, g2, i1, i2)
    t = lambda v: torch.from_numpy(v).to(dev())
    og1, og2 = ops.chamfer_backward(t(a), t(b), t(g1), t(g2), t(i1), t(i2))
    # atomics (ours and the reference's) sum in arbitrary order: 1e-4 rel as north_star states
    np.testing.assert_allclose(og1.cpu().numpy(), rg1, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(og2.cpu().numpy(), rg2, rtol=1e-4, atol=1e-6)


def test_modules_surface_and_autograd():
    """chamferDist (dist_chamfer.py:48-53): 4-tuple, int32 indices, keyword call of val.py:302;
    ChamferDistance (GRNet __init__.py:28-42): scalar, ignore_zeros; gradients against autograd of
    the closed form (the check GRNet/extensions/chamfer_dist/test.py:22-28 does with gradcheck)."""
    import softpool_b200 as spb
    a, b = clouds(4, 64, 128, seed=11)
    ta = torch.from_numpy(a).to(dev()).requires_grad_(True)
    tb = torch.from_numpy(b).to(dev()).requires_grad_(True)
    cd = spb.chamferDist()
    dist1, dist2, idx1, idx2 = cd.forward(input1=ta, input2=tb)
    assert idx1.dtype == torch.int32 and idx2.dtype == torch.int32
    assert dist1.shape == (4, 64) and dist2.shape == (4, 128)
    loss = (torch.mean(dist1, 1) + torch.mean(dist2, 1)).mean(0)       # train.py:68-69,84
    loss.backward()
    ra = torch.from_numpy(a).to(dev()).requires_grad_(True)
    rb = torch.from_numpy(b).to(dev()).requires_grad_(True)
    D = ((ra[:, :, None] - rb[:, None]) ** 2).sum(-1)
    rloss = (D.min(2)[0].mean(1) + D.min(1)[0].mean(1)).mean(0)
    rloss.backward()
    torch.testing.assert_close(loss, rloss, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(ta.grad, ra.grad, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(tb.grad, rb.grad, rtol=1e-4, atol=1e-7)
    assert torch.equal(idx1.long(), D.min(2)[1]) and torch.equal(idx2.long(), D.min(1)[1])

    s = spb.ChamferDistance()(ta.detach(), tb.detach())
    torch.testing.assert_close(s, D.min(2)[0].mean() + D.min(1)[0].mean(), rtol=1e-5, atol=1e-8)
    # ignore_zeros with batch 1 drops points whose coordinates sum to 0
    za = ta.detach()[:1].clone(); zb = tb.detach()[:1].clone()
    za[0, :10] = 0; zb[0, 5:9] = 0
    s0 = spb.ChamferDistance(ignore_zeros=True)(za, zb)
    s1 = spb.ChamferDistance()(za[:, 10:], torch.cat([zb[:, :5], zb[:, 9:]], 1))
    torch.testing.assert_close(s0, s1)


def test_loss_epilogue():
    from softpool_b200 import ops
    a, b = clouds(32, 2048, 2048, seed=3)
    d1, d2, _, _ = ops.chamfer_forward(torch.from_numpy(a).to(dev()), torch.from_numpy(b).to(dev()))
    loss = ops.chamfer_loss(d1, d2)
    torch.testing.assert_close(loss, d1.mean(1) + d2.mean(1), rtol=1e-5, atol=1e-9)


def test_empty_clouds_leave_zeros():
    from softpool_b200 import ops
    a = torch.zeros(2, 0, 3, device=dev()); b = torch.rand(2, 7, 3, device=dev())
    d1, d2, i1, i2 = ops.chamfer_forward(a, b)
    assert d1.shape == (2, 0) and (d2 == 0).all() and (i2 == 0).all()


def _load_ref():
    from oracle import build_ref
    return build_ref.load_ref()


def test_c_oracle_matches_reference_kernel():
    """Pins oracle/chamfer_oracle.c (and through it our kernels) against the reference's own
    chamfer.cu, compiled unmodified into oracle/_ref by oracle/build_ref.py."""
    ref = _load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built (reference sources absent at build time)")
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    for ci, (B, n, m) in enumerate([(4, 64, 128), (2, 700, 1100), (2, 2048, 2048), (1, 2048, 4096)]):
        a, b = clouds(B, n, m, seed=100 + ci)
        if ci == 1:
            b[:, 600:700] = b[:, 0:100]      # duplicates across the reference's 512-point chunks
        ta, tb = torch.from_numpy(a).to(dev()), torch.from_numpy(b).to(dev())
        dist1 = torch.zeros(B, n, device=dev()); dist2 = torch.zeros(B, m, device=dev())
        idx1 = torch.zeros(B, n, dtype=torch.int32, device=dev()); idx2 = torch.zeros(B, m, dtype=torch.int32, device=dev())
        torch.cuda.synchronize()
        ref.forward(ta, tb, dist1, dist2, idx1, idx2)      # legacy default stream (chamfer.cu:142)
        torch.cuda.synchronize()
        r = co.forward(a, b)
        assert np.array_equal(idx1.cpu().numpy(), r[2]) and np.array_equal(idx2.cpu().numpy(), r[3])
        assert np.array_equal(dist1.cpu().numpy().view(np.uint32), r[0].view(np.uint32))
        assert np.array_equal(dist2.cpu().numpy().view(np.uint32), r[1].view(np.uint32))
        rng = np.random.default_rng(ci)
        g1 = rng.random((B, n), dtype=np.float32); g2 = rng.random((B, m), dtype=np.float32)
        gx1 = torch.zeros(B, n, 3, device=dev()); gx2 = torch.zeros(B, m, 3, device=dev())
        ref.backward(ta, tb, gx1, gx2, torch.from_numpy(g1).to(dev()), torch.from_numpy(g2).to(dev()), idx1, idx2)
        torch.cuda.synchronize()
        rg1, rg2 = co.backward(a, b, g1, g2, r[2], r[3])
        np.testing.assert_allclose(gx1.cpu().numpy(), rg1, rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(gx2.cpu().numpy(), rg2, rtol=1e-4, atol=1e-6)
        if ci < 2:   # small ones are brought back and committed as tests/golden/chamfer_ref_*.npz
            np.savez_compressed(os.path.join(out_dir, "chamfer_ref_%d.npz" % ci), xyz1=a, xyz2=b,
                                dist1=dist1.cpu().numpy(), dist2=dist2.cpu().numpy(),
                                idx1=idx1.cpu().numpy(), idx2=idx2.cpu().numpy(), g1=g1, g2=g2,
                                grad_xyz1=gx1.cpu().numpy(), grad_xyz2=gx2.cpu().numpy())


# ---------------------------------------------------------------------------------------------
# tensor-core path (chamfer_tc.cu): its fp16-split distance block only FILTERS candidates; every
# case below is built to stress that filter (near-ties at and below its error budget, poor
# conditioning, exact duplicates across chunk / half / super-block boundaries)
# ---------------------------------------------------------------------------------------------
def _check_exact(a, b):
    r = co.forward(a, b)
    o = cuda_forward(a, b)
    for x, y, name in zip(o, r, ("dist1", "dist2", "idx1", "idx2")):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), name
    return r


@pytest.mark.parametrize("jitter", [0.0, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4])
def test_tensor_path_lattice_near_ties(jitter):
    """Targets on a 12^3 lattice, queries at cell centres (+ jitter): 8 corners at (near-)equal distance."""
    rng = np.random.default_rng(42)
    g = (np.arange(12, dtype=np.float32) / 12.0 - 0.5)
    lat = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)            # 1728 targets
    b = np.stack([lat[rng.permutation(len(lat))] for _ in range(3)]).astype(np.float32)
    cells = lat[rng.integers(0, len(lat), size=(3, 1500))] + np.float32(0.5 / 12.0)
    a = (cells + rng.standard_normal(cells.shape) * jitter).astype(np.float32)
    r = _check_exact(a, b)
    if jitter == 0.0:
        # every query really has several exact minima: the lowest index must have been returned
        d = ((a[0, :64, None, :] - b[0, None, :, :]) ** 2).sum(-1)
        assert ((d <= d.min(1, keepdims=True) * (1 + 1e-6)).sum(1) >= 2).all()


@pytest.mark.parametrize("offset,spread", [(1000.0, 1.0), (1000.0, 1e-3), (-3e4, 10.0), (0.0, 1e-6), (5.0, 5e4)])
def test_tensor_path_poor_conditioning(offset, spread):
    """Clouds far from the origin / tiny / huge: the approximate block degrades, the result may not."""
    rng = np.random.default_rng(7)
    a = (offset + spread * (rng.random((2, 600, 3)) - 0.5)).astype(np.float32)
    b = (offset + spread * (rng.random((2, 1300, 3)) - 0.5)).astype(np.float32)
    _check_exact(a, b)


def test_tensor_path_duplicates_across_boundaries():
    """Exact duplicates 16, 64, 128, 1024 and 2048 targets apart (chunk, column half, tile, super-block):
    the FIRST one wins; also clustered data with far outliers."""
    rng = np.random.default_rng(9)
    b = (rng.random((2, 4096, 3), dtype=np.float32) - 0.5)
    for gap in (16, 64, 128, 1024, 2048):
        b[:, gap:gap + 8] = b[:, 0:8]
    b[:, 3000] = b[:, 5]
    a = b[:, rng.integers(0, 4096, size=1000)] + np.float32(0.0)
    a[:, :8] = b[:, 0:8]
    r = _check_exact(a, b)
    assert (r[2][:, :8] == np.arange(8)).all() and (r[0][:, :8] == 0).all()
    c = (rng.standard_normal((2, 2000, 3)) * 0.01).astype(np.float32)
    c[:, ::97] += 50.0
    _check_exact(c, (c[:, ::-1] * np.float32(1.0000001)).astype(np.float32).copy())


def test_tensor_and_fma_paths_agree(monkeypatch):
    """The A/B switch: SPK_CHAMFER_EXACT=1 forces the plain float32 FMA kernel; both paths are bit-identical."""
    from softpool_b200 import _lib
    assert _lib.lib().chamfer_fwd_workspace_bytes(4, 2048, 2048) > 0
    a, b = clouds(4, 2048, 3000, seed=77)
    tc = cuda_forward(a, b)
    monkeypatch.setenv("SPK_CHAMFER_EXACT", "1")
    fma = cuda_forward(a, b)
    for x, y in zip(tc, fma):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


@pytest.mark.parametrize("B,n,m", [(2, 100, 700), (2, 400, 1400), (1, 2100, 1500)])
def test_non_finite_inputs_match_the_reference_order(B, n, m):
    """NaN / inf coordinates: the reference's answer follows from its loop order -- targets in batches of 512, the batch's
    first target initialises the batch minimum, batches merge with a strict `>` (chamfer.cu:16-129; restated in
    oracle/chamfer_oracle.c).  Samples with a non-finite coordinate are evaluated in exactly that order by every path
    (plain kernel below 256 x 256 pairs, dense and sorted tensor paths above), so dist AND idx match the oracle:
    NaN where the oracle has NaN (payloads aside), bit-exact elsewhere."""
    a, b = clouds(B, n, m, seed=5 + n)
    a[0, 3] = np.nan                        # a NaN query: every distance NaN -> (NaN, 0)
    b[0, 0] = np.nan                        # the very first target initialises the result with NaN: nothing is < NaN
    if m > 1024:
        b[0, 512, 1] = np.nan               # a NaN at a batch start hides the rest of that batch
        b[0, 1030, 2] = np.nan              # a NaN inside a batch is simply never chosen
    if B > 1:
        b[1, 10, 1] = np.inf                # inf - finite = inf distances, inf - inf = NaN
        a[1, 7, 0] = -np.inf
    r = co.forward(a, b)
    o = cuda_forward(a, b)
    for x, y, name in zip(o, r, ("dist1", "dist2", "idx1", "idx2")):
        assert np.array_equal(x, y, equal_nan=(x.dtype == np.float32)), name
        if x.dtype == np.float32:
            fin = np.isfinite(y)
            assert np.array_equal(x[fin].view(np.uint32), y[fin].view(np.uint32)), name
    assert np.isnan(r[0][0]).any()          # the case really exercises NaN results


@pytest.mark.parametrize("P,B,n,m", [(3, 4, 2048, 2048), (2, 3, 700, 1100), (4, 2, 100, 90), (3, 1, 4096, 5000)])
def test_multi_prediction_call_matches_separate_calls(P, B, n, m):
    """chamfer_fwd_multi_f32 / ops.chamfer_multi: P predictions against one shared ground truth (train.py:68-86) in one call
    -- bit-identical to P separate calls (which are bit-identical to the oracle), loss within float rounding, gradients of
    the predictions as in the separate calls and the ground truth's gradient = their sum."""
    from softpool_b200 import ops
    rng = np.random.default_rng(P * 100 + n)
    preds = (rng.random((P, B, n, 3), dtype=np.float32) - 0.5)
    gt = (rng.random((B, m, 3), dtype=np.float32) - 0.5)
    preds[1, 0, :5] = gt[0, :5]                               # zero distances / ties
    tp = torch.from_numpy(preds).to(dev()).requires_grad_(True)
    tg = torch.from_numpy(gt).to(dev()).requires_grad_(True)
    loss, d1, d2, i1, i2 = ops.chamfer_multi(tp, tg)
    w = torch.from_numpy(rng.random((P, B), dtype=np.float32)).to(dev())
    (loss * w).sum().backward()
    g_gt = torch.zeros_like(tg)
    for p in range(P):
        r = co.forward(preds[p], gt)
        assert np.array_equal(i1[p].cpu().numpy(), r[2]) and np.array_equal(i2[p].cpu().numpy(), r[3])
        assert np.array_equal(d1[p].detach().cpu().numpy().view(np.uint32), r[0].view(np.uint32))
        assert np.array_equal(d2[p].detach().cpu().numpy().view(np.uint32), r[1].view(np.uint32))
        np.testing.assert_allclose(loss[p].detach().cpu().numpy(), r[0].mean(1) + r[1].mean(1), rtol=1e-5, atol=1e-9)
        a = torch.from_numpy(preds[p]).to(dev()).requires_grad_(True)
        g = torch.from_numpy(gt).to(dev()).requires_grad_(True)
        (ops.chamfer_mean_loss(a, g) * w[p]).sum().backward()
        torch.testing.assert_close(tp.grad[p], a.grad, rtol=1e-4, atol=1e-7)
        g_gt += g.grad
    torch.testing.assert_close(tg.grad, g_gt, rtol=1e-4, atol=1e-7)


def test_selfcheck_switch_runs_clean():
    """SPK_CHAMFER_SELFCHECK=1 (read once per process, so a subprocess): every tensor-path call re-evaluates sample 0 with the
    plain float32 kernel and fails loudly on a difference -- the run-time guard for the filter's hardware assumption."""
    import subprocess
    import sys
    code = ("import torch, sys; sys.path.insert(0, %r); from softpool_b200 import ops\n"
            "g = torch.Generator().manual_seed(1)\n"
            "for (B, n, m) in [(4, 2048, 2048), (2, 700, 1100), (32, 4096, 4096), (1, 2048, 16384)]:\n"
            "    a = (torch.rand(B, n, 3, generator=g) - 0.5).cuda(); b = (torch.rand(B, m, 3, generator=g) - 0.5).cuda()\n"
            "    ops.chamfer_forward(a, b); torch.cuda.synchronize()\n"
            "print('selfcheck ok')\n") % ROOT
    env = dict(os.environ, SPK_CHAMFER_SELFCHECK="1")
    env.pop("SPK_CHAMFER_PATH", None)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "selfcheck ok" in r.stdout, r.stderr[-2000:]


def test_random_shapes_vs_oracle():
    """40 random cloud sizes on both sides of the tensor-path threshold, ragged against the 128/1024 tiling."""
    rng = np.random.default_rng(77)
    for trial in range(40):
        B = int(rng.integers(1, 5)); n = int(rng.integers(1, 1500)); m = int(rng.integers(1, 2600))
        a, b = clouds(B, n, m, seed=1000 + trial, scale=float(rng.choice([1.0, 0.01, 37.0])))
        if trial % 4 == 0:
            b[:, : min(m, 50)] = a[:, :1]               # many exact ties at distance 0 for the first query
        r = co.forward(a, b)
        o = cuda_forward(a, b)
        tag = "trial %d: B=%d n=%d m=%d" % (trial, B, n, m)
        for x, y in zip(o, r):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), tag


@pytest.mark.parametrize("split", [1, 2, 5, 64])
def test_split_jobs_are_bit_identical(split, monkeypatch):
    """Sub-jobs over target ranges (used when the grid is underfilled) merge through a 64-bit atomicMin:
    the result is the same first minimum, whatever the split."""
    a, b = clouds(2, 3000, 5000, seed=split)
    b[:, 4000:4100] = b[:, 100:200]                   # duplicates in different sub-jobs: the lower index wins
    r = co.forward(a, b)
    monkeypatch.setenv("SPK_TC_SPLIT", str(split))
    o = cuda_forward(a, b)
    for x, y in zip(o, r):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


@pytest.mark.parametrize("kernel,ctas", [("auto", 1), ("auto", 2), ("auto", 4), ("generic", 1), ("generic", 2), ("generic", 4)])
@pytest.mark.parametrize("B,n,m", [(2, 3001, 5000), (1, 700, 16384), (3, 4097, 4096)])
def test_sort_kernels_and_cluster_sizes_are_bit_identical(kernel, ctas, B, n, m, monkeypatch):
    """The sorted search's prep step runs as a cluster of 2 x {1, 2, 4} CTAs per sample, in a register form (slices of up to
    4096 points) or the generic form: whichever sorts the clouds, dist / idx are the oracle's bit for bit (ragged sizes,
    slices that end inside a chunk, an empty last slice, duplicates that land in different slices)."""
    a, b = clouds(B, n, m, seed=ctas + n)
    b[:, m - 100:] = b[:, 100:200]
    r = co.forward(a, b)
    monkeypatch.setenv("SPK_CHAMFER_PATH", "sorted")
    monkeypatch.setenv("SPK_SORT_CTAS", str(ctas))
    if kernel != "auto":
        monkeypatch.setenv("SPK_SORT_KERNEL", kernel)
    o = cuda_forward(a, b)
    for x, y in zip(o, r):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


def test_cluster_backward_matches_two_kernel_form(monkeypatch):
    """One cluster launch per call (default) against the two-launch form (pass A, pass B), on a size whose
    points exceed what a cluster's threads hold in registers, and against the oracle."""
    from softpool_b200 import ops
    B, n, m = 2, 21000, 17001
    a, b = clouds(B, n, m, seed=9)
    rng = np.random.default_rng(2)
    g1 = rng.random((B, n), dtype=np.float32); g2 = rng.random((B, m), dtype=np.float32)
    t = lambda v: torch.from_numpy(v).to(dev())
    d1, d2, i1, i2 = ops.chamfer_forward(t(a), t(b))
    one = ops.chamfer_backward(t(a), t(b), t(g1), t(g2), i1, i2)
    monkeypatch.setenv("SPK_CH_BWD_TWO_KERNELS", "1")
    two = ops.chamfer_backward(t(a), t(b), t(g1), t(g2), i1, i2)
    for x, y in zip(one, two):
        torch.testing.assert_close(x, y, rtol=1e-4, atol=1e-7)
    rg1, rg2 = co.backward(a, b, g1, g2, i1.cpu().numpy(), i2.cpu().numpy())
    np.testing.assert_allclose(one[0].cpu().numpy(), rg1, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(one[1].cpu().numpy(), rg2, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("B,n,m", [(32, 2048, 2048), (1, 4096, 16384), (3, 700, 1100), (2, 100, 90), (5, 2049, 1023)])
def test_fused_loss_matches_mean_of_distances(B, n, m):
    """chamfer_fwd_loss_f32: mean(dist1,1) + mean(dist2,1) from the forward launch itself (tensor path, split
    jobs, and the small-shape path), dist/idx unchanged."""
    from softpool_b200 import ops
    a, b = clouds(B, n, m, seed=B + n)
    t = lambda v: torch.from_numpy(v).to(dev())
    d1, d2, i1, i2 = ops.chamfer_forward(t(a), t(b))
    e1, e2, j1, j2, loss = ops.chamfer_forward(t(a), t(b), want_loss=True)
    assert torch.equal(d1, e1) and torch.equal(d2, e2) and torch.equal(i1, j1) and torch.equal(i2, j2)
    ref = d1.double().mean(1) + d2.double().mean(1)
    torch.testing.assert_close(loss.double(), ref, rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(loss, ops.chamfer_loss(d1, d2), rtol=1e-5, atol=1e-9)
    for _ in range(3):                                        # the accumulator is re-zeroed by every call
        torch.testing.assert_close(ops.chamfer_forward(t(a), t(b), want_loss=True)[4], loss, rtol=1e-5, atol=1e-9)


def test_mean_loss_function_backward():
    """ops.chamfer_mean_loss == mean(dist1,1)+mean(dist2,1) of chamferDist, forward and gradients."""
    import softpool_b200 as spb
    from softpool_b200 import ops
    a, b = clouds(4, 1500, 2048, seed=21)
    ta = torch.from_numpy(a).to(dev()).requires_grad_(True); tb = torch.from_numpy(b).to(dev()).requires_grad_(True)
    w = torch.tensor([1.0, 0.5, 2.0, -1.0], device=dev())
    (ops.chamfer_mean_loss(ta, tb) * w).sum().backward()
    ua = torch.from_numpy(a).to(dev()).requires_grad_(True); ub = torch.from_numpy(b).to(dev()).requires_grad_(True)
    d1, d2, _, _ = spb.chamferDist()(ua, ub)
    ((d1.mean(1) + d2.mean(1)) * w).sum().backward()
    torch.testing.assert_close(ta.grad, ua.grad, rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(tb.grad, ub.grad, rtol=1e-4, atol=1e-8)


def test_real_clouds_from_the_reference_artefacts():
    """Partial scans (2048 points, > 1200 exact duplicates each from resample_pcd) against their 16384-point ground
    truths, and network outputs against ground-truth subsets: the reference's own PLY artefacts
    (tests/golden/make_chamfer_real.py).  Bit-exact, batched and one cloud at a time (the val.py shape, B=1)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "chamfer_real.npz"))
    o = cuda_forward(g["a"], g["b"])
    for x, y in zip(o, (g["d1"], g["d2"], g["i1"], g["i2"])):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
    o = cuda_forward(g["c"], np.ascontiguousarray(g["b"][:, :4096]))
    for x, y in zip(o, (g["e1"], g["e2"], g["j1"], g["j2"])):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
    for s in range(g["a"].shape[0]):
        o = cuda_forward(g["a"][s:s + 1], g["b"][s:s + 1])
        for x, y in zip(o, (g["d1"][s:s + 1], g["d2"][s:s + 1], g["i1"][s:s + 1], g["i2"][s:s + 1])):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


def test_tensor_path_randomized_stress(monkeypatch):
    """Clustered, surface-like and near-degenerate clouds at several scales and offsets: the tensor path (fp16
    accumulators, packed-minima filter, exact refinement) against the plain float32 kernel (SPK_CHAMFER_EXACT=1),
    bit for bit.  This is the adversary of the filter's error model: many targets within 1e-7..1e-3 of the minimum."""
    rng = np.random.default_rng(424242)
    for trial in range(14):
        B = int(rng.integers(1, 4)); n = int(rng.integers(300, 5000)); m = int(rng.integers(300, 6000))
        kind = trial % 4
        if kind == 0:      # points on a few planes + tiny normal noise
            base = rng.random((B, n + m, 3), dtype=np.float32) - 0.5
            base[..., 2] = np.round(base[..., 2] * 3) / 3 + rng.normal(0, 1e-6, (B, n + m)).astype(np.float32)
        elif kind == 1:    # tight clusters
            cent = rng.random((B, 12, 3), dtype=np.float32) - 0.5
            pick = rng.integers(0, 12, (B, n + m))
            base = np.take_along_axis(cent, pick[..., None].repeat(3, -1), 1) + rng.normal(0, 10.0 ** -int(rng.integers(3, 8)), (B, n + m, 3)).astype(np.float32)
        elif kind == 2:    # a line (rank-1 geometry) with jitter
            t = rng.random((B, n + m, 1), dtype=np.float32)
            base = t * np.array([0.7, -0.2, 0.4], np.float32) + rng.normal(0, 1e-5, (B, n + m, 3)).astype(np.float32)
        else:              # uniform with many exact duplicates
            base = rng.random((B, n + m, 3), dtype=np.float32) - 0.5
            base[:, ::3] = base[:, 1::3][:, : base[:, ::3].shape[1]]
        scale = float(rng.choice([1.0, 1e-3, 250.0])); off = float(rng.choice([0.0, 17.0, -900.0]))
        base = (base * scale + off).astype(np.float32)
        a, b = np.ascontiguousarray(base[:, :n]), np.ascontiguousarray(base[:, n:])
        monkeypatch.delenv("SPK_CHAMFER_EXACT", raising=False)
        tc = cuda_forward(a, b)
        monkeypatch.setenv("SPK_CHAMFER_EXACT", "1")
        fma = cuda_forward(a, b)
        for x, y in zip(tc, fma):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), "trial %d kind %d B=%d n=%d m=%d scale=%g off=%g" % (trial, kind, B, n, m, scale, off)
