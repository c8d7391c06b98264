"""CPU: the numpy oracle (oracle/softpool_oracle.py) against the golden outputs of the reference
`softpool.py` itself (tests/golden/make_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases
from oracle import softpool_oracle as so


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def load(case):
    g = np.load(os.path.join(GOLDEN, "softpool_%s.npz" % case))
    B, C, N, R, sp_ratio, cab, k, tie_free = [int(v) for v in g["meta"]]
    return g, (B, C, N, R, sp_ratio, cab, k), bool(tie_free)


def check_against(out, grad_x, g, prefix):
    # indices: bit-exact (sp_idx is the float32, (R+3)-fold replicated cube of softpool.py:146)
    assert np.array_equal(out["sp_idx"], g[prefix + "sp_idx"])
    # features are copies / maxima of x: bit-exact, NaN payloads included
    assert np.array_equal(bits(out["sp_cube"]), bits(g[prefix + "sp_cube"]))
    assert np.array_equal(bits(out["cabins"]), bits(g[prefix + "cabins"]))
    # autograd sums the same terms in another order: float32 tolerance, stated here (1e-4 rel)
    np.testing.assert_allclose(grad_x, g[prefix + "grad_x"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("case", golden_cases("softpool"))
def test_oracle_matches_reference(case):
    """Tie-free fixtures: the UNMODIFIED reference pins every output.  Tie fixtures: the reference
    run with torch.sort(stable=True) (arrays st_*) pins the stable tie order."""
    g, (B, C, N, R, sp_ratio, cab, k), tie_free = load(case)
    out = so.softpool_forward(g["x"], g["keys"], k, cab)
    grad_x = so.softpool_backward(g["g_cube"], g["g_cabins"], out["idx"], out["cab_arg"], N)
    assert np.array_equal(out["id_activa"], g["id_activa"]) and out["id_activa"].dtype == np.int64
    check_against(out, grad_x, g, "" if tie_free else "st_")


@pytest.mark.parametrize("case", [c for c in golden_cases("softpool")])
def test_oracle_vs_unmodified_reference_on_ties(case):
    """With ties the unmodified reference (stable=False) may order equal keys differently: then
    (i) the selected key VALUES must still agree slot by slot, and (ii) given the reference's own
    indices, gather / window max / backward must reproduce its outputs."""
    g, (B, C, N, R, sp_ratio, cab, k), tie_free = load(case)
    ref_idx = g["sp_idx"][:, 0].astype(np.int64)                       # (B,R,k)
    assert np.array_equal(g["sp_idx"], np.broadcast_to(ref_idx[:, None].astype(np.float32), g["sp_idx"].shape))
    mine = so.topk_indices(g["keys"], k)
    ku = so.order_key(g["keys"])
    assert np.array_equal(np.take_along_axis(ku, mine, -1), np.take_along_axis(ku, ref_idx, -1))
    out = so.softpool_forward(g["x"], g["keys"], k, cab, idx=ref_idx)
    grad_x = so.softpool_backward(g["g_cube"], g["g_cabins"], out["idx"], out["cab_arg"], N)
    check_against(out, grad_x, g, "")


def test_order_key_rules():
    k = np.array([np.nan, np.inf, 1.0, 0.0, -0.0, -1.0, -np.inf], np.float32)
    u = so.order_key(k)
    assert u[0] == 0xFFFFFFFF and u[3] == u[4]
    assert list(np.argsort(~u, kind="stable")) == [0, 1, 2, 3, 4, 5, 6]
    # ties keep ascending index
    idx = so.topk_indices(np.array([[[1, 2, 2, 1, 2]]], np.float32), 5)
    assert idx.tolist() == [[[1, 2, 4, 0, 3]]]


@pytest.mark.parametrize("case", [c for c in golden_cases("softpool")])
def test_torch_port_matches_reference(case):
    """The timed CPU baseline (oracle/softpool_torch_port.py) reproduces the unmodified reference."""
    import torch
    from oracle import softpool_torch_port as port
    g, (B, C, N, R, sp_ratio, cab, k), tie_free = load(case)
    x = torch.from_numpy(g["x"])
    grad = port.forward_backward(x, torch.from_numpy(g["keys"]), k, cab,
                                 torch.from_numpy(g["g_cube"]), torch.from_numpy(g["g_cabins"]))
    sp_cube, sp_idx, cabins, id_activa = port.forward(x, torch.from_numpy(g["keys"]), k, cab)
    assert np.array_equal(sp_idx.numpy(), g["sp_idx"])
    assert np.array_equal(id_activa.numpy(), g["id_activa"])
    assert np.array_equal(bits(sp_cube.numpy()), bits(g["sp_cube"]))
    assert np.array_equal(bits(cabins.numpy()), bits(g["cabins"]))
    np.testing.assert_allclose(grad.numpy(), g["grad_x"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("case", golden_cases("chamfer_ref"))
def test_chamfer_oracle_matches_reference_kernel_outputs(case):
    """tests/golden/chamfer_ref_*.npz are outputs of the reference's own chamfer.cu (compiled
    unmodified for sm_100 into oracle/_ref, run on a B200 by
    tests/test_chamfer_gpu.py::test_c_oracle_matches_reference_kernel): distances and indices of
    the C restatement must match them bit for bit, gradients within float32 atomics noise."""
    from oracle import chamfer_oracle as co
    g = np.load(os.path.join(GOLDEN, "chamfer_ref_%s.npz" % case))
    d1, d2, i1, i2 = co.forward(g["xyz1"], g["xyz2"])
    assert np.array_equal(i1, g["idx1"]) and np.array_equal(i2, g["idx2"])
    assert np.array_equal(bits(d1), bits(g["dist1"])) and np.array_equal(bits(d2), bits(g["dist2"]))
    gx1, gx2 = co.backward(g["xyz1"], g["xyz2"], g["g1"], g["g2"], i1, i2)
    np.testing.assert_allclose(gx1, g["grad_xyz1"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(gx2, g["grad_xyz2"], rtol=1e-4, atol=1e-6)


def test_real_cloud_fixture_is_what_the_oracle_computes():
    """tests/golden/chamfer_real.npz (reference PLY artefacts): the stored answers are the C oracle's, and a plain
    numpy brute force agrees on a slice (first-minimum rule on the many exact duplicates included)."""
    import numpy as np
    from oracle import chamfer_oracle as co
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chamfer_real.npz"))
    a, b = g["a"][:1, :256], g["b"][:1, :4096]
    d1, d2, i1, i2 = co.forward(np.ascontiguousarray(a), np.ascontiguousarray(b))
    diff = b[0][None, :, :].astype(np.float64) - a[0][:, None, :].astype(np.float64)
    dd = (diff * diff).sum(-1)
    assert np.allclose(np.take_along_axis(dd, i1[0][:, None].astype(np.int64), 1)[:, 0], dd.min(1), rtol=1e-5, atol=1e-12)
    assert np.allclose(d1[0], dd.min(1), rtol=1e-5, atol=1e-12)
    # first minimum: no earlier target is (float32-)closer than the chosen one
    chosen = np.take_along_axis(dd, i1[0][:, None].astype(np.int64), 1)
    earlier = np.where(np.arange(dd.shape[1])[None, :] < i1[0][:, None], dd, np.inf)
    assert (earlier.min(1) >= chosen[:, 0] * (1 - 1e-6)).all()
    full = co.forward(g["a"], g["b"])
    for x, y in zip(full, (g["d1"], g["d2"], g["i1"], g["i2"])):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
