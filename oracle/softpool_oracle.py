"""CPU oracle for the SoftPool sort/top-k + gather hot path (numpy restatement).

TEST INFRASTRUCTURE ONLY -- never imported by ``softpool_b200/`` (the product path
fails loudly without its CUDA library).  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this.

What it restates (reference = wangyida/softpool @ 31a2d18):

* ``Sorter.forward``            softpool.py:93-96    argmax over the R activation rows
* region loop of ``SoftPool``   softpool.py:139-147  descending sort -> first k indices ->
                                                     gather of all C channels, float32 index cube
* ``train2cabins``              softpool.py:71-85    window max over k // cab consecutive slots
* backward                      autograd of softpool.py:139-151 (no hand-written backward exists)

Third-party arithmetic the reference calls and does not vendor: ``torch.sort``,
``torch.gather``, ``torch.max``, ``torch.argmax`` (PyTorch, un-pinned by the reference;
README.md:42 recommends 1.2.0).  Their CPU behaviour in torch 2.11.0 is restated here:

* ``torch.sort(descending=True)``: NaN compares greater than everything (sorts first);
  -0.0 == +0.0.  The reference passes the default ``stable=False`` (softpool.py:140), so the
  order of EQUAL keys is unspecified by torch (this image's AVX-512 CPU build is deterministic
  but NOT stable; SURVEY.md section 7 claimed otherwise -- corrected in DESIGN.md).  The contract
  restated here is the stable order -- equal keys keep ascending original index -- which is
  what the reference itself produces with ``stable=True`` and coincides with its unmodified
  output whenever the keys of a row are distinct.
* ``torch.argmax`` / ``torch.max(dim)``: first maximum wins, the first NaN wins outright.

PINNED: ``tests/golden/softpool_*.npz`` hold outputs of the reference ``softpool.py``
itself, imported in the build container by ``tests/golden/make_golden.py`` (unmodified, and
with ``torch.sort`` forced stable for the tie fixtures); ``tests/test_oracle_cpu.py`` checks
this file against every one of them.
"""
import numpy as np


def order_key(keys):
    """float32 -> uint32 whose DESCENDING unsigned order is torch's descending sort order.

    -0.0 is canonicalised to +0.0 (they tie), every NaN maps to 0xFFFFFFFF (NaN first,
    NaNs tie with each other).  The CUDA kernel uses the identical transform.
    """
    k = np.ascontiguousarray(keys, dtype=np.float32)
    bits = k.view(np.uint32).copy()
    bits[bits == np.uint32(0x80000000)] = 0                      # -0 -> +0
    neg = (bits >> 31).astype(bool)
    u = np.where(neg, ~bits, bits | np.uint32(0x80000000)).astype(np.uint32)
    u[np.isnan(k)] = np.uint32(0xFFFFFFFF)
    return u


def topk_indices(keys, k):
    """(B,R,N) float32 -> (B,R,k) int64: first k entries of the stable descending argsort
    of every row (softpool.py:140-142)."""
    u = order_key(keys)
    order = np.argsort(~u, axis=-1, kind="stable")               # ascending ~u == descending u
    return order[..., :k].astype(np.int64)


def argmax_regions(keys):
    """(B,R,N) -> (B,N) int64, softpool.py:95 (first max over R; first NaN wins)."""
    u = order_key(keys)
    return np.argmax(u, axis=1).astype(np.int64)                 # np.argmax returns the first maximum


def window_argmax(sp_cube, cab):
    """train2cabins (softpool.py:71-85) -> (cabins, cab_arg).  cab_arg is the slot j inside the
    region (0 <= j < k) of each window's first maximum; trailing k mod cab slots are ignored."""
    B, C, R, k = sp_cube.shape
    w = k // cab
    win = sp_cube[..., : w * cab].reshape(B, C, R, cab, w)
    u = order_key(win).reshape(win.shape)
    a = np.argmax(u, axis=-1)
    cabins = np.take_along_axis(win, a[..., None], axis=-1)[..., 0]
    cab_arg = (a + np.arange(cab)[None, None, None, :] * w).astype(np.int32)
    return cabins.astype(np.float32), cab_arg


def softpool_forward(x, keys, k, cab=8, idx=None):
    """x (B,C,N) f32, keys (B,R,N) f32  ->  dict with the reference's four outputs
    (softpool.py:171) plus the two integer tensors the backward needs.  ``idx`` (B,R,k) overrides
    the top-k selection (used to check gather / window max / backward against a reference run
    whose tie order differs)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    B, C, N = x.shape
    R = keys.shape[1]
    idx = topk_indices(keys, k) if idx is None else np.asarray(idx, dtype=np.int64)   # (B,R,k)
    sp_cube = np.empty((B, C, R, k), np.float32)
    for r in range(R):                                            # softpool.py:139-145
        sp_cube[:, :, r, :] = np.take_along_axis(x, np.broadcast_to(idx[:, None, r, :], (B, C, k)), axis=2)
    sp_idx = np.broadcast_to(idx[:, None, :, :].astype(np.float32), (B, R + 3, R, k)).copy()
    cabins, cab_arg = window_argmax(sp_cube, cab)
    return dict(sp_cube=sp_cube, sp_idx=sp_idx, cabins=cabins, id_activa=argmax_regions(keys),
                idx=idx.astype(np.int32), cab_arg=cab_arg)


def softpool_backward(g_cube, g_cabins, idx, cab_arg, N):
    """grad_x (B,C,N): every selected slot sends (g_cube + its window's g_cabins if it is the
    window arg-max) back to point idx[b,r,j]; regions accumulate in ascending r in float32
    (the order the CUDA kernel uses; autograd's own order differs by < 1 ulp per add)."""
    B, C, R, k = g_cube.shape
    g = np.array(g_cube, dtype=np.float32, copy=True)
    if g_cabins is not None:
        cab = g_cabins.shape[-1]
        hit = np.zeros_like(g)
        np.put_along_axis(hit, cab_arg.astype(np.int64), g_cabins.astype(np.float32), axis=-1)
        touched = np.zeros(g.shape, bool)
        np.put_along_axis(touched, cab_arg.astype(np.int64), True, axis=-1)
        g = np.where(touched, g + hit, g)
    grad_x = np.zeros((B, C, N), np.float32)
    for r in range(R):
        dst = np.broadcast_to(idx[:, None, r, :].astype(np.int64), (B, C, k))
        cur = np.take_along_axis(grad_x, dst, axis=2)
        np.put_along_axis(grad_x, dst, cur + g[:, :, r, :], axis=2)   # indices unique within a region
    return grad_x
