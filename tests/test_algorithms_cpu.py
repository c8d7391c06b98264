"""numpy emulations of the two non-obvious algorithms of the CUDA kernels, checked against the oracle on the CPU:
the radix-select / early-finish / single-scan compaction of sp_topk_kernel and the inverse-table ("pull") backward
of sp_gather_bwd_pull_kernel.  They document the algorithms and catch a logic regression without a GPU."""
import numpy as np
import pytest

from oracle import softpool_oracle as so


def emulate_topk_row(keys_row, k):
    """One (b, r) row the way sp_topk_kernel does it: 4-bit radix select with an early finish when the chosen bucket
    holds <= 32 keys, selection of everything > V plus the first `need` keys == V in index order, sort of the survivors
    as unique (key, ~index) words."""
    u = so.order_key(keys_row[None, None, :])[0, 0].astype(np.uint64)
    V, want, fin_sh = 0, k, -1
    for rnd in range(8):
        sh = 28 - 4 * rnd
        cand = np.ones(len(u), bool) if rnd == 0 else (u >> np.uint64(sh + 4)) == np.uint64(V >> (sh + 4))
        digit = ((u >> np.uint64(sh)) & np.uint64(15)).astype(int)
        c = np.bincount(digit[cand], minlength=16)
        S = np.cumsum(c[::-1])[::-1]                                   # suffix sums: # candidates with digit >= d
        dsel = max(d for d in range(16) if S[d] >= want)
        want -= S[dsel] - c[dsel]
        V |= dsel << sh
        if rnd < 7 and c[dsel] <= 32:
            fin_sh = sh
            break
    if fin_sh >= 0:                                                    # one warp ranks the bucket's keys directly
        bucket = u[(u >> np.uint64(fin_sh)) == np.uint64(V >> fin_sh)]
        V = int(np.sort(bucket)[::-1][want - 1])
    gt, eq = u > V, u == V
    need = k - int(gt.sum())
    take = gt | (eq & (np.cumsum(eq) - eq < need))                     # the first `need` ties in index order
    # single-scan output position of the kernel: #(>V) before + min(#(==V) before, need)
    pos = (np.cumsum(gt) - gt) + np.minimum(np.cumsum(eq) - eq, need)
    assert sorted(pos[take]) == list(range(k))
    idx = np.nonzero(take)[0]
    words = (u[idx] << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - idx.astype(np.uint64))
    return idx[np.argsort(words)[::-1]]


@pytest.mark.parametrize("seed,N,k,quant", [(0, 2048, 32, 0), (1, 2048, 256, 0), (2, 500, 77, 8), (3, 64, 64, 2), (4, 1000, 1, 0), (5, 300, 299, 1)])
def test_topk_select_emulation_matches_stable_argsort(seed, N, k, quant):
    rng = np.random.default_rng(seed)
    keys = rng.standard_normal(N).astype(np.float32)
    if quant:
        keys = np.round(keys * quant) / quant                          # many ties, +-0
    if seed == 2:
        keys[[3, 77, 200]] = np.nan; keys[5] = -0.0; keys[6] = 0.0; keys[9] = np.inf; keys[10] = -np.inf
    ref = so.topk_indices(keys[None, None, :], k)[0, 0]
    assert np.array_equal(emulate_topk_row(keys, k), ref)


def emulate_pull_backward(g_cube, g_cabins, idx, cab_arg, N):
    """grad_x the way sp_gather_bwd_pull_kernel forms it: window-max gradient folded into the winning slot, inverse
    table built by prepending in DESCENDING region order, chains walked from `first` (ascending regions)."""
    B, C, R, k = g_cube.shape
    cab = g_cabins.shape[-1]
    NONE = 0xFFFF
    out = np.zeros((B, C, N), np.float32)
    for b in range(B):
        first = np.full(N, NONE, np.int64); link = np.full(R * k, NONE, np.int64)
        for r in range(R - 1, -1, -1):
            for j in range(k):
                s = r * k + j; n = idx[b, r, j]
                link[s] = first[n]; first[n] = s
        g = g_cube[b].reshape(C, R * k).copy()
        for c in range(C):
            for r in range(R):
                for w in range(cab):
                    g[c, r * k + cab_arg[b, c, r, w]] += g_cabins[b, c, r, w]       # the oracle's cab_arg is region-relative
        for n in range(N):
            s = first[n]
            acc = np.zeros(C, np.float32)
            while s != NONE:
                acc = (acc + g[:, s]).astype(np.float32)
                s = link[s]
            out[b, :, n] = acc
    return out


def test_pull_backward_emulation_is_bit_identical_to_the_oracle():
    rng = np.random.default_rng(7)
    B, C, N, R, k, cab = 2, 5, 96, 4, 24, 8                            # R*k = N: every point selected about once
    x = rng.standard_normal((B, C, N), dtype=np.float32)
    keys = np.round(rng.standard_normal((B, R, N)).astype(np.float32) * 4) / 4
    f = so.softpool_forward(x, keys, k, cab)
    g_cube = rng.standard_normal((B, C, R, k), dtype=np.float32)
    g_cab = rng.standard_normal((B, C, R, cab), dtype=np.float32)
    ref = so.softpool_backward(g_cube, g_cab, f["idx"], f["cab_arg"], N)
    emu = emulate_pull_backward(g_cube, g_cab, f["idx"], f["cab_arg"], N)
    assert np.array_equal(emu.view(np.uint32), ref.view(np.uint32))
    multi = np.bincount(f["idx"][0].ravel(), minlength=N)
    assert multi.max() >= 2 and multi.min() == 0                       # chains and empty points both occur
