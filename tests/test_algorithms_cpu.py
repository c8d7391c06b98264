"""numpy emulations of the non-obvious algorithms of the CUDA kernels, checked against the oracle on the CPU:
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


def emulate_fused_bitonic(words, NT=256):
    """The survivor sort of sp_topk_kernel (softpool_topk.cu: block_stages / ce_stage): blocks of 32 * EPL words sorted in
    a warp's registers (strides >= 32 between a lane's own words, smaller ones lane to lane), then per merge size the strides
    that cross blocks through shared memory and the rest again inside the blocks.  `words` unique, length a power of two >= 64."""
    buf = np.array(words, dtype=np.uint64)
    K2s = len(buf)
    EPL = 4 if K2s >= 1024 else 2
    BL = 32 * EPL

    def block_stages(base, size_lo, size_hi):
        e = [buf[base + 32 * u: base + 32 * u + 32].copy() for u in range(EPL)]          # e[u][lane]
        lane = np.arange(32)
        size = size_lo
        while size <= size_hi:
            desc = [((base + lane + 32 * u) & size) == 0 for u in range(EPL)]
            rs = EPL // 2
            while rs > 0:                                                               # stride 32 * rs: the lane's own registers
                if size >= 64 * rs:
                    for u in range(EPL):
                        if (u & rs) == 0:
                            sw = (e[u] < e[u + rs]) == desc[u]
                            a = np.where(sw, e[u + rs], e[u]); c = np.where(sw, e[u], e[u + rs])
                            e[u], e[u + rs] = a, c
                rs >>= 1
            stride = min(size >> 1, 16)
            while stride > 0:                                                           # lane to lane (shuffle xor)
                lower = (lane & stride) == 0
                for u in range(EPL):
                    o = e[u][lane ^ stride]
                    e[u] = np.where(lower == desc[u], np.maximum(o, e[u]), np.minimum(o, e[u]))
                stride >>= 1
            size <<= 1
        for u in range(EPL):
            buf[base + 32 * u: base + 32 * u + 32] = e[u]

    def ce_stage(size, stride):
        for p in range(K2s >> 1):
            i = ((p & ~(stride - 1)) << 1) | (p & (stride - 1)); j = i + stride
            desc = (i & size) == 0
            if (buf[i] < buf[j]) == desc:
                buf[i], buf[j] = buf[j], buf[i]

    if K2s == 64:
        EPL, BL = 2, 64
    for base in range(0, K2s, BL):
        block_stages(base, 2, BL)
    size = 2 * BL
    while size <= K2s:
        stride = size >> 1
        while stride >= BL:
            ce_stage(size, stride)
            stride >>= 1
        for base in range(0, K2s, BL):
            block_stages(base, size, size)
        size <<= 1
    return buf


@pytest.mark.parametrize("K2s", [64, 128, 256, 512, 1024, 2048])
def test_fused_bitonic_network_sorts_descending(K2s):
    rng = np.random.default_rng(K2s)
    keys = rng.integers(0, 1 << 20, K2s).astype(np.uint64)                              # ties in the key half
    words = (keys << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - np.arange(K2s, dtype=np.uint64))
    words[rng.permutation(K2s)[: K2s // 5]] = 0                                         # empty slots (fewer survivors than slots)
    words[0] = (np.uint64(7) << np.uint64(32)) | np.uint64(1)                           # keep at least a few distinct
    out = emulate_fused_bitonic(words)
    assert np.array_equal(out, np.sort(words)[::-1])


def hilbert_of_cell(x0, x1, x2, bits=4):
    """Skilling's axes-to-transpose, then bit interleave: what chamfer_tc.cu evaluates at compile time into its table."""
    M = 1 << (bits - 1)
    Q = M
    while Q > 1:
        P = Q - 1
        if x0 & Q: x0 ^= P
        if x1 & Q: x0 ^= P
        else: t = (x0 ^ x1) & P; x0 ^= t; x1 ^= t
        if x2 & Q: x0 ^= P
        else: t = (x0 ^ x2) & P; x0 ^= t; x2 ^= t
        Q >>= 1
    x1 ^= x0; x2 ^= x1
    t = 0; Q = M
    while Q > 1:
        if x2 & Q: t ^= Q - 1
        Q >>= 1
    x0 ^= t; x1 ^= t; x2 ^= t
    code = 0
    for b in range(bits):
        code |= ((x0 >> b) & 1) << (3 * b + 2) | ((x1 >> b) & 1) << (3 * b + 1) | ((x2 >> b) & 1) << (3 * b)
    return code


def test_hilbert_cell_order_is_a_face_adjacent_walk():
    """The property the sorted Chamfer search relies on: the cell index is a bijection onto 0..4095 and
    consecutive indices are face-adjacent cells, so 16 consecutive sorted points never straddle a jump of the curve."""
    cells = {}
    for x in range(16):
        for y in range(16):
            for z in range(16):
                cells[hilbert_of_cell(x, y, z)] = (x, y, z)
    assert sorted(cells) == list(range(4096))
    for h in range(4095):
        a, b = cells[h], cells[h + 1]
        assert sum(abs(p - q) for p, q in zip(a, b)) == 1, (h, a, b)


@pytest.mark.parametrize("n,CL", [(8192, 2), (3001, 1), (3001, 2), (5000, 4), (700, 4), (130, 4), (16384, 4)])
def test_cluster_counting_sort_slices_form_a_permutation(n, CL):
    """Index arithmetic of chamfer_sort_reg_kernel / chamfer_sort_kernel: CTA h of a cloud histograms the h-th slice of the
    points; the base of (slice, cell) = exclusive scan of the cloud's cell totals + the counts of the lower slices in that
    cell; position = base + rank inside (slice, cell).  Positions must be a permutation ordered by cell, and the chunk
    slices (multiples of 8 chunks) must tile the tile-padded chunk range, every position having exactly one owner."""
    rng = np.random.default_rng(n + CL)
    pts = rng.random((n, 3), dtype=np.float32)
    lo, hi = pts.min(0), pts.max(0)
    inv = np.where(hi > lo, 16.0 / (hi - lo), 0.0).astype(np.float32)
    cell = np.minimum(((pts - lo) * inv).astype(np.int64), 15)
    code = np.array([hilbert_of_cell(int(a), int(b), int(c)) for a, b, c in cell])
    pps = (((n + CL - 1) // CL) + 3) & ~3
    hist = np.zeros((CL, 4096), np.int64); rank = np.zeros(n, np.int64)
    for h in range(CL):
        i0, i1 = min(n, h * pps), min(n, h * pps + pps)
        for i in range(i0, i1):                                                        # (atomicAdd order: any order inside a slice)
            rank[i] = hist[h, code[i]]; hist[h, code[i]] += 1
    total = hist.sum(0)
    start = np.concatenate([[0], np.cumsum(total)[:-1]])
    lower = np.cumsum(hist, 0) - hist                                                  # counts of the lower slices
    pos = np.empty(n, np.int64)
    for h in range(CL):
        i0, i1 = min(n, h * pps), min(n, h * pps + pps)
        idx = np.arange(i0, i1)
        pos[idx] = start[code[idx]] + lower[h, code[idx]] + rank[idx]
    assert sorted(pos) == list(range(n))
    order = np.empty(n, np.int64); order[pos] = np.arange(n)
    assert (np.diff(code[order]) >= 0).all()                                           # sorted by cell
    same = np.diff(code[order]) == 0
    assert (np.diff(order)[same] > 0).all()          # inside a cell: slice order, then (in this sequential emulation) index order
    n_pad = (n + 127) // 128 * 128
    nch = n_pad // 16
    cps = (((nch + CL - 1) // CL) + 7) & ~7
    owners = np.zeros(nch, np.int64)
    for h in range(CL):
        c0 = min(nch, h * cps); c1 = min(nch, c0 + cps)
        assert (c1 - c0) % 8 == 0
        owners[c0:c1] += 1
    assert (owners == 1).all()
    span = cps * 16
    assert ((pos // span) < CL).all()                                                  # the DSMEM scatter's owner CTA exists
