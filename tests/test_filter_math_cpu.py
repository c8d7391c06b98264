"""The integer identities the tensor-core Chamfer epilogue relies on (softpool_b200/csrc/chamfer_tc.cu), checked on
the CPU with numpy: they are what makes the packed 16-bit filter conservative, independent of any GPU."""
import numpy as np


def test_positive_fp16_bit_patterns_order_like_their_values():
    bits = np.arange(0, 0x7C01, dtype=np.uint16)                 # +0 .. +inf
    vals = bits.view(np.float16).astype(np.float64)
    assert (np.diff(vals) > 0).all()


def test_packed_compare_keeps_bit15_exactly_when_c_le_t():
    """Per 16-bit lane, (0x8000 | t) - c keeps bit 15 iff c <= t for 15-bit c, t; two lanes in one 32-bit subtraction
    never borrow across the lane boundary."""
    rng = np.random.default_rng(1)
    t = rng.integers(0, 0x8000, 200000, dtype=np.uint32)
    c_lo = rng.integers(0, 0x8000, 200000, dtype=np.uint32)
    c_hi = rng.integers(0, 0x8000, 200000, dtype=np.uint32)
    # edge cases
    t[:4] = [0, 0x7FFF, 0x1234, 0x1234]; c_lo[:4] = [0, 0x7FFF, 0x1234, 0x1235]; c_hi[:4] = [0x7FFF, 0, 0x1233, 0x1234]
    T2 = ((t * np.uint32(0x10001)) | np.uint32(0x80008000)).astype(np.uint32)
    w = (c_lo | (c_hi << np.uint32(16))).astype(np.uint32)
    x = (T2 - w).astype(np.uint32)                               # wraps like the GPU's 32-bit subtraction
    assert (((x >> 15) & 1).astype(bool) == (c_lo <= t)).all()
    assert (((x >> 31) & 1).astype(bool) == (c_hi <= t)).all()


def test_mask_assembly_places_word_i_at_bits_i_and_16_plus_i():
    rng = np.random.default_rng(2)
    for _ in range(200):
        t = np.uint32(rng.integers(0, 0x8000))
        cm = rng.integers(0, 0x8000, 32).astype(np.uint32) | (rng.integers(0, 0x8000, 32).astype(np.uint32) << np.uint32(16))
        T2 = np.uint32((int(t) * 0x10001) | 0x80008000)
        m = [0, 0]
        for half in range(2):
            for i in range(16):
                x = np.uint32((int(T2) - int(cm[16 * half + i])) & 0xFFFFFFFF)
                m[half] |= (int(x) >> (15 - i)) & (0x10001 << i)
        mask = (m[1] << 32) | m[0]
        for pbit in range(64):
            w = ((pbit >> 5) << 4) | (pbit & 15); hi = (pbit >> 4) & 1           # the kernel's decode
            c = (int(cm[w]) >> (16 * hi)) & 0xFFFF
            assert bool((mask >> pbit) & 1) == (c <= int(t))


def test_b_row_permutation_makes_the_lanes_contiguous_chunks():
    """b_row_of: target u of an aligned group of 32 sits in accumulator column ((u & 15) << 1) | (u >> 4): a permutation,
    even columns (low 16-bit lane of a packed register) = targets 0..15, odd columns (high lane) = targets 16..31."""
    def b_row_of(r):
        return (r & ~31) | ((r & 15) << 1) | ((r >> 4) & 1)
    cols = [b_row_of(r) for r in range(96)]
    assert sorted(cols) == list(range(96))
    for r in range(96):
        c = b_row_of(r)
        assert c // 32 == r // 32
        lane_hi = c & 1                                                        # odd column -> high half of the register
        assert lane_hi == ((r & 31) >> 4)
        assert (c & 31) >> 1 == (r & 15)                                       # register index inside the 32-column load


def test_threshold_rounding_is_upwards():
    """The filter threshold is rounded UP to fp16 (cvt.rp), so an fp16 value <= the fp32 threshold never fails the compare."""
    rng = np.random.default_rng(3)
    thr = (rng.random(100000) * 12 + 0.0157).astype(np.float32)
    h = thr.astype(np.float16)                                                 # round to nearest
    up = np.where(h.astype(np.float32) < thr, np.nextafter(h, np.float16(np.inf)), h)   # emulate round-up
    assert (up.astype(np.float32) >= thr).all()
    c = (rng.random(100000) * 12 + 0.0157).astype(np.float16)
    passes_f32 = c.astype(np.float32) <= thr
    passes_bits = c.view(np.uint16) <= up.view(np.uint16)
    assert (passes_bits | ~passes_f32).all()                                   # never stricter than the fp32 compare
