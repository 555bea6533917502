"""Executable model of the 16x2-lane integer arithmetic of the streaming kernels (csrc/ff_stream.cu), checked on the
CPU against the reference's per-pixel definitions - exhaustively where the domain allows it.

The kernels keep two pixels per 32-bit register and mix per-lane DPX operations (VIADDMNMX / VIMNMX: 16-bit lanes,
wrapping adds) with plain 32-bit adds and multiply-adds that run over BOTH lanes at once.  DESIGN.md argues that no
lane can carry or borrow into its neighbour and that the results equal the reference's float64 expressions
(scripts/process_videos.py:670-674 background subtraction, :397-399 thresholded frame difference, :759-763 count of
pixels above the noise threshold).  This file states that argument as code:

  * signed lanes (8- and 12-bit pixels): sub = relu(x - bg); carry Pn = 0x4000 - sub (one IMAD over both lanes);
    E = sub + Pn (one 32-bit add); r = relu(E + k); m = min(r, 1); out = r + m * (thr - 1) (one IMAD); a frame
    without a valid difference uses k = 0x8001 per lane and must produce zeros;
  * unsigned lanes (16-bit pixels): every subtraction is "max, then subtract";
  * the in-place compare of packed 12-bit fields (count12x8_simd): A lanes carry a junk nibble below the field;
  * lane -> float32 / float64 through the mantissa of 2^23 / 2^52.

Nothing here touches the GPU or the library; the GPU tests compare the kernels themselves with the oracle."""
import numpy as np
import pytest

M16 = 0xFFFF
M32 = 0xFFFFFFFF


# ---- the instruction set, on int64 arrays holding 32-bit register values -----------------------------------------
def lanes(w):
    return w & M16, (w >> 16) & M16


def pack(lo, hi):
    return (lo & M16) | ((hi & M16) << 16)


def s16(v):                      # 16-bit two's complement -> signed
    v = v & M16
    return np.where(v >= 0x8000, v - 0x10000, v)


def viaddmnmx_s16x2(a, b, c, *, use_max, relu):
    """Per signed 16-bit lane: max|min(wrap16(a + b), c), then max(., 0) with relu."""
    out = []
    for la, lb, lc in zip(lanes(a), lanes(b), lanes(c)):
        t = s16(la + lb)                                   # wrapping add
        t = np.maximum(t, s16(lc)) if use_max else np.minimum(t, s16(lc))
        if relu:
            t = np.maximum(t, 0)
        out.append(t & M16)
    return pack(*out)


def vimax3_u16x2(a, b, c):
    return pack(*[np.maximum(np.maximum(x, y), z) for x, y, z in zip(lanes(a), lanes(b), lanes(c))])


def vimin3_u16x2(a, b, c):
    return pack(*[np.minimum(np.minimum(x, y), z) for x, y, z in zip(lanes(a), lanes(b), lanes(c))])


def viaddmin_u16x2(a, b, c):
    return pack(*[np.minimum((x + y) & M16, z) for x, y, z in zip(lanes(a), lanes(b), lanes(c))])


def add32(a, b):
    return (a + b) & M32


def imad32(a, b, c):
    return (a * b + c) & M32


K_LANE_MIN2 = 0x80008000
ONE2 = 0x00010001


def dup(v):
    return (int(v) & M16) * ONE2


# ---- the reference's definitions ------------------------------------------------------------------------------------
def ref_diff(x, x_prev, bg, thr):
    sub = np.maximum(x - bg, 0)
    prev = np.maximum(x_prev - bg, 0)
    d = sub - prev
    return np.where(d >= thr, d, 0)


def signed_lane_diff(x2, prev_sub2_as_pn, bg, thr, valid=True):
    """streamx_kernel, signed lanes: returns (out word, new carry word)."""
    nbg2 = dup(-min(bg, 4095))
    tm1 = min(max(thr, 0) - 1, 8190)
    k2 = dup(-(0x4000 + tm1))
    kd2 = k2 if valid else 0x80018001
    sub2 = viaddmnmx_s16x2(x2, nbg2, K_LANE_MIN2, use_max=True, relu=True)
    e2 = add32(sub2, prev_sub2_as_pn)
    r2 = viaddmnmx_s16x2(e2, kd2, K_LANE_MIN2, use_max=True, relu=True)
    m2 = viaddmnmx_s16x2(r2, 0, ONE2, use_max=False, relu=True)          # VIMNMX.RELU: min(r, 1)
    out = imad32(m2, tm1 & M32, r2)
    pn = imad32(sub2, M32, 0x40004000)                                     # sub * neg_one + 0x4000...
    return out, pn


@pytest.mark.parametrize("thr", [0, 1, 5, 100, 4095, 4096, 9000])
@pytest.mark.parametrize("bg", [0, 59, 4095])
def test_signed_lane_difference_is_the_reference_difference_for_every_pixel_pair(thr, bg):
    rng = np.random.default_rng(thr * 7 + bg)
    if (thr, bg) in ((5, 59), (0, 0), (4096, 4095)):         # every (current, previous) 12-bit pixel value
        v = np.arange(4096, dtype=np.int64)
        x, xp = np.meshgrid(v, v, indexing="ij")
        x, xp = x.ravel(), xp.ravel()
    else:                                                    # a million random pairs + every pair of edge values
        edge = np.array([0, 1, bg, min(bg + 1, 4095), max(bg - 1, 0), min(bg + max(thr, 0), 4095), 4094, 4095])
        x = np.concatenate([rng.integers(0, 4096, 1_000_000), np.repeat(edge, edge.size)])
        xp = np.concatenate([rng.integers(0, 4096, 1_000_000), np.tile(edge, edge.size)])
    # the neighbouring lane holds something unrelated: a permutation of the same values
    perm = rng.permutation(x.size)
    x_hi, xp_hi = x[perm], xp[perm]
    sub_prev = pack(np.maximum(xp - bg, 0), np.maximum(xp_hi - bg, 0))
    pn = (0x40004000 - sub_prev) & M32
    out, pn_new = signed_lane_diff(pack(x, x_hi), pn, bg, thr)
    lo, hi = lanes(out)
    assert np.array_equal(lo, ref_diff(x, xp, bg, thr))
    assert np.array_equal(hi, ref_diff(x_hi, xp_hi, bg, thr))
    new_lo, new_hi = lanes(pn_new)
    assert np.array_equal(new_lo, 0x4000 - np.maximum(x - bg, 0)) and np.array_equal(new_hi, 0x4000 - np.maximum(x_hi - bg, 0))
    # a frame without a valid difference (first frame of the clip, skip_frames entry): zeros, carry still updated
    out0, pn0 = signed_lane_diff(pack(x, x_hi), pn, bg, thr, valid=False)
    assert not out0.any() and np.array_equal(pn0, pn_new)


def unsigned_lane_diff(x2, prev2, bg, thr):
    """streamx_kernel, unsigned lanes (16-bit pixels): returns (out word, new carry word = sub)."""
    bg2 = dup(min(bg, 65535))
    tm1 = min(max(thr - 1, 0), 65535)
    t2 = dup(tm1)
    sub2 = (vimax3_u16x2(x2, bg2, bg2) - bg2) & M32
    rd2 = (vimax3_u16x2(sub2, prev2, prev2) - prev2) & M32
    r2 = (vimax3_u16x2(rd2, t2, t2) - t2) & M32
    m2 = vimin3_u16x2(r2, ONE2, ONE2)
    return add32(r2, (m2 * tm1) & M32), sub2


@pytest.mark.parametrize("thr", [0, 1, 5, 40000, 65535, 65536])
@pytest.mark.parametrize("bg", [0, 59, 30000, 65535])
def test_unsigned_lane_difference_never_borrows_across_lanes(thr, bg):
    rng = np.random.default_rng(thr + bg)
    n = 500_000
    edge = np.array([0, 1, bg, min(bg + 1, 65535), max(bg - 1, 0), 65534, 65535], dtype=np.int64)
    x = np.concatenate([rng.integers(0, 65536, n), np.repeat(edge, edge.size)])
    xp = np.concatenate([rng.integers(0, 65536, n), np.tile(edge, edge.size)])
    perm = rng.permutation(x.size)
    x_hi, xp_hi = x[perm], xp[perm]
    prev2 = pack(np.maximum(xp - bg, 0), np.maximum(xp_hi - bg, 0))
    out, carry = unsigned_lane_diff(pack(x, x_hi), prev2, bg, thr)
    lo, hi = lanes(out)
    assert np.array_equal(lo, ref_diff(x, xp, bg, thr)) and np.array_equal(hi, ref_diff(x_hi, xp_hi, bg, thr))
    c_lo, c_hi = lanes(carry)
    assert np.array_equal(c_lo, np.maximum(x - bg, 0)) and np.array_equal(c_hi, np.maximum(x_hi - bg, 0))


# ---- counts: packed 12-bit fields compared in place ---------------------------------------------------------------
def count12_pair(b0, b1, b2, c):
    """One byte triple = two pixels (pyMRAW layout: hi = b0<<4 | b1>>4, lo = (b1&15)<<8 | b2) the way count12x8_simd
    looks at it: an A lane (b0<<8 | b1: the field with a junk nibble below it) and a B lane ((b1<<8 | b2) & 0xFFF)."""
    k_a = (c << 4) | 15
    a = pack(b0 << 8 | b1, 0)
    b = pack(((b1 << 8) | b2) & 0x0FFF, 0)
    fa = viaddmin_u16x2(vimax3_u16x2(a, dup(k_a), dup(k_a)), dup(0x10000 - k_a), ONE2)
    fb = viaddmnmx_s16x2(b, dup(-c), ONE2, use_max=False, relu=True)
    return (fa & M16) + (fb & M16)


@pytest.mark.parametrize("c", [0, 1, 69, 2047, 2048, 4094, 4095])
def test_in_place_compare_of_packed_12_bit_fields_counts_like_the_decoded_pixels(c):
    if c in (69, 2048):                                      # every pair of 12-bit pixels
        v = np.arange(4096, dtype=np.int64)
        hi, lo = np.meshgrid(v, v, indexing="ij")
        hi, lo = hi.ravel(), lo.ravel()
    else:                                                    # a million random pairs + every pair of values around c
        rng = np.random.default_rng(c)
        edge = np.array([0, 1, max(c - 1, 0), c, min(c + 1, 4095), 4094, 4095])
        hi = np.concatenate([rng.integers(0, 4096, 1_000_000), np.repeat(edge, edge.size)])
        lo = np.concatenate([rng.integers(0, 4096, 1_000_000), np.tile(edge, edge.size)])
    b0, b1, b2 = hi >> 4, ((hi & 15) << 4) | (lo >> 8), lo & 255
    got = count12_pair(b0, b1, b2, c)
    assert np.array_equal(got, (hi > c).astype(np.int64) + (lo > c).astype(np.int64))


def test_sixteen_bit_and_eight_bit_count_lanes():
    v = np.arange(65536, dtype=np.int64)
    for c in (0, 1, 255, 32767, 32768, 65534, 65535):       # unsigned compare: max, add, min
        got = viaddmin_u16x2(vimax3_u16x2(pack(v, v[::-1]), dup(c), dup(c)), dup(0x10000 - c), ONE2)
        lo, hi = lanes(got)
        assert np.array_equal(lo, (v > c).astype(np.int64)) and np.array_equal(hi, (v[::-1] > c).astype(np.int64))
    b = np.arange(256, dtype=np.int64)
    for c in (0, 1, 69, 254, 255):                           # 8-bit pixels widened to lanes, signed compare
        got = viaddmnmx_s16x2(pack(b, b[::-1]), dup(-c), ONE2, use_max=False, relu=True)
        lo, hi = lanes(got)
        assert np.array_equal(lo, (b > c).astype(np.int64)) and np.array_equal(hi, (b[::-1] > c).astype(np.int64))


def test_flag_sums_as_multiply_adds_are_plain_sums():
    """FF_COUNT_FMA_ADDS: acc = flag * one + acc with one == 1; the accumulator lanes never overflow 16 bits in an
    item (a thread adds at most 32 flags per lane and item before the lanes are folded)."""
    rng = np.random.default_rng(5)
    flags = rng.integers(0, 2, (32, 2, 100000))
    acc = np.zeros(100000, dtype=np.int64)
    for k in range(32):
        acc = imad32(pack(flags[k, 0], flags[k, 1]), 1, acc)
    lo, hi = lanes(acc)
    assert np.array_equal(lo, flags[:, 0].sum(0)) and np.array_equal(hi, flags[:, 1].sum(0))


# ---- lanes widened to float on the way out ---------------------------------------------------------------------------
def test_lane_to_float_and_double_are_exact_for_every_lane_value():
    v = np.arange(65536, dtype=np.uint32)
    f = (np.uint32(0x4B000000) | v).view(np.float32) - np.float32(8388608.0)       # PRMT + FADD
    assert f.dtype == np.float32 and np.array_equal(f, v.astype(np.float32))
    d = ((np.uint64(0x43300000) << np.uint64(32)) | v.astype(np.uint64)).view(np.float64) - 4503599627370496.0
    assert np.array_equal(d, v.astype(np.float64))


# ---- byte permutes: the selector constants of decode12x8_16x2 and count12x8_simd ------------------------------------
def byte_perm(a, b, sel):
    """__byte_perm / PRMT: result byte k = byte (sel >> 4k) & 7 of the 8-byte value {b, a} (a = bytes 0-3)."""
    both = a | (b << 32)
    out = 0
    for k in range(4):
        idx = (sel >> (4 * k)) & 7
        out = out | (((both >> (8 * idx)) & 0xFF) << (8 * k))
    return out


def test_decode_and_count_selectors_on_whole_12_byte_groups():
    """Eight packed 12-bit pixels = 12 bytes = three little-endian words, as the kernels read them from shared
    memory.  decode12x8_16x2 must deliver the eight pixels (pixel 2j in the low lane of word j), count12x8_simd the
    number of pixels above c; the pixel values come from the oracle's own unpack."""
    from oracle import flame_oracle as fo
    rng = np.random.default_rng(11)
    n = 200_000
    raw = rng.integers(0, 256, (n, 12), dtype=np.uint8)
    px = fo.unpack12(raw.reshape(-1)).reshape(n, 8).astype(np.int64)
    w = raw.view("<u4").astype(np.int64)                     # [n, 3]
    w0, w1, w2 = w[:, 0], w[:, 1], w[:, 2]
    # decode12x8_16x2 (ff_common.cuh)
    p = [byte_perm(w0, 0, 0x1201), byte_perm(w0, w1, 0x4534), byte_perm(w1, w2, 0x3423), byte_perm(w2, 0, 0x2312)]
    for j in range(4):
        x = ((p[j] >> 4) & 0x00000FFF) | (p[j] & 0x0FFF0000)
        lo, hi = lanes(x)
        assert np.array_equal(lo, px[:, 2 * j]) and np.array_equal(hi, px[:, 2 * j + 1]), j
    # count12x8_simd (ff_stream.cu)
    for c in (0, 69, 2048, 4095):
        k_a = (c << 4) | 15
        a0, a1 = byte_perm(w0, w1, 0x3401), byte_perm(w1, w2, 0x5623)
        b0, b1 = byte_perm(w0, w1, 0x4512) & 0x0FFF0FFF, byte_perm(w1, w2, 0x6734) & 0x0FFF0FFF
        total = np.zeros(n, dtype=np.int64)
        for a in (a0, a1):
            total = add32(total, viaddmin_u16x2(vimax3_u16x2(a, dup(k_a), dup(k_a)), dup(0x10000 - k_a), ONE2))
        for b in (b0, b1):
            total = add32(total, viaddmnmx_s16x2(b, dup(-c), ONE2, use_max=False, relu=True))
        lo, hi = lanes(total)
        # one call covers two of the eight pixels per lane pair: four words x two lanes = the eight pixels
        assert np.array_equal(lo + hi, (px > c).sum(1))
