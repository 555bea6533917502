"""Executable model of the range kernel's work decomposition (csrc/ff_stream.cu: launch_range_fused + the consumer /
producer loops of range_kernel), checked on the CPU for thousands of clip shapes - including the small-frame regime
(one to four 12-KiB items per frame) that BASELINE config 1 runs in.

The kernel never divides inside its loops: the items (frame-major: item i = tile i % T of frame i // T) are cut into
slabs of 2^shift items, the slabs are dealt to the CTAs round-robin, and every CTA reaches its next slab by STEPPING
(frame, tile) by a precomputed (slab_step_frames, slab_step_tiles).  A segment - the part of a slab inside one frame -
ends with the frame's last tile or with the slab; the detector side expects exactly
((f + 1) T - 1 >> shift) - (f T >> shift) + 1 segments per frame before it decides the frame.  If the stepping or the
expected count were off by one for some shape, a frame would never be decided (or decided early on partial counts).

What is modelled is the index arithmetic only; the GPU tests run the kernel itself against the oracle."""
import numpy as np
import pytest


def launch_plan(tiles_per_frame: int, n_frames: int, wave: int):
    """launch_range_fused: slab size (power of two, at most 32, every CTA >= 64 slabs when the clip allows), grid."""
    total = tiles_per_frame * n_frames
    slab = min(total // (wave * 64), 32)
    slab = max(slab, 1)
    shift = 0
    while (2 << shift) <= slab:
        shift += 1
    n_slabs = (total + (1 << shift) - 1) >> shift
    grid = min(n_slabs, wave)
    return shift, n_slabs, grid, ((grid << shift) // tiles_per_frame, (grid << shift) % tiles_per_frame)


def walk_cta(b: int, T: int, n_frames: int, shift: int, n_slabs: int, grid: int, step):
    """The consumer loop of one CTA: yields (item index, frame, tile, segment_ends) in the order the kernel visits."""
    total = T * n_frames
    S = 1 << shift
    first = b << shift
    fs, ts = first // T, first % T                      # the only division: once per CTA
    slab = b
    while slab < n_slabs:
        i0 = slab << shift
        n_it = min(i0 + S, total) - i0
        f, tile = fs, ts
        fs += step[0]
        ts += step[1]
        if ts >= T:
            ts -= T
            fs += 1
        for k in range(n_it):
            last_of_frame = tile + 1 == T
            yield i0 + k, f, tile, last_of_frame or k + 1 == n_it
            if last_of_frame:
                tile = 0
                f += 1
            else:
                tile += 1
        slab += grid


def check_shape(T: int, n_frames: int, wave: int):
    shift, n_slabs, grid, step = launch_plan(T, n_frames, wave)
    assert 1 <= grid <= n_slabs and step[1] < T
    seen = np.zeros(T * n_frames, dtype=np.int32)
    segments = np.zeros(n_frames, dtype=np.int64)
    for b in range(grid):
        for i, f, tile, seg_end in walk_cta(b, T, n_frames, shift, n_slabs, grid, step):
            assert (f, tile) == (i // T, i % T), (T, n_frames, wave, b, i)
            seen[i] += 1
            segments[f] += seg_end
    assert (seen == 1).all(), "every item exactly once"
    f = np.arange(n_frames, dtype=np.int64)
    expected = (((f + 1) * T - 1) >> shift) - ((f * T) >> shift) + 1      # what the detector warps wait for
    assert np.array_equal(segments, expected), (T, n_frames, wave, shift)


@pytest.mark.parametrize("wave", [296, 148, 7, 1])
def test_stepping_and_segment_counts_for_random_clip_shapes(wave):
    rng = np.random.default_rng(wave)
    for _ in range(120):
        T = int(rng.choice([1, 2, 3, 4, 5, 7, 16, 17, 32, 33, 128]))
        n_frames = int(rng.integers(1, 4000 if T < 8 else 700))
        check_shape(T, n_frames, wave)


@pytest.mark.parametrize("T,n_frames", [(16, 20000), (32, 20000), (4, 500), (1, 7), (1, 1), (3, 1), (128, 300), (17, 90)])
def test_baseline_and_edge_shapes(T, n_frames):
    """C2 (16 items per frame), C3 (32), C1 (4), frames smaller than an item (T = 1), single-frame clips, C4-sized
    frames, the ragged 130-row frame of the GPU tests (16.25 tiles -> 17)."""
    check_shape(T, n_frames, 296)
