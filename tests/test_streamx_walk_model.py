"""Executable model of streamx_kernel's walk over its items (csrc/ff_stream.cu: launch_streamx + the item loop),
checked on the CPU: the three orders in which the kernel visits (frame, tile) items, and the hand strength-reduced
running state of the loop (ring slot and phase, pixel offset of the output, index of the partial counts, source
address of the producer) against the closed forms they replaced.

  * retained difference image, long clips: frame-synchronous UNITS - unit u = frames [seg * len, ...) of tile
    u % tiles, dealt to the CTAs round-robin; every unit starts with a halo item (the frame before it) unless it
    starts the clip and the caller gave no halo frame;
  * retained difference image, short clips: tile-major contiguous runs per CTA, a run that crosses a tile boundary is
    processed as two segments;
  * no difference image (decoded pixels, counts): frame-major contiguous runs, tiles of a frame are contiguous in
    memory and the last one may be ragged.

Every (frame, tile) must be produced exactly once, halo items must carry the previous frame of the SAME tile, and
the running values must equal f * px_per_frame + t * 8192 etc. at every item.  Index arithmetic only; the GPU tests
run the kernel itself against the oracle."""
import numpy as np
import pytest

TILE_GROUPS = 1024                    # 8-pixel groups per tile (4 * kThreads)
GROUP_PX = 8
WARPS = 8


def plan(tiles: int, n_frames: int, wave: int, diff: bool):
    """launch_streamx: items per CTA / grid, or the unit length when the frame-synchronous order applies."""
    total = tiles * n_frames
    items_per_cta = max(1, (total + wave - 1) // wave)
    grid = (total + items_per_cta - 1) // items_per_cta
    unit_frames, n_units = 0, 0
    if diff and n_frames >= 96:
        best_nseg, best_cost = 1, 1e30
        nseg = (n_frames + 511) // 512
        while nseg <= n_frames // 48:
            length = (n_frames + nseg - 1) // nseg
            units = tiles * ((n_frames + length - 1) // length)
            rounds = (units + wave - 1) // wave
            cost = (rounds * wave) / units * (1.0 + 1.0 / length)
            if cost < best_cost - 1e-9:
                best_cost, best_nseg = cost, nseg
            if cost < 1.012:
                break
            nseg += 1
        unit_frames = (n_frames + best_nseg - 1) // best_nseg
        n_units = tiles * ((n_frames + unit_frames - 1) // unit_frames)
        grid = min(n_units, wave)
    return items_per_cta, grid, unit_frames, n_units


def walk(cta: int, tiles: int, n_frames: int, px_per_frame: int, bits: int, stages: int, diff: bool, have_halo_frame: bool,
         items_per_cta: int, grid: int, unit_frames: int, n_units: int):
    """The consumer's view of one CTA: yields dicts of the running state at every item, exactly as the kernel
    updates it (no divisions or 64-bit products inside the item loop)."""
    tile_bytes = TILE_GROUPS * bits
    frame_bytes = px_per_frame * bits // 8
    groups_per_frame = px_per_frame // GROUP_PX
    total = tiles * n_frames
    work = cta * items_per_cta
    work_end = min(work + items_per_cta, total)
    units = diff and unit_frames > 0
    unit = cta
    s, ph, n_item = 0, 0, 0
    last_tile = tiles - 1
    last_groups = groups_per_frame - last_tile * TILE_GROUPS
    if not units and work >= work_end:
        return
    while (unit < n_units) if units else (work < work_end):
        has_halo = 0
        if diff:
            if units:
                seg, tile = unit // tiles, unit % tiles
                f0 = seg * unit_frames
                n_seg = min(unit_frames, n_frames - f0)
                unit += grid
            else:
                tile, f0 = work // n_frames, work % n_frames
                n_seg = min(n_frames - f0, work_end - work)
            has_halo = 1 if (f0 - 1 >= 0 or have_halo_frame) else 0
        else:
            f0, tile = work // tiles, work % tiles
            n_seg = work_end - work
        work += n_seg
        n_items = n_seg + has_halo
        f = f0 - has_halo if diff else f0
        t = tile
        px0 = f * px_per_frame + t * TILE_GROUPS * GROUP_PX
        pidx = (f * tiles + t) * WARPS
        src = f * frame_bytes + t * tile_bytes
        for it in range(n_items):
            is_halo = it < has_halo
            tile_groups = last_groups if t == last_tile else TILE_GROUPS
            yield dict(f=f, t=t, is_halo=is_halo, px0=px0, pidx=pidx, src=src, s=s, ph=ph, n=n_item,
                       bytes=tile_groups * bits, first_of_clip_halo=is_halo and f < 0)
            if diff:
                f += 1
                px0 += px_per_frame
                pidx += tiles * WARPS
                src += frame_bytes
            else:
                px0 += tile_groups * GROUP_PX
                pidx += WARPS
                src += tile_groups * bits
                t += 1
                if t == tiles:
                    t = 0
                    f += 1
            s += 1
            if s == stages:
                s = 0
                ph ^= 1
            n_item += 1


def check(h: int, w: int, n_frames: int, bits: int, stages: int, wave: int, diff: bool, have_halo_frame: bool):
    px = h * w
    assert px % 32 == 0
    tiles = (px // GROUP_PX + TILE_GROUPS - 1) // TILE_GROUPS
    items_per_cta, grid, unit_frames, n_units = plan(tiles, n_frames, wave, diff)
    frame_bytes = px * bits // 8
    seen = np.zeros((n_frames, tiles), dtype=np.int32)
    for cta in range(grid):
        prev = None
        for st in walk(cta, tiles, n_frames, px, bits, stages, diff, have_halo_frame, items_per_cta, grid, unit_frames, n_units):
            f, t = st["f"], st["t"]
            assert st["s"] == st["n"] % stages and st["ph"] == (st["n"] // stages) & 1          # ring slot = item number mod stages
            if st["is_halo"]:
                assert diff and 0 <= t < tiles and f >= -1
                if f >= 0:                                       # halo = frame f of the same tile, inside the range
                    assert st["src"] == f * frame_bytes + t * TILE_GROUPS * bits
                else:                                            # the caller's halo frame stands in for frame -1
                    assert have_halo_frame
            else:
                assert 0 <= f < n_frames and 0 <= t < tiles, (f, t)
                seen[f, t] += 1
                assert st["px0"] == f * px + t * TILE_GROUPS * GROUP_PX
                assert st["pidx"] == (f * tiles + t) * WARPS
                assert st["src"] == f * frame_bytes + t * TILE_GROUPS * bits
                groups = min(TILE_GROUPS, px // GROUP_PX - t * TILE_GROUPS)
                assert st["bytes"] == groups * bits and st["src"] + st["bytes"] <= (f + 1) * frame_bytes
                if diff:
                    # the carry in registers must be frame f - 1 of this tile: the item right before (a halo item or a
                    # real one) - or there is no previous frame at all (clip start without a caller's halo frame)
                    carried = prev is not None and prev["t"] == t and prev["f"] == f - 1
                    assert carried or (f == 0 and not have_halo_frame), (f, t, prev)
            prev = st
    assert (seen == 1).all(), "every (frame, tile) exactly once"


@pytest.mark.parametrize("diff", [True, False])
@pytest.mark.parametrize("have_halo_frame", [False, True])
def test_walk_orders_and_running_state(diff, have_halo_frame):
    rng = np.random.default_rng(3 + diff * 2 + have_halo_frame)
    shapes = [(128, 1024), (130, 1024), (1024, 1024), (256, 1024), (256, 512), (129, 1024)]
    for _ in range(40):
        h, w = shapes[int(rng.integers(len(shapes)))]
        n = int(rng.choice([1, 2, 5, 23, 95, 96, 130, 400, 1000]))
        if h * w >= 1 << 20:
            n = min(n, 130)
        bits = int(rng.choice([8, 12, 16]))
        stages = int(rng.choice([2, 3, 4, 5, 6]))
        wave = int(rng.choice([148, 296, 5, 1]))
        check(h, w, n, bits, stages, wave, diff, have_halo_frame and diff)


def test_config4_shape_units_are_balanced():
    """C4 (1024 x 1024 x 5000, 128 tiles per frame, 148 CTAs): the unit length keeps the halo overhead below 2.5 % and
    every CTA within one unit of the others."""
    tiles, n_frames, wave = 128, 5000, 148
    _, grid, unit_frames, n_units = plan(tiles, n_frames, wave, True)
    assert grid == wave and 48 <= unit_frames <= 512
    per_cta = [len(range(c, n_units, grid)) for c in range(grid)]
    assert max(per_cta) - min(per_cta) <= 1
    assert 1.0 / unit_frames < 0.025
