"""The chained-speculation tracker (csrc/ff_head.cu: head_track_spec / fixup / resolve kernels +
commit_walk) as an executable model, checked against the plain sequential walk on thousands of
random clips.  The CUDA kernels themselves are compared with the sequential kernel on the GPU
(tests/test_gpu_head.py); this model checks the ALGORITHM - the induction over segment guesses, the
give-up rule, exit handling, the fallback - on far more state patterns than a few clips can reach:
carried-in states, frames without detections, several fronts, exits anywhere, guesses that miss.

The per-frame search is abstracted: every active frame holds a few candidate positions with
strengths; a search over the window [s0, s1) returns the strongest candidate inside it (first on
ties) - like the real search, a deterministic function of (frame data, state) only.
"""
from dataclasses import dataclass
from typing import List, Optional, Tuple

import pytest
from hypothesis import given, settings, strategies as st

SEG = 4            # frames per segment (32 on the GPU)
MAX_REPAIR = 2     # kMaxRepair (16 on the GPU)
W, MARGIN, MAX_DISP, WINDOW, EXIT_MARGIN = 64, 2, 3, 6, 4
NONE = (-1, -1)
STATS = {"repaired": 0, "met": 0, "exit_in_repair": 0, "gave_up": 0, "all_rerun": 0, "fallback": 0,
         "exit": 0, "skipped_exit_guess": 0}      # which branches the clips reached


@dataclass
class Frame:
    flag: int                                   # 0 inactive, 1 active, 2 active without a difference image
    cands: List[Tuple[int, int]]                # (position, strength)


def track_frame(fr: Frame, f: int, state):
    """(final, s0, s1) for one frame from a tracker state - ff_head.cu track_frame, abstracted."""
    last_f, last_p = state
    if last_p < 0:
        s0, s1 = MARGIN, W - MARGIN
    else:
        s0, s1 = last_p, min(W - MARGIN, last_p + MAX_DISP * max(1, f - last_f) + WINDOW)
    final = -1
    if fr.flag == 1 and s1 > s0:
        best = None
        for pos, strength in fr.cands:
            if s0 <= pos < s1 and (best is None or strength > best[1]):
                best = (pos, strength)
        if best is not None:
            final = best[0]
    return final, s0, s1


def sequential(frames, init):
    """head_track_generic_kernel: the reference loop."""
    out = [None] * len(frames)
    state, exit_f = init, None
    for f, fr in enumerate(frames):
        if fr.flag == 0:
            continue
        final, s0, s1 = track_frame(fr, f, state)
        out[f] = (final, s0, s1)
        if final >= 0:
            state = (f, final)
            if final >= W - EXIT_MARGIN:
                exit_f = f
                break
    return out, exit_f, state


def commit_walk(frames, out, seg_begin, cur):
    """One warp validating the speculative rows segment by segment (ff_head.cu commit_walk)."""
    n_seg = -(-len(frames) // SEG)
    for s in range(seg_begin, n_seg):
        lo, hi = s * SEG, min(len(frames), (s + 1) * SEG)
        act = [f for f in range(lo, hi) if frames[f].flag]
        spec = NONE                              # (callers never pass segment 0 with a carried-in state here)
        spec_final = {f: out[f][0] for f in act}
        k = 0
        while k < len(act) and spec != cur:
            f = act[k]
            k += 1
            final, s0, s1 = track_frame(frames[f], f, cur)
            if spec_final[f] >= 0:
                spec = (f, spec_final[f])
            if final >= 0:
                cur = (f, final)
            out[f] = (final, s0, s1)
            if final >= W - EXIT_MARGIN and final >= 0:
                return f, cur
        for f in act[k:]:                        # the rest stands as speculated
            if spec_final[f] >= 0:
                cur = (f, spec_final[f])
                if spec_final[f] >= W - EXIT_MARGIN:
                    return f, cur
    return None, cur


def chained(frames, init):
    n = len(frames)
    n_seg = -(-n // SEG)
    out = [None] * n
    # ---- full-width answers + speculative walks (head_track_fullwidth / spec kernels) ----
    E1 = []
    for s in range(n_seg):
        state = init if s == 0 else NONE
        for f in range(s * SEG, min(n, (s + 1) * SEG)):
            if frames[f].flag == 0:
                continue
            final, s0, s1 = track_frame(frames[f], f, state)
            out[f] = (final, s0, s1)
            if final >= 0:
                state = (f, final)
        E1.append(state)
    with_state = [s for s in range(n_seg) if E1[s][1] >= 0]
    first_with_state = with_state[0] if with_state else 10 ** 9
    # ---- fix-up (head_track_fixup_kernel) ----
    rec = []
    for s in range(n_seg):
        lo, hi = s * SEG, min(n, (s + 1) * SEG)
        act = [f for f in range(lo, hi) if frames[f].flag]
        r = {"E1": E1[s], "E2": None, "n_fixed": 0, "exit": None, "G": None, "saved": {}}
        if act:
            cur = NONE
            for t in range(s - 1, first_with_state - 1, -1):
                if E1[t][1] >= 0:
                    cur = E1[t]
                    break
            finals = {f: out[f][0] for f in act}
            if s != 0 and cur[1] >= W - EXIT_MARGIN and cur[0] >= 0:
                STATS["skipped_exit_guess"] += 1
            if s != 0 and cur[1] >= 0 and (cur[1] < W - EXIT_MARGIN or cur[0] < 0):
                guess, spec, k, exit_hit = cur, NONE, 0, False
                spec_final = dict(finals)
                while k < len(act) and r["n_fixed"] < MAX_REPAIR and spec != cur:
                    f = act[k]
                    k += 1
                    final, s0, s1 = track_frame(frames[f], f, cur)
                    if spec_final[f] >= 0:
                        spec = (f, spec_final[f])
                    if final >= 0:
                        cur = (f, final)
                    r["saved"][f] = out[f]
                    out[f] = (final, s0, s1)
                    finals[f] = final
                    r["n_fixed"] += 1
                    if final >= 0 and final >= W - EXIT_MARGIN:
                        exit_hit = True
                        break
                met = spec == cur
                own = E1[s][1] >= 0
                STATS["repaired"] += 1
                if met or exit_hit:
                    STATS["exit_in_repair" if exit_hit else "met"] += 1
                    r["E2"] = E1[s] if own else guess
                elif k == len(act):
                    STATS["all_rerun"] += 1
                    r["E2"] = cur
                else:
                    STATS["gave_up"] += 1
                    r["E2"] = (-2, -2)
                r["G"] = guess
            ex = [f for f in act if finals[f] >= 0 and finals[f] >= W - EXIT_MARGIN]
            r["exit"] = ex[0] if ex else None
        rec.append(r)
    # ---- resolve (head_track_resolve_kernel) ----
    u = n_seg
    for s, r in enumerate(rec):
        if r["n_fixed"] > 0:
            target = r["E1"] if r["E1"][1] >= 0 else r["G"]
            if r["E2"] != target:
                u = s
                break
    exits = [r["exit"] for r in rec[:u] if r["exit"] is not None]
    exit_f = min(exits) if exits else None
    cur = init
    if exit_f is not None:
        cur = (exit_f, out[exit_f][0])
    elif u >= n_seg:
        for s in range(n_seg - 1, -1, -1):
            end = rec[s]["E2"] if rec[s]["n_fixed"] > 0 else rec[s]["E1"]
            if end[1] >= 0:
                cur = end
                break
    else:
        STATS["fallback"] += 1
        for r in rec[u:]:                        # the speculation's rows back, then one warp validates
            for f, row in r["saved"].items():
                out[f] = row
        exit_f, cur = commit_walk(frames, out, u, rec[u]["G"])
    if exit_f is not None:
        STATS["exit"] += 1
        for f in range(exit_f + 1, n):
            out[f] = None
    return out, exit_f, cur


frame_st = st.builds(
    Frame,
    flag=st.sampled_from([0, 0, 1, 1, 1, 1, 2]),
    cands=st.lists(st.tuples(st.integers(0, W - 1), st.integers(1, 5)), min_size=0, max_size=3))


def front_clip(draw):
    """A front that moves right with noise candidates - the typical, mostly-converging case."""
    n = draw(st.integers(1, 40))
    x = draw(st.integers(0, 20))
    frames = []
    for _ in range(n):
        x += draw(st.integers(0, 3))
        flag = draw(st.sampled_from([0, 1, 1, 1, 1, 2]))
        cands = [(min(x, W - 1), 4)] + draw(st.lists(st.tuples(st.integers(0, W - 1), st.integers(1, 6)), max_size=2))
        frames.append(Frame(flag, cands))
    return frames


@settings(max_examples=1500, deadline=None)
@given(st.one_of(st.lists(frame_st, min_size=1, max_size=40), st.composite(front_clip)()),
       st.sampled_from([NONE, NONE, (-4, 5), (-1, 30), (-2, W - 3)]))      # carried-in: detected before the range
def test_chained_speculation_equals_sequential_walk(frames, init):
    want_out, want_exit, want_state = sequential(frames, init)
    got_out, got_exit, got_state = chained(frames, init)
    assert got_exit == want_exit
    assert got_out == want_out
    assert got_state == want_state


def test_model_exercises_every_path():
    """Hand-made clips that take the give-up, missed-guess fallback and exit-in-repair branches."""
    # a guess that misses: a second, stronger front appears inside the window of the true state only
    frames = [Frame(1, [(10, 3)]), Frame(1, [(11, 3)]), Frame(1, [(12, 3)]), Frame(1, [(13, 3)]),
              Frame(1, [(14, 3), (40, 5)]), Frame(1, [(15, 3), (41, 5)]), Frame(1, [(16, 3), (42, 5)]),
              Frame(1, [(17, 3), (43, 5)]), Frame(1, [(18, 3), (44, 5)]), Frame(1, [(19, 3), (61, 5)]),
              Frame(1, [(20, 3)]), Frame(1, [(21, 3)])]
    for init in (NONE, (-1, 9)):
        assert chained(frames, init) == sequential(frames, init)
    # exit on a repaired frame
    frames = [Frame(1, [(50, 3)])] * 4 + [Frame(1, [(61, 3), (5, 9)])] + [Frame(1, [(7, 9)])] * 5
    assert chained(frames, NONE) == sequential(frames, NONE)


def test_random_clips_reach_every_branch_of_the_model():
    """The random comparison is only worth something if it reaches the rare branches: count them."""
    import random
    rng = random.Random(5)
    for k in STATS:
        STATS[k] = 0
    for _ in range(3000):
        frames, x = [], rng.randint(0, 20)
        drift = rng.random() < 0.5
        for _ in range(rng.randint(1, 48)):
            x += rng.randint(0, 3)
            cands = [(rng.randint(0, W - 1), rng.randint(1, 6)) for _ in range(rng.randint(0, 3))]
            if drift:
                cands.insert(0, (min(x, W - 1), 4))
            frames.append(Frame(rng.choice([0, 1, 1, 1, 1, 2]), cands))
        init = rng.choice([NONE, NONE, (-4, 5), (-1, 30), (-2, W - 3)])
        assert chained(frames, init) == sequential(frames, init)
    assert all(v > 20 for v in STATS.values()), STATS
