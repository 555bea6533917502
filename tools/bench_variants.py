#!/usr/bin/env python
"""Per-kernel timings of the other BASELINE configurations (C1, C3, C4) and of every
streaming-kernel variant - these are parity-test cases, not bench.py's headline, but
DESIGN.md quotes their roofline fractions.  Prints one JSON object per line.

    python tools/bench_variants.py [--frames-scale 1.0] [--reps 5]
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

import torch  # noqa: E402

from high_speed_image_processing_b200 import synthetic as syn  # noqa: E402
from high_speed_image_processing_b200.engine import DetectionParams, FlameFrontEngine  # noqa: E402


def time_call(fn, reps: int) -> float:
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1))
    best.sort()
    return best[len(best) // 2]


def bench_head(eng, name, spec, packed, n, reps) -> None:
    """HEAD-parity detector: streaming kernel (counts) + band kernel (difference -> 3x3 opening
    -> Gaussian -> Sobel/gradient centre rows, float64) + sequential tracker."""
    from high_speed_image_processing_b200.head import HeadParams
    h, w = spec.height, spec.width
    hp = HeadParams()
    out = {}

    def run():
        out["res"] = eng.process_head(packed, n, h, w, 12, hp, float(spec.record_rate), 0.000833333)
    ms = time_call(run, reps)
    # back to back, as when one clip follows another: the host's launch work for a clip overlaps the
    # previous clip's streaming kernel instead of preceding an idle GPU
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4 * reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms_pipelined = e0.elapsed_time(e1) / (4 * reps)
    res = out.pop("res")
    flags = res.flags.cpu().numpy()
    track = res.track.cpu().numpy()
    stop = res.stop.cpu().numpy()
    print(json.dumps({
        "config": name, "frames": n, "shape": [h, w], "method": "head (FlameDetector parity)",
        "whole_path_ms": ms, "whole_path_frames_per_s": n / ms * 1e3,
        "whole_path_gbs": n * (spec.frame_bytes + 8) / ms / 1e6,
        "back_to_back_ms_per_clip": ms_pipelined, "back_to_back_frames_per_s": n / ms_pipelined * 1e3,
        "frames_through_band_kernel": int((flags == 1).sum()), "detections": int((track[:, 0] >= 0).sum()),
        "exit_frame": int(stop[0])}), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames-scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="", help="comma list of config:diff pairs, e.g. C4:uint16,C2:None")
    args = ap.parse_args()
    eng = FlameFrontEngine(0)
    peak = 6547.5
    pk = REPO / "MEASURED_PEAKS.json"
    if pk.exists():
        peak = float(json.loads(pk.read_text())["hbm_gbs"])

    cases = [
        ("C1", "half_maximum", None, False),
        ("C2", "half_maximum", None, False),
        ("C3", "threshold", None, False),
        ("C4", "gradient", "uint16", False),
        ("C4", "gradient", "float32", False),
        ("C4", "gradient", "float64", False),
        ("C2", "half_maximum", None, True),       # also materialise decoded uint16 frames
        ("C2/16", "half_maximum", None, False),    # same clip stored as 16-bit little-endian MRAW
        ("C2/8", "half_maximum", None, False),     # ... and as 8-bit
        ("C4/16", "gradient", "uint16", False),
        ("C4/16", "gradient", "float32", False),
        ("C4/8", "gradient", "uint16", False),
        ("C4/8", "gradient", "float32", False),
        ("C3", "unpack", None, False),            # stage 1 alone: ff_unpack, packed 12-bit -> uint16
        ("C2", "head", None, False),              # the detector the reference runs at HEAD
        ("C4", "head", None, False),
    ]
    if args.only:
        want = {tuple(x.split(":")) for x in args.only.split(",")}
        cases = [c for c in cases if (((c[0], str(c[2])) in want or (c[0], c[1]) in want) and not c[3])
                 or (c[3] and (c[0], "decoded") in want)]
    cache = {}
    for name, method, diff, decoded in cases:
        cfg_name, _, bits_s = name.partition("/")
        bits = int(bits_s) if bits_s else 12
        base = syn.config_spec(cfg_name)
        if bits != 12:
            base = syn.SyntheticSpec(**{**base.__dict__, "bits": bits})
        n = max(8, int(base.n_frames * args.frames_scale))
        if diff == "float64":
            n = min(n, 2500)                       # 8 B/px x 1 Mpx x 2500 = 21 GB retained
        spec = base
        if n != base.n_frames:
            spec = syn.config_spec(cfg_name, n_frames=n)
            if bits != 12:
                spec = syn.SyntheticSpec(**{**spec.__dict__, "bits": bits})
        key = (name, n)
        if key not in cache:
            cache.clear()
            cache[key] = syn.render_packed_torch(spec, eng.device)
        packed = cache[key]
        h, w, fb = spec.height, spec.width, spec.frame_bytes
        params = DetectionParams(method=method if method not in ("head", "unpack") else "gradient")
        scalars, bg_dev = eng.clip_scalars(packed[:fb], h, w, bits)

        if method == "head":
            bench_head(eng, name, spec, packed, n, args.reps)
            torch.cuda.empty_cache()
            continue
        if method == "unpack":
            keep = {}

            def run_unpack():
                keep["out"] = eng.unpack(packed, n, h, w, 12)
            ms = time_call(run_unpack, args.reps)
            alg_u = n * (fb + 2 * h * w)
            print(json.dumps({"config": name, "frames": n, "shape": [h, w], "method": "ff_unpack (decode only)",
                              "algorithmic_bytes_per_frame": fb + 2 * h * w, "kernel_ms": ms,
                              "kernel_gbs": alg_u / ms / 1e6, "frac_of_measured_peak": alg_u / ms / 1e6 / peak}),
                  flush=True)
            keep.clear()
            torch.cuda.empty_cache()
            continue

        eng._stream_events = []
        eng._stream_events_every = 1
        out = {}

        def run():
            out["res"] = eng.process_range(packed, n, h, w, bits, params, scalars, bg_dev, diff_dtype=diff,
                                           keep_decoded=decoded)
        whole_ms = time_call(run, args.reps)
        ev = eng._stream_events
        torch.cuda.synchronize()
        stream_ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev[2:])
        stream_ms = stream_ms[len(stream_ms) // 2]
        eng._stream_events = None
        px = h * w
        out_bytes = {None: 0, "uint16": 2, "float32": 4, "float64": 8}[diff] * px + (2 * px if decoded else 0)
        alg = n * (fb + out_bytes + 8)
        res = out.pop("res")
        fe = int(res.first_exit.cpu().item())
        ndet = int((res.pos >= 0).sum().item())
        del res
        print(json.dumps({
            "config": name, "frames": n, "shape": [h, w], "method": method, "diff_dtype": diff, "decoded_out": decoded,
            "algorithmic_bytes_per_frame": fb + out_bytes + 8,
            "stream_kernel_ms": stream_ms, "stream_kernel_gbs": alg / stream_ms / 1e6,
            "stream_kernel_frac_of_measured_peak": alg / stream_ms / 1e6 / peak,
            "whole_path_ms": whole_ms, "whole_path_frames_per_s": n / whole_ms * 1e3,
            "whole_path_gbs": alg / whole_ms / 1e6, "first_exit": fe, "detections": ndet}), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
