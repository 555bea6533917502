"""Timing of the frame-level API (seam B3) on one GPU: FlameDetector.detect per call, the
element-wise frame functions, and ff_head_images over a packed range.  Prints one JSON line.

    python tools/bench_frame_api.py [--frames 64] [--height 128] [--width 1024]
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from high_speed_image_processing_b200 import process_videos as pv  # noqa: E402
from high_speed_image_processing_b200 import synthetic as syn  # noqa: E402
from high_speed_image_processing_b200.engine import get_engine  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--height", type=int, default=128)
    ap.add_argument("--width", type=int, default=1024)
    a = ap.parse_args()
    eng = get_engine(0)
    spec = syn.SyntheticSpec(width=a.width, height=a.height, n_frames=a.frames, bits=12, style="nova", t_enter=1.0,
                             velocity=a.width / (a.frames + 8.0), seed=11)
    frames = syn.render_frames(spec)
    bg = float(np.max(frames[0]))
    out = {"shape": [a.height, a.width], "frames": a.frames}

    for mode in ("host", "none"):
        det = pv.FlameDetector(pv.FlameDetectorConfig(use_spline_estimator=False), 160000.0, 0.000833333, engine=eng,
                               intermediates=mode, keep_results=False)
        det.detect(frames[0], 0, bg)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(1, a.frames):
            det.detect(frames[i], i, bg)
        torch.cuda.synchronize()
        out[f"detect_ms_per_call_intermediates_{mode}"] = (time.perf_counter() - t0) * 1e3 / (a.frames - 1)

    packed = torch.from_numpy(syn.pack_frames(frames, 12)).to(eng.device)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for want in (eng.HEAD_IMAGES, ("sobel_output", "gradient_output")):
        eng.head_images(packed, a.frames, a.height, a.width, 12, int(bg), want=want)
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(5):
            eng.head_images(packed, a.frames, a.height, a.width, 12, int(bg), want=want)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 5
        written = len(want) * 8 * a.frames * a.height * a.width
        out[f"head_images_{len(want)}_outputs"] = {"ms": ms, "frames_per_s": a.frames / ms * 1e3,
                                                   "output_gbs": written / ms / 1e6}

    dev = torch.from_numpy(frames[1]).to(eng.device)
    prior = torch.from_numpy(frames[0]).to(eng.device)
    for name, call in (("subtract_background", lambda: eng.frame_op("subtract_background", [dev], bg)),
                       ("difference", lambda: eng.frame_op("difference", [dev, prior], 5.0)),
                       ("count_above", lambda: eng.frame_count_above(dev, 50.0))):
        call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(200):
            call()
        torch.cuda.synchronize()
        out[f"{name}_us_per_call_device_resident"] = (time.perf_counter() - t0) * 1e6 / 200
    print(json.dumps(out))


if __name__ == "__main__":
    main()
