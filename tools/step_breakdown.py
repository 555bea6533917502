#!/usr/bin/env python
"""Where the C2 step's time goes outside the streaming kernel: times the device-resident step
with parts of it removed (CUDA events over 30 back-to-back steps, clip resident in HBM)."""
from __future__ import annotations

import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

import torch  # noqa: E402

from high_speed_image_processing_b200 import synthetic as syn  # noqa: E402
from high_speed_image_processing_b200.engine import DetectionParams, FlameFrontEngine  # noqa: E402


def timed(fn, steps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main() -> None:
    eng = FlameFrontEngine(0)
    spec = syn.config_spec("C2")
    h, w, fb, n = spec.height, spec.width, spec.frame_bytes, spec.n_frames
    packed = syn.render_packed_torch(spec, eng.device)
    params = DetectionParams(method="half_maximum")
    scalars, bg_dev = eng.clip_scalars(packed[:fb], h, w, 12)
    partial = torch.empty(n * 128, dtype=torch.int32, device=eng.device)
    pos = torch.empty(n, dtype=torch.int32, device=eng.device)
    counts = torch.empty(n, dtype=torch.int32, device=eng.device)
    fe = torch.full((1,), 2**31 - 1, dtype=torch.int32, device=eng.device)
    out = {}
    out["full step (frame0 -> bg kernel -> async stats -> stream -> detect -> truncate)"] = timed(
        lambda: eng.process_range(packed, n, h, w, 12, params, frame0=packed[:fb]))
    out["scalars given (no bg kernel, no host sync): fill + stream + detect + truncate"] = timed(
        lambda: eng.process_range(packed, n, h, w, 12, params, scalars, bg_dev))
    out["scalars given, outputs preallocated (no fill, no allocations)"] = timed(
        lambda: eng.process_range(packed, n, h, w, 12, params, scalars, bg_dev, partial=partial, pos_out=pos,
                                  counts_out=counts, first_exit=fe))
    out["same without truncate"] = timed(
        lambda: eng.process_range(packed, n, h, w, 12, params, scalars, bg_dev, partial=partial, pos_out=pos,
                                  counts_out=counts, first_exit=fe, truncate=False))
    lib, st = eng._lib, torch.cuda.current_stream().cuda_stream
    out["stream kernel alone (ff_stream_frames back to back)"] = timed(
        lambda: lib.ff_stream_frames(packed.data_ptr(), None, n, h, w, 12, bg_dev.data_ptr(), -1, 5, None,
                                     partial.data_ptr(), None, 0, None, st))
    print(json.dumps({k: round(v, 4) for k, v in out.items()}, indent=1))


if __name__ == "__main__":
    main()
