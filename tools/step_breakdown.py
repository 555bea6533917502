#!/usr/bin/env python
"""Where a step's time goes outside the range kernel, on 1..N GPUs (CUDA events, clip resident in HBM).

    python tools/step_breakdown.py                                   # one GPU
    torchrun --nproc-per-node N tools/step_breakdown.py [--config C2|C3]

Per rank, 40 back-to-back steps each:
  plain        ff_process_range on the rank's 20000-frame range, no exchange (what every rank does alone)
  hooks        the same with the exchange's hooks (ack wait in prep, exit propagation, publish from the last CTA)
               but NO merge kernel - never valid as a product step, it isolates what the hooks cost
  side         hooks + merge kernel on the side stream (the product's step)
  main         hooks + merge kernel on the main stream (the exchange serialised behind the range kernel)
and the rank skew: per-rank times are gathered, so `max - min` of `plain` is what a barrier would cost anyway.
Rank 0 prints one JSON object.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from high_speed_image_processing_b200 import synthetic as syn  # noqa: E402
from high_speed_image_processing_b200.engine import DetectionParams, FlameFrontEngine  # noqa: E402
from high_speed_image_processing_b200.sharding import RangeExchange, contiguous_range  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2", choices=["C2", "C3"])
    ap.add_argument("--steps", type=int, default=40)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    eng = FlameFrontEngine(local)
    ex = RangeExchange(engine=eng)
    if args.config == "C2":          # weak scaling: 20000 frames per rank, flame at the end of the recording
        fpr = 20000
        total = fpr * world
        base = syn.config_spec("C2")
        spec = syn.SyntheticSpec(**{**base.__dict__, "n_frames": total, "t_enter": float(total - 1100)})
        a, b = rank * fpr, (rank + 1) * fpr
        method = "half_maximum"
    else:                            # strong scaling: one 20000-frame clip
        spec = syn.config_spec("C3")
        total = spec.n_frames
        a, b = contiguous_range(total, rank, world)
        method = "threshold"
    h, w, n = spec.height, spec.width, b - a
    packed = syn.render_packed_torch(spec, device, a, b)
    frame0 = syn.render_packed_torch(spec, device, 0, 1)
    halo = syn.render_packed_torch(spec, device, a - 1, a) if a else None
    params = DetectionParams(method=method)
    host_ms = []

    def timed(step, join=None):
        for _ in range(5):
            step()
        if join:
            join()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            step()
        if join:
            join()
        e1.record()
        host_ms.append((time.perf_counter() - h0) / args.steps * 1e3)      # what the host needs to ENQUEUE a step
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    def plain():
        eng.process_range(packed, n, h, w, 12, params, frame0=frame0, first_frame=a, halo=halo, want_scalars=False)

    out = {"plain": timed(plain)}
    if world > 1:
        lib = eng._lib

        def hooked(merge: str):
            blk = ex.begin(total)
            eng.process_range(packed, n, h, w, 12, params, frame0=frame0, first_frame=a, halo=halo, want_scalars=False,
                              **ex.range_kwargs(blk))
            if merge == "side":
                return ex.finish(blk)
            if merge == "main":
                pos = torch.empty(total, dtype=torch.int32, device=device)
                cnt = torch.empty(total, dtype=torch.int32, device=device)
                fe = torch.empty(1, dtype=torch.int32, device=device)
                st = lib.ff_exchange_finish(ex._xchg, total, pos.data_ptr(), cnt.data_ptr(), fe.data_ptr(),
                                            torch.cuda.current_stream(device).cuda_stream)
                assert st == 0
                return pos
            return None

        out["side"] = timed(lambda: hooked("side"), ex.join)
        out["main"] = timed(lambda: hooked("main"))
        ex.check()
    t = torch.tensor([out.get(k, 0.0) for k in ("plain", "side", "main")], dtype=torch.float64, device=device)
    if world > 1:
        allt = torch.empty(world * 3, dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(allt, t)
        allt = allt.view(world, 3).cpu().tolist()
    else:
        allt = [t.cpu().tolist()]
    if rank == 0:
        rep = {"config": args.config, "n_gpus": world, "frames_per_rank": n, "steps": args.steps, "transport": ex.transport if world > 1 else None,
               "host_enqueue_ms_per_step_rank0": [round(v, 4) for v in host_ms]}
        for k, name in enumerate(("plain", "side", "main")):
            col = [r[k] for r in allt]
            if world == 1 and name != "plain":
                continue
            rep[name] = {"per_rank_ms": [round(v, 4) for v in col], "max_ms": round(max(col), 4), "min_ms": round(min(col), 4)}
        if world > 1:
            rep["rank_skew_ms_of_plain_steps"] = round(rep["plain"]["max_ms"] - rep["plain"]["min_ms"], 4)
            rep["exchange_cost_ms_side_stream"] = round(rep["side"]["max_ms"] - rep["plain"]["max_ms"], 4)
            rep["exchange_cost_ms_main_stream"] = round(rep["main"]["max_ms"] - rep["plain"]["max_ms"], 4)
        print(json.dumps(rep), flush=True)
    ex.close()
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
