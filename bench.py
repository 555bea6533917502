#!/usr/bin/env python
"""Benchmark of the flame-front hot path (decode + detect) on BASELINE.json's configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # CPU reference arm

Headline workload: C2 = Nova-style synthetic 1024x128 x 20000 frames per GPU, packed 12-bit,
half_maximum, frame-difference profile.  A step is one pass of the whole hot path over that clip: prep
kernel (background scalar of frame 0), ONE range kernel (stream every frame from HBM once, above-noise
counts, empty-frame decision, warp-per-profile detection, first-exit min, truncation).  For N>1 (weak
scaling) the recording is N x 20000 frames; every rank owns a contiguous 20000-frame range plus a
one-frame halo, publishes its block from the range kernel and merges the ranks' blocks (exit-frame min,
truncation, position gather) with one kernel on a side stream.

`value`  = frames/s with the packed clip already resident in HBM (CUDA events, max over ranks).
`e2e`    = frames/s through ff_process_host_range on pinned HOST buffers: chunked H2D double-buffered
           against the kernels, results copied back to the host, every step.

The other BASELINE configurations ride in the same JSON line as extra keys, each checked against the
oracle on sampled frames of every rank:
  `c3_strong`  config 3: ONE 1024x256 x 20000 threshold clip, flame exit at frame ~15000, split into
               contiguous ranges over the N GPUs (strong scaling), device-resident and end to end; the
               ranks share exit frames while they stream and stop uploading behind the first one.
  `c4`         config 4: 1024x1024 x 5000 gradient with the full-frame difference retained on the device
               (uint16, lossless; and the reference's float64), 5000/N frames per GPU.
  `c5`         config 5: a VideoCollection of 8 x N Nova+Mini recordings (64 at N=8), mixed methods and
               calibrations, sharded by video.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

METRIC = "frames/sec decode+detect"
UNIT = "frames/s"
FRAMES_PER_GPU = 20000
FLAME_FRAMES = 1100          # the front enters this many frames before the end of the recording


def workload_spec(total_frames: int):
    from high_speed_image_processing_b200 import synthetic as syn
    base = syn.config_spec("C2")
    return syn.SyntheticSpec(**{**base.__dict__, "n_frames": total_frames,
                                "t_enter": float(max(2, total_frames - FLAME_FRAMES))})


def config_dict(world: int, frames_per_gpu: int, chunk_mb: int, sample_frames: int) -> dict:
    return {
        "workload": "C2: Nova-style synthetic 1024x128, packed 12-bit MRAW, half_maximum on the "
                    "frame-difference centre-row profile, per-file calibration",
        "frames_per_gpu": frames_per_gpu, "total_frames": frames_per_gpu * world,
        "width": 1024, "height": 128, "bits": 12, "detection_method": "half_maximum",
        "sharding": "single GPU" if world == 1 else f"contiguous frame ranges + 1-frame halo over {world} GPUs; "
                    "blocks published by the range kernel, one merge kernel per rank and step on a side stream "
                    "(exit-frame min + truncation + position gather)",
        "l2": "inputs larger than L2 (3.93 GB per GPU per step >> 126 MB); no flush needed",
        "e2e_chunk_mb": chunk_mb,
        "cpu_arm_sample": f"the CPU arms (cpu_baseline, --impl reference) time a {sample_frames}-frame sample per "
                          "step - lead-in and flame frames in the clip's own proportions - not the whole clip",
    }


# ----------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, period_ms: int = 100):
        self.gpu = gpu_index
        self.period_ms = max(10, int(period_ms))
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", str(self.period_ms),
                 "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 8:
                    continue
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     parts[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ----------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference's NumPy path)
# ----------------------------------------------------------------------------------------
def baseline_sample_ranges(total_frames: int, t_enter: int, n_sample: int):
    """Two contiguous blocks in the clip's own proportions: empty lead-in and flame frames."""
    flame_share = min(1.0, FLAME_FRAMES / total_frames)
    n_flame = max(8, int(round(n_sample * flame_share)))
    n_lead = max(8, n_sample - n_flame)
    lead0 = min(1000, max(1, t_enter - n_lead - 1))
    flame0 = min(total_frames - n_flame, t_enter + 200)
    return (lead0, lead0 + n_lead), (flame0, flame0 + n_flame)


def cpu_baseline_serial(packed_blocks, frame0_packed, h, w, method: str) -> dict:
    """Serial oracle (1 core): decode + per-frame loop, exactly the reference's numeric path."""
    from oracle import flame_oracle as fo
    t0 = time.perf_counter()
    n_total = 0
    checks = []
    frame0 = fo.frames_from_bytes(frame0_packed, 1, h, w, 12)[0]
    for first, blk in packed_blocks:           # blk holds [halo, frames...]
        n = blk.size // (h * w * 3 // 2)
        frames = fo.frames_from_bytes(blk, n, h, w, 12)
        r = fo.process_clip(frames[1:], fo.ClipParams(method=method), frame0=frame0, first_index=first,
                            prior_frame=frames[0])
        checks.append((first, r.pos_px, r.nonempty))
        n_total += n - 1
    dt = time.perf_counter() - t0
    return {"frames": n_total, "seconds": dt, "value": n_total / dt, "checks": checks}


_POOL_STATE = {}


def _pool_worker(job):
    from oracle import flame_oracle as fo
    rank, size = job
    st = _POOL_STATE
    h, w, fb = st["h"], st["w"], st["fb"]
    idx = list(range(rank, st["n"], size))      # round-robin, src/photron/parallel.py:99-100
    if not idx:
        return 0
    frames = fo.unpack12(st["packed"].reshape(-1, fb)[idx].reshape(-1)).reshape(len(idx), h, w)
    fo.process_clip(frames, fo.ClipParams(method=st["method"]), frame0=st["frame0"], first_index=0)
    return len(idx)


def _head_pool_worker(job):
    """The reference's own FlameDetector loop on this worker's round-robin share of the frames."""
    from oracle import ref_stage
    rank, size = job
    st = _POOL_STATE
    pv = ref_stage.load()
    idx = list(range(rank, st["n"], size))
    if not idx:
        return 0
    ref_stage.run_head_loop(pv, st["frames"][idx], st["rate"], st["cal"], st["off"], lambda i: i / st["rate"],
                            frame0=st["frame0"])
    return len(idx)


def reference_arm(args) -> None:
    """The reference's CPU implementation of the path on all host cores.

    Default workload: the headline's NumPy numeric path (oracle port - the README-era detection methods
    have no reference code) under the reference's own round-robin frame decomposition, one process per
    core (mpiexec is not installed, so the ranks are multiprocessing workers).
    ``--workload head``: the reference's OWN code - ``FlameDetector`` and the frame functions of
    scripts/process_videos.py, staged into oracle/_ref by build() - on the same clip."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    import numpy as np
    import torch
    from high_speed_image_processing_b200 import synthetic as syn
    from oracle import flame_oracle as fo

    world = max(1, args.gpus)
    total = args.frames * world
    spec = workload_spec(total)
    h, w, fb = spec.height, spec.width, spec.frame_bytes
    head = args.workload == "head"
    n_sample = min(args.sample_frames, 660) if head else args.sample_frames
    (l0, l1), (f0, f1) = baseline_sample_ranges(total, int(spec.t_enter), n_sample)
    dev = "cuda" if torch.cuda.is_available() else "cpu"      # torch ops only generate the synthetic input
    blocks = [syn.render_packed_torch(spec, dev, a, b).cpu().numpy() for a, b in ((l0, l1), (f0, f1))]
    packed = np.concatenate(blocks)
    frame0 = fo.frames_from_bytes(syn.render_packed_torch(spec, dev, 0, 1).cpu().numpy(), 1, h, w, 12)[0]
    n = packed.size // fb
    cores = os.cpu_count() or 1
    kind = "port"
    if head:
        from oracle import ref_stage
        if ref_stage.load() is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref holds no staged reference "
                              "(run __graft_entry__.build() where /root/reference exists)"}), flush=True)
            return
        kind = "reference"
        _POOL_STATE.update(frames=fo.frames_from_bytes(packed, n, h, w, 12), frame0=frame0, n=n,
                           rate=float(spec.record_rate), cal=0.000833333, off=1.347567)
        worker = _head_pool_worker
    else:
        _POOL_STATE.update(packed=packed, frame0=frame0, h=h, w=w, fb=fb, n=n, method="half_maximum")
        worker = _pool_worker
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        jobs = [(r, cores) for r in range(cores)]
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            done = sum(pool.map(worker, jobs, chunksize=1))
            dt = time.perf_counter() - t0
            assert done == n
            if it >= args.warmup:
                times.append(dt)
    sec = sum(times) / len(times)
    value = n / sec
    what = ("the reference's own FlameDetector loop (scripts/process_videos.py staged in oracle/_ref)" if head
            else "decode + NumPy per-frame path (oracle port)")
    sample = (f"{n} frames per step ({l1 - l0} lead-in + {f1 - f0} flame frames, the clip's proportions), "
              f"{what}, round-robin over {cores} worker processes")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(world, args.frames, args.chunk_mb, args.sample_frames),
        "workload": args.workload, "sample_frames_per_step": n,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if head:          # (the default workload keeps the b200 arm's config verbatim; this one names what it ran)
        line["config"].update(
            workload="C2 clip (Nova-style synthetic 1024x128, packed 12-bit MRAW) through the detector the reference "
                     "executes at HEAD: FlameDetector (3x3 opening, Gaussian, Sobel / gradient, windowed tracker)",
            detection_method="head (FlameDetector)")
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------
class Ctx:
    """What every leg needs: ranks, device, engine, exchange, timing helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from high_speed_image_processing_b200.engine import FlameFrontEngine
        from high_speed_image_processing_b200.sharding import RangeExchange, bind_to_gpu_numa_node
        self.args = args
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.device)
        bind_to_gpu_numa_node(self.local_rank)      # before any pinned allocation
        self.eng = FlameFrontEngine(self.local_rank, host_chunk_bytes=args.chunk_mb << 20)
        self.exchange = RangeExchange(engine=self.eng, transport=args.exchange)
        peaks_path = REPO / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            self.peak, self.peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        else:
            self.peak, self.peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_ints(self, x: int):
        if self.world == 1:
            return [int(x)]
        t = self.torch.tensor([x], dtype=self.torch.int64, device=self.device)
        out = self.torch.empty(self.world, dtype=self.torch.int64, device=self.device)
        self.dist.all_gather_into_tensor(out, t)
        return [int(v) for v in out.tolist()]

    def all_ok(self, ok: bool, what: str):
        """Every rank must have passed its oracle check; raises on all ranks otherwise."""
        flag = 1 if ok else 0
        if self.world > 1:
            t = self.torch.tensor([flag], dtype=self.torch.int32, device=self.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
            flag = int(t.item())
        assert flag == 1, f"{what}: the CUDA results differ from the oracle on at least one rank"

    def timed_steps(self, step, steps: int, warmup: int, join=None):
        """warmup untimed steps, then `steps` steps between CUDA events behind a barrier; max over ranks.
        Returns (ms per step, result of the last step)."""
        torch = self.torch
        out = None
        for _ in range(warmup):
            out = step()
        self.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            out = step()
        if join is not None:
            join()
        ev1.record()
        self.barrier()
        return self.max_over_ranks(ev0.elapsed_time(ev1)) / steps, out

    def pinned_copy(self, dev_tensor):
        host = self.torch.empty(dev_tensor.numel(), dtype=self.torch.uint8, pin_memory=True)
        host.copy_(dev_tensor)
        return host

    def close(self):
        self.exchange.close()
        self.eng.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def oracle_windows(fo, packed_dev, a: int, halo_dev, frame0_np, h, w, fb, params, windows, keep_diffs=False):
    """Run the oracle on windows [g0, g1) (clip-global frame indices) of this rank's range.  `packed_dev`
    holds frames [a, ...) of the clip; the frame before a window comes from the range or from the halo."""
    out = []
    for g0, g1 in windows:
        lo = g0 - 1
        if lo >= a:
            raw = packed_dev[(lo - a) * fb:(g1 - a) * fb].cpu().numpy()
            frames = fo.frames_from_bytes(raw, g1 - lo, h, w, 12)
            prior, cur = frames[0], frames[1:]
        else:
            raw = packed_dev[(g0 - a) * fb:(g1 - a) * fb].cpu().numpy()
            cur = fo.frames_from_bytes(raw, g1 - g0, h, w, 12)
            prior = None if halo_dev is None else fo.frames_from_bytes(halo_dev.cpu().numpy(), 1, h, w, 12)[0]
        p = fo.ClipParams(**{**params.__dict__, "keep_diffs": keep_diffs})
        out.append((g0, fo.process_clip(cur, p, frame0=frame0_np, first_index=g0, prior_frame=prior)))
    return out


def check_against_oracle(results, pos, counts, first_exit, no_exit) -> bool:
    """Untruncated oracle positions below the exit frame, DROPPED at/after it; counts everywhere."""
    import numpy as np
    ok = True
    fe = len(pos) if first_exit == no_exit else first_exit
    for g0, r in results:
        n = len(r.pos_px)
        want = r.pos_px.copy()
        got = pos[g0:g0 + n]
        cut = max(0, min(n, fe - g0))
        ok &= bool(np.array_equal(got[:cut], want[:cut])) and bool((got[cut:] == -2).all())
        ok &= bool(np.array_equal(counts[g0:g0 + n], r.nonempty.astype(np.int32)))
    return ok


# ---------------------------------------------------------------------------------------- headline: C2
def leg_c2(cx: Ctx, line: dict) -> None:
    import numpy as np
    torch, dist, eng, exchange, args = cx.torch, cx.dist, cx.eng, cx.exchange, cx.args
    from high_speed_image_processing_b200 import synthetic as syn
    from high_speed_image_processing_b200._cabi import FF_NO_EXIT
    from high_speed_image_processing_b200.engine import DetectionParams
    from oracle import flame_oracle as fo
    world, rank, device = cx.world, cx.rank, cx.device

    fpr = args.frames
    total = fpr * world
    spec = workload_spec(total)
    h, w, fb = spec.height, spec.width, spec.frame_bytes
    a, b = rank * fpr, (rank + 1) * fpr
    packed = syn.render_packed_torch(spec, device, a, b)
    frame0 = syn.render_packed_torch(spec, device, 0, 1)
    halo = syn.render_packed_torch(spec, device, a - 1, a) if a > 0 else None
    params = DetectionParams(method="half_maximum")
    alg_bytes_per_frame = fb + 8                      # packed input once + pos_px + count

    # ---------------- device-resident: `value` ------------------------------------------------
    def step_device():
        if world > 1:       # the range kernel writes straight into this rank's block and publishes it
            blk = exchange.begin(total)
            res = eng.process_range(packed, fpr, h, w, 12, params, frame0=frame0, first_frame=a, halo=halo,
                                    **exchange.range_kwargs(blk))
            return exchange.finish(blk), res
        res = eng.process_range(packed, fpr, h, w, 12, params, frame0=frame0, first_frame=a, halo=halo)
        return res, res

    # the clock sampler comes up BEFORE the warm-up steps: nothing idles the GPU between warm-up and timed region
    sampler = ClockSampler(cx.local_rank, period_ms=args.clock_period_ms)
    if rank == 0 and args.clock_period_ms > 0:
        sampler.start()
        time.sleep(0.25)
    cx.barrier()
    for _ in range(args.warmup):
        step_device()
    cx.barrier()
    eng._stream_events, eng._stream_event_tick = [], 0
    launches0 = eng.launches
    ms_per_step, (out, res) = cx.timed_steps(step_device, args.steps, 0, join=exchange.join if world > 1 else None)
    exchange.check()
    launches = eng.launches - launches0
    stream_ms = [e0.elapsed_time(e1) for e0, e1 in eng._stream_events]
    eng._stream_events = None
    value = total / (ms_per_step * 1e-3)
    kernel_ms = sum(stream_ms) / len(stream_ms)        # prep + range kernel of this rank (events around ff_process_range)
    scalars = res.scalars                               # host float64 statistics of the last step (NumPy)

    # sanity on the result of the last step (not timed)
    pos = out.pos.cpu().numpy()
    counts_np = out.counts.cpu().numpy()
    first_exit = int((out.first_exit_t if world > 1 else out.first_exit).cpu().item())
    det = np.nonzero(pos >= 0)[0]
    assert det.size > 200, "bench workload produced no detections"
    ideal = np.array([spec.front_position(float(f)) for f in det])
    assert np.abs(pos[det] - ideal).max() < 20, "detected front does not follow the synthetic front"
    assert first_exit != FF_NO_EXIT and abs(first_exit - spec.exit_frame(10)) < 12
    assert scalars.background >= 40

    # oracle check on sampled frames of EVERY rank's range (the N=1 CPU baseline below covers 4000 more)
    frame0_np = fo.frames_from_bytes(frame0.cpu().numpy(), 1, h, w, 12)[0]
    windows = [(a + (1 if a == 0 else 0), a + 25)]
    if b > int(spec.t_enter) + 300:                    # this rank holds flame frames
        g = max(a + 1, int(spec.t_enter) + 240)
        windows.append((g, min(b, g + 24)))
    if a <= first_exit < b:
        windows.append((max(a + 1, first_exit - 12), min(b, first_exit + 12)))
    got = oracle_windows(fo, packed, a, halo, frame0_np, h, w, fb, fo.ClipParams(method="half_maximum"), windows)
    cx.all_ok(check_against_oracle(got, pos, counts_np, first_exit, FF_NO_EXIT), "C2")
    oracle_frames = sum(len(r.pos_px) for _, r in got)

    # ---------------- end to end from pinned host memory: `e2e` -------------------------------
    host = cx.pinned_copy(packed)
    frame0_host = cx.pinned_copy(frame0)
    halo_host = cx.pinned_copy(halo) if halo is not None else None
    torch.cuda.synchronize()
    moved = [0]

    def step_e2e():
        f0 = frame0_host.to(device, non_blocking=True)
        sc, _ = eng.clip_scalars(f0, h, w, 12)
        if world > 1:
            blk = exchange.begin(total)
            hres = eng.process_host(host, fpr, h, w, 12, params, sc, first_frame=a, halo=halo_host, block=blk,
                                    hooks=blk.hooks, to_host=False)
            moved[0] = hres.bytes_uploaded
            g = exchange.finish(blk)
            return g.pos.cpu().numpy(), g.first_exit
        hres = eng.process_host(host, fpr, h, w, 12, params, sc, first_frame=a, halo=halo_host)
        moved[0] = hres.bytes_uploaded
        return hres.pos, hres.first_exit

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(min(2, args.warmup)):
        pos_h, fe_h = step_e2e()
    cx.barrier()
    launches_e0 = eng.launches
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pos_h, fe_h = step_e2e()
    torch.cuda.synchronize()
    e2e_sec = cx.max_over_ranks(time.perf_counter() - t0) / e2e_steps
    launches_e2e = (eng.launches - launches_e0) // e2e_steps
    cx.barrier()
    clocks = sampler.stop() if rank == 0 else {}
    assert fe_h == first_exit and np.array_equal(pos_h, pos), "host-streamed result differs from device-resident"
    e2e_value = total / e2e_sec
    moved_per_rank = cx.gather_ints(moved[0] + fb)     # + frame 0 for the clip scalars
    # the same call on a PAGEABLE copy of the clip (what np.memmap of the .mraw file is): the library
    # stages it through pinned bounce buffers with a thread pool.  Informational, rank 0 at N=1.
    pageable = None
    if world == 1 and not args.no_pageable:
        host_np = np.empty(fpr * fb, dtype=np.uint8)
        host_np[:] = host.numpy()
        sc, _ = eng.clip_scalars(frame0_host.to(device), h, w, 12)
        eng.process_host(host_np, fpr, h, w, 12, params, sc)
        tp = time.perf_counter()
        for _ in range(2):
            hp = eng.process_host(host_np, fpr, h, w, 12, params, sc)
        tp = (time.perf_counter() - tp) / 2
        assert hp.first_exit == first_exit and np.array_equal(hp.pos, pos)
        pageable = {"value": fpr / tp, "unit": UNIT, "h2d_gbs": fpr * fb / tp / 1e9,
                    "copy_threads": int(os.environ.get("FF_HOST_COPY_THREADS", (os.cpu_count() or 2) // 2))}
        del host_np
    # PCIe / host-memory roofline for the end-to-end path: plain pinned-host -> device copies of the same
    # buffer by ALL ranks at once behind a barrier - what the box sustains when every GPU pulls from host
    # memory together; at N=1 it is the plain pinned-copy rate.
    h2d_peak = 0.0
    scratch = torch.empty_like(packed)
    for _ in range(3):
        cx.barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        scratch.copy_(host, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        ms = cx.max_over_ranks(c0.elapsed_time(c1))
        h2d_peak = max(h2d_peak, host.numel() / (ms * 1e-3) / 1e9)
    del scratch
    # context for the roofline: what a library pure-read kernel (torch.sum over the same buffer)
    # reaches on this GPU - MEASURED_PEAKS' figure is a 50/50 read+write copy
    read_probe = 0.0
    words = packed[: (packed.numel() // 8) * 8].view(torch.int64)
    for _ in range(4):
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        words.sum()
        c1.record()
        torch.cuda.synchronize()
        read_probe = max(read_probe, words.numel() * 8 / (c0.elapsed_time(c1) * 1e-3) / 1e9)
    h2d = fpr * fb + fb + (fb if halo is not None else 0)
    d2h = (2 * 4 * fpr + 4 + 4 + 2 * w) if world == 1 else (2 * 4 * total + 4 + 4 + 2 * w)

    # ---------------- CPU baseline (rank 0, N=1 only) --------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        (l0, l1), (f0_, f1) = baseline_sample_ranges(total, int(spec.t_enter), args.sample_frames)
        blocks = [(s, packed[(s - 1) * fb:e * fb].cpu().numpy()) for s, e in ((l0, l1), (f0_, f1))]
        r = cpu_baseline_serial(blocks, frame0.cpu().numpy(), h, w, "half_maximum")
        # the baseline's own outputs double as a checker for the timed GPU result on those frames
        # (untruncated positions: compare below the exit frame only)
        for first, o_pos, o_cnt in r["checks"]:
            hi = min(first + len(o_pos), first_exit)
            assert np.array_equal(pos[first:hi], o_pos[:hi - first]), "GPU positions differ from the oracle"
            assert np.array_equal(counts_np[first:first + len(o_cnt)], o_cnt.astype(np.int32)), \
                "GPU above-noise counts differ from the oracle"
        cpu = {"value": r["value"], "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{r['frames']} frames ({l1 - l0} lead-in + {f1 - f0_} flame frames, the clip's "
                         f"proportions) in {r['seconds']:.1f} s: NumPy decode + per-frame path, serial"}
        # ... and on every host core, the reference's round-robin decomposition (parallel.py:99-100),
        # one worker process per core (what `mpiexec -n <cores>` would run; mpi4py is not installed)
        try:
            import multiprocessing as mp
            sample = np.concatenate([blk[fb:] for _, blk in blocks])
            cores = os.cpu_count() or 1
            _POOL_STATE.update(packed=sample, frame0=frame0_np, h=h, w=w, fb=fb, n=sample.size // fb,
                               method="half_maximum")
            with mp.get_context("fork").Pool(cores) as pool:
                jobs = [(k, cores) for k in range(cores)]
                pool.map(_pool_worker, jobs, chunksize=1)                      # warm-up
                t_all = time.perf_counter()
                for _ in range(2):
                    pool.map(_pool_worker, jobs, chunksize=1)
                t_all = (time.perf_counter() - t_all) / 2
            cpu["all_cores"] = {"value": (sample.size // fb) / t_all, "unit": UNIT, "cores": cores,
                                "how": "same sample, round-robin over one worker process per core"}
        except Exception as exc:                                               # never fail the GPU line over it
            cpu["all_cores"] = {"error": repr(exc)}

    # ---------------- the detector the reference executes at HEAD, same clip (rank 0, N=1 only) -------
    head = None
    if rank == 0 and world == 1 and not args.no_head:
        head = leg_head(cx, spec, packed, frame0, fpr, h, w, fb, host=host)

    if rank == 0:
        achieved = fpr * alg_bytes_per_frame / (kernel_ms * 1e-3) / 1e9
        traffic, traffic_note = measured_traffic()
        line.update({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": config_dict(world, fpr, args.chunk_mb, args.sample_frames),
            "exchange_transport": exchange.transport if world > 1 else None,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": cx.peak, "unit": "GB/s",
                         "frac": achieved / cx.peak, "traffic": traffic, "traffic_source": traffic_note,
                         "kernel": "ff::range_kernel<12, 4 stages> (+ prep_kernel)",
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": fpr * alg_bytes_per_frame,
                         "peak_source": cx.peak_src + "; a 50/50 read+write copy - this kernel only reads, "
                                        "so frac can exceed 1", "torch_sum_pure_read_gbs": read_probe,
                         "whole_step_gbs": fpr * alg_bytes_per_frame / (ms_per_step * 1e-3) / 1e9},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": e2e_sec * 1e3,
                    "h2d_gbs_per_gpu": h2d / e2e_sec / 1e9, "h2d_peak_gbs_measured": h2d_peak,
                    "h2d_peak_how": "pinned copy of the same buffer, all ranks concurrently, slowest rank",
                    "frac_of_h2d_peak": (h2d / e2e_sec / 1e9) / h2d_peak if h2d_peak else None,
                    "h2d_bytes_moved_per_rank": moved_per_rank,
                    "launches_per_step": launches_e2e, "pageable_source": pageable},
            "head_detector": head,
            "gpu_launches": launches,
            "clocks": clocks,
            "result": {"first_exit_frame": first_exit, "detections": int(det.size),
                       "oracle_checked_frames_per_rank": oracle_frames,
                       "clip_scalars": {"background": scalars.background, "flame_threshold": scalars.flame_threshold}},
        })
    del packed, host
    torch.cuda.empty_cache()


def measured_traffic():
    """DRAM bytes per launch of the range kernel from the round's `ncu --set full` capture, only if that capture
    was taken on a library whose ff_stream.cu (where range_kernel lives) was compiled from the same source, headers
    and flags as the loaded one (`build.unit_fingerprint`: nvcc's output is not byte-reproducible, so the binary's
    own hash would change with every rebuild); else null."""
    tpath = REPO / "profiles" / "range_kernel_traffic.json"
    if not tpath.exists():
        return None, "no capture committed"
    rec = json.loads(tpath.read_text())
    from high_speed_image_processing_b200 import build as ffbuild
    unit = rec.get("unit", "ff_stream.cu")
    built, now = ffbuild.built_unit_fingerprint(unit), ffbuild.unit_fingerprint(unit)
    if not built or built != now:
        return None, "the loaded libflamefront.so was not built from the sources in the tree"
    if rec.get("unit_fingerprint") != built:
        return None, (f"profiles/range_kernel_traffic.json was captured on a build of other sources "
                      f"({str(rec.get('unit_fingerprint', '?'))[:12]})")
    return rec.get("dram_bytes_per_launch"), f"ncu --set full, {rec.get('captured', '?')}, {unit} built from the same sources"


class _HostClip:
    """The part of PhotonVideo that process_video reads, over a clip that already sits in (pinned) host memory."""

    def __init__(self, raw, n, h, w, bits, spec):
        self.raw, self.n, self.frame_shape, self.storage_bits, self.spec = raw, n, (h, w), bits, spec
        self.fb = h * w * bits // 8
        self.frame_rate = int(spec.record_rate)

    def __len__(self):
        return self.n

    def raw_frames(self, a, b):
        return self.raw[a * self.fb:b * self.fb]

    def get_absolute_time(self, i):
        return (self.spec.start_frame + i * self.spec.skip_frame) / self.spec.record_rate

    get_time = get_absolute_time


def leg_head(cx: Ctx, spec, packed, frame0, fpr, h, w, fb, host=None) -> dict:
    """The same clip through the FlameDetector parity path, next to the reference's own code on the CPU."""
    import numpy as np
    torch, eng, args = cx.torch, cx.eng, cx.args
    from high_speed_image_processing_b200.head import HeadParams, finish_head_track
    hp = HeadParams()
    cal_h, off_h, rate_h = 0.000833333, 1.347567, float(spec.record_rate)
    for _ in range(3):
        hres = eng.process_head(packed, fpr, h, w, 12, hp, rate_h, cal_h)
    torch.cuda.synchronize()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for _ in range(args.steps):
        hres = eng.process_head(packed, fpr, h, w, 12, hp, rate_h, cal_h)
    h1.record()
    torch.cuda.synchronize()
    head_ms = h0.elapsed_time(h1) / args.steps
    time_of = lambda i: (spec.start_frame + i * spec.skip_frame) / spec.record_rate      # noqa: E731
    got = finish_head_track(hres.track.cpu().numpy(), hres.flags.cpu().numpy(), 0, w, rate_h, cal_h, off_h,
                            time_of, hp)
    head = {"what": "FlameDetector parity path (3x3 opening, Gaussian, Sobel/gradient in float64, windowed "
                    "tracker) on the same clip, device-resident, steps back to back",
            "value": fpr / (head_ms * 1e-3), "unit": UNIT, "ms_per_clip": head_ms, "rows": len(got.rows),
            "stop": list(got.stop) if got.stop else None}
    if host is not None:
        # end to end through the public call: process_video(detection_method="head") on the clip in pinned host
        # memory - chunked upload, streaming + band + tracker kernels, rows and velocities built on the host
        from high_speed_image_processing_b200.process_videos import VideoSourceConfig, process_video
        clip_h = _HostClip(host.numpy(), fpr, h, w, 12, spec)
        cfg = VideoSourceConfig(name="bench", enabled=True, calibration=cal_h, position_offset=off_h,
                                detection_method="head", head_params=hp)
        vres = process_video(clip_h, cfg, cal_h, off_h, engine=eng)
        n_e2e = max(1, min(args.steps, args.e2e_steps))
        t_e = time.perf_counter()
        for _ in range(n_e2e):
            vres = process_video(clip_h, cfg, cal_h, off_h, engine=eng)
        torch.cuda.synchronize()
        t_e = (time.perf_counter() - t_e) / n_e2e
        assert [list(r) for r in vres.rows] == [list(r) for r in got.rows] and vres.stop == got.stop, \
            "HEAD rows from the host-resident clip differ from the device-resident run"
        stop_f = got.stop[1] if got.stop else fpr
        head["e2e"] = {"value": fpr / t_e, "unit": UNIT, "ms_per_clip": t_e * 1e3, "steps": n_e2e,
                       "h2d_bytes_per_step": int(min(fpr, stop_f + 1) * fb + fb),
                       "d2h_bytes_per_step": int(fpr * 21),
                       "how": "process_video(detection_method='head') on the clip in pinned host memory: uploads in "
                              "chunks of FF_HEAD_CHUNK_MB, kernels, result rows + velocities on the host"}
    if args.no_cpu_baseline:
        return head
    # frame 0 + a window that starts in the empty lead-in and runs into the flame, through the reference's OWN
    # loop (oracle/_ref) when it was staged, else the oracle loop; its rows double as a checker of the GPU rows
    from oracle import flame_oracle as fo
    from oracle import head_oracle as ho
    from oracle import ref_stage
    n_lead, n_flame = 600, 60
    a0 = int(spec.t_enter) - n_lead
    win = fo.frames_from_bytes(packed[a0 * fb:(a0 + n_lead + n_flame) * fb].cpu().numpy(), n_lead + n_flame, h, w, 12)
    f0np = fo.frames_from_bytes(frame0.cpu().numpy(), 1, h, w, 12)
    shift = a0 - 1
    clip = np.concatenate([f0np, win])
    pv = ref_stage.load()
    t_cpu = time.perf_counter()
    if pv is not None:
        rows, _ = ref_stage.run_head_loop(pv, clip, rate_h, cal_h, off_h, lambda i: time_of(i + shift))
        kind = "reference"
    else:
        rows = ho.run_head(clip, rate_h, cal_h, off_h, lambda i: time_of(i + shift)).rows
        kind = "port"
    t_cpu = time.perf_counter() - t_cpu
    mine = [list(r) for r in got.rows if r[0] < a0 + n_lead + n_flame]
    assert mine == [[r[0] + shift] + r[1:] for r in rows], "GPU HEAD rows differ from the reference loop"
    head["cpu_baseline"] = {"value": (n_lead + n_flame) / t_cpu, "unit": UNIT, "cores": 1, "kind": kind,
                            "sample": f"{n_lead} lead-in + {n_flame} flame frames in {t_cpu:.1f} s "
                                      f"(the clip holds ~{FLAME_FRAMES} flame frames in {fpr}), serial; "
                                      + ("the reference's own FlameDetector / frame functions (oracle/_ref)"
                                         if kind == "reference" else "oracle loop (reference not staged)"),
                            "rows_checked": len(mine)}
    if pv is not None:
        try:
            import multiprocessing as mp
            cores = os.cpu_count() or 1
            _POOL_STATE.update(frames=clip[1:], frame0=clip[0], n=len(clip) - 1, rate=rate_h, cal=cal_h, off=off_h)
            with mp.get_context("fork").Pool(cores) as pool:
                jobs = [(k, cores) for k in range(cores)]
                pool.map(_head_pool_worker, jobs, chunksize=1)
                t_all = time.perf_counter()
                pool.map(_head_pool_worker, jobs, chunksize=1)
                t_all = time.perf_counter() - t_all
            head["cpu_baseline"]["all_cores"] = {
                "value": (len(clip) - 1) / t_all, "unit": UNIT, "cores": cores, "kind": "reference",
                "how": "same window, the reference's round-robin frame decomposition (parallel.py:99-100), one "
                       "worker process per core, each with its own FlameDetector as under mpiexec"}
            if "e2e" in head:
                head["e2e_ratio_vs_reference_all_cores"] = head["e2e"]["value"] / head["cpu_baseline"]["all_cores"]["value"]
        except Exception as exc:
            head["cpu_baseline"]["all_cores"] = {"error": repr(exc)}
    return head


# ---------------------------------------------------------------------------------------- config 3
def leg_c3(cx: Ctx) -> dict:
    """BASELINE config 3: ONE clip split over the ranks (strong scaling), threshold, exit at ~15000."""
    import numpy as np
    torch, eng, exchange, args = cx.torch, cx.eng, cx.exchange, cx.args
    from high_speed_image_processing_b200 import synthetic as syn
    from high_speed_image_processing_b200._cabi import FF_NO_EXIT
    from high_speed_image_processing_b200.engine import DetectionParams
    from high_speed_image_processing_b200.sharding import contiguous_range
    from oracle import flame_oracle as fo
    world, rank, device = cx.world, cx.rank, cx.device
    spec = syn.config_spec("C3")
    total = spec.n_frames if args.c3_frames <= 0 else args.c3_frames
    if total != spec.n_frames:
        spec = syn.config_spec("C3", n_frames=total)
    h, w, fb = spec.height, spec.width, spec.frame_bytes
    a, b = contiguous_range(total, rank, world)
    n = b - a
    packed = syn.render_packed_torch(spec, device, a, b)
    frame0 = syn.render_packed_torch(spec, device, 0, 1)
    halo = syn.render_packed_torch(spec, device, a - 1, a) if a > 0 else None
    params = DetectionParams(method="threshold")
    steps, warm = args.steps, max(3, args.warmup)

    def step_device():
        if world > 1:
            blk = exchange.begin(total)
            eng.process_range(packed, n, h, w, 12, params, frame0=frame0, first_frame=a, halo=halo,
                              want_scalars=False, **exchange.range_kwargs(blk))
            return exchange.finish(blk)
        return eng.process_range(packed, n, h, w, 12, params, frame0=frame0, first_frame=a, halo=halo,
                                 want_scalars=False)

    eng._stream_events, eng._stream_event_tick = [], 0
    ms_dev, out = cx.timed_steps(step_device, steps, warm, join=exchange.join if world > 1 else None)
    ev = eng._stream_events[1:]            # (every 4th call is timed; the first sample is a warm-up step)
    eng._stream_events = None
    kernel_ms = cx.max_over_ranks(sum(e0.elapsed_time(e1) for e0, e1 in ev) / len(ev))
    exchange.check()
    pos = out.pos.cpu().numpy()
    counts = out.counts.cpu().numpy()
    first_exit = int((out.first_exit_t if world > 1 else out.first_exit).cpu().item())
    assert first_exit != FF_NO_EXIT and abs(first_exit - spec.exit_frame(10)) < 12, first_exit

    frame0_np = fo.frames_from_bytes(frame0.cpu().numpy(), 1, h, w, 12)[0]
    windows = [(a + (1 if a == 0 else 0), min(b, a + 13))]
    t_in = int(spec.t_enter)
    if a <= t_in + 200 < b:
        windows.append((t_in + 200, min(b, t_in + 212)))
    if a <= first_exit < b:
        windows.append((max(a + 1, first_exit - 8), min(b, first_exit + 8)))
    got = oracle_windows(fo, packed, a, halo, frame0_np, h, w, fb, fo.ClipParams(method="threshold"), windows)
    ok = check_against_oracle(got, pos, counts, first_exit, FF_NO_EXIT)
    for g0, r in got:              # the window around the exit must hold the oracle's own first exit frame
        if g0 <= first_exit < g0 + len(r.pos_px):
            ok &= (g0 + r.first_exit == first_exit)
    cx.all_ok(ok, "C3")

    # the same clip on ONE GPU of this box (rank 0 alone), for the strong-scaling efficiency of this run
    ms_one = None
    if world > 1:
        if rank == 0:
            whole = syn.render_packed_torch(spec, device, 0, total)
            for _ in range(3):
                r1 = eng.process_range(whole, total, h, w, 12, params, frame0=frame0, want_scalars=False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                r1 = eng.process_range(whole, total, h, w, 12, params, frame0=frame0, want_scalars=False)
            e1.record()
            torch.cuda.synchronize()
            ms_one = e0.elapsed_time(e1) / steps
            assert np.array_equal(r1.pos.cpu().numpy(), pos), "range-sharded C3 differs from the single-GPU run"
            del whole, r1
            torch.cuda.empty_cache()
        cx.barrier()

    # end to end: every rank streams ITS range from pinned host memory; exit frames are shared while streaming
    host = cx.pinned_copy(packed)
    frame0_host = cx.pinned_copy(frame0)
    halo_host = cx.pinned_copy(halo) if halo is not None else None
    moved = [0, 0]

    def step_e2e():
        sc, _ = eng.clip_scalars(frame0_host.to(device, non_blocking=True), h, w, 12)
        if world > 1:
            blk = exchange.begin(total)
            hres = eng.process_host(host, n, h, w, 12, params, sc, first_frame=a, halo=halo_host, block=blk,
                                    hooks=blk.hooks, to_host=False)
            moved[0], moved[1] = hres.bytes_uploaded, hres.frames_done
            g = exchange.finish(blk)
            return g.pos.cpu().numpy(), g.first_exit
        hres = eng.process_host(host, n, h, w, 12, params, sc, first_frame=a, halo=halo_host)
        moved[0], moved[1] = hres.bytes_uploaded, hres.frames_done
        return hres.pos, hres.first_exit

    e2e_steps = max(2, min(steps, args.e2e_steps))
    for _ in range(2):
        pos_h, fe_h = step_e2e()
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pos_h, fe_h = step_e2e()
    torch.cuda.synchronize()
    e2e_sec = cx.max_over_ranks(time.perf_counter() - t0) / e2e_steps
    cx.barrier()
    assert fe_h == first_exit and np.array_equal(pos_h, pos), "C3: host-streamed result differs from device-resident"
    moved_per_rank = cx.gather_ints(moved[0])
    frames_done_per_rank = cx.gather_ints(moved[1])
    alg = fb + 8
    res = {
        "config": "BASELINE config 3: Mini-style synthetic 1024x256 x %d frames, threshold detection, flame exit at "
                  "frame %d, ONE clip split into contiguous frame ranges over %d GPU(s)" % (total, first_exit, world),
        "scaling": "strong", "total_frames": total, "frames_per_gpu": [contiguous_range(total, r, world)[1] -
                                                                         contiguous_range(total, r, world)[0]
                                                                         for r in range(world)],
        "value": total / (ms_dev * 1e-3), "unit": UNIT, "ms_per_step": ms_dev, "steps": steps,
        "range_kernel_ms_slowest_rank": kernel_ms,
        "range_kernel_gbs_per_gpu": max(1, n) * alg / (kernel_ms * 1e-3) / 1e9 if n else None,
        "frac_of_hbm_peak_per_gpu": (n * alg / (kernel_ms * 1e-3) / 1e9) / cx.peak if n else None,
        "single_gpu_same_box_ms": ms_one,
        "strong_scaling_efficiency": (ms_one / (world * ms_dev)) if ms_one else (1.0 if world == 1 else None),
        "e2e": {"value": total / e2e_sec, "unit": UNIT, "ms_per_step": e2e_sec * 1e3, "steps": e2e_steps,
                "h2d_bytes_moved_per_rank": moved_per_rank, "frames_uploaded_per_rank": frames_done_per_rank,
                "h2d_bytes_if_nothing_stopped": total * fb,
                "fraction_of_clip_uploaded": sum(moved_per_rank) / float(total * fb),
                "early_exit": "ranks share exit frames through peer memory while they stream; nothing at or behind "
                              "the smallest one is uploaded (scripts/process_videos.py:1494 across ranks)"},
        "first_exit_frame": first_exit, "detections": int((pos >= 0).sum()),
        "oracle_checked_frames_per_rank": sum(len(r.pos_px) for _, r in got),
    }
    del packed, host
    torch.cuda.empty_cache()
    return res


# ---------------------------------------------------------------------------------------- config 4
def leg_c4(cx: Ctx) -> dict:
    """BASELINE config 4: 1024x1024 x 5000, gradient, full-frame difference retained on the device."""
    import numpy as np
    torch, eng, exchange, args = cx.torch, cx.eng, cx.exchange, cx.args
    from high_speed_image_processing_b200 import synthetic as syn
    from high_speed_image_processing_b200._cabi import FF_NO_EXIT
    from high_speed_image_processing_b200.engine import DetectionParams
    from high_speed_image_processing_b200.sharding import contiguous_range
    from oracle import flame_oracle as fo
    world, rank, device = cx.world, cx.rank, cx.device
    spec = syn.config_spec("C4")
    total = spec.n_frames if args.c4_frames <= 0 else args.c4_frames
    if total != spec.n_frames:
        spec = syn.config_spec("C4", n_frames=total)
    h, w, fb = spec.height, spec.width, spec.frame_bytes
    a, b = contiguous_range(total, rank, world)
    n = b - a
    packed = syn.render_packed_torch(spec, device, a, b)
    frame0 = syn.render_packed_torch(spec, device, 0, 1)
    halo = syn.render_packed_torch(spec, device, a - 1, a) if a > 0 else None
    params = DetectionParams(method="gradient")
    frame0_np = fo.frames_from_bytes(frame0.cpu().numpy(), 1, h, w, 12)[0]
    out = {"config": "BASELINE config 4: synthetic 1024x1024 x %d frames, gradient detection, full-frame difference "
                     "retained on the device, contiguous ranges over %d GPU(s)" % (total, world),
           "scaling": "strong", "total_frames": total, "variants": {}}
    keep = {}
    for dtype, steps in (("uint16", max(3, min(args.steps, 10))), ("float64", 3)):
        px_bytes = {"uint16": 2, "float64": 8}[dtype]
        if n * h * w * px_bytes > 100e9:       # leave room next to the clip on a 180 GB part
            out["variants"][dtype] = {"skipped": "retained image would exceed 100 GB on one GPU"}
            continue

        def step():
            if world > 1:
                blk = exchange.begin(total)
                keep["res"] = eng.process_range(packed, n, h, w, 12, params, frame0=frame0, first_frame=a, halo=halo,
                                                diff_dtype=dtype, want_scalars=False, **exchange.range_kwargs(blk))
                return exchange.finish(blk)
            keep["res"] = eng.process_range(packed, n, h, w, 12, params, frame0=frame0, first_frame=a, halo=halo,
                                            diff_dtype=dtype, want_scalars=False)
            return keep["res"]

        eng._stream_events, eng._stream_events_every = [], 1      # multi-millisecond steps: time every one of them
        ms, g = cx.timed_steps(step, steps, 3, join=exchange.join if world > 1 else None)
        ev = eng._stream_events[3:]
        eng._stream_events, eng._stream_events_every = None, 4
        kernel_ms = cx.max_over_ranks(sum(e0.elapsed_time(e1) for e0, e1 in ev) / len(ev))
        exchange.check()
        pos = g.pos.cpu().numpy()
        counts = g.counts.cpu().numpy()
        first_exit = int((g.first_exit_t if world > 1 else g.first_exit).cpu().item())
        diff = keep["res"].diff
        # oracle: positions, counts and the retained difference frames themselves, on sampled frames of this rank
        windows = [(a + (1 if a == 0 else 0), min(b, a + 4))]
        t_in = int(spec.t_enter)
        if a <= t_in + 300 < b:
            windows.append((t_in + 300, min(b, t_in + 303)))
        got = oracle_windows(fo, packed, a, halo, frame0_np, h, w, fb, fo.ClipParams(method="gradient"), windows,
                             keep_diffs=True)
        ok = check_against_oracle(got, pos, counts, first_exit, FF_NO_EXIT)
        for g0, r in got:
            mine = diff[g0 - a:g0 - a + len(r.pos_px)].cpu().numpy().astype(np.float64)
            ok &= bool(np.array_equal(mine, r.diffs))
        cx.all_ok(ok, f"C4 {dtype}")
        alg = fb + h * w * px_bytes + 8
        out["variants"][dtype] = {
            "value": total / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "kernels_ms_slowest_rank": kernel_ms, "algorithmic_bytes_per_frame": alg,
            "kernels_gbs_per_gpu": n * alg / (kernel_ms * 1e-3) / 1e9,
            "frac_of_hbm_peak_per_gpu": (n * alg / (kernel_ms * 1e-3) / 1e9) / cx.peak,
            "retained_gb_per_gpu": n * h * w * px_bytes / 1e9, "first_exit_frame": None if first_exit == FF_NO_EXIT else first_exit,
            "detections": int((pos >= 0).sum()),
            "oracle_checked_frames_per_rank": sum(len(r.pos_px) for _, r in got),
            "oracle_check": "positions, counts and every pixel of the retained difference frames",
        }
        del diff, g
        keep.clear()
        torch.cuda.empty_cache()
    # end to end for the uint16 variant: upload of the rank's range from pinned memory + the kernels
    host = cx.pinned_copy(packed)
    frame0_host = cx.pinned_copy(frame0)

    def step_e2e():
        f0 = frame0_host.to(device, non_blocking=True)
        dev_frames = eng.upload(host)
        if world > 1:
            blk = exchange.begin(total)
            keep["res"] = eng.process_range(dev_frames, n, h, w, 12, params, frame0=f0, first_frame=a, halo=halo,
                                            diff_dtype="uint16", want_scalars=False, **exchange.range_kwargs(blk))
            g = exchange.finish(blk)
            return g.pos.cpu().numpy()
        keep["res"] = eng.process_range(dev_frames, n, h, w, 12, params, frame0=f0, first_frame=a, halo=halo,
                                        diff_dtype="uint16", want_scalars=False)
        return keep["res"].pos.cpu().numpy()

    step_e2e()
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        step_e2e()
    torch.cuda.synchronize()
    e2e_sec = cx.max_over_ranks(time.perf_counter() - t0) / 2
    out["e2e_uint16"] = {"value": total / e2e_sec, "unit": UNIT, "ms_per_step": e2e_sec * 1e3,
                         "h2d_bytes_per_rank": n * fb + fb, "how": "upload of the rank's range from pinned memory "
                         "(ff_host_upload), then the kernels; the difference image stays on the device"}
    keep.clear()
    del packed, host
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------- config 5
def leg_c5(cx: Ctx) -> dict:
    """BASELINE config 5: a VideoCollection of Nova+Mini recordings sharded by video (8 per GPU: 64 at N=8)."""
    import numpy as np
    args = cx.args
    from tools import bench_collection as bc
    from oracle import flame_oracle as fo
    clips = args.c5_clips_per_gpu * cx.world
    vdir = Path(tempfile.gettempdir()) / f"ff_bench_c5_{os.environ.get('MASTER_PORT', 'solo')}"
    rep = bc.run_collection(cx.eng, cx.exchange, cx.device, cx.rank, cx.world, clips, args.c5_frames, vdir,
                            reps=2, residency="auto", pin=True)
    # oracle: two windows (lead-in, flame) of up to two of this rank's clips, rows and exit frame
    ok = True
    checked = 0
    for i in rep["mine"][:2]:
        spec, cfg, res = rep["specs"][i], rep["configs"][i], rep["results"][i]
        h, w, fb = spec.height, spec.width, spec.frame_bytes
        raw = np.fromfile(vdir / f"run-{i:02d}-.mraw", dtype=np.uint8)
        frame0_np = fo.frames_from_bytes(raw[:fb], 1, h, w, 12)[0]
        t_in = int(spec.t_enter)
        fe = spec.n_frames if res.first_exit is None else res.first_exit
        for g0, g1 in ((1, 9), (t_in + 150, t_in + 158)):
            g1 = min(g1, spec.n_frames)
            if g0 >= g1:
                continue
            frames = fo.frames_from_bytes(raw[(g0 - 1) * fb:g1 * fb], g1 - g0 + 1, h, w, 12)
            r = fo.process_clip(frames[1:], fo.ClipParams(method=cfg.detection_method), frame0=frame0_np,
                                first_index=g0, prior_frame=frames[0])
            cut = max(0, min(len(r.pos_px), fe - g0))
            ok &= bool(np.array_equal(res.pos_px[g0:g0 + cut], r.pos_px[:cut]))
            ok &= bool(np.array_equal(res.nonempty_counts[g0:g0 + len(r.nonempty)], r.nonempty.astype(np.int32)))
            checked += len(r.pos_px)
        cal, off = cfg.get_calibration_for_file(f"run-{i:02d}-.cihx")
        ok &= all(p_m == px * cal + off for _, _, px, p_m, _ in res.rows[:50])
    cx.all_ok(ok, "C5")
    out = rep["summary"]
    out["oracle_checked_frames_per_rank"] = checked
    cx.barrier()
    if cx.rank == 0:
        shutil.rmtree(vdir, ignore_errors=True)
    return out


def own_arm(args) -> None:
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:       # launched by hand: re-exec under torchrun
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                                   "--master-port", "29531", __file__, *sys.argv[1:]])
    cx = Ctx(args)
    line: dict = {}
    legs = [x for x in args.legs.split(",") if x]
    t_legs = {}
    t0 = time.perf_counter()
    leg_c2(cx, line)
    t_legs["c2"] = time.perf_counter() - t0
    for name, fn in (("c3_strong", leg_c3), ("c4", leg_c4), ("c5", leg_c5)):
        if name not in legs:
            continue
        t0 = time.perf_counter()
        try:
            res = fn(cx)
        except AssertionError:
            raise
        except Exception as exc:           # an extra leg must not cost the headline line (parity failures do raise)
            res = {"error": repr(exc)}
            cx.torch.cuda.empty_cache()
        t_legs[name] = time.perf_counter() - t0
        if cx.rank == 0:
            line[name] = res
    if cx.rank == 0:
        head = line.get("head_detector")
        if head and head.get("cpu_baseline", {}).get("all_cores", {}).get("value"):
            head["ratio_vs_reference_all_cores"] = head["value"] / head["cpu_baseline"]["all_cores"]["value"]
        line["leg_seconds"] = {k: round(v, 1) for k, v in t_legs.items()}
        print(json.dumps(line), flush=True)
    cx.close()


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100,
                    help="timed steps; 100 x 0.6 ms: a region long enough that its start-up and the clock sampler's queries do not weigh on the step time")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=["c2", "head"], default="c2",
                    help="--impl reference only: the headline's NumPy path (port) or the reference's own FlameDetector loop")
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU, help="frames per GPU (BASELINE: 20000)")
    ap.add_argument("--chunk-mb", type=int, default=128, help="H2D chunk size of the end-to-end path")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--sample-frames", type=int, default=4000, help="CPU baseline sample size")
    ap.add_argument("--legs", default="c3_strong,c4,c5", help="extra BASELINE configurations to run (comma list; empty = none)")
    ap.add_argument("--c3-frames", type=int, default=0, help="override config 3's frame count (tests)")
    ap.add_argument("--c4-frames", type=int, default=0, help="override config 4's frame count (tests)")
    ap.add_argument("--c5-clips-per-gpu", type=int, default=8)
    ap.add_argument("--c5-frames", type=int, default=2000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pageable", action="store_true", help="skip the pageable-source end-to-end leg")
    ap.add_argument("--no-head", action="store_true", help="skip the HEAD-detector leg (N=1 only)")
    ap.add_argument("--clock-period-ms", type=int, default=100, help="nvidia-smi sampling period (0 = off)")
    ap.add_argument("--exchange", choices=["auto", "peer", "gathered"], default="auto",
                    help="multi-GPU block transport: peer memory over NVLink (CUDA IPC) or one NCCL all-gather")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3                      # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        reference_arm(args)
    else:
        own_arm(args)


if __name__ == "__main__":
    main()
