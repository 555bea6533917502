#!/usr/bin/env python
"""Benchmark of the flame-front hot path (decode + detect) on BASELINE.json's configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # CPU reference arm

Workload (N=1): C2 = Nova-style synthetic 1024x128 x 20000 frames, packed 12-bit,
half_maximum, frame-difference profile.  A step is one pass of the whole hot path over that
clip: background reduction of frame 0, fused streaming front end (one HBM read per frame),
warp-per-profile detection, first-exit min, truncation.  For N>1 (weak scaling) the recording
is N x 20000 frames; every rank owns a contiguous 20000-frame range plus a one-frame halo and
the ranks exchange the exit-frame min (all-reduce) and the positions (all-gather) each step.

`value`  = frames/s with the packed clip already resident in HBM (CUDA events, max over ranks).
`e2e`    = frames/s through ff_process_host on pinned HOST buffers: chunked H2D double-buffered
           against the kernels, results copied back to the host, every step.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
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


def config_dict(world: int, frames_per_gpu: int, chunk_mb: int) -> dict:
    return {
        "workload": "C2: Nova-style synthetic 1024x128, packed 12-bit MRAW, half_maximum on the "
                    "frame-difference centre-row profile, per-file calibration",
        "frames_per_gpu": frames_per_gpu, "total_frames": frames_per_gpu * world,
        "width": 1024, "height": 128, "bits": 12, "detection_method": "half_maximum",
        "sharding": "single GPU" if world == 1 else f"contiguous frame ranges + 1-frame halo over {world} GPUs; "
                    "one exchange kernel per rank and step (exit-frame min + truncation + position gather)",
        "l2": "inputs larger than L2 (3.93 GB per GPU per step >> 126 MB); no flush needed",
        "e2e_chunk_mb": chunk_mb,
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


def reference_arm(args) -> None:
    """The reference's CPU implementation of the path on all host cores: its NumPy numeric path
    (oracle port; the Python reference itself cannot travel to the GPU box) under its own
    round-robin frame decomposition, one process per core (mpiexec is not installed, so the
    ranks are multiprocessing workers)."""
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
    n_sample = args.sample_frames
    (l0, l1), (f0, f1) = baseline_sample_ranges(total, int(spec.t_enter), n_sample)
    dev = "cuda" if torch.cuda.is_available() else "cpu"      # torch ops only generate the synthetic input
    blocks = [syn.render_packed_torch(spec, dev, a, b).cpu().numpy() for a, b in ((l0, l1), (f0, f1))]
    packed = np.concatenate(blocks)
    frame0 = fo.frames_from_bytes(syn.render_packed_torch(spec, dev, 0, 1).cpu().numpy(), 1, h, w, 12)[0]
    n = packed.size // fb
    cores = os.cpu_count() or 1
    _POOL_STATE.update(packed=packed, frame0=frame0, h=h, w=w, fb=fb, n=n, method="half_maximum")
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        jobs = [(r, cores) for r in range(cores)]
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            done = sum(pool.map(_pool_worker, jobs, chunksize=1))
            dt = time.perf_counter() - t0
            assert done == n
            if it >= args.warmup:
                times.append(dt)
    sec = sum(times) / len(times)
    value = n / sec
    sample = (f"{n} frames per step ({l1 - l0} lead-in + {f1 - f0} flame frames, the clip's proportions), "
              f"decode + NumPy per-frame path, round-robin over {cores} worker processes")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(world, args.frames, args.chunk_mb),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------
def own_arm(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist
    from high_speed_image_processing_b200 import synthetic as syn
    from high_speed_image_processing_b200._cabi import FF_NO_EXIT
    from high_speed_image_processing_b200.engine import DetectionParams, FlameFrontEngine
    from high_speed_image_processing_b200.sharding import RangeExchange

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:       # launched by hand: re-exec under torchrun
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                                   "--master-port", "29531", __file__, *sys.argv[1:]])
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from high_speed_image_processing_b200.sharding import bind_to_gpu_numa_node
    numa_cpus = bind_to_gpu_numa_node(local_rank)      # before any pinned allocation
    eng = FlameFrontEngine(local_rank, host_chunk_bytes=args.chunk_mb << 20)
    exchange = RangeExchange(engine=eng, transport=args.exchange)

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

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident: `value` ------------------------------------------------
    def step_device():
        if world > 1:       # ff_detect writes straight into this rank's range block; one kernel finishes
            blk = exchange.begin(total)
            eng.process_range(packed, fpr, h, w, 12, params, frame0=frame0, first_frame=a, halo=halo,
                              truncate=False, pos_out=blk.pos, counts_out=blk.counts, first_exit=blk.first_exit)
            g = exchange.finish(blk)
            return g.pos, g.first_exit_t, g.counts
        res = eng.process_range(packed, fpr, h, w, 12, params, frame0=frame0, first_frame=a, halo=halo)
        return res.pos, res.first_exit, res.counts

    for _ in range(args.warmup):
        pos_t, fe_t, cnt_t = step_device()
    barrier()
    sampler = ClockSampler(local_rank, period_ms=args.clock_period_ms)
    if rank == 0 and args.clock_period_ms > 0:
        sampler.start()
        time.sleep(0.25)
    eng._stream_events = []
    launches0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        pos_t, fe_t, cnt_t = step_device()
    ev1.record()
    barrier()
    exchange.check()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = eng.launches - launches0
    stream_ms = [e0.elapsed_time(e1) for e0, e1 in eng._stream_events]
    eng._stream_events = None
    ms_per_step = ms_total / args.steps
    value = total / (ms_per_step * 1e-3)
    kernel_ms = sum(stream_ms) / len(stream_ms)

    # sanity on the result of the last step (not timed)
    pos = pos_t.cpu().numpy()
    counts_np = cnt_t.cpu().numpy()
    first_exit = int(fe_t.cpu().item())
    det = np.nonzero(pos >= 0)[0]
    assert det.size > 200, "bench workload produced no detections"
    ideal = np.array([spec.front_position(float(f)) for f in det])
    assert np.abs(pos[det] - ideal).max() < 20, "detected front does not follow the synthetic front"
    assert first_exit != FF_NO_EXIT and abs(first_exit - spec.exit_frame(10)) < 12

    # ---------------- end to end from pinned host memory: `e2e` -------------------------------
    host = torch.empty(fpr * fb, dtype=torch.uint8, pin_memory=True)
    host.copy_(packed)
    frame0_host = torch.empty(fb, dtype=torch.uint8, pin_memory=True)
    frame0_host.copy_(frame0)
    halo_host = None
    if halo is not None:
        halo_host = torch.empty(fb, dtype=torch.uint8, pin_memory=True)
        halo_host.copy_(halo)
    torch.cuda.synchronize()

    def step_e2e():
        f0 = frame0_host.to(device, non_blocking=True)
        scalars, _ = eng.clip_scalars(f0, h, w, 12)
        hres = eng.process_host(host, fpr, h, w, 12, params, scalars, first_frame=a, halo=halo_host)
        if world > 1:
            g = exchange.finish_arrays(torch.from_numpy(hres.pos).to(device),
                                       torch.tensor([hres.first_exit], dtype=torch.int32, device=device), total)
            return g.pos.cpu().numpy(), g.first_exit
        return hres.pos, hres.first_exit

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(min(2, args.warmup)):
        pos_h, fe_h = step_e2e()
    barrier()
    launches_e0 = eng.launches
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pos_h, fe_h = step_e2e()
    torch.cuda.synchronize()
    e2e_sec = max_over_ranks(time.perf_counter() - t0) / e2e_steps
    launches_e2e = (eng.launches - launches_e0) // e2e_steps
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    assert fe_h == first_exit and np.array_equal(pos_h, pos), "host-streamed result differs from device-resident"
    e2e_value = total / e2e_sec
    # the same call on a PAGEABLE copy of the clip (what np.memmap of the .mraw file is): the library
    # stages it through pinned bounce buffers with a thread pool.  Informational, rank 0 at N=1.
    pageable = None
    if world == 1 and not args.no_pageable:
        host_np = np.empty(fpr * fb, dtype=np.uint8)
        host_np[:] = host.numpy()
        f0 = frame0_host.to(device)
        sc, _ = eng.clip_scalars(f0, h, w, 12)
        eng.process_host(host_np, fpr, h, w, 12, params, sc)
        tp = time.perf_counter()
        for _ in range(2):
            hp = eng.process_host(host_np, fpr, h, w, 12, params, sc)
        tp = (time.perf_counter() - tp) / 2
        assert hp.first_exit == first_exit and np.array_equal(hp.pos, pos)
        pageable = {"value": fpr / tp, "unit": UNIT, "h2d_gbs": fpr * fb / tp / 1e9,
                    "copy_threads": int(os.environ.get("FF_HOST_COPY_THREADS", (os.cpu_count() or 2) // 2))}
        del host_np
    # PCIe / host-memory roofline for the end-to-end path: plain pinned-host -> device copies of the
    # same buffer, (a) this rank alone is not separable under torchrun, so (b) ALL ranks at once
    # behind a barrier - what the box sustains when every GPU pulls from host memory together -
    # is the denominator; at N=1 the two coincide.
    h2d_peak = 0.0
    scratch = torch.empty_like(packed)
    for _ in range(3):
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        scratch.copy_(host, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(c0.elapsed_time(c1))
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
    d2h = 2 * 4 * fpr + 4 + 4 + 2 * w

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
            from oracle import flame_oracle as fo
            sample = np.concatenate([b[fb:] for _, b in blocks])
            cores = os.cpu_count() or 1
            _POOL_STATE.update(packed=sample, frame0=fo.frames_from_bytes(frame0.cpu().numpy(), 1, h, w, 12)[0],
                               h=h, w=w, fb=fb, n=sample.size // fb, method="half_maximum")
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
        if not args.no_cpu_baseline:
            # the oracle loop (the reference's SciPy calls) on frame 0 + a window that starts in the empty
            # lead-in and runs into the flame; its rows double as a checker of the GPU rows
            from oracle import flame_oracle as fo
            from oracle import head_oracle as ho
            n_lead, n_flame = 600, 60
            a0 = int(spec.t_enter) - n_lead
            win = fo.frames_from_bytes(packed[a0 * fb:(a0 + n_lead + n_flame) * fb].cpu().numpy(), n_lead + n_flame,
                                       h, w, 12)
            f0np = fo.frames_from_bytes(frame0.cpu().numpy(), 1, h, w, 12)
            shift = a0 - 1
            t_cpu = time.perf_counter()
            want = ho.run_head(np.concatenate([f0np, win]), rate_h, cal_h, off_h, lambda i: time_of(i + shift))
            t_cpu = time.perf_counter() - t_cpu
            mine = [list(r) for r in got.rows if r[0] < a0 + n_lead + n_flame]
            assert mine == [[r[0] + shift] + r[1:] for r in want.rows], "GPU HEAD rows differ from the oracle loop"
            head["cpu_baseline"] = {"value": (n_lead + n_flame) / t_cpu, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{n_lead} lead-in + {n_flame} flame frames in {t_cpu:.1f} s "
                                              f"(the clip holds ~{FLAME_FRAMES} flame frames in {total}), serial",
                                    "rows_checked": len(mine)}

    if rank == 0:
        peaks_path = REPO / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = fpr * alg_bytes_per_frame / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tpath = REPO / "profiles" / "stream_kernel_traffic.json"
        if tpath.exists():
            traffic = json.loads(tpath.read_text()).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": config_dict(world, fpr, args.chunk_mb),
            "exchange_transport": exchange.transport if world > 1 else None,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "kernel": "ff::count12_kernel<4 stages>",
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": fpr * alg_bytes_per_frame,
                         "peak_source": peak_src + "; a 50/50 read+write copy - this kernel only reads, "
                                        "so frac can exceed 1", "torch_sum_pure_read_gbs": read_probe,
                         "whole_step_gbs": fpr * alg_bytes_per_frame / (ms_per_step * 1e-3) / 1e9},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": e2e_sec * 1e3,
                    "h2d_gbs_per_gpu": h2d / e2e_sec / 1e9, "h2d_peak_gbs_measured": h2d_peak,
                    "h2d_peak_how": "pinned copy of the same buffer, all ranks concurrently, slowest rank",
                    "frac_of_h2d_peak": (h2d / e2e_sec / 1e9) / h2d_peak if h2d_peak else None,
                    "launches_per_step": launches_e2e, "pageable_source": pageable},
            "head_detector": head,
            "gpu_launches": launches,
            "clocks": clocks,
            "result": {"first_exit_frame": first_exit, "detections": int(det.size)},
        }
        print(json.dumps(line), flush=True)
    exchange.close()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100,
                    help="timed steps; 100 x 0.6 ms: a region long enough that its start-up and the clock sampler's queries do not weigh on the step time (10 steps: 0.617 ms, 50+: 0.581-0.585 ms)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU, help="frames per GPU (BASELINE: 20000)")
    ap.add_argument("--chunk-mb", type=int, default=128, help="H2D chunk size of the end-to-end path")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--sample-frames", type=int, default=4000, help="CPU baseline sample size")
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
