#!/usr/bin/env python
"""What this box's memory system can actually do, by access mix - the denominators behind
the roofline fractions DESIGN.md quotes for the write-heavy streaming variants.

MEASURED_PEAKS.json's HBM figure is a 50/50 read+write copy.  The C4 variants of the
streaming kernel write more than they read (uint16 diff 43/57, float64 diff 16/84), so this
tool also measures pure-write and pure-read rates with library kernels, plus the pinned
host->device rate that bounds the end-to-end path.  One JSON object per line.

    python tools/measure_mem_peaks.py [--gib 4] [--reps 10]
"""
from __future__ import annotations

import argparse
import json

import torch


def best_ms(fn, reps: int) -> float:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=4.0)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    n = int(args.gib * (1 << 30)) // 8
    a = torch.empty(n, dtype=torch.int64, device=dev)
    b = torch.empty(n, dtype=torch.int64, device=dev)
    a.fill_(3)
    nbytes = n * 8

    def emit(name, moved, ms, **kw):
        print(json.dumps({"test": name, "bytes_moved": moved, "ms": ms, "gbs": moved / ms / 1e6, **kw}), flush=True)

    emit("copy (read+write, 50/50)", 2 * nbytes, best_ms(lambda: b.copy_(a), args.reps), gib_each=args.gib)
    emit("fill_ (pure write)", nbytes, best_ms(lambda: b.fill_(7), args.reps))
    emit("zero_ (pure write)", nbytes, best_ms(lambda: b.zero_(), args.reps))
    emit("sum int64 (pure read)", nbytes, best_ms(lambda: a.sum(), args.reps))
    a32 = a.view(torch.int32)
    emit("max int32 (pure read)", nbytes, best_ms(lambda: a32.max(), args.reps))
    # write-heavy mixes: out (wide) = f(in (narrow)); same r/w ratios as the C4 variants
    src16 = a.view(torch.int16)[: n]            # n int16 = 2n bytes read
    dst64 = b.view(torch.float64)               # n f64   = 8n bytes written
    emit("int16 -> float64 convert (20/80 r/w)", 2 * n + 8 * n,
         best_ms(lambda: dst64.copy_(src16), args.reps))
    dst32 = b.view(torch.float32)[: n]
    emit("int16 -> float32 convert (33/67 r/w)", 2 * n + 4 * n, best_ms(lambda: dst32.copy_(src16), args.reps))
    del a32, src16, dst64, dst32

    # pinned host -> device, by transfer size
    for mib in (16, 64, 256, 1024, int(args.gib * 1024)):
        sz = mib << 20
        host = torch.empty(sz, dtype=torch.uint8, pin_memory=True)
        host.fill_(1)
        dst = b.view(torch.uint8)[:sz]
        ms = best_ms(lambda: dst.copy_(host, non_blocking=True), max(3, args.reps // 2))
        emit(f"pinned H2D {mib} MiB", sz, ms)
        # device -> pinned host for completeness
        if mib == 256:
            ms = best_ms(lambda: host.copy_(dst, non_blocking=True), max(3, args.reps // 2))
            emit(f"pinned D2H {mib} MiB", sz, ms)
        del host


if __name__ == "__main__":
    main()
