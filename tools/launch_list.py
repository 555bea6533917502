#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (one row per kernel launch) as a Markdown
table: kernel, grid, block, launches, total / average / minimum duration and the share of all kernel time.

    # on the GPU box (gpurun), after the same command has exited 0 without ncu:
    ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \\
        python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-pageable --e2e-steps 1 --legs c3_strong,c4
    # here:
    python tools/launch_list.py gpurun_out/launches.csv > profiles/r02_launch_list.md
"""
from __future__ import annotations

import csv
import re
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::|unnamed>::|ff::|^void ", "", name)
    return re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", name)


def main() -> None:
    path = sys.argv[1]
    rows = [r for r in csv.reader(line for line in open(path) if line.startswith('"'))]
    head = rows[0]
    col = {h: i for i, h in enumerate(head)}
    groups: "OrderedDict[tuple, list]" = OrderedDict()
    for r in rows[1:]:
        if r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        ns = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        ns *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}.get(unit, 1.0)
        key = (short(r[col["Kernel Name"]]), r[col["Grid Size"]], r[col["Block Size"]])
        groups.setdefault(key, []).append(ns)
    total = sum(sum(v) for v in groups.values())
    print("| kernel | grid | block | launches | total us | avg us | min us | share of kernel time |")
    print("|---|---|---|---|---|---|---|---|")
    for (name, grid, block), v in groups.items():
        print(f"| `{name}` | {grid} | {block} | {len(v)} | {sum(v) / 1e3:.1f} | {sum(v) / len(v) / 1e3:.2f} | "
              f"{min(v) / 1e3:.2f} | {100.0 * sum(v) / total:.1f} % |")
    print(f"\n{sum(len(v) for v in groups.values())} launches, {total / 1e6:.2f} ms of kernel time in total.")


if __name__ == "__main__":
    main()
