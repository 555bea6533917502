#!/usr/bin/env python
"""Turn an `ncu --set full` report of the range kernel into profiles/range_kernel_traffic.json, stamped with the
fingerprint of the translation unit (ff_stream.cu + headers + nvcc flags) the range kernel of the libflamefront.so it
was captured on was compiled from (bench.py reports `roofline.traffic` only when the loaded library's unit was built
from the same sources; nvcc's output itself is not byte-reproducible).

    # on the GPU box (gpurun):
    ncu --set full --clock-control none --import-source on -k regex:range_kernel -c 1 -o gpurun_out/range_c2 \\
        python tools/bench_variants.py --only C2:None --reps 1
    # here:
    python tools/capture_traffic.py gpurun_out/range_c2.ncu-rep [--details profiles/r02_range_kernel_c2_details.txt]
"""
from __future__ import annotations

import csv
import io
import json
import subprocess
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from high_speed_image_processing_b200 import build as ffbuild  # noqa: E402


def main() -> None:
    rep = Path(sys.argv[1])
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, vals = rows[0], rows[1], rows[2]
    m = {h: (v, u) for h, u, v in zip(head, units, vals)}

    def to_bytes(key):
        v, u = m[key]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
        return float(v.replace(",", "")) * scale

    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    dur, dur_u = m["gpu__time_duration.sum"]
    dur_us = float(dur.replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}[dur_u]
    out = {
        "kernel": m["Kernel Name"][0],
        "dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
        "duration_us_under_ncu": dur_us,
        "dram_throughput_pct_of_peak": float(m["dram__throughput.avg.pct_of_peak_sustained_elapsed"][0]) if
        m.get("dram__throughput.avg.pct_of_peak_sustained_elapsed", ("", ""))[0] else None,
        "registers_per_thread": int(float(m["launch__registers_per_thread"][0])),
        "grid": m.get("launch__grid_size", ("?", ""))[0], "block": m.get("launch__block_size", ("?", ""))[0],
        "unit": "ff_stream.cu", "unit_fingerprint": ffbuild.built_unit_fingerprint("ff_stream.cu"),
        "captured": time.strftime("%Y-%m-%d"), "report": rep.name,
        "how": "ncu --set full --clock-control none, one launch of the C2 range kernel (20000 frames 1024x128, packed 12-bit)",
    }
    (REPO / "profiles" / "range_kernel_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))
    if "--details" in sys.argv:
        dst = Path(sys.argv[sys.argv.index("--details") + 1])
        text = subprocess.run(["ncu", "-i", str(rep), "--page", "details"], stdout=subprocess.PIPE, text=True, check=True).stdout
        dst.write_text(text)
        print(dst)


if __name__ == "__main__":
    main()
