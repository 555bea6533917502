"""Build libflamefront.so (sm_100a) in-tree with nvcc.

    python -m high_speed_image_processing_b200.build [--force] [--verbose]

The shared library lands in ``high_speed_image_processing_b200/lib/`` (git-ignored, but it
travels with the repo snapshot to the GPU box).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libflamefront.so"
SOURCES = ["ff_stream.cu", "ff_detect.cu", "ff_range.cu", "ff_head.cu", "ff_frameops.cu", "ff_exchange.cu", "ff_api.cu",
           "ff_hostcopy.cpp"]       # .cpp: host-only, handed to g++ by nvcc
HEADERS = [CSRC / "ff_common.cuh", CSRC / "ff_detect_core.cuh", CSRC / "ff_internal.h",
           PKG_DIR.parent / "include" / "flamefront.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--cudart", "static",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libflamefront.so cannot be built")


def is_stale() -> bool:
    if not LIB_PATH.exists():
        return True
    built = LIB_PATH.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + HEADERS + [Path(__file__)]
    return any(d.stat().st_mtime > built for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB_PATH
    LIB_DIR.mkdir(parents=True, exist_ok=True)
    nvcc = find_nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = LIB_DIR / (Path(src).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("FF_NVCC_EXTRA", "").split(), "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for src, proc in procs:
        out, _ = proc.communicate()
        if verbose and out:
            print(out)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    tmp = LIB_PATH.with_suffix(".so.tmp")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "--cudart", "static",
            "-Xcompiler", "-fPIC", "-o", str(tmp), *objs]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
