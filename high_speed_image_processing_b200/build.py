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


STAMP_PATH = LIB_DIR / "libflamefront.so.sources"


def unit_fingerprint(src: str) -> str:
    """SHA-256 over one translation unit: its source, the headers and the nvcc flags - what decides the SASS of the
    kernels defined in it.  Measurements of one kernel (profiles/range_kernel_traffic.json: range_kernel lives in
    ff_stream.cu) are stamped with this, so that a change to another kernel's file does not orphan them."""
    import hashlib
    h = hashlib.sha256()
    for f in [CSRC / src] + sorted(HEADERS, key=lambda q: q.name):
        h.update(f.name.encode() + b"\0" + f.read_bytes() + b"\0")
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def source_fingerprint() -> str:
    """SHA-256 over everything the library is compiled from (every unit_fingerprint).  nvcc's output is not
    byte-reproducible (the mangled names of anonymous namespaces differ from run to run), so a build is identified
    by this instead of a hash of the binary; build() records next to the .so which sources it was made from."""
    import hashlib
    return hashlib.sha256("".join(unit_fingerprint(s) for s in SOURCES).encode()).hexdigest()


def _stamp() -> dict:
    import json
    if not (STAMP_PATH.exists() and LIB_PATH.exists()):
        return {}
    try:
        return json.loads(STAMP_PATH.read_text())
    except ValueError:
        return {}


def built_fingerprint() -> str:
    """The fingerprint of the sources the library on disk was built from ('' if unknown)."""
    return _stamp().get("all", "")


def built_unit_fingerprint(src: str) -> str:
    return _stamp().get("units", {}).get(src, "")


def is_stale() -> bool:
    """True when the library on disk was not built from the sources in the tree: by the recorded fingerprint (copies
    of the tree - the snapshot on the GPU box - do not keep modification times), by modification time only for a
    library without a stamp."""
    if not LIB_PATH.exists():
        return True
    built = built_fingerprint()
    if built:
        return built != source_fingerprint()
    mtime = LIB_PATH.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + HEADERS + [Path(__file__)]
    return any(d.stat().st_mtime > mtime for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB_PATH
    LIB_DIR.mkdir(parents=True, exist_ok=True)
    import json
    # before compiling: what the compiler is about to read
    fingerprint = json.dumps({"all": source_fingerprint(), "units": {s: unit_fingerprint(s) for s in SOURCES}}, indent=1)
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
    STAMP_PATH.write_text(fingerprint + "\n")
    return LIB_PATH


SASS_MNEMONICS = [
    ("UBLKCP", r"\bUBLKCP", "1-D bulk async copy global -> shared (TMA, cp.async.bulk)"),
    ("SYNCS", r"\bSYNCS\.", "mbarrier operations"),
    ("DPX 16x2", r"\bVI(ADD)?MNMX3?\.[US]16x2", "16x2 SIMD min/max/add (VIMNMX3 / VIADDMNMX .U16x2/.S16x2)"),
    ("PRMT", r"\bPRMT\b", "byte permutes (12-bit field extraction)"),
    ("ATOMG/REDG", r"\b(ATOMG|REDG)\.", "global atomics / reductions"),
    ("ATOMS", r"\bATOMS\.", "shared-memory atomics"),
    ("sys-scope", r"\.STRONG\.SYS", "system-scope loads / stores / reductions (peer memory flags)"),
    ("ACQBULK/PDL", r"\bACQBULK\b", "griddepcontrol.wait (programmatic dependent launch)"),
    ("NANOSLEEP", r"\bNANOSLEEP", "back-off of idle detector warps / bounded spins"),
    ("DADD/DMUL/DFMA", r"\bD(ADD|MUL|FMA)\b", "float64 arithmetic (HEAD detector, device-side clip statistics)"),
    ("UTMALDG", r"\bUTMALDG", "tensor-map TMA loads (none expected: tiles are a flat byte stream)"),
    ("tensor core", r"\b(UTCMMA|HMMA|IMMA|UTCHMMA|TCGEN05)", "tensor-core instructions (none expected: no contraction on this path)"),
]


def sass_summary(out_path=None) -> str:
    """Per-kernel counts of the SASS mnemonics that show what the kernels are made of (cuobjdump -sass of the
    built library).  Written to profiles/r02_sass_summary.md by ``--verbose``."""
    import re
    dump = subprocess.run([str(Path(find_nvcc()).parent / "cuobjdump"), "-sass", str(LIB_PATH)], stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True, check=True).stdout
    try:
        filt = shutil.which("cu++filt") or shutil.which("c++filt")
    except Exception:
        filt = None
    kernels, name, arch = {}, None, set()
    for line in dump.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = {k: 0 for k, _, _ in SASS_MNEMONICS}
            kernels[name]["instructions"] = 0
            continue
        m = re.match(r"\s*arch = (sm_\w+)", line)
        if m:
            arch.add(m.group(1))
        if name is None or "/*" not in line or ";" not in line:
            continue
        kernels[name]["instructions"] += 1
        for key, pattern, _ in SASS_MNEMONICS:
            if re.search(pattern, line):
                kernels[name][key] += 1
    names = list(kernels)
    pretty = dict(zip(names, names))
    if filt and names:
        res = subprocess.run([filt], input="\n".join(names), stdout=subprocess.PIPE, text=True)
        if res.returncode == 0:
            pretty = dict(zip(names, res.stdout.splitlines()))
    cols = [k for k, _, _ in SASS_MNEMONICS]
    lines = ["# SASS summary of libflamefront.so (round 2)", "",
             f"`cuobjdump -sass` of `{LIB_PATH.relative_to(PKG_DIR.parent)}` (built from sources `{built_fingerprint()[:16]}`), "
             f"{len(kernels)} kernels, architectures: {', '.join(sorted(arch)) or '?'}.  Regenerate with "
             "`python -m high_speed_image_processing_b200.build --force --verbose`.", "",
             "| kernel | instr | " + " | ".join(cols) + " |", "|---|---|" + "---|" * len(cols)]
    def short(n):
        n = re.sub(r"\(anonymous namespace\)::|ff::|void ", "", n)
        return re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", n)[:90]
    for n in sorted(names, key=lambda k: short(pretty[k])):
        k = kernels[n]
        lines.append(f"| `{short(pretty[n])}` | {k['instructions']} | " + " | ".join(str(k[c]) for c in cols) + " |")
    tot = {c: sum(k[c] for k in kernels.values()) for c in cols}
    lines.append("| **all kernels** | " + str(sum(k["instructions"] for k in kernels.values())) + " | " +
                 " | ".join(f"**{tot[c]}**" for c in cols) + " |")
    lines += ["", "Columns:"] + [f"* **{k}** - {desc}" for k, _, desc in SASS_MNEMONICS]
    text = "\n".join(lines) + "\n"
    if out_path is not None:
        Path(out_path).write_text(text)
    return text


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
    if "--verbose" in sys.argv or "--sass" in sys.argv:
        out = PKG_DIR.parent / "profiles" / "r02_sass_summary.md"
        sass_summary(out)
        print(out)
