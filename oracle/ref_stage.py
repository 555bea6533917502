"""Stage the reference's own Python sources next to the oracle so that they can be TIMED on the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``/root/reference`` exists only in the build container.  ``stage()`` - called by
``__graft_entry__.build()`` there - copies ``scripts/process_videos.py`` and ``src/photron/*.py``
unmodified into ``oracle/_ref/reference/`` (git-ignored, so the history stays free of reference code;
not gpurun-ignored, so the directory travels to the GPU box like a built ``.so``) and records their
SHA-256.  ``load()`` imports that copy with the two stubs ``oracle/make_golden.py`` uses (``pyMRAW`` ->
the oracle's restated decoder, ``matplotlib`` -> empty module) and returns the reference's
``process_videos`` module, or ``None`` when nothing was staged.

``run_head_loop`` is the reference's frame loop (scripts/process_videos.py:1441-1516) with plotting and
printing removed, every numeric step a call into the reference's own functions - what ``bench.py`` times
as ``cpu_baseline.kind == "reference"`` beside the GPU HEAD detector.
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
import types
from pathlib import Path
from typing import Callable, Optional

import numpy as np

HERE = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference")
STAGE = HERE / "_ref" / "reference"
MANIFEST = HERE / "_ref" / "STAGED.json"
FILES = ["scripts/process_videos.py", "src/__init__.py", "src/photron/__init__.py", "src/photron/collection.py",
         "src/photron/metadata.py", "src/photron/parallel.py", "src/photron/video.py"]


def stage(force: bool = False) -> Optional[Path]:
    """Copy the reference's sources into ``oracle/_ref/reference`` (no-op without ``/root/reference``)."""
    if not REF_SRC.exists():
        return STAGE if MANIFEST.exists() else None
    manifest = {}
    for rel in FILES:
        src, dst = REF_SRC / rel, STAGE / rel
        data = src.read_bytes()
        manifest[rel] = hashlib.sha256(data).hexdigest()
        if force or not dst.exists() or dst.read_bytes() != data:
            dst.parent.mkdir(parents=True, exist_ok=True)
            shutil.copyfile(src, dst)
    MANIFEST.write_text(json.dumps({"source": str(REF_SRC), "sha256": manifest}, indent=1))
    return STAGE


def staged() -> bool:
    return MANIFEST.exists() and all((STAGE / rel).exists() for rel in FILES)


def _install_stubs() -> None:
    from oracle import flame_oracle as fo
    if "pyMRAW" not in sys.modules:
        pm = types.ModuleType("pyMRAW")

        def load_video(path):      # decode layout restated in oracle/flame_oracle.py (parity unpinned there)
            from high_speed_image_processing_b200 import mraw as our_mraw
            info = our_mraw.get_cih(path)
            raw = np.fromfile(str(Path(path).with_suffix(".mraw")), dtype=np.uint8)
            images = fo.frames_from_bytes(raw, int(info["Total Frame"]), int(info["Image Height"]),
                                          int(info["Image Width"]), int(info["Color Bit"]))
            return images, info

        pm.load_video = load_video
        sys.modules["pyMRAW"] = pm
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


_module = None


def load():
    """The reference's ``process_videos`` module imported from the staged copy, or ``None``."""
    global _module
    if _module is not None:
        return _module
    if not staged():
        return None
    want = json.loads(MANIFEST.read_text())["sha256"]
    for rel in FILES:
        if hashlib.sha256((STAGE / rel).read_bytes()).hexdigest() != want[rel]:
            raise RuntimeError(f"oracle/_ref/reference/{rel} differs from what was staged")
    _install_stubs()
    for path in (str(STAGE), str(STAGE / "scripts")):       # the script itself does `from src import ...`
        if path not in sys.path:
            sys.path.insert(0, path)
    import importlib
    _module = importlib.import_module("process_videos")
    if not str(Path(_module.__file__).resolve()).startswith(str(STAGE.resolve())):
        raise RuntimeError(f"imported {_module.__file__}, not the staged reference")
    return _module


def run_head_loop(pv, frames: np.ndarray, frame_rate: float, calibration: float, offset: float,
                  time_of: Callable[[int], float], first_index: int = 0, frame0: Optional[np.ndarray] = None,
                  exit_margin: int = 15):
    """scripts/process_videos.py:1441-1516 on ``frames[N,H,W]`` (decoded, as ``video[i]`` returns them):
    background scalar from frame 0 (:1357-1358), subtract / empty test / ``FlameDetector.detect`` per frame,
    the two stop rules, result tuples.  Returns ``(rows, stop)`` with rows as lists
    ``[frame, time_s, px, pos_m, is_post_ddt]`` (what ``oracle.head_oracle.run_head`` returns, too)."""
    f0 = frames[0] if frame0 is None else frame0
    width = frames.shape[2]
    background_scalar = float(np.max(f0))
    cfg = pv.FlameDetectorConfig(gaussian_sigma=1.5, morphology_kernel_size=3, max_velocity_change_m_s=200.0)
    det = pv.FlameDetector(config=cfg, frame_rate=frame_rate, calibration_m_per_px=calibration)
    rows, stop = [], None
    for i in range(frames.shape[0]):
        frame_idx = first_index + i
        frame = frames[i]
        time_s = time_of(frame_idx)
        sub = pv.subtract_scalar_background(frame, background_scalar)
        noise_thresh = max(10.0, background_scalar * 0.5)
        if pv.is_empty_frame(sub, noise_threshold=noise_thresh, min_signal_fraction=0.0005):
            det._prior_frame = sub.copy()
            continue
        r = det.detect(frame=frame, frame_idx=frame_idx, background_scalar=background_scalar)
        pos = r.final_position
        velocity = det.last_velocity
        if pos is not None and pos >= width - exit_margin:
            det.clear_last_central_difference()
            stop = ("exit", frame_idx)
            break
        vh = det.get_velocity_history()
        if velocity is not None and len(vh) >= 2:
            prev_v1 = vh[-2][1]
            if prev_v1 is not None and prev_v1 > 100 and (prev_v1 - velocity) / prev_v1 > 0.5:
                det.clear_last_central_difference()
                stop = ("velocity_drop", frame_idx)
                break
        if pos is not None:
            post = det.ddt_detected and frame_idx >= det.ddt_frame
            rows.append([frame_idx, time_s, int(pos), pos * calibration + offset, bool(post)])
    return rows, stop


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
