"""Drop-in for the reference's ``scripts/process_videos.py`` on the flame-front path.

Same configuration objects (``VideoSourceConfig`` - with the README's ``detection_method``
switch, README.md:53-63 -, ``FileCalibration``), same ``process_video_source(config,
processor)`` / ``main()`` entry points, same result tuples and text output; the frame loop of
the reference (scripts/process_videos.py:1441-1516) is replaced by the B200 engine:

    frame 0 -> ff_background -> host float64 statistics            (:1357-1370)
    packed frames -> ff_stream_frames (one HBM read per frame)     (:1455-1463, :397-399)
                  -> ff_detect (warp per profile) -> exit atomicMin (:1488-1494)
    [multi-GPU: all-reduce(min) of the exit frame, all-gather of positions]
    -> ff_truncate -> host: Time_s, Position_m, text files         (:1449-1452, :1512, :1561-1619)

The reference's frame-level names (``FlameDetector``, ``FlameDetectorConfig``,
``FlameDetectionResult``, ``subtract_scalar_background``, ``subtract_prior_frame``,
``three_frame_difference``, ``is_empty_frame``, ``write_results``) are importable from here as from
the reference script; they live in ``detector.py`` and run on the GPU frame by frame.

Matplotlib diagnostics (:783-1270) live in ``diagnostics.py`` (optional, off this path;
``process_video_source(..., diagnostics=True)``).
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .detector import (FlameDetectionResult, FlameDetector, FlameDetectorConfig, is_empty_frame,  # noqa: F401
                       subtract_prior_frame, subtract_scalar_background, three_frame_difference, write_results)
from .engine import ClipScalars, DetectionParams, DETECTION_METHODS
from .head import HeadParams, finish_head_track
from .photron import MPIVideoProcessor, PhotonVideo, SpatialCalibration, open_video

_REPO_ROOT = Path(__file__).resolve().parent.parent


# --------------------------------------------------------------------------------------
# configuration (reference: scripts/process_videos.py:49-161)
# --------------------------------------------------------------------------------------
@dataclass
class FileCalibration:
    """Calibration for a subset of files.  ``files`` entries are substrings of the file name
    or ``"a:b"`` ranges compared on the LAST integer in each name (bug-compatible with the
    reference, :84-101: ``run-5-_C001H001S0001.cihx`` compares as 1)."""
    calibration: float
    position_offset: float = 0.0
    files: List[str] = field(default_factory=list)

    def matches(self, filename: str) -> bool:
        for pattern in self.files:
            if ":" in pattern:
                lo, hi = pattern.split(":", 1)
                if self._matches_range(filename, lo.strip(), hi.strip()):
                    return True
            elif pattern in filename:
                return True
        return False

    def _matches_range(self, filename: str, start: str, end: str) -> bool:
        nums = [re.findall(r"\d+", s) for s in (start, end, filename)]
        if not all(nums):
            return False
        lo, hi, cur = (int(n[-1]) for n in nums)
        return lo <= cur <= hi


@dataclass
class VideoSourceConfig:
    name: str
    enabled: bool = False
    calibration: float = 1.0
    position_offset: float = 0.0
    trigger_frame: Optional[int] = None
    use_frame_diff: bool = True
    use_absolute_time: bool = True
    skip_frames: List[int] = field(default_factory=list)
    file_calibrations: List[FileCalibration] = field(default_factory=list)
    # "head" (default) = what the reference executes: FlameDetector (:220-663) with its FlameDetectorConfig
    # defaults (exit margin 15, :193), the 7-column velocity file and the pre-/post-DDT files (:1561-1619).
    # "threshold" / "gradient" / "half_maximum" = the README-era switch (README.md:55,62,132-141), absent from
    # the reference's HEAD dataclass: opt-in extensions with the 4-column file of README.md:90-97.
    detection_method: str = "head"
    exit_margin_px: int = 10                 # README-era methods only: README.md:146
    frame_diff_threshold: float = 5.0        # FlameDetectorConfig.frame_diff_threshold (:169)
    min_gradient_strength: float = 10.0      # FlameDetectorConfig.min_gradient_strength (:174)
    min_run_px: int = 1
    # detection_method = "head": the detector the reference executes at HEAD (FlameDetector,
    # :220-663) with its own FlameDetectorConfig defaults (exit margin 15, :193)
    head_params: HeadParams = field(default_factory=HeadParams)

    _video_path: Optional[str] = field(default=None, init=False, repr=False)
    _output_dir: Optional[str] = field(default=None, init=False, repr=False)

    @property
    def video_path(self) -> Optional[str]:
        return self._video_path

    @video_path.setter
    def video_path(self, path: Optional[str]) -> None:
        self._video_path = self._resolve_path(path)

    @property
    def output_dir(self) -> Optional[str]:
        return self._output_dir

    @output_dir.setter
    def output_dir(self, path: Optional[str]) -> None:
        self._output_dir = self._resolve_path(path)

    def _resolve_path(self, path: Optional[str]) -> Optional[str]:
        if path is None or os.path.isabs(path):
            return path
        return str((_REPO_ROOT / path).resolve())     # relative paths hang off the repo root (:136-143)

    def get_calibration_for_file(self, filename: str) -> Tuple[float, float]:
        for rule in self.file_calibrations:           # first matching rule wins (:158-161)
            if rule.matches(filename):
                return (rule.calibration, rule.position_offset)
        return (self.calibration, self.position_offset)

    def detection_params(self) -> DetectionParams:
        if self.detection_method not in DETECTION_METHODS:
            raise ValueError(f"unknown detection_method {self.detection_method!r}; "
                             f"options: {', '.join(DETECTION_METHODS + ('head',))}")
        return DetectionParams(method=self.detection_method, use_frame_diff=self.use_frame_diff,
                               frame_diff_threshold=self.frame_diff_threshold,
                               min_gradient_strength=self.min_gradient_strength,
                               min_run_px=self.min_run_px, exit_margin_px=self.exit_margin_px)


# --------------------------------------------------------------------------------------
# per-video processing
# --------------------------------------------------------------------------------------
ResultRow = Tuple[int, float, int, float, bool]     # (frame, time_s, px, pos_m, is_post_ddt)  (:1516)


@dataclass
class VideoResult:
    scalars: ClipScalars
    pos_px: np.ndarray            # int32[N]: >= 0 position, -1 none, -2 dropped by exit truncation
    nonempty_counts: np.ndarray   # int32[N]
    first_exit: Optional[int]     # first exit frame (not recorded), None if the flame never exits
    rows: List[ResultRow]
    empty_frames: int
    velocity_history: List[list] = field(default_factory=list)   # [frame, v_backward1, v_backward2, v_central]
    ddt_frame: Optional[int] = None
    stop: Optional[Tuple[str, int]] = None                       # HEAD mode: ("exit"|"velocity_drop", frame)

    @property
    def frames(self) -> List[int]:
        return [r[0] for r in self.rows]


def build_rows(video: PhotonVideo, pos_px: np.ndarray, calibration: float, offset: float,
               use_absolute_time: bool) -> List[ResultRow]:
    """Host-side scalars, with the reference's own expressions so they are bit-identical:
    time (:1449-1452 -> src/photron/video.py:220,240-241) and pos_m = px*cal + off (:1512)."""
    rows: List[ResultRow] = []
    for frame_idx in np.nonzero(pos_px >= 0)[0].tolist():
        px = int(pos_px[frame_idx])
        time_s = video.get_absolute_time(frame_idx) if use_absolute_time else video.get_time(frame_idx)
        rows.append((frame_idx, time_s, px, px * calibration + offset, False))
    return rows


def process_video(video: PhotonVideo, config: VideoSourceConfig, calibration: float, offset: float,
                  engine=None, exchange=None, residency: str = "auto") -> VideoResult:
    """Run the flame-front path on one recording.

    ``residency``: "host" (= "auto") streams chunks from host memory through ``ff_process_host``
    - the memory-mapped file via threaded pinned bounce buffers, or a ``pin_memory()``-staged
    recording in place; "device" uploads the (rank's) packed frames once and runs the
    device-resident kernels.
    ``exchange`` (a ``sharding.RangeExchange``) splits the clip into contiguous frame ranges
    across ranks; results are identical on every rank.
    """
    import torch
    from .engine import get_engine
    from ._cabi import FF_NO_EXIT

    eng = engine if engine is not None else get_engine()
    n, (h, w), bits = len(video), video.frame_shape, video.storage_bits
    if n == 0:
        raise ValueError("video has no frames")
    if config.detection_method == "head":
        return _process_video_head(video, config, calibration, offset, eng, exchange)
    params = config.detection_params()

    # per-clip scalars from frame 0 (every rank reads frame 0 itself: 1 frame of H2D)
    frame0 = eng.upload(video.raw_frames(0, 1))
    scalars, bg_dev = eng.clip_scalars(frame0, h, w, bits)

    skip_np = None
    if config.skip_frames:
        skip_np = np.zeros(n, dtype=np.uint8)
        for s in config.skip_frames:
            if 0 <= s < n:
                skip_np[s] = 1

    a, b = (0, n) if exchange is None else exchange.my_range(n)
    halo_idx = a - 1
    if skip_np is not None:
        while halo_idx >= 0 and skip_np[halo_idx]:
            halo_idx -= 1
    halo_np = video.raw_frames(halo_idx, halo_idx + 1) if halo_idx >= 0 else None

    multi = exchange is not None and exchange.size > 1
    if residency == "auto":
        # stream from host memory chunk by chunk: copies overlap the kernels, nothing is copied past
        # the exit frame, pinned recordings are DMA'd in place and memory-mapped (pageable) ones go
        # through ff_process_host's threaded bounce buffers
        residency = "host"
    if residency not in ("device", "host"):
        raise ValueError("residency must be 'auto', 'device' or 'host'")
    skip_range = None if skip_np is None else skip_np[a:b]

    if multi:
        # every rank handles its contiguous range; one exchange kernel per rank finishes the clip
        if exchange.engine is None:
            exchange.engine = eng
        blk = exchange.begin(n, eng.device)
        if b - a == 0:                  # more ranks than frames: an empty block, still part of the exchange
            exchange.acquire(blk)
            exchange.publish(blk)
        elif residency == "device":     # range uploaded once; the range kernel fills the block in place
            frames_dev = eng.upload(video.raw_frames(a, b))
            halo_dev = None if halo_np is None else eng.upload(halo_np)
            skip_dev = None if skip_range is None else torch.from_numpy(skip_range.copy()).to(eng.device)
            eng.process_range(frames_dev, b - a, h, w, bits, params, scalars, bg_dev, first_frame=a,
                              halo=halo_dev, skip=skip_dev, **exchange.range_kwargs(blk))
        elif blk.hooks is not None:     # range streamed from host memory straight into the block; the ranks share
            #                             exit frames while they stream and stop uploading behind the first one
            eng.process_host(video.raw_frames(a, b), b - a, h, w, bits, params, scalars, first_frame=a,
                             halo=halo_np, skip=skip_range, block=blk, hooks=blk.hooks, to_host=False)
        else:                           # gathered transport: results come back to the host first
            hres = eng.process_host(video.raw_frames(a, b), b - a, h, w, bits, params, scalars, first_frame=a,
                                    halo=halo_np, skip=skip_range)
            blk.pos[:b - a].copy_(torch.from_numpy(hres.pos))
            blk.counts[:b - a].copy_(torch.from_numpy(hres.counts))
            blk.first_exit.fill_(hres.first_exit)
        g = exchange.finish(blk)
        pos_np, cnt_np, first_exit = g.pos.cpu().numpy(), g.counts.cpu().numpy(), g.first_exit
        exchange.check()
    elif residency == "device":
        frames_dev = eng.upload(video.raw_frames(a, b))
        skip_dev = None if skip_np is None else torch.from_numpy(skip_np).to(eng.device)
        res = eng.process_range(frames_dev, n, h, w, bits, params, scalars, bg_dev, skip=skip_dev)
        pos_np, cnt_np = res.pos.cpu().numpy(), res.counts.cpu().numpy()
        first_exit = int(res.first_exit.cpu().item())
    else:
        hres = eng.process_host(video.raw_frames(a, b), b - a, h, w, bits, params, scalars, first_frame=a,
                                halo=halo_np, skip=skip_range)
        pos_np, cnt_np, first_exit = hres.pos, hres.counts, hres.first_exit

    kb_min = eng_min_signal(params, h * w)
    processed = np.ones(n, dtype=bool) if skip_np is None else skip_np == 0
    limit = n if first_exit == FF_NO_EXIT else first_exit
    empty_frames = int(np.count_nonzero((cnt_np[:limit] < kb_min) & processed[:limit]))
    rows = build_rows(video, pos_np, calibration, offset, config.use_absolute_time)
    return VideoResult(scalars, pos_np, cnt_np, None if first_exit == FF_NO_EXIT else first_exit, rows,
                       empty_frames)


def _head_chunk_frames(frame_bytes: int) -> int:
    """Frames per upload of the HEAD path: bounded device memory whatever the recording's size
    (FF_HEAD_CHUNK_MB, default 2048 MiB of stored bytes; the float64 line buffers add 16 W bytes per frame)."""
    mb = int(os.environ.get("FF_HEAD_CHUNK_MB", "2048"))
    return max(2, (mb << 20) // max(1, frame_bytes))


def _head_walk_range(eng, video: PhotonVideo, hp: HeadParams, calibration: float, a: int, b: int,
                     skip_np: Optional[np.ndarray], frame0_dev, state: Tuple[int, int, int],
                     stopped=None):
    """Frames [a, b) of a recording through the HEAD detector in chunks: upload (+ the halo frame, the
    latest non-skipped frame before the chunk), image pipeline, tracker with the state carried from
    chunk to chunk.  Nothing is uploaded past the chunk in which the walk ends - on the exit frame
    (:1488-1494), or where ``stopped(track, flags)`` (the host-side stop rules) says so.
    ``state`` = (exit frame or FF_NO_EXIT, last detection frame, last position) on entry;
    returns ``(track int32[b-a,5], flags uint8[b-a], state, pending scalars)``."""
    import torch
    from ._cabi import FF_NO_EXIT
    from .head import max_displacement_px
    n, (h, w), bits = len(video), video.frame_shape, video.storage_bits
    fb = video.raw_frames(0, 1).size
    track = np.full((b - a, 5), -1, dtype=np.int32)
    flags = np.zeros(b - a, dtype=np.uint8)
    pending = None
    step = _head_chunk_frames(fb)
    prev_dev, prev_hi = None, -1
    for c0 in range(a, b, step):
        if state[0] != FF_NO_EXIT:
            break
        c1 = min(b, c0 + step)
        halo_idx = c0 - 1
        if skip_np is not None:
            while halo_idx >= 0 and skip_np[halo_idx]:
                halo_idx -= 1
        if halo_idx < 0:
            halo_dev = None
        elif prev_dev is not None and halo_idx == prev_hi - 1:
            halo_dev = prev_dev[-fb:].clone()                    # last frame of the previous chunk, already here
        else:
            halo_dev = eng.upload(video.raw_frames(halo_idx, halo_idx + 1))
        prev_dev = None                                          # release the previous chunk before the next upload
        frames_dev = eng.upload(video.raw_frames(c0, c1))
        skip_dev = None if skip_np is None else torch.from_numpy(skip_np[c0:c1].copy()).to(eng.device)
        lines, flags_dev, pend = eng.head_lines(frames_dev, c1 - c0, h, w, bits, hp, frame0=frame0_dev,
                                                first_frame=c0, halo=halo_dev, skip=skip_dev)
        pending = pending or pend
        track_dev, stop_dev = eng.head_track_lines(lines, flags_dev, c0, w, hp,
                                                   max_displacement_px(video.frame_rate, calibration, hp),
                                                   (state[1], state[2]))
        track[c0 - a:c1 - a] = track_dev.cpu().numpy()
        flags[c0 - a:c1 - a] = flags_dev.cpu().numpy()
        state = tuple(int(v) for v in stop_dev.tolist())
        prev_dev, prev_hi = frames_dev, c1
        if stopped is not None and stopped(track[:c1 - a], flags[:c1 - a]):
            break
    return track, flags, state, pending


def _head_lines_range(eng, video: PhotonVideo, hp: HeadParams, a: int, b: int, skip_np: Optional[np.ndarray],
                      frame0_dev, exit_known=None):
    """The image part of the HEAD detector for frames [a, b) in chunks - upload (+ halo), streaming kernel,
    band kernel - keeping only the two float64 centre rows and the flag of every frame on the device
    (16 W + 1 bytes per frame).  Everything that costs PCIe or HBM bandwidth happens here, and none of it
    depends on the tracker's state.  ``exit_known()`` (polled between chunks) ends the uploads early.
    Returns ``([(c0, c1, lines, flags_dev), ...], pending scalars)``."""
    import torch
    (h, w), bits = video.frame_shape, video.storage_bits
    fb = video.raw_frames(0, 1).size
    step = _head_chunk_frames(fb)
    chunks, pending = [], None
    prev_dev, prev_hi = None, -1
    for c0 in range(a, b, step):
        if exit_known is not None and exit_known():
            break
        c1 = min(b, c0 + step)
        halo_idx = c0 - 1
        if skip_np is not None:
            while halo_idx >= 0 and skip_np[halo_idx]:
                halo_idx -= 1
        if halo_idx < 0:
            halo_dev = None
        elif prev_dev is not None and halo_idx == prev_hi - 1:
            halo_dev = prev_dev[-fb:].clone()
        else:
            halo_dev = eng.upload(video.raw_frames(halo_idx, halo_idx + 1))
        prev_dev = None
        frames_dev = eng.upload(video.raw_frames(c0, c1))
        skip_dev = None if skip_np is None else torch.from_numpy(skip_np[c0:c1].copy()).to(eng.device)
        lines, flags_dev, pend = eng.head_lines(frames_dev, c1 - c0, h, w, bits, hp, frame0=frame0_dev,
                                                first_frame=c0, halo=halo_dev, skip=skip_dev)
        pending = pending or pend
        chunks.append((c0, c1, lines, flags_dev))
        prev_dev, prev_hi = frames_dev, c1
    return chunks, pending


def _process_video_head(video: PhotonVideo, config: VideoSourceConfig, calibration: float, offset: float,
                        eng, exchange=None) -> VideoResult:
    """HEAD-parity mode: the reference loop :1441-1516 with ``FlameDetector.detect`` on the GPU.

    The recording goes through the device in bounded chunks (``_head_walk_range``), and nothing is
    uploaded past the chunk in which the walk stops.  With an ``exchange`` over several ranks the clip
    is split into contiguous frame ranges like the other methods: every rank FIRST runs the image pipeline
    on its range (``_head_lines_range`` - all of the HBM and PCIe traffic, in parallel on all ranks) and keeps
    the centre rows on the device; the search, sequential only through (last frame, last position), then
    runs range after range: a rank receives that 3-int state from its predecessor (requested before the
    image pipeline, waited for after it), walks its stored rows (``ff_head_track``, ~0.1 ms) and passes the
    state on; one all-gather of the int32 result rows finishes the clip.  A rank stops uploading as soon
    as its predecessor reports that the flame left before its range.  Results are identical on every rank."""
    import torch
    import torch.distributed as dist
    from ._cabi import FF_NO_EXIT

    hp = config.head_params
    n, (h, w), bits = len(video), video.frame_shape, video.storage_bits
    skip_np = None
    if config.skip_frames:
        skip_np = np.zeros(n, dtype=np.uint8)
        for s in config.skip_frames:
            if 0 <= s < n:
                skip_np[s] = 1
    time_of = video.get_absolute_time if config.use_absolute_time else video.get_time

    def finish(track, flags):
        return finish_head_track(track, flags, 0, w, video.frame_rate, calibration, offset, time_of, hp)

    frame0 = eng.upload(video.raw_frames(0, 1))
    start = (FF_NO_EXIT, -1, -1)
    multi = exchange is not None and exchange.size > 1
    summary = None
    if not multi:
        # the host-side stop rules (exit, velocity drop: :1486-1509) end the uploads as well; the bookkeeping of
        # the last look is the clip's (frames behind it were never walked: flags 0), so it is not repeated below
        last_look = []

        def stopped(t, f) -> bool:
            last_look[:] = [finish(t, f)]
            return last_look[0].stop is not None

        track, flags, _, pending = _head_walk_range(eng, video, hp, calibration, 0, n, skip_np, frame0, start,
                                                    stopped=stopped)
        summary = last_look[0] if last_look else None
    else:
        a, b = exchange.my_range(n)
        rank, size, group = exchange.rank, exchange.size, exchange.group
        peer = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
        from .head import max_displacement_px
        state = torch.tensor(start, dtype=torch.int32, device=eng.device)
        # The predecessor's tracker state is asked for FIRST and waited for LAST: every rank runs its image
        # pipeline (all PCIe and HBM traffic) at once; only the search (~0.1 ms per range) goes rank by rank.
        req = dist.irecv(state, src=peer(rank - 1), group=group) if rank > 0 else None

        def exit_before_my_range() -> bool:      # the predecessor is done and the flame has left already
            return req is not None and req.is_completed() and int(state[0].item()) != FF_NO_EXIT

        chunks, pending = _head_lines_range(eng, video, hp, a, b, skip_np, frame0, exit_before_my_range)
        if req is not None:
            req.wait()
        after = tuple(int(v) for v in state.tolist())
        track = np.full((b - a, 5), -1, dtype=np.int32)
        flags = np.zeros(b - a, dtype=np.uint8)
        maxdisp = max_displacement_px(video.frame_rate, calibration, hp)
        for c0, c1, lines, flags_dev in chunks:
            if after[0] != FF_NO_EXIT:
                break
            track_dev, stop_dev = eng.head_track_lines(lines, flags_dev, c0, w, hp, maxdisp, (after[1], after[2]))
            track[c0 - a:c1 - a] = track_dev.cpu().numpy()
            flags[c0 - a:c1 - a] = flags_dev.cpu().numpy()
            after = tuple(int(v) for v in stop_dev.tolist())
        del chunks
        if rank + 1 < size:
            dist.send(torch.tensor(after, dtype=torch.int32, device=eng.device), dst=peer(rank + 1), group=group)
        cap = exchange.block_cap(n)
        block = torch.full((cap, 6), -1, dtype=torch.int32, device=eng.device)      # 5 result columns + flag
        block[:, 5] = 0
        if b > a:
            block[:b - a, :5] = torch.from_numpy(track).to(eng.device)
            block[:b - a, 5] = torch.from_numpy(flags.astype(np.int32)).to(eng.device)
        gathered = torch.empty((size * cap, 6), dtype=torch.int32, device=eng.device)
        dist.all_gather_into_tensor(gathered, block, group=group)
        gathered = gathered.view(size, cap, 6).cpu().numpy()
        from .sharding import contiguous_range
        rows_all = np.concatenate([gathered[r, :hi - lo] for r in range(size)
                                   for lo, hi in [contiguous_range(n, r, size)]])
        track, flags = np.ascontiguousarray(rows_all[:, :5]), rows_all[:, 5].astype(np.uint8)
        if pending is None:              # this rank uploaded nothing: it still reports the clip's scalars
            _, _, pending = eng.head_lines(frame0, 1, h, w, bits, hp)
    scalars = eng.head_scalars(pending)
    if summary is None:
        summary = finish(track, flags)
    pos = np.full(n, -1, dtype=np.int32)
    for frame_idx, _, px, _, _ in summary.rows:
        pos[frame_idx] = px
    stop_frame = summary.stop[1] if summary.stop else n
    pos[stop_frame:] = -2
    processed = np.ones(n, dtype=bool) if skip_np is None else skip_np == 0
    empty_frames = int(np.count_nonzero((flags[:stop_frame] == 0) & processed[:stop_frame]))
    first_exit = summary.stop[1] if summary.stop and summary.stop[0] == "exit" else None
    return VideoResult(scalars, pos, np.zeros(n, dtype=np.int32), first_exit, list(summary.rows), empty_frames,
                       summary.velocity_history, summary.ddt_frame, summary.stop)


def eng_min_signal(params: DetectionParams, n_pixels: int) -> int:
    from .engine import min_signal_count
    return min_signal_count(n_pixels, params.min_signal_fraction)


# --------------------------------------------------------------------------------------
# output files
# --------------------------------------------------------------------------------------
def write_position_file(rows: Sequence[ResultRow], filepath) -> str:
    """The README's 4-column result file (README.md:90-97): space separated, ``.9f`` floats."""
    with open(filepath, "w") as fh:
        fh.write("#Frame Time_s Position_px Position_m\n")
        for frame_idx, t_s, px, p_m, _ in rows:
            fh.write(f"{frame_idx} {t_s:.9f} {px} {p_m:.9f}\n")
    return str(filepath)


_VELOCITY_HEADER = [
    "# Flame Position and Velocity Data",
    "#",
    "# Velocity Extraction Methods:",
    "#   Vel_Backward1: First-order backward difference",
    "#                  v_n = (x_n - x_{n-1}) / dt",
    "#                  Evaluates velocity at current time step",
    "#",
    "#   Vel_Backward2: Second-order backward difference",
    "#                  v_n = (3*x_n - 4*x_{n-1} + x_{n-2}) / (2*dt)",
    "#                  Higher accuracy at current time, requires 3 points",
    "#",
    "#   Vel_Central:   Second-order central difference",
    "#                  v_{n-1} = (x_n - x_{n-2}) / (2*dt)",
    "#                  Most accurate, but evaluates at PRIOR time step",
    "#",
]


def merge_velocities(rows: Sequence[ResultRow], velocity_history: Sequence[Sequence]) -> List[tuple]:
    """(frame, t, px, m, v1, v2, vc, is_post_ddt) - the merge at :1548-1555."""
    vel = {e[0]: (e[1], e[2], e[3]) for e in velocity_history}
    return [(f, t, px, m, *vel.get(f, (None, None, None)), post) for f, t, px, m, post in rows]


def write_velocity_file(data: Sequence[tuple], filepath) -> str:
    """The 7-column file of the reference's HEAD writer (:1561-1604), byte for byte: 15 comment
    lines, a column line, then space-separated rows with ``.9f`` time/position and ``.3f``
    velocities (empty field when a velocity is undefined)."""
    with open(filepath, "w") as fh:
        for line in _VELOCITY_HEADER:
            fh.write(line + "\n")
        fh.write(" ".join(["#Frame", "Time_s", "Position_px", "Position_m", "Vel_Backward1", "Vel_Backward2",
                           "Vel_Central"]) + "\n")
        for f_idx, t_s, px, p_m, v1, v2, vc in data:
            fields = [str(f_idx), f"{t_s:.9f}", str(px), f"{p_m:.9f}",
                      f"{v1:.3f}" if v1 is not None else "", f"{v2:.3f}" if v2 is not None else "",
                      f"{vc:.3f}" if vc is not None else ""]
            fh.write(" ".join(fields) + "\n")
    return str(filepath)


def write_head_outputs(res: "VideoResult", out_dir, stem: str) -> List[str]:
    """All / pre-DDT / post-DDT files exactly as :1606-1619."""
    merged = merge_velocities(res.rows, res.velocity_history)
    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    written = [write_velocity_file([m[:7] for m in merged], out_dir / f"{stem}-flame-position.txt")]
    pre = [m[:7] for m in merged if not m[7]]
    post = [m[:7] for m in merged if m[7]]
    if pre:
        written.append(write_velocity_file(pre, out_dir / f"{stem}-flame-position-pre-DDT.txt"))
    if post:
        written.append(write_velocity_file(post, out_dir / f"{stem}-flame-position-post-DDT.txt"))
    return written


# --------------------------------------------------------------------------------------
# driver (reference: scripts/process_videos.py:1277-1629, :1633-1703)
# --------------------------------------------------------------------------------------
def process_video_source(config: VideoSourceConfig, processor: Optional[MPIVideoProcessor] = None,
                         engine=None, exchange=None, verbose: bool = True,
                         diagnostics: bool = False) -> Dict[str, VideoResult]:
    """Process every ``*.cihx`` under ``config.video_path`` and write the result files.

    ``diagnostics=True`` additionally renders, on the root rank and after the timed work, the
    figures the reference driver draws for every recording (stacked sequences and one 12-panel
    figure per frame that reaches the detector, :1385-1418, :1474-1480) into
    ``<output_dir>/<stem>-frames/`` - needs Matplotlib (``diagnostics.py``)."""
    is_root = (processor is None or processor.is_root) and (exchange is None or exchange.rank == 0)
    say = print if (verbose and is_root) else (lambda *a, **k: None)
    say(f"\n{'=' * 60}\nProcessing: {config.name}\nVideo path: {config.video_path}")
    say(f"Detection method: {config.detection_method}")
    say(f"Default calibration: {config.calibration} m/pixel, offset: {config.position_offset} m")

    cihx_files = sorted(Path(config.video_path).rglob("*.cihx"))
    results: Dict[str, VideoResult] = {}
    if not cihx_files:
        say(f"No CIHX files found in {config.video_path}")
        return results

    for cihx_file in cihx_files:
        cal, off = config.get_calibration_for_file(cihx_file.name)
        say(f"\nLoading: {cihx_file.name}\n  Using calibration: {cal} m/pixel, offset: {off} m")
        video = open_video(str(cihx_file), trigger_frame=config.trigger_frame,
                           calibration=SpatialCalibration(scale=cal, units="m"))
        try:
            say(f"  Frames: {len(video)}  rate: {video.frame_rate} fps  shape: {video.frame_shape}")
            res = process_video(video, config, cal, off, engine=engine, exchange=exchange)
            sc = res.scalars
            say(f"  Background scalar: {sc.background}")
            say(f"  Centerline noise (from frame 0): mean={sc.centerline_mean:.1f}, "
                f"std={sc.centerline_std:.1f}, max={sc.centerline_max:.1f}")
            say(f"  Centerline flame threshold: {sc.flame_threshold:.1f}")
            if res.first_exit is not None:
                say(f"  Wave exited domain at frame {res.first_exit} (not recorded)")
            say(f"  Skipped {res.empty_frames} empty/noise-only frames; {len(res.rows)} detections")
            if is_root and res.rows and config.output_dir:
                out_dir = Path(config.output_dir)
                out_dir.mkdir(parents=True, exist_ok=True)
                if config.detection_method == "head":
                    for path in write_head_outputs(res, out_dir, cihx_file.stem):
                        say(f"  Results: {path}")
                    if res.ddt_frame is not None:
                        say(f"  DDT detected at frame {res.ddt_frame}")
                else:
                    path = write_position_file(res.rows, out_dir / f"{cihx_file.stem}-flame-position.txt")
                    say(f"  All results: {path} ({len(res.rows)} points)")
            if diagnostics and is_root and config.output_dir:
                from .diagnostics import render_video_diagnostics
                n_fig = render_video_diagnostics(video, config.name, cihx_file.stem, cal,
                                                 Path(config.output_dir) / f"{cihx_file.stem}-frames",
                                                 skip_frames=config.skip_frames, engine=engine)
                say(f"  Diagnostics: {n_fig} frame figures in {cihx_file.stem}-frames/")
            results[cihx_file.name] = res
        finally:
            video.close()
    return results


def process_collection(collection, configs, engine=None, exchange=None, balance: bool = True,
                       per_video=None) -> Dict[int, "VideoResult"]:
    """BASELINE config 5: a ``VideoCollection`` sharded by WHOLE VIDEOS across ranks
    (the GPU analogue of ``MPIVideoProcessor.process_videos``, src/photron/parallel.py:173-208).

    ``configs`` is one ``VideoSourceConfig`` for all videos or a sequence with one per video
    (mixed methods / calibrations).  Each rank processes its videos end to end on its own GPU
    (no data-path collective); the per-video results (a few KB each) are all-gathered so every
    rank returns the full ``{video_index: VideoResult}`` map, ordered by index like the
    reference's gather + sort.  ``per_video(video, cfg, cal, off)`` overrides the per-video
    function (tests of the host logic pass a stub)."""
    import torch.distributed as dist
    from .sharding import assign_videos

    n = len(collection)
    cfgs = list(configs) if isinstance(configs, (list, tuple)) else [configs] * n
    if len(cfgs) != n:
        raise ValueError("configs must be a single VideoSourceConfig or one per video")
    rank = 0 if exchange is None else exchange.rank
    size = 1 if exchange is None else exchange.size
    weights = [len(v) * v.frame_shape[0] * v.frame_shape[1] for v in collection] if balance else None
    mine = assign_videos(n, rank, size, weights)
    run = per_video if per_video is not None else (
        lambda video, cfg, cal, off: process_video(video, cfg, cal, off, engine=engine, exchange=None))
    local = {}
    for vi in mine:
        video, cfg = collection[vi], cfgs[vi]
        cal, off = cfg.get_calibration_for_file(video.filepath.name)
        local[vi] = run(video, cfg, cal, off)
    if size == 1:
        return dict(sorted(local.items()))
    parts = [None] * size
    dist.all_gather_object(parts, local, group=exchange.group)
    merged = {}
    for part in parts:
        merged.update(part)
    return dict(sorted(merged.items()))


def default_configs(readme_methods: bool = False) -> List[VideoSourceConfig]:
    """The two sources hard-coded in the reference's main() (:1646-1685), processed like the reference does
    (``detection_method = "head"``).  ``readme_methods=True`` selects the README's per-camera methods instead
    (README.md:55,62: half_maximum for Nova, threshold for Mini)."""
    nova = VideoSourceConfig(name="Nova")
    nova.enabled = True
    nova.detection_method = "half_maximum" if readme_methods else "head"
    nova.video_path = "./Nova-Video-Files"
    nova.output_dir = "./Processed-Photos/Nova-Output"
    nova.file_calibrations = [
        FileCalibration(calibration=0.000833333, position_offset=1.0159, files=["run-1-"]),
        FileCalibration(calibration=0.000833333, position_offset=1.197565, files=["run-2-"]),
        FileCalibration(calibration=0.000833333, position_offset=1.347567, files=["run-3-:run-10-"]),
    ]
    mini = VideoSourceConfig(name="Mini")
    mini.enabled = True
    mini.detection_method = "threshold" if readme_methods else "head"
    mini.video_path = "./Mini-Video-Files"
    mini.output_dir = "./Processed-Photos/Mini-Output"
    mini.file_calibrations = [
        FileCalibration(calibration=0.000869565, position_offset=0.050237, files=["run-1-:run-10-"]),
    ]
    return [nova, mini]


def main() -> None:
    """One process per GPU under torchrun (NCCL), or a single process on cuda:0."""
    import torch
    import torch.distributed as dist
    from .sharding import RangeExchange

    exchange = None
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
        from .engine import get_engine
        exchange = RangeExchange(engine=get_engine())
        if exchange.rank == 0:
            print(f"Running on {exchange.size} GPUs")
    for cfg in default_configs():
        if cfg.enabled and cfg.video_path and Path(cfg.video_path).exists():
            process_video_source(cfg, None, exchange=exchange)
    if exchange is not None:
        exchange.close()
        dist.barrier()
        dist.destroy_process_group()
    if exchange is None or exchange.rank == 0:
        print("\nProcessing complete!")


if __name__ == "__main__":
    main()
