"""ctypes binding of libflamefront.so (C-ABI declared in include/flamefront.h).

There is deliberately no fallback: if the shared library is missing or does not load, every
entry point raises ``FlameFrontLibraryError``.  The hot path has no CPU implementation in
this package.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "lib" / "libflamefront.so"

FF_OK = 0
FF_ERR_INVALID = -1
FF_ERR_UNSUPPORTED = -2
FF_ERR_CUDA = -3
FF_ERR_NO_DEVICE = -4
FF_ERR_ALIGNMENT = -5

FF_METHOD = {"threshold": 0, "gradient": 1, "half_maximum": 2}
FF_DIFF_NONE, FF_DIFF_U16, FF_DIFF_F32, FF_DIFF_F64 = 0, 1, 2, 3
FF_POS_NONE = -1
FF_POS_DROPPED = -2
FF_NO_EXIT = 2147483647
FF_PX_U8, FF_PX_U16, FF_PX_F64 = 0, 1, 3
FF_HOOK_WAIT, FF_HOOK_PUBLISH = 1, 2
FF_ABI_VERSION = 6


class FlameFrontLibraryError(RuntimeError):
    """libflamefront.so is missing, stale or failed to load."""


class FlameFrontError(RuntimeError):
    """A C-ABI call returned a non-zero status."""

    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64
_int = C.c_int


class RangeHooks(C.Structure):
    """``ff_range_hooks``: ties one rank's range to its peers (filled by ``ff_exchange_begin``)."""
    _fields_ = [("table_dev", _vp), ("epoch", _i32), ("world", _i32), ("rank", _i32), ("flags", _i32),
                ("spin_limit", _i64), ("exit_word_dev", _vp)]


class RangeArgs(C.Structure):
    """``ff_range_args`` of ``ff_process_range``."""
    _fields_ = [("frames_dev", _vp), ("halo_dev", _vp), ("frame0_dev", _vp),
                ("n_frames", _i64), ("first_frame", _i64),
                ("height", _i32), ("width", _i32), ("bits", _i32),
                ("method", _i32), ("use_frame_diff", _i32), ("min_run_px", _i32), ("exit_margin_px", _i32),
                ("diff_thr", _i32), ("grad2_bound", _i32), ("empty_thr", _i32), ("threshold_floor", _i32),
                ("min_signal_count", _i64),
                ("skip_dev", _vp), ("scalars_dev", _vp), ("centerline_dev", _vp),
                ("pos_out_dev", _vp), ("count_out_dev", _vp), ("first_exit_dev", _vp),
                ("init_first_exit", _i32), ("truncate", _i32),
                ("diff_out_dev", _vp), ("diff_dtype", _i32), ("reserved", _i32),
                ("decoded_out_dev", _vp), ("profile_out_dev", _vp), ("partial_dev", _vp),
                ("workspace_dev", _vp), ("hooks", C.POINTER(RangeHooks))]


class HostArgs(C.Structure):
    """``ff_host_args`` of ``ff_process_host_range``."""
    _fields_ = [("frames_host", _vp), ("halo_host", _vp), ("n_frames", _i64), ("first_frame", _i64),
                ("height", _i32), ("width", _i32), ("bits", _i32), ("bg", _i32), ("empty_thr", _i32),
                ("method", _i32), ("use_frame_diff", _i32), ("diff_thr", _i32), ("threshold_floor", _i32),
                ("grad2_bound", _i32), ("min_run_px", _i32), ("exit_margin_px", _i32),
                ("min_signal_count", _i64), ("skip_host", _vp),
                ("pos_out_host", _vp), ("count_out_host", _vp),
                ("pos_block_dev", _vp), ("count_block_dev", _vp), ("first_exit_block_dev", _vp),
                ("hooks", C.POINTER(RangeHooks)),
                ("frames_done_out", C.POINTER(_i64)), ("first_exit_out", C.POINTER(_i32)),
                ("bytes_uploaded_out", C.POINTER(_i64))]


# name -> (restype, argtypes); must list every function declared in include/flamefront.h
SIGNATURES = {
    "ff_abi_version": (_int, []),
    "ff_strerror": (C.c_char_p, [_int]),
    "ff_last_cuda_error": (C.c_char_p, []),
    "ff_device_count": (_int, [C.POINTER(_int)]),
    "ff_device_sm_count": (_int, [_int, C.POINTER(_int)]),
    "ff_partial_len": (_int, [_i64, _int, _int, _int, C.POINTER(_i64), C.POINTER(_int)]),
    "ff_unpack": (_int, [_vp, _vp, _i64, _int, _int, _int, _vp]),
    "ff_background": (_int, [_vp, _int, _int, _int, _vp, _vp, _vp]),
    "ff_stream_frames": (_int, [_vp, _vp, _i64, _int, _int, _int, _vp, _i32, _i32, _vp, _vp, _vp, _int, _vp, _vp]),
    "ff_detect": (_int, [_vp, _vp, _i64, _i64, _int, _int, _int, _vp, _vp, _i64, _int, _int, _i32, _i32, _i32,
                         _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ff_truncate": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "ff_process_range_plan": (_int, [_i64, _int, _int, _int, _int, _int, _int, C.POINTER(_i64), C.POINTER(_i64),
                                     C.POINTER(_int)]),
    "ff_process_range": (_int, [C.POINTER(RangeArgs), _vp]),
    "ff_range_block_len": (_int, [_i64, C.POINTER(_i64)]),
    "ff_merge_ranges": (_int, [_vp, _int, _i64, _i64, _vp, _vp, _vp, _vp]),
    "ff_exchange_create": (_int, [_int, _int, _int, _i64, C.POINTER(_vp)]),
    "ff_exchange_handle_bytes": (_int, []),
    "ff_exchange_get_handle": (_int, [_vp, _vp]),
    "ff_exchange_open_peers": (_int, [_vp, _vp]),
    "ff_exchange_begin": (_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(RangeHooks)]),
    "ff_exchange_acquire": (_int, [_vp, _vp]),
    "ff_exchange_publish": (_int, [_vp, _vp]),
    "ff_exchange_finish": (_int, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "ff_exchange_status": (_int, [_vp, C.POINTER(_i32), _vp]),
    "ff_exchange_destroy": (_int, [_vp]),
    "ff_head_lines": (_int, [_vp, _vp, _i64, _int, _int, _int, _vp, _vp, _i64, _i32, _int, C.POINTER(C.c_double), _int,
                             _vp, _vp, _vp, _vp, _vp]),
    "ff_head_track_scratch_len": (_int, [_i64, C.POINTER(_i64)]),
    "ff_head_track": (_int, [_vp, _vp, _i64, _i64, _int, _i32, _i32, _i32, C.c_double, C.c_double, _i32, _i32, _i32,
                             _vp, _vp, _vp, _vp]),
    "ff_frame_subtract_background": (_int, [_vp, _int, _i64, C.c_double, _vp, _vp]),
    "ff_frame_difference": (_int, [_vp, _vp, _int, _i64, C.c_double, _vp, _vp]),
    "ff_frame_three_difference": (_int, [_vp, _vp, _vp, _int, _i64, C.c_double, _vp, _vp]),
    "ff_frame_count_above": (_int, [_vp, _int, _i64, C.c_double, _vp, _vp]),
    "ff_head_images": (_int, [_vp, _vp, _i64, _int, _int, _int, _i32, _i32, _i32, _int, C.POINTER(C.c_double), _int,
                              _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ff_host_ctx_create": (_int, [_int, _i64, C.POINTER(_vp)]),
    "ff_host_ctx_destroy": (_int, [_vp]),
    "ff_host_ctx_set_copy_threads": (_int, [_vp, _int]),
    "ff_host_upload": (_int, [_vp, _vp, _vp, _i64]),
    "ff_process_host": (_int, [_vp, _vp, _vp, _i64, _i64, _int, _int, _int, _i32, _i32, _i64, _int, _int, _i32,
                               _i32, _i32, _i32, _i32, _vp, _vp, _vp, C.POINTER(_i64), C.POINTER(_i32)]),
    "ff_process_host_range": (_int, [_vp, C.POINTER(HostArgs)]),
}

_lib: Optional[C.CDLL] = None


def load(path: Optional[Path] = None) -> C.CDLL:
    """Load (once) and return the library with argtypes/restypes set."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path is not None else Path(os.environ.get("FF_LIB_PATH") or LIB_PATH)   # A/B knob for tools/
    if not p.exists():
        raise FlameFrontLibraryError(
            f"{p} not found - build it with `python -m high_speed_image_processing_b200.build` "
            "(needs nvcc; there is no CPU fallback for the flame-front path)")
    try:
        lib = C.CDLL(str(p))
    except OSError as exc:  # pragma: no cover - depends on the machine
        raise FlameFrontLibraryError(f"could not load {p}: {exc}") from exc
    for name, (restype, argtypes) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise FlameFrontLibraryError(f"{p} does not export {name}; rebuild it") from exc
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.ff_abi_version() != FF_ABI_VERSION:
        raise FlameFrontLibraryError(
            f"{p} has ABI version {lib.ff_abi_version()}, binding expects {FF_ABI_VERSION}; rebuild it")
    if path is None:
        _lib = lib
    return lib


def check(status: int, what: str) -> None:
    """Map a status code onto the exception the reference API would raise."""
    if status == FF_OK:
        return
    lib = load()
    msg = lib.ff_strerror(status).decode()
    if status == FF_ERR_CUDA:
        msg += ": " + lib.ff_last_cuda_error().decode()
    text = f"{what}: {msg}"
    if status in (FF_ERR_INVALID, FF_ERR_UNSUPPORTED, FF_ERR_ALIGNMENT):
        raise ValueError(text)
    raise FlameFrontError(status, text)
