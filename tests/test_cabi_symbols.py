"""The C-ABI shared library loads on a CPU-only box and exports exactly what
include/flamefront.h declares.  No compute calls here."""
import ctypes
import re
from pathlib import Path

import pytest

from high_speed_image_processing_b200 import _cabi

HEADER = Path(__file__).resolve().parent.parent / "include" / "flamefront.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(ff_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_binding_surface():
    assert declared_functions() == sorted(_cabi.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(str(_cabi.LIB_PATH))
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} missing from {_cabi.LIB_PATH.name}"


def test_load_and_version():
    lib = _cabi.load()
    assert lib.ff_abi_version() == _cabi.FF_ABI_VERSION
    assert lib.ff_strerror(0) == b"ok"
    assert b"invalid" in lib.ff_strerror(_cabi.FF_ERR_INVALID)
    n = ctypes.c_int64(0)
    tiles = ctypes.c_int(0)
    assert lib.ff_partial_len(20000, 128, 1024, 12, ctypes.byref(n), ctypes.byref(tiles)) == 0
    assert tiles.value == 16 * 8 and n.value == 20000 * 16 * 8      # 16 tiles x 8 warps
    assert lib.ff_partial_len(10, 5, 7, 12, ctypes.byref(n), ctypes.byref(tiles)) == 0 and tiles.value == 1
    assert lib.ff_partial_len(10, 5, 7, 10, ctypes.byref(n), ctypes.byref(tiles)) == _cabi.FF_ERR_UNSUPPORTED


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(_cabi.FlameFrontLibraryError):
        _cabi.load(tmp_path / "libflamefront.so")


def test_status_mapping():
    _cabi.check(0, "x")
    with pytest.raises(ValueError):
        _cabi.check(_cabi.FF_ERR_INVALID, "x")
    with pytest.raises(ValueError):
        _cabi.check(_cabi.FF_ERR_UNSUPPORTED, "x")
    with pytest.raises(_cabi.FlameFrontError):
        _cabi.check(_cabi.FF_ERR_NO_DEVICE, "x")


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from high_speed_image_processing_b200.engine import FlameFrontEngine, get_engine
    with pytest.raises(RuntimeError):
        FlameFrontEngine()
    with pytest.raises(RuntimeError):
        get_engine()
