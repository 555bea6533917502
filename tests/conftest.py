import json
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The suite tests the built library (symbols, structs, stamp; every GPU test calls through it): build it here if
    # the tree holds none or a stale one and nvcc is around (it cross-compiles without a GPU).  No nvcc, no build -
    # the tests that need the library then fail loudly, as the product does.
    from high_speed_image_processing_b200 import build as ffbuild
    try:
        if ffbuild.is_stale():
            ffbuild.build()
    except RuntimeError as exc:
        print(f"[conftest] libflamefront.so not rebuilt: {exc}", file=sys.stderr)


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return json.loads((GOLDEN / "reference_golden.json").read_text())


@pytest.fixture(scope="session")
def clip_small():
    """The committed small recording: packed MRAW bytes + the CIHX file bytes."""
    z = np.load(GOLDEN / "clip_small.npz")
    return {"packed": z["packed"], "cihx": z["cihx"].tobytes()}


@pytest.fixture(scope="session")
def clip_small_profiles():
    return np.load(GOLDEN / "clip_small_profiles.npz")["diff_profiles"]


@pytest.fixture()
def clip_small_on_disk(tmp_path, clip_small, golden):
    stem = golden["clip_small"]["stem"]
    (tmp_path / f"{stem}.cihx").write_bytes(clip_small["cihx"])
    (tmp_path / f"{stem}.mraw").write_bytes(clip_small["packed"].tobytes())
    return tmp_path / f"{stem}.cihx"


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from high_speed_image_processing_b200.engine import FlameFrontEngine
    return FlameFrontEngine(0)


class _Recorder:
    """Stands in for a Matplotlib object: records every call made on it or its children."""

    def __init__(self, log, name):
        self._log, self._name = log, name

    def __getattr__(self, attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Recorder(self._log, f"{self._name}.{attr}")

    def __call__(self, *args, **kwargs):
        self._log.append((self._name, args, kwargs))
        if self._name.endswith("subplots"):
            import numpy as _np
            rows = args[0] if len(args) >= 1 else kwargs.get("nrows", 1)
            cols = args[1] if len(args) >= 2 else kwargs.get("ncols", 1)
            fig = _Recorder(self._log, "fig")
            if kwargs.get("squeeze", True) and rows == 1 and cols == 1:
                return fig, _Recorder(self._log, "ax")
            grid = _np.empty((rows, cols), dtype=object)
            for i in range(rows):
                for j in range(cols):
                    grid[i, j] = _Recorder(self._log, f"ax[{i},{j}]")
            return fig, grid
        if self._name.endswith("add_subplot"):
            return _Recorder(self._log, f"ax{sum(1 for n, _, _ in self._log if n.endswith('add_subplot')) - 1}")
        return _Recorder(self._log, self._name + "()")

    def __getitem__(self, key):
        return (self._name, key)


@pytest.fixture()
def fake_pyplot(monkeypatch):
    """A recording stand-in for matplotlib / matplotlib.pyplot (not installed in this image): the
    diagnostics are checked by WHAT they hand to Matplotlib, not by pixels.  Yields the call log."""
    import types
    log = []
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    rec = _Recorder(log, "plt")
    for name in ("figure", "subplots", "savefig", "close", "subplots_adjust"):
        setattr(plt, name, getattr(rec, name))
    mpl.pyplot = plt
    monkeypatch.setitem(sys.modules, "matplotlib", mpl)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    return log
