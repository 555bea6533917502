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
