"""The stamp that ties measurements to a build: `build.source_fingerprint()` / `unit_fingerprint()` (sources + headers +
nvcc flags; nvcc's output itself is not byte-reproducible) are recorded next to the library by `build()`, and `bench.py`
quotes the ncu traffic capture of `profiles/range_kernel_traffic.json` only for a library whose ff_stream.cu was built
from the same sources."""
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from high_speed_image_processing_b200 import build as ffbuild  # noqa: E402


def test_library_in_tree_was_built_from_the_sources_in_the_tree():
    fp = ffbuild.source_fingerprint()
    assert len(fp) == 64 and fp == ffbuild.source_fingerprint()
    assert ffbuild.LIB_PATH.exists(), "build() runs before the tests"
    assert ffbuild.built_fingerprint() == fp
    for src in ffbuild.SOURCES:
        assert ffbuild.built_unit_fingerprint(src) == ffbuild.unit_fingerprint(src)
    assert ffbuild.unit_fingerprint("ff_stream.cu") != ffbuild.unit_fingerprint("ff_head.cu")
    assert not ffbuild.is_stale()


def test_traffic_is_quoted_only_for_the_build_it_was_captured_on(monkeypatch):
    import bench
    rec = json.loads((REPO / "profiles" / "range_kernel_traffic.json").read_text())
    traffic, note = bench.measured_traffic()
    if rec.get("unit_fingerprint") == ffbuild.unit_fingerprint(rec.get("unit", "ff_stream.cu")):
        assert traffic == rec["dram_bytes_per_launch"] and "same sources" in note
    else:
        assert traffic is None and "other sources" in note
    # any change to what the compiler reads - here a flag - invalidates the stamp
    monkeypatch.setattr(ffbuild, "NVCC_FLAGS", ffbuild.NVCC_FLAGS + ["-DSOMETHING_ELSE"])
    assert ffbuild.source_fingerprint() != ffbuild.built_fingerprint()
    assert ffbuild.unit_fingerprint("ff_stream.cu") != ffbuild.built_unit_fingerprint("ff_stream.cu")
    traffic, note = bench.measured_traffic()
    assert traffic is None and "not built from the sources" in note
