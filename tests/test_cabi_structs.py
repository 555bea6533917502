"""The ctypes structures of the binding have the layout of the C structs in include/flamefront.h
(checked with the C compiler), and the summation order the prep kernel assumes is NumPy's."""
import ctypes as C
import math
import shutil
import subprocess
import textwrap
from pathlib import Path

import numpy as np
import pytest

from high_speed_image_processing_b200 import _cabi

INCLUDE = Path(__file__).resolve().parent.parent / "include"


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_struct_layouts_match_the_header(tmp_path):
    pairs = [("ff_range_hooks", _cabi.RangeHooks), ("ff_range_args", _cabi.RangeArgs), ("ff_host_args", _cabi.HostArgs)]
    lines = []
    for cname, cls in pairs:
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for field, _ in cls._fields_:
            lines.append(f'printf("{cname}.{field} %zu\\n", offsetof({cname}, {field}));')
    src = tmp_path / "layout.c"
    src.write_text(textwrap.dedent("""
        #include <stdio.h>
        #include <stddef.h>
        #include "flamefront.h"
        int main(void) {
        %s
        return 0; }
        """) % "\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", f"-I{INCLUDE}", str(src), "-o", str(exe)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in pairs:
        assert int(got[cname]) == C.sizeof(cls)
        for field, _ in cls._fields_:
            assert int(got[f"{cname}.{field}"]) == getattr(cls, field).offset, f"{cname}.{field}"


def pairwise_sum(a):
    """The summation order prep_kernel implements (csrc/ff_detect.cu): NumPy's pairwise add.reduce."""
    n = len(a)
    if n < 8:
        r = 0.0
        for v in a:
            r += v
        return r
    if n <= 128:
        r = [a[j] for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] += a[i + j]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res += a[i]
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return pairwise_sum(a[:n2]) + pairwise_sum(a[n2:])


def test_prep_kernel_summation_order_is_this_numpys():
    """If this fails, NumPy changed how add.reduce sums float64 and the device-side flame threshold may
    differ from NumPy's in the last bit (the engine would raise at run time); see engine.PendingScalars."""
    rng = np.random.default_rng(0)
    for n in list(range(1, 300)) + [511, 512, 640, 1000, 1024, 1280, 2048, 4096]:
        line = np.clip(np.rint(rng.normal(40, 9, n)), 0, 4095).astype(np.uint16).astype(np.float64)
        xs = [float(v) for v in line]
        mean = pairwise_sum(xs) / n
        sq = [(v - mean) * (v - mean) for v in xs]
        std = math.sqrt(pairwise_sum(sq) / n)
        assert (mean, std) == (float(np.mean(line)), float(np.std(line))), n
