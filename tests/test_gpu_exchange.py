"""The multi-GPU finishing step (csrc/ff_exchange.cu) on real devices.

* ff_merge_ranges on one GPU, fed with hand-built range blocks for several world sizes and
  ragged totals, against the NumPy statement of the same step;
* a 2-rank torchrun run (needs >= 2 GPUs, skipped otherwise) of the whole range-sharded path
  with BOTH block transports - peer memory over NVLink (CUDA IPC + epoch flags) and one NCCL
  all-gather - compared with the oracle's serial answer.
"""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from high_speed_image_processing_b200._cabi import FF_NO_EXIT, FF_POS_DROPPED
from high_speed_image_processing_b200.sharding import BLOCK_HEADER, RangeExchange, contiguous_range

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent


def build_blocks(rng, world, total, cap, exits):
    blocks = np.zeros((world, BLOCK_HEADER + 2 * cap), dtype=np.int32)
    pos = rng.integers(-1, 1000, size=total, dtype=np.int32)
    cnt = rng.integers(0, 5000, size=total, dtype=np.int32)
    blocks[:, 1:BLOCK_HEADER] = 12345                 # header padding is ignored
    blocks[:, BLOCK_HEADER:] = -77                    # so is the tail beyond a rank's range
    for r in range(world):
        a, b = contiguous_range(total, r, world)
        blocks[r, 0] = exits[r]
        blocks[r, BLOCK_HEADER:BLOCK_HEADER + (b - a)] = pos[a:b]
        blocks[r, BLOCK_HEADER + cap:BLOCK_HEADER + cap + (b - a)] = cnt[a:b]
    return blocks, pos, cnt


@pytest.mark.parametrize("world,total,slack", [(1, 17, 0), (2, 61, 0), (3, 64, 5), (8, 160000, 0), (8, 5, 0),
                                               (5, 3, 2), (64, 1000, 1)])
def test_merge_ranges_kernel(engine, world, total, slack):
    rng = np.random.default_rng(world * 100003 + total)
    cap = max(1, -(-total // world)) + slack
    for case in range(3):
        exits = np.full(world, FF_NO_EXIT, dtype=np.int64)
        if case >= 1:                                  # one or several ranks saw an exit
            for r in rng.choice(world, size=min(world, case), replace=False):
                a, b = contiguous_range(total, int(r), world)
                if b > a:
                    exits[r] = rng.integers(a, b)
        blocks, pos, cnt = build_blocks(rng, world, total, cap, exits)
        fe = int(exits.min())
        want = pos.copy()
        want[min(fe, total):] = FF_POS_DROPPED
        g = torch.from_numpy(blocks.reshape(-1)).to(engine.device)
        pos_out = torch.full((total,), 999, dtype=torch.int32, device=engine.device)
        cnt_out = torch.full((total,), 999, dtype=torch.int32, device=engine.device)
        fe_out = torch.zeros(1, dtype=torch.int32, device=engine.device)
        engine.merge_ranges(g, world, cap, total, pos_out, cnt_out, fe_out)
        assert int(fe_out.item()) == fe
        assert np.array_equal(pos_out.cpu().numpy(), want)
        assert np.array_equal(cnt_out.cpu().numpy(), cnt)
        engine.merge_ranges(g, world, cap, total, pos_out, None, fe_out)      # counts are optional
        assert np.array_equal(pos_out.cpu().numpy(), want)


def test_merge_ranges_rejects_bad_arguments(engine):
    g = torch.zeros(2 * (BLOCK_HEADER + 8), dtype=torch.int32, device=engine.device)
    out = torch.zeros(8, dtype=torch.int32, device=engine.device)
    fe = torch.zeros(1, dtype=torch.int32, device=engine.device)
    with pytest.raises(ValueError):
        engine.merge_ranges(g, 2, 3, 8, out, None, fe)          # ceil(total / world) exceeds the block capacity
    with pytest.raises(ValueError):
        engine.merge_ranges(g, 65, 1, 8, out, None, fe)         # more ranks than the kernel's table
    with pytest.raises(ValueError):
        engine.merge_ranges(g.cpu(), 2, 4, 8, out, None, fe)    # not on the device


def test_single_process_exchange_runs_the_merge_kernel(engine):
    ex = RangeExchange(engine=engine)
    blk = ex.begin(6)
    assert blk.pos.device == engine.device and int(blk.first_exit.item()) == FF_NO_EXIT
    blk.pos[:6] = torch.arange(6, dtype=torch.int32, device=engine.device)
    blk.counts[:6] = 3
    blk.first_exit.fill_(4)
    before = engine.launches
    g = ex.finish(blk)
    assert engine.launches == before + 1
    assert g.first_exit == 4 and g.pos.tolist() == [0, 1, 2, 3, FF_POS_DROPPED, FF_POS_DROPPED]
    assert g.counts.tolist() == [3] * 6


@pytest.mark.parametrize("nproc", [2])
def test_two_rank_run_both_transports(tmp_path, nproc):
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    out = tmp_path / "exchange.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", "29547", str(REPO / "tests" / "check_exchange.py"),
           "--out", str(out), "--steps", "5"]
    env = dict(os.environ, PYTHONPATH=str(REPO))
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    report = json.loads(out.read_text())
    assert report["ok"] and set(report["transports"]) == {"peer", "gathered"}
