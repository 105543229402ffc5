"""CPU tests of the multi-GPU host logic: block planning, halo handling, ordered gather over gloo (world_size 2).
The filter run on each rank is the oracle's NLMeans restatement on tiny frames: sharded == unsharded, bit for bit."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from handbrake_b200 import sharding, synth
from oracle_port import OraclePort

REPO = Path(__file__).resolve().parent.parent


def test_plan_blocks_covers_clip_once():
    for n, world, block in ((10, 2, 3), (7, 4, 2), (16, 8, 2), (5, 3, 8), (0, 2, 4)):
        blocks = sharding.plan_blocks(n, world, block, halo_before=1, halo_after=2)
        owned = [t for b in blocks for t in range(b.start, b.stop)]
        assert owned == list(range(n))
        for b in blocks:
            assert b.rank == b.index % world
            assert b.load_start == max(0, b.start - 1) and b.load_stop == min(n, b.stop + 2)


def test_sharded_nlmeans_equals_unsharded_single_process():
    port = OraclePort()
    w, h, n, nf = 48, 32, 9, 3
    clip = synth.progressive_clip(synth.PIX_FMT_YUV420P, w, h, n, seed=8)
    params = [dict(strength=6, nframes=nf)] * 3
    whole = port.nlmeans_clip(clip, w, h, 8, params)
    blocks = sharding.plan_blocks(n, 2, 2, halo_after=nf - 1)
    parts = {}
    for rank in range(2):
        parts.update(sharding.run_rank(blocks, rank, lambda a, b: clip[a:b], lambda fr: port.nlmeans_clip(fr, w, h, 8, params)))
    got = np.concatenate([parts[b.index] for b in blocks])
    assert np.array_equal(got, whole)


WORKER = r'''
import os, sys
sys.path.insert(0, {repo!r}); sys.path.insert(0, {tests!r})
import numpy as np, torch, torch.distributed as dist
from handbrake_b200 import sharding, synth
from oracle_port import OraclePort
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
port = OraclePort()
w, h, n, nf = 48, 32, 7, 2
clip = synth.progressive_clip(synth.PIX_FMT_YUV420P, w, h, n, seed=9)
params = [dict(strength=6, nframes=nf)] * 3
blocks = sharding.plan_blocks(n, world, 2, halo_after=nf - 1)
res = sharding.run_rank(blocks, rank, lambda a, b: clip[a:b], lambda fr: port.nlmeans_clip(fr, w, h, 8, params))
out = sharding.ordered_gather(res, blocks, rank, world, dist=dist)
if rank == 0:
    whole = port.nlmeans_clip(clip, w, h, 8, params)
    assert out.shape == whole.shape and np.array_equal(out, whole), "sharded != unsharded"
    print("GATHER_OK", out.shape[0])
else:
    assert out is None
dist.barrier()
dist.destroy_process_group()
'''


def test_ordered_gather_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(repo=str(REPO), tests=str(REPO / "tests")))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                       capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GATHER_OK 7" in r.stdout
