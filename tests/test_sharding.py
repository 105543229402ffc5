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


@pytest.mark.parametrize("prefilter", [1, 2, 1024 + 1])
def test_sharded_nlmeans_with_prefilter_needs_a_leading_halo(prefilter):
    """The stream's first frame alone takes the unfiltered plane as patch source: a block that restarts the filter must
    restart one frame early (nlmeans_halo) or its first owned frame repeats that start-of-stream rule."""
    port = OraclePort()
    w, h, n, nf = 48, 32, 9, 2
    clip = synth.progressive_clip(synth.PIX_FMT_YUV420P, w, h, n, seed=11)
    params = [dict(strength=8, nframes=nf, prefilter=prefilter)] * 3
    run = lambda fr: port.nlmeans_clip(fr, w, h, 8, params)
    whole = run(clip)
    before, after = sharding.nlmeans_halo(nf, [prefilter] * 3)
    assert (before, after) == (1, nf - 1)
    good = _sharded(sharding.plan_blocks(n, 2, 2, halo_before=before, halo_after=after), 2, lambda a, b: clip[a:b], run)
    assert np.array_equal(good, whole)
    naive = _sharded(sharding.plan_blocks(n, 2, 2, halo_before=0, halo_after=after), 2, lambda a, b: clip[a:b], run)
    assert not np.array_equal(naive, whole), "the leading halo is what makes the difference"
    assert sharding.nlmeans_halo(nf, [0, 0, 0]) == (0, nf - 1) and sharding.nlmeans_halo(1, [1, 1, 1]) == (0, 0)


def _sharded(blocks, world, clip_of, run, k=1):
    parts = {}
    for rank in range(world):
        parts.update(sharding.run_rank(blocks, rank, clip_of, run, outputs_per_frame=k))
    return np.concatenate([parts[b.index] for b in blocks])


@pytest.mark.parametrize("world,block", [(2, 3), (4, 2), (3, 1)])
def test_sharded_comb_detect_and_decomb_equal_unsharded(ref, world, block):
    """comb-detect and decomb (without EEDI2) are pure functions of (prev, cur, next) plus per-frame tags, so
    a one-frame halo either side reproduces the unsharded stream -- the first-frame / end-of-stream rules of the filters
    (duplicated neighbour, exhaustive check) fall on halo frames whose outputs are dropped.  Run with the reference's own
    filter objects; the CUDA objects have the same stream semantics (tests/test_*_gpu.py)."""
    from test_oracle import decomb_inputs
    w, h, depth = 96, 64, 8
    fmt = synth.PIX_FMT_YUV420P
    clip, flags, combed = decomb_inputs(depth, w, h, 8, seed=3)
    n = clip.shape[0]
    blocks = sharding.plan_blocks(n, world, block, halo_before=1, halo_after=1)
    window = {}

    def clip_of(a, b):
        window["range"] = (a, b)
        return clip[a:b]

    # comb-detect: the verdict travels with the frame
    def comb(frames):
        a, b = window["range"]
        r = ref.run("hb_filter_comb_detect", "mode=3:motion-thresh=1:spatial-thresh=1", frames, fmt, w, h, flags=flags[a:b])
        return np.asarray(r.combed, np.uint8).reshape(-1, 1)
    whole = np.asarray(ref.run("hb_filter_comb_detect", "mode=3:motion-thresh=1:spatial-thresh=1", clip, fmt, w, h, flags=flags).combed, np.uint8)
    assert len(set(whole.tolist())) > 1
    assert np.array_equal(_sharded(blocks, world, clip_of, comb).ravel(), whole)

    # decomb: one picture per frame (mode 7, selective mode 39 on the tags) and two per frame (bob, mode 23)
    for mode, k, tags in ((7, 1, None), (39, 1, combed), (23, 2, None)):
        def decomb(frames):
            a, b = window["range"]
            return ref.run("hb_filter_decomb", f"mode={mode}", frames, fmt, w, h, flags=flags[a:b], combed=None if tags is None else tags[a:b]).frames
        whole = ref.run("hb_filter_decomb", f"mode={mode}", clip, fmt, w, h, flags=flags, combed=tags).frames
        assert whole.shape[0] == n * k
        assert np.array_equal(_sharded(blocks, world, clip_of, decomb, k), whole), mode


def test_sharded_product_host_filters_equal_unsharded_reference(ref):
    """the same property with the PRODUCT's filter objects (their host code over the CPU stand-ins for the device calls,
    tests/test_hostlogic.py) on the sharded side and the reference, unsharded, on the other: what bench.py --gpus N and a
    multi-GPU libhb would run per rank"""
    from handbrake_b200.hblib import FilterLib
    from test_hostlogic import HOSTLOGIC_SO
    from test_oracle import decomb_inputs
    if not HOSTLOGIC_SO.exists():
        pytest.skip("oracle/_ref/libhostlogic.so not built")
    prod = FilterLib(HOSTLOGIC_SO)
    w, h, fmt = 96, 64, synth.PIX_FMT_YUV420P
    clip, flags, combed = decomb_inputs(8, w, h, 8, seed=3)
    n, world = clip.shape[0], 3
    window = {}

    def clip_of(a, b):
        window["range"] = (a, b)
        return clip[a:b]

    blocks = sharding.plan_blocks(n, world, 2, halo_before=1, halo_after=1)
    for mode, k in ((7, 1), (23, 2)):
        def decomb(frames):
            a, b = window["range"]
            return prod.run("hb_filter_decomb_cuda", f"mode={mode}", frames, fmt, w, h, flags=flags[a:b]).frames
        whole = ref.run("hb_filter_decomb", f"mode={mode}", clip, fmt, w, h, flags=flags).frames
        assert np.array_equal(_sharded(blocks, world, clip_of, decomb, k), whole), mode

    def comb(frames):
        a, b = window["range"]
        return np.asarray(prod.run("hb_filter_comb_detect_cuda", None, frames, fmt, w, h, flags=flags[a:b]).combed, np.uint8).reshape(-1, 1)
    whole = np.asarray(ref.run("hb_filter_comb_detect", None, clip, fmt, w, h, flags=flags).combed, np.uint8)
    assert np.array_equal(_sharded(blocks, world, clip_of, comb).ravel(), whole)

    nf = 3
    blocks = sharding.plan_blocks(n, world, 2, halo_after=nf - 1)
    s = f"y-strength=6:y-patch-size=3:y-range=3:y-frame-count={nf}"
    nlm = lambda fr: prod.run("hb_filter_nlmeans_cuda", s, fr, fmt, w, h).frames
    assert np.array_equal(_sharded(blocks, world, lambda a, b: clip[a:b], nlm), ref.run("hb_filter_nlmeans", s, clip, fmt, w, h).frames)


def test_sharded_lapsharp_needs_no_halo(ref):
    w, h = 96, 64
    clip = synth.progressive_clip(synth.PIX_FMT_YUV420P, w, h, 7, seed=4, noise=15)
    run = lambda fr: ref.run("hb_filter_lapsharp", "y-strength=0.3:y-kernel=isolap", fr, synth.PIX_FMT_YUV420P, w, h).frames
    blocks = sharding.plan_blocks(7, 3, 2)
    assert np.array_equal(_sharded(blocks, 3, lambda a, b: clip[a:b], run), run(clip))


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
