"""GPU parity: hb_filter_lapsharp_cuda vs the reference's hb_filter_lapsharp (bit-exact)."""
import numpy as np
import pytest

from handbrake_b200 import synth
from test_oracle import LAPSHARP_CASES

pytestmark = pytest.mark.gpu
FMT = {8: synth.PIX_FMT_YUV420P, 10: synth.PIX_FMT_YUV420P10}


@pytest.mark.parametrize("settings,strengths,kernels", LAPSHARP_CASES)
@pytest.mark.parametrize("depth,w,h", [(8, 200, 90), (8, 256, 64), (10, 200, 90), (10, 192, 66), (8, 640, 360)])
def test_matches_reference(ref, cuda_filters, settings, strengths, kernels, depth, w, h):
    clip = synth.progressive_clip(FMT[depth], w, h, 9, seed=31, noise=20)
    r = ref.run("hb_filter_lapsharp_mt", settings, clip, FMT[depth], w, h)      # as libhb runs it (mt_frame wrapper)
    g = cuda_filters.run("hb_filter_lapsharp_cuda", settings, clip, FMT[depth], w, h)
    assert g.saw_eof and g.frames.shape == r.frames.shape
    assert np.array_equal(g.frames, r.frames)
    assert np.array_equal(g.start, r.start)
    assert cuda_filters.buffers_alive() == 0


def test_full_chain_order(ref, cuda_filters):
    """libhb's enforced order (hb.c:1701-1720): comb_detect -> decomb -> nlmeans -> lapsharp"""
    w, h = 256, 144
    from test_oracle import decomb_inputs
    clip, flags, _ = decomb_inputs(8, w, h, 8, seed=3)
    names_r = ["hb_filter_comb_detect", "hb_filter_decomb", "hb_filter_nlmeans", "hb_filter_lapsharp_mt"]
    names_g = ["hb_filter_comb_detect_cuda", "hb_filter_decomb_cuda", "hb_filter_nlmeans_cuda", "hb_filter_lapsharp_cuda"]
    s = [None, "mode=39", "y-strength=6", "y-strength=0.2:y-kernel=isolap"]
    r = ref.run(names_r, [s[0], s[1], s[2] + ":threads=2", s[3]], clip, FMT[8], w, h, flags=flags)
    g = cuda_filters.run(names_g, s, clip, FMT[8], w, h, flags=flags)
    assert g.frames.shape == r.frames.shape and np.array_equal(g.frames, r.frames)
    assert np.array_equal(g.start, r.start) and list(g.combed) == list(r.combed)
