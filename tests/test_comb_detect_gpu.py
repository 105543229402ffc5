"""GPU parity: hb_filter_comb_detect_cuda vs the reference's hb_filter_comb_detect (verdicts identical,
frames untouched), and the device masks vs the oracle restatement through the C-ABI."""
import ctypes as C

import numpy as np
import pytest

import handbrake_b200
from handbrake_b200 import synth
from oracle_port import OraclePort, comb_params
from test_oracle import COMB_SETTINGS, mixed_interlaced_clip

pytestmark = pytest.mark.gpu
FMT = {8: synth.PIX_FMT_YUV420P, 10: synth.PIX_FMT_YUV420P10}


@pytest.mark.parametrize("settings", COMB_SETTINGS)
@pytest.mark.parametrize("depth", [8, 10])
def test_verdicts_match_reference(ref, cuda_filters, settings, depth):
    w, h = 336, 208
    clip = mixed_interlaced_clip(FMT[depth], w, h, 11)
    flags = np.full(clip.shape[0], synth.PIC_FLAG_TOP_FIELD_FIRST, np.uint16)
    r = ref.run("hb_filter_comb_detect", settings, clip, FMT[depth], w, h, flags=flags)
    g = cuda_filters.run("hb_filter_comb_detect_cuda", settings, clip, FMT[depth], w, h, flags=flags)
    assert g.saw_eof and g.frames.shape == r.frames.shape
    assert np.array_equal(g.frames, clip)                      # frames pass through untouched
    assert list(g.combed) == list(r.combed), (settings, depth)
    assert np.array_equal(g.start, r.start) and np.array_equal(g.flags, r.flags)
    assert cuda_filters.buffers_alive() == 0


def test_short_clips(ref, cuda_filters):
    """1, 2 and 3 frame clips: first/last-frame duplication and the exhaustive check"""
    w, h = 176, 112
    for n in (1, 2, 3):
        clip = synth.interlaced_clip(FMT[8], w, h, n, static_every=0)
        r = ref.run("hb_filter_comb_detect", None, clip, FMT[8], w, h)
        g = cuda_filters.run("hb_filter_comb_detect_cuda", None, clip, FMT[8], w, h)
        assert list(g.combed) == list(r.combed) and g.frames.shape[0] == n


class CombConfig(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("depth", C.c_int), ("device", C.c_int), ("slots", C.c_int),
                ("mode", C.c_int), ("spatial_metric", C.c_int), ("filter_mode", C.c_int),
                ("motion_threshold", C.c_int), ("spatial_threshold", C.c_int),
                ("block_threshold", C.c_int), ("block_width", C.c_int), ("block_height", C.c_int),
                ("gamma_motion_threshold", C.c_float), ("gamma_spatial_threshold", C.c_float),
                ("gamma_spatial_threshold6", C.c_float), ("comb32detect_min", C.c_int), ("comb32detect_max", C.c_int),
                ("gamma_lut", C.c_void_p)]


@pytest.mark.parametrize("settings", [None, COMB_SETTINGS[1], COMB_SETTINGS[2], COMB_SETTINGS[3], COMB_SETTINGS[5]])
@pytest.mark.parametrize("depth", [8, 10])
def test_masks_match_oracle_through_c_abi(cuda_filters, settings, depth):
    """raw mask and scored mask, bit for bit, incl. ragged geometry (not a multiple of the 64x32 filter tile)"""
    core = C.CDLL(str(handbrake_b200.LIBHBCU))
    core.hbcu_last_error.restype = C.c_char_p
    core.hbcu_comb_detect_upload.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
    core.hbcu_comb_detect_run.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int]
    core.hbcu_comb_detect_result.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int)]
    core.hbcu_comb_detect_masks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    core.hbcu_comb_detect_destroy.argtypes = [C.c_void_p]
    port = OraclePort()
    w, h = 203, 131
    clip = mixed_interlaced_clip(FMT[depth], w, h, 4, seed=77)
    planes = [synth.split_planes(f, FMT[depth], w, h)[0] for f in clip]
    p = comb_params(settings)
    maxv = (1 << depth) - 1
    mth, sth = p.motion_threshold << (depth - 8), p.spatial_threshold << (depth - 8)
    lut = np.zeros(maxv + 1, np.float32)
    port.lib.oracle_comb_gamma_lut(depth, C.c_void_p(lut.ctypes.data))     # the C expression, not numpy's pow
    cfg = CombConfig(w, h, depth, 0, 8, p.mode, p.spatial_metric, p.filter_mode, mth, sth, p.block_threshold,
                     min(p.block_width, w), min(p.block_height, h),
                     np.float32(mth) / np.float32(maxv), np.float32(sth) / np.float32(maxv),
                     np.float32(6) * (np.float32(sth) / np.float32(maxv)), 10 << (depth - 8), 15 << (depth - 8),
                     lut.ctypes.data)
    hnd = C.c_void_p()
    assert core.hbcu_comb_detect_create(C.byref(hnd), C.byref(cfg)) == 0, core.hbcu_last_error()
    bps = 2 if depth > 8 else 1
    for i, pl in enumerate(planes):
        pl = np.ascontiguousarray(pl)
        planes[i] = pl
        assert core.hbcu_comb_detect_upload(hnd, i, pl.ctypes.data, w * bps) == 0, core.hbcu_last_error()
    for (a, b, c, force) in ((0, 0, 1, 1), (0, 1, 2, 0), (1, 2, 3, 0), (2, 3, 3, 1)):
        assert core.hbcu_comb_detect_run(hnd, a, b, c, force) == 0, core.hbcu_last_error()
        verdict = C.c_int(-1)
        assert core.hbcu_comb_detect_result(hnd, b, C.byref(verdict)) == 0
        raw = np.zeros((h, w), np.uint8)
        scored = np.zeros((h, w), np.uint8)
        assert core.hbcu_comb_detect_masks(hnd, raw.ctypes.data, scored.ctypes.data) == 0
        v, m, f = port.comb_detect_masks(planes[a], planes[b], planes[c], w, h, depth, settings, force)
        assert np.array_equal(raw, m), (settings, depth, (a, b, c), int(np.count_nonzero(raw != m)))
        assert np.array_equal(scored, f)
        assert verdict.value == v
    core.hbcu_comb_detect_destroy(hnd)
