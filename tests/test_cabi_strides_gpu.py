"""The C-ABI with frame buffers that are NOT laid out like hb_frame_buffer_init's (decoder-owned AVFRAME memory with an
arbitrary linesize per plane, planes in separate allocations, libhb/hbffmpeg.c:182-239): every entry point takes explicit
strides, the one-copy-per-frame fast path must not be taken, results must not change."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest

import handbrake_b200
from handbrake_b200 import synth

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bench import NlmConfig, nlm_config, fmt_of   # noqa: E402

pytestmark = pytest.mark.gpu


def padded_planes(frame, dims, bps, pad):
    """each plane in its own allocation, row stride = width*bps + pad bytes; returns (arrays, ptr array, stride array)"""
    arrs, off = [], 0
    for (w, h) in dims:
        a = np.full((h, w * bps + pad), 0xA5, np.uint8)
        a[:, : w * bps] = frame[off: off + w * h * bps].reshape(h, w * bps)
        arrs.append(a)
        off += w * h * bps
    ptrs = (C.c_void_p * 3)(*[a.ctypes.data for a in arrs])
    strides = (C.c_int * 3)(*[a.shape[1] for a in arrs])
    return arrs, ptrs, strides


@pytest.mark.parametrize("depth,pad", [(8, 48), (10, 16), (8, 1)])
def test_nlmeans_c_abi_with_foreign_strides(cuda_filters, depth, pad):
    w, h, n = 200, 120, 5
    fmt = fmt_of(depth)
    bps = 2 if depth > 8 else 1
    clip = synth.progressive_clip(fmt, w, h, n, seed=101)
    want = cuda_filters.run("hb_filter_nlmeans_cuda", "y-strength=6", clip, fmt, w, h).frames      # itself checked against the reference elsewhere
    core = C.CDLL(str(handbrake_b200.LIBHBCU))
    core.hbcu_last_error.restype = C.c_char_p
    cfg = nlm_config(cuda_filters.lib, dict(settings="y-strength=6", depth=depth, width=w, height=h), 0, ring=8, out_slots=4)
    hnd = C.c_void_p()
    assert core.hbcu_nlmeans_create(C.byref(hnd), C.byref(cfg)) == 0, core.hbcu_last_error()
    core.hbcu_nlmeans_upload.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    core.hbcu_nlmeans_filter.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    core.hbcu_nlmeans_wait.argtypes = [C.c_void_p, C.c_int64]
    dims = synth.plane_dims(w, h)
    keep = []
    for t in range(n):
        arrs, ptrs, strides = padded_planes(clip[t], dims, bps, pad)
        keep.append(arrs)
        assert core.hbcu_nlmeans_upload(hnd, t, ptrs, strides) == 0, core.hbcu_last_error()
    for t in range(n):
        out, optrs, ostrides = padded_planes(np.zeros_like(clip[t]), dims, bps, pad + 3 if pad > 1 else pad)
        assert core.hbcu_nlmeans_filter(hnd, t, min(2, n - t), optrs, ostrides) == 0, core.hbcu_last_error()
        assert core.hbcu_nlmeans_wait(hnd, t) == 0, core.hbcu_last_error()
        got = np.concatenate([a[:, : dw * bps].reshape(-1) for a, (dw, dh) in zip(out, dims)])
        assert np.array_equal(got, want[t]), f"frame {t}"
        for a, (dw, dh) in zip(out, dims):
            assert np.all(a[:, dw * bps:] == 0xA5)           # the caller's stride padding is left alone
    core.hbcu_nlmeans_destroy.argtypes = [C.c_void_p]
    core.hbcu_nlmeans_destroy(hnd)


def test_unsharp_c_abi_with_foreign_strides(ref, cuda_filters):
    """unsharp through hbcu_unsharp_filter_frames with separately allocated, padded planes on both sides"""
    w, h = 200, 120
    fmt = synth.PIX_FMT_YUV420P
    clip = synth.progressive_clip(fmt, w, h, 2, seed=103)
    want = ref.run("hb_filter_unsharp_mt", None, clip, fmt, w, h).frames

    class UnsharpConfig(C.Structure):
        _fields_ = [("width", C.c_int), ("height", C.c_int), ("depth", C.c_int), ("chroma_shift_w", C.c_int), ("chroma_shift_h", C.c_int),
                    ("device", C.c_int), ("slots", C.c_int), ("smooth", C.c_int), ("amount", C.c_int * 3), ("steps", C.c_int * 3)]

    core = C.CDLL(str(handbrake_b200.LIBHBCU))
    core.hbcu_last_error.restype = C.c_char_p
    cfg = UnsharpConfig(w, h, 8, 1, 1, 0, 4, 0, (C.c_int * 3)(16384, 16384, 16384), (C.c_int * 3)(3, 3, 3))
    hnd = C.c_void_p()
    assert core.hbcu_unsharp_create(C.byref(hnd), C.byref(cfg)) == 0, core.hbcu_last_error()
    core.hbcu_unsharp_filter_frames.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    core.hbcu_unsharp_wait.argtypes = [C.c_void_p, C.c_int64]
    dims = synth.plane_dims(w, h)
    for t in range(2):
        src, sp, ss = padded_planes(clip[t], dims, 1, 40)
        out, op, os_ = padded_planes(np.zeros_like(clip[t]), dims, 1, 24)
        assert core.hbcu_unsharp_filter_frames(hnd, t, None, sp, ss, None, op, os_) == 0, core.hbcu_last_error()
        assert core.hbcu_unsharp_wait(hnd, t) == 0
        got = np.concatenate([a[:, :dw].reshape(-1) for a, (dw, dh) in zip(out, dims)])
        assert np.array_equal(got, want[t])
    core.hbcu_unsharp_destroy.argtypes = [C.c_void_p]
    core.hbcu_unsharp_destroy(hnd)
