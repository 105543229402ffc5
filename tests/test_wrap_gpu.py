"""GPU tests of the NVDEC / NVENC seam: device memory somebody else owns -- what an
AVFrame of AV_PIX_FMT_CUDA carries (libhb/nvenc_common.c:329-336, libhb/hwaccel.c:15-60) -- enters the filter path as an
HBCU_DEVICE frame without a copy, and an external consumer's stream can read any device frame in stream order."""
import ctypes as C

import numpy as np
import pytest

import handbrake_b200
from handbrake_b200 import synth

pytestmark = pytest.mark.gpu

FMT8 = synth.PIX_FMT_YUV420P
UP, DOWN = "hb_filter_hbcu_upload", "hb_filter_hbcu_download"


def core_lib():
    core = C.CDLL(str(handbrake_b200.LIBHBCU))
    core.hbcu_last_error.restype = C.c_char_p
    core.hbcu_frames_alive.restype = C.c_long
    core.hbcu_frame_plane.restype = C.c_void_p
    return core


def test_wrapped_surfaces_through_the_filter_chain(ref, cuda_filters, monkeypatch):
    """the upload adapter plays a hardware decoder (HBCU_UPLOAD_EXTERNAL): downstream filters read wrapped surfaces"""
    w, h, n = 333, 211, 9
    clip = synth.progressive_clip(FMT8, w, h, n, seed=19)
    r = ref.run(["hb_filter_nlmeans", "hb_filter_lapsharp_mt"], ["y-strength=6:threads=2", "y-strength=0.3"], clip, FMT8, w, h)
    monkeypatch.setenv("HBCU_UPLOAD_EXTERNAL", "1")
    cuda_filters.lib.hbcu_test_surfaces_returned.restype = C.c_long
    before = cuda_filters.lib.hbcu_test_surfaces_returned()
    g = cuda_filters.run([UP, "hb_filter_nlmeans_cuda", "hb_filter_lapsharp_cuda", DOWN], [None, "y-strength=6", "y-strength=0.3", None],
                         clip, FMT8, w, h)
    assert g.saw_eof and np.array_equal(g.frames, r.frames) and np.array_equal(g.start, r.start)
    assert cuda_filters.lib.hbcu_test_surfaces_returned() - before == n
    assert cuda_filters.buffers_alive() == 0 and core_lib().hbcu_frames_alive() == 0


class NlmPlane(C.Structure):
    _fields_ = [("patch_size", C.c_int), ("range", C.c_int), ("nframes", C.c_int), ("bypass", C.c_int),
                ("origin_tune", C.c_double), ("weight_fact", C.c_float), ("diff_max", C.c_int),
                ("exptable", C.c_float * 128), ("prefilter", C.c_int)]


class NlmConfig(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("depth", C.c_int), ("chroma_shift_w", C.c_int),
                ("chroma_shift_h", C.c_int), ("device", C.c_int), ("ring_frames", C.c_int), ("out_slots", C.c_int),
                ("plane", NlmPlane * 3)]


def test_torch_owned_pitched_memory_wrapped_at_the_c_abi(cuda_filters):
    """planes allocated by ANOTHER allocator (torch) with a decoder-style pitch, written on ITS stream: wrap, denoise,
    compare with the host-buffer path; release fires once, only after the device is done with the surface"""
    import torch
    core = core_lib()
    flt = cuda_filters.lib
    w, h, n = 640, 360, 4
    dims = synth.plane_dims(w, h)
    clip = synth.progressive_clip(FMT8, w, h, n, seed=5)

    cfg = NlmConfig()
    flt.hb_parse_filter_settings.restype = C.c_void_p
    flt.hb_parse_filter_settings.argtypes = [C.c_char_p]
    flt.hb_nlmeans_cuda_build_config.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(NlmConfig), C.c_void_p, C.c_void_p, C.c_void_p]
    assert flt.hb_nlmeans_cuda_build_config(flt.hb_parse_filter_settings(b"y-strength=6"), FMT8, w, h, C.byref(cfg), None, None, None) == 0
    cfg.device, cfg.ring_frames, cfg.out_slots = 0, 8, 4

    def run(wrapped):
        hdl = C.c_void_p()
        assert core.hbcu_nlmeans_create(C.byref(hdl), C.byref(cfg)) == 0, core.hbcu_last_error()
        outs, keep, released = [], [], []
        REL = C.CFUNCTYPE(None, C.c_void_p)
        rel = REL(lambda opaque: released.append(int(opaque or 0)))
        side = torch.cuda.Stream()
        for t in range(n):
            off, planes = 0, []
            for (pw, ph) in dims:
                planes.append(clip[t, off:off + pw * ph].reshape(ph, pw)); off += pw * ph
            if wrapped:
                pitch = [1024, 512, 512]                        # a decoder's pitch, not hb_image_stride's
                surf = [torch.zeros((ph + 8) * p + 256, dtype=torch.uint8, device="cuda") for (pw, ph), p in zip(dims, pitch)]
                with torch.cuda.stream(side):                   # the "decoder" writes on its own stream
                    for s_, pl, (pw, ph), p in zip(surf, planes, dims, pitch):
                        s_[:ph * p].view(ph, p)[:, :pw].copy_(torch.from_numpy(np.ascontiguousarray(pl)), non_blocking=False)
                keep.append(surf)
                ptrs = (C.c_void_p * 3)(*[s_.data_ptr() for s_ in surf])
                rb = (C.c_int * 3)(*[pw for pw, _ in dims]); rows = (C.c_int * 3)(*[ph for _, ph in dims]); st = (C.c_int * 3)(*pitch)
                fr = C.c_void_p()
                core.hbcu_frame_wrap.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                                 C.c_void_p, REL, C.c_void_p]
                assert core.hbcu_frame_wrap(C.byref(fr), 0, ptrs, rb, rows, st, 256, C.c_void_p(side.cuda_stream), rel, C.c_void_p(t + 1)) == 0, core.hbcu_last_error()
                assert core.hbcu_nlmeans_upload_frame(hdl, C.c_int64(t), fr) == 0, core.hbcu_last_error()
                core.hbcu_frame_release(fr)                     # our reference goes; the surface returns when the border kernels are done
            else:
                hp = [np.ascontiguousarray(pl) for pl in planes]
                keep.append(hp)
                ptrs = (C.c_void_p * 3)(*[a.ctypes.data for a in hp]); st = (C.c_int * 3)(*[pw for pw, _ in dims])
                assert core.hbcu_nlmeans_upload(hdl, C.c_int64(t), ptrs, st) == 0, core.hbcu_last_error()
        for t in range(n):
            out = [np.zeros((ph, pw), np.uint8) for pw, ph in dims]
            ptrs = (C.c_void_p * 3)(*[a.ctypes.data for a in out]); st = (C.c_int * 3)(*[pw for pw, _ in dims])
            assert core.hbcu_nlmeans_filter(hdl, C.c_int64(t), n - t, ptrs, st) == 0, core.hbcu_last_error()
            assert core.hbcu_nlmeans_wait(hdl, C.c_int64(t)) == 0
            outs.append(np.concatenate([a.ravel() for a in out]))
        core.hbcu_nlmeans_destroy(hdl)
        return np.stack(outs), released

    host, _ = run(False)
    dev, released = run(True)
    assert np.array_equal(host, dev)
    assert sorted(released) == [1, 2, 3, 4]
    assert core.hbcu_frames_alive() == 0

    # refusals: unaligned pitch, no readable tail
    fr = C.c_void_p()
    buf = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
    ptrs = (C.c_void_p * 3)(buf.data_ptr(), buf.data_ptr() + 4096 * 64, buf.data_ptr() + 4096 * 96)
    rb = (C.c_int * 3)(64, 32, 32); rows = (C.c_int * 3)(48, 24, 24)
    assert core.hbcu_frame_wrap(C.byref(fr), 0, ptrs, rb, rows, (C.c_int * 3)(100, 64, 64), 256, None, C.cast(None, C.CFUNCTYPE(None, C.c_void_p)), None) != 0
    assert core.hbcu_frame_wrap(C.byref(fr), 0, ptrs, rb, rows, (C.c_int * 3)(128, 64, 64), 0, None, C.cast(None, C.CFUNCTYPE(None, C.c_void_p)), None) != 0


def test_external_stream_reads_a_device_frame(cuda_filters):
    """the NVENC side: an outside stream acquires a pooled device frame behind its producer and marks its read done"""
    import torch
    core = core_lib()
    w, h = 256, 144
    dims = synth.plane_dims(w, h)
    frame = synth.progressive_clip(FMT8, w, h, 1, seed=2)[0]
    rb = (C.c_int * 3)(*[pw for pw, _ in dims]); rows = (C.c_int * 3)(*[ph for _, ph in dims]); st = (C.c_int * 3)(*[(pw + 63) // 64 * 64 for pw, _ in dims])
    fr = C.c_void_p()
    assert core.hbcu_frame_alloc(C.byref(fr), 0, rb, rows, st) == 0
    x = C.c_void_p()
    assert core.hbcu_xfer_create(C.byref(x), 0, 4) == 0
    off, hp = 0, []
    for (pw, ph) in dims:
        hp.append(np.ascontiguousarray(frame[off:off + pw * ph].reshape(ph, pw))); off += pw * ph
    ptrs = (C.c_void_p * 3)(*[a.ctypes.data for a in hp]); hst = (C.c_int * 3)(*[pw for pw, _ in dims])
    assert core.hbcu_xfer_upload(x, C.c_int64(0), fr, ptrs, hst) == 0                 # producer: queued, not waited for
    enc = torch.cuda.Stream()
    assert core.hbcu_frame_acquire(fr, C.c_void_p(enc.cuda_stream)) == 0
    import cuda.bindings.runtime as rt
    got = []
    for p, (pw, ph) in enumerate(dims):
        n = st[p] * ph
        t = torch.empty(n, dtype=torch.uint8, device="cuda")
        # the "encoder" reads the plane on ITS stream: device-to-device copy queued behind the acquire
        err, = rt.cudaMemcpyAsync(t.data_ptr(), core.hbcu_frame_plane(fr, p), n, rt.cudaMemcpyKind.cudaMemcpyDeviceToDevice, enc.cuda_stream)
        assert int(err) == 0
        got.append((t, pw, ph, st[p]))
    assert core.hbcu_frame_done(fr, C.c_void_p(enc.cuda_stream)) == 0
    core.hbcu_frame_release(fr)
    enc.synchronize()
    for (t, pw, ph, s_), want in zip(got, hp):
        assert np.array_equal(t.cpu().numpy().reshape(ph, s_)[:, :pw], want)
    core.hbcu_xfer_destroy(x)
    core.hbcu_frame_trim()
