"""tools/bench_filters.py -- the same measurement bench.py makes for NLMeans, for the other filters of the path
(BASELINE.json configs[2]: 4K 10-bit comb_detect + decomb; lapsharp).  One JSON line per workload:
  value    outputs/s with inputs resident in HBM (kernels only, CUDA events on the compute stream)
  e2e      input frames/s through the hb_filter_*_cuda object with pinned host buffers (H2D + kernels + D2H)
  roofline algorithmic HBM bytes / device time, against MEASURED_PEAKS.json
  cpu      the unmodified reference filter (oracle/_ref) through the same hb_bench protocol
usage: python tools/bench_filters.py [--frames 48] [--cpu-frames 6] [--only NAME]"""
import argparse
import ctypes as C
import json
import os
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import handbrake_b200  # noqa: E402
from handbrake_b200 import synth  # noqa: E402
from bench import BenchStats, bind_bench, fmt_of, measured_peaks  # noqa: E402

W, H = 3840, 2160


class DecombConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("width", "height", "depth", "chroma_shift_w", "chroma_shift_h", "device", "slots", "out_slots",
                                      "mode", "magnitude_threshold", "variance_threshold", "laplacian_threshold", "dilation_threshold",
                                      "erosion_threshold", "noise_threshold", "maximum_search_distance", "post_processing")]


class LapsharpConfig(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("depth", C.c_int), ("chroma_shift_w", C.c_int), ("chroma_shift_h", C.c_int),
                ("device", C.c_int), ("slots", C.c_int), ("strength", C.c_double * 3), ("kernel", C.c_int * 3)]


def plane_ptrs(t, dims, bps):
    base, off, ptrs, strides = t.data_ptr(), 0, [], []
    for (w, h) in dims:
        ptrs.append(base + off); strides.append(w * bps); off += w * h * bps
    return (C.c_void_p * 3)(*ptrs), (C.c_int * 3)(*strides)


def e2e_and_cpu(flt, ref, sym_cuda, sym_ref, settings, fmt, host, n, n_cpu):
    st = BenchStats()
    out = {}
    s = settings.encode() if settings else None
    for rep in range(2):
        b = flt.hb_bench_open(C.addressof(C.c_char.in_dll(flt, sym_cuda)), s, fmt, W, H)
        assert b and flt.hb_bench_run(b, host.ctypes.data, host.shape[0], n, C.byref(st)) == 0
    out["e2e"] = {"value": round(n / st.seconds, 1), "unit": "input frames/s", "frames_out": int(st.frames_out),
                  "h2d_bytes": int(st.bytes_in), "d2h_bytes": int(st.bytes_out)}
    if ref is not None:
        b = ref.hb_bench_open(C.addressof(C.c_char.in_dll(ref, sym_ref)), s, fmt, W, H)
        assert b and ref.hb_bench_run(b, host.ctypes.data, host.shape[0], n_cpu, C.byref(st)) == 0
        out["cpu_baseline"] = {"value": round(n_cpu / st.seconds, 3), "unit": "input frames/s", "kind": "reference",
                               "cores": ref.hb_get_cpu_count(), "sample": f"{n_cpu} frames"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=48)
    ap.add_argument("--cpu-frames", type=int, default=6)
    ap.add_argument("--only", default="")
    args = ap.parse_args()

    class _LazyTorch:                   # only the kernel-only arms need torch (device tensors); its import costs a minute on a fresh box
        def __getattr__(self, name):
            import torch as t
            return getattr(t, name)
    torch = _LazyTorch()
    core = C.CDLL(str(handbrake_b200.LIBHBCU))
    core.hbcu_last_error.restype = C.c_char_p
    flt = C.CDLL(str(handbrake_b200.LIBHBCU_FILTERS))
    bind_bench(flt)
    flt.hbcu_use_pinned_buffers(1)
    ref_so = REPO / "oracle" / "_ref" / "libhbref.so"
    ref = None
    if ref_so.exists():
        ref = C.CDLL(str(ref_so)); bind_bench(ref)
    peak, peak_src = measured_peaks()
    n = args.frames
    core.hbcu_host_reserve.argtypes = [C.c_size_t, C.c_int]

    def ck(rc):
        if rc != 0:
            raise RuntimeError(core.hbcu_last_error().decode())

    jobs = []

    def decomb_job(name, mode, depth, desc, alg_factor):
        fmt = fmt_of(depth); bps = 2 if depth > 8 else 1
        fb = synth.frame_bytes(fmt, W, H)
        host = np.stack([synth.interlaced_frame(fmt, W, H, t) for t in range(4)])
        core.hbcu_host_reserve(fb + 4096, 3 * n + 24)
        cfg = DecombConfig(W, H, depth, 1, 1, 0, 8, 8, mode, 10, 20, 20, 4, 2, 50, 24, 1)
        h = C.c_void_p(); ck(core.hbcu_decomb_create(C.byref(h), C.byref(cfg)))
        dims = synth.plane_dims(W, H)
        dev = [torch.from_numpy(host[i % 4]).cuda() for i in range(8)]
        pp = [plane_ptrs(t, dims, bps) for t in dev]
        core.hbcu_decomb_upload_device.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        core.hbcu_decomb_filter_device.argtypes = [C.c_void_p] + [C.c_int64] * 4 + [C.c_int] * 3 + [C.c_void_p, C.c_void_p]
        fields = 2 if mode & 16 else 1

        def run(count, base):
            tk = base * 2
            for i in range(count):
                idx = base + i
                ck(core.hbcu_decomb_upload_device(h, idx, pp[idx % 8][0], pp[idx % 8][1]))
                if idx >= 2:
                    for f in range(fields):
                        ck(core.hbcu_decomb_filter_device(h, tk, idx - 2, idx - 1, idx, mode & ~32, f, 1, None, None)); tk += 1
        run(8, 0)
        ck(core.hbcu_decomb_sync(h)); ck(core.hbcu_decomb_mark(h, 0))
        run(n, 8)
        ck(core.hbcu_decomb_mark(h, 1))
        ms = C.c_float(); ck(core.hbcu_decomb_elapsed_ms(h, C.byref(ms)))
        core.hbcu_decomb_destroy(h)
        outs = n * fields
        alg = alg_factor * fb * outs                       # per output picture: read prev+cur+next, write one
        r = {"workload": name, "desc": desc, "value": round(outs / (ms.value / 1e3), 1), "unit": "output pictures/s",
             "roofline": {"bound": "hbm", "achieved": round(alg / (ms.value / 1e3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                          "frac": round(alg / (ms.value / 1e3) / 1e9 / peak, 4), "algorithmic_bytes_per_output": alg_factor * fb,
                          "peak_source": peak_src}}
        r.update(e2e_and_cpu(flt, ref, "hb_filter_decomb_cuda", "hb_filter_decomb", f"mode={mode}", fmt, host, n, args.cpu_frames))
        return r

    jobs.append(("4k10_decomb_yadif", lambda: decomb_job("4k10_decomb_yadif", 7, 10, "3840x2160 yuv420p10, decomb default (yadif+cubic, every frame)", 4)))
    jobs.append(("4k10_decomb_eedi2bob", lambda: decomb_job("4k10_decomb_eedi2bob", 31, 10, "3840x2160 yuv420p10, decomb eedi2bob (mode 31)", 4)))

    def lapsharp_job():
        depth = 8; fmt = fmt_of(depth)
        fb = synth.frame_bytes(fmt, W, H)
        host = np.stack([synth.progressive_frame(fmt, W, H, t) for t in range(4)])
        core.hbcu_host_reserve(fb + 4096, 3 * n + 24)
        cfg = LapsharpConfig(W, H, depth, 1, 1, 0, 8, (C.c_double * 3)(0.2, 0.2, 0.2), (C.c_int * 3)(1, 1, 1))
        h = C.c_void_p(); ck(core.hbcu_lapsharp_create(C.byref(h), C.byref(cfg)))
        dims = synth.plane_dims(W, H)
        dev = [torch.from_numpy(host[i % 4]).cuda() for i in range(16)]      # 16 x 12.4 MB > L2
        pp = [plane_ptrs(t, dims, 1) for t in dev]
        core.hbcu_lapsharp_filter_device.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        for i in range(8):
            ck(core.hbcu_lapsharp_filter_device(h, i, pp[i % 16][0], pp[i % 16][1], None, None))
        ck(core.hbcu_lapsharp_sync(h)); ck(core.hbcu_lapsharp_mark(h, 0))
        for i in range(n):
            ck(core.hbcu_lapsharp_filter_device(h, 100 + i, pp[i % 16][0], pp[i % 16][1], None, None))
        ck(core.hbcu_lapsharp_mark(h, 1))
        ms = C.c_float(); ck(core.hbcu_lapsharp_elapsed_ms(h, C.byref(ms)))
        core.hbcu_lapsharp_destroy(h)
        alg = 2 * fb * n
        r = {"workload": "4k_lapsharp", "desc": "3840x2160 yuv420p 8-bit, lapsharp medium (isolap 0.2)",
             "value": round(n / (ms.value / 1e3), 1), "unit": "frames/s",
             "roofline": {"bound": "hbm", "achieved": round(alg / (ms.value / 1e3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                          "frac": round(alg / (ms.value / 1e3) / 1e9 / peak, 4), "algorithmic_bytes_per_output": 2 * fb, "peak_source": peak_src}}
        r.update(e2e_and_cpu(flt, ref, "hb_filter_lapsharp_cuda", "hb_filter_lapsharp_mt", "y-strength=0.2:y-kernel=isolap", fmt, host, n, max(args.cpu_frames, 16)))
        return r
    jobs.append(("4k_lapsharp", lapsharp_job))

    def unsharp_job(name, smooth, settings, ref_name, desc):
        depth = 8; fmt = fmt_of(depth)
        fb = synth.frame_bytes(fmt, W, H)
        host = np.stack([synth.progressive_frame(fmt, W, H, t) for t in range(4)])
        core.hbcu_host_reserve(fb + 4096, 3 * n + 24)

        class UnsharpConfig(C.Structure):
            _fields_ = [("width", C.c_int), ("height", C.c_int), ("depth", C.c_int), ("chroma_shift_w", C.c_int), ("chroma_shift_h", C.c_int),
                        ("device", C.c_int), ("slots", C.c_int), ("smooth", C.c_int), ("amount", C.c_int * 3), ("steps", C.c_int * 3)]

        class Frame(C.Structure):
            pass

        # kernel-only: device frames in, device frames out (the device-resident chain's view of the filter)
        amount = (C.c_int * 3)(0 if smooth else 16384, 16384, 16384)        # strength 0.25 (defaults); chroma_smooth copies luma
        cfg = UnsharpConfig(W, H, depth, 1, 1, 0, 8, int(smooth), amount, (C.c_int * 3)(3, 3, 3))
        h = C.c_void_p(); ck(core.hbcu_unsharp_create(C.byref(h), C.byref(cfg)))
        dims = synth.plane_dims(W, H)
        rb = (C.c_int * 3)(*[d[0] for d in dims]); rows = (C.c_int * 3)(*[d[1] for d in dims]); st = (C.c_int * 3)(*[d[0] for d in dims])
        core.hbcu_frame_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        core.hbcu_frame_plane.restype = C.c_void_p
        core.hbcu_frame_plane.argtypes = [C.c_void_p, C.c_int]
        core.hbcu_frame_release.argtypes = [C.c_void_p]
        core.hbcu_unsharp_filter_frames.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        fin, fout = [], []
        x = C.c_void_p(); ck(core.hbcu_xfer_create(C.byref(x), 0, 8))
        core.hbcu_xfer_upload.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        core.hbcu_xfer_wait.argtypes = [C.c_void_p, C.c_int64]
        core.hbcu_xfer_destroy.argtypes = [C.c_void_p]
        for i in range(16):                                              # 16 x 12.4 MB in + out > L2
            a, b = C.c_void_p(), C.c_void_p()
            ck(core.hbcu_frame_alloc(C.byref(a), 0, rb, rows, st)); ck(core.hbcu_frame_alloc(C.byref(b), 0, rb, rows, st))
            base = host[i % 4].ctypes.data
            offs = [0, dims[0][0] * dims[0][1], dims[0][0] * dims[0][1] + dims[1][0] * dims[1][1]]
            planes = (C.c_void_p * 3)(*[base + o for o in offs])
            ck(core.hbcu_xfer_upload(x, i, a, planes, st)); ck(core.hbcu_xfer_wait(x, i))
            fin.append(a); fout.append(b)
        core.hbcu_xfer_destroy(x)
        for i in range(8):
            ck(core.hbcu_unsharp_filter_frames(h, i, fin[i % 16], None, None, fout[i % 16], None, None))
        ck(core.hbcu_unsharp_sync(h)); ck(core.hbcu_unsharp_mark(h, 0))
        for i in range(n):
            ck(core.hbcu_unsharp_filter_frames(h, 100 + i, fin[i % 16], None, None, fout[i % 16], None, None))
        ck(core.hbcu_unsharp_mark(h, 1))
        ms = C.c_float(); ck(core.hbcu_unsharp_elapsed_ms(h, C.byref(ms)))
        core.hbcu_unsharp_destroy(h)
        for f in fin + fout: core.hbcu_frame_release(f)
        alg = 2 * fb * n
        r = {"workload": name, "desc": desc, "value": round(n / (ms.value / 1e3), 1), "unit": "frames/s",
             "roofline": {"bound": "hbm", "achieved": round(alg / (ms.value / 1e3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                          "frac": round(alg / (ms.value / 1e3) / 1e9 / peak, 4), "algorithmic_bytes_per_output": 2 * fb, "peak_source": peak_src}}
        r.update(e2e_and_cpu(flt, ref, f"hb_filter_{ref_name}_cuda", f"hb_filter_{ref_name}_mt", settings, fmt, host, n, max(args.cpu_frames, 16)))
        return r
    jobs.append(("4k_unsharp", lambda: unsharp_job("4k_unsharp", False, None, "unsharp", "3840x2160 yuv420p 8-bit, unsharp defaults (0.25, size 7)")))
    jobs.append(("4k_chroma_smooth", lambda: unsharp_job("4k_chroma_smooth", True, None, "chroma_smooth", "3840x2160 yuv420p 8-bit, chroma_smooth defaults (0.25, size 7)")))

    def hqdn3d_job():
        depth = 8; fmt = fmt_of(depth)
        fb = synth.frame_bytes(fmt, W, H)
        host = np.stack([synth.progressive_frame(fmt, W, H, t) for t in range(4)])
        core.hbcu_host_reserve(fb + 4096, 3 * n + 24)
        r = {"workload": "4k_hqdn3d", "desc": "3840x2160 yuv420p 8-bit, hqdn3d defaults (4:3:6)"}
        r.update(e2e_and_cpu(flt, ref, "hb_filter_denoise_cuda", "hb_filter_denoise", None, fmt, host, n, max(args.cpu_frames, 8)))
        return r
    jobs.append(("4k_hqdn3d", hqdn3d_job))

    def detelecine_job():
        depth = 8; fmt = fmt_of(depth)
        fb = synth.frame_bytes(fmt, W, H)
        # one 2:3 cadence cycle (4 film frames -> 5 pictures), bottom field first: the bench harness sets no TFF flag
        host, _ = synth.telecined_clip(fmt, W, H, 4, tff=False)
        core.hbcu_host_reserve(fb + 4096, 3 * n + 24)
        r = {"workload": "4k_detelecine", "desc": "3840x2160 yuv420p 8-bit, hard 2:3 pulldown, detelecine defaults"}
        r.update(e2e_and_cpu(flt, ref, "hb_filter_detelecine_cuda", "hb_filter_detelecine", None, fmt, host, n, max(args.cpu_frames, 20)))
        return r
    jobs.append(("4k_detelecine", detelecine_job))

    def comb_job():
        depth = 10; fmt = fmt_of(depth)
        progressive = os.environ.get("HBCU_BENCH_COMB_PROGRESSIVE") == "1"       # content that rarely passes the spatial gate
        host = np.stack([(synth.progressive_frame(fmt, W, H, t, noise=2) if progressive else synth.interlaced_frame(fmt, W, H, t)) for t in range(4)])
        core.hbcu_host_reserve(synth.frame_bytes(fmt, W, H) + 4096, 3 * n + 24)
        r = {"workload": "4k10_comb_detect", "desc": "3840x2160 yuv420p10, comb_detect preset default" + (" (progressive content)" if progressive else " (interlaced content)")}
        # kernel-only: luma planes resident in HBM, verdict per frame
        sys.path.insert(0, str(REPO / "tests"))
        from test_comb_detect_gpu import CombConfig
        maxv = 1023
        lut = (np.arange(maxv + 1, dtype=np.float32) / np.float32(maxv)).astype(np.float64) ** 2.2
        lut = lut.astype(np.float32)              # throughput only; the filter object builds the exact table in C
        cfg = CombConfig(W, H, depth, 0, 8, 3, 2, 2, 4, 4, 40, 16, 16, np.float32(4 / maxv), np.float32(4 / maxv),
                         np.float32(24 / maxv), 40, 60, lut.ctypes.data)
        h = C.c_void_p(); ck(core.hbcu_comb_detect_create(C.byref(h), C.byref(cfg)))
        core.hbcu_comb_detect_upload_device.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        core.hbcu_comb_detect_run.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int]
        dev = [torch.from_numpy(host[i % 4]).cuda() for i in range(8)]

        def run(count, base):
            for i in range(count):
                idx = base + i
                ck(core.hbcu_comb_detect_upload_device(h, idx, dev[idx % 8].data_ptr(), W * 2))
                if idx >= 2:
                    ck(core.hbcu_comb_detect_run(h, idx - 2, idx - 1, idx, 0))
        run(8, 0)
        ck(core.hbcu_comb_detect_sync(h)); ck(core.hbcu_comb_detect_mark(h, 0))
        run(n, 8)
        ck(core.hbcu_comb_detect_mark(h, 1))
        ms = C.c_float(); ck(core.hbcu_comb_detect_elapsed_ms(h, C.byref(ms)))
        core.hbcu_comb_detect_destroy(h)
        alg = 3 * W * H * 2 * n                                   # prev, cur, next luma per verdict
        r["value"] = round(n / (ms.value / 1e3), 1); r["unit"] = "verdicts/s"
        r["roofline"] = {"bound": "hbm", "achieved": round(alg / (ms.value / 1e3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(alg / (ms.value / 1e3) / 1e9 / peak, 4), "algorithmic_bytes_per_output": 3 * W * H * 2,
                         "peak_source": peak_src}
        s = "mode=3:spatial-metric=2:motion-thresh=1:spatial-thresh=1:filter-mode=2:block-thresh=40:block-width=16:block-height=16"
        r.update(e2e_and_cpu(flt, ref, "hb_filter_comb_detect_cuda", "hb_filter_comb_detect", s, fmt, host, n, max(args.cpu_frames, 16)))
        return r
    jobs.append(("4k10_comb_detect", comb_job))

    for name, fn in jobs:
        if args.only and args.only != name:
            continue
        print(json.dumps(fn()), flush=True)


if __name__ == "__main__":
    main()
