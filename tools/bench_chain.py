#!/usr/bin/env python
"""tools/bench_chain.py -- a filter CHAIN end to end (host hb_buffer_t in, host hb_buffer_t out), three ways:
  host     every CUDA filter uploads and downloads its own frames (what libhb does today between CPU filters)
  device   hb_filter_hbcu_upload -> CUDA filters handing HBCU_DEVICE buffers on -> hb_filter_hbcu_download
  cpu      the reference's own filter objects (oracle/_ref/libhbref.so) on the host cores
One JSON line.  usage: python tools/bench_chain.py [--config 5] [--frames 24] [--cpu-frames 3] [--width W --height H]"""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import handbrake_b200  # noqa: E402
from handbrake_b200 import synth  # noqa: E402
from bench import BenchStats, fmt_of  # noqa: E402

CHAINS = {
    # BASELINE.json configs[4], in libhb's enforced filter order (hb.c:1701-1720)
    "5": dict(width=7680, height=4320, depth=10, interlaced=True,
              filters=["decomb", "nlmeans", "lapsharp"], settings=["mode=7", "y-strength=6", "y-strength=0.2:y-kernel=isolap"],
              desc="7680x4320 yuv420p10: decomb -> NLMeans medium -> lapsharp"),
    # configs[2]
    "3": dict(width=3840, height=2160, depth=10, interlaced=True,
              filters=["comb_detect", "decomb"], settings=["mode=3:spatial-metric=2:motion-thresh=1:spatial-thresh=1:filter-mode=2:block-thresh=40:block-width=16:block-height=16", "mode=63"],
              desc="3840x2160 yuv420p10: comb_detect -> decomb EEDI2 bob (selective)"),
    "4k": dict(width=3840, height=2160, depth=10, interlaced=True,
               filters=["comb_detect", "decomb", "nlmeans", "lapsharp"], settings=[None, "mode=39", "y-strength=6", "y-strength=0.2:y-kernel=isolap"],
               desc="3840x2160 yuv420p10: comb_detect -> decomb -> NLMeans medium -> lapsharp"),
}


def bind(lib):
    lib.hb_bench_run_chain.restype = C.c_int
    lib.hb_bench_run_chain.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_int, C.c_int, C.POINTER(BenchStats)]


def run(lib, names, settings, fmt, w, h, flags, host, n):
    protos = (C.c_void_p * len(names))(*[C.addressof(C.c_char.in_dll(lib, x)) for x in names])
    sets = (C.c_char_p * len(names))(*[(s.encode() if s else None) for s in settings])
    st = BenchStats()
    rc = lib.hb_bench_run_chain(len(names), protos, sets, fmt, w, h, flags, host.ctypes.data, host.shape[0], n, C.byref(st))
    if rc != 0:
        raise RuntimeError(f"chain {names} failed rc={rc}")
    return st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="5", choices=sorted(CHAINS))
    ap.add_argument("--frames", type=int, default=24)
    ap.add_argument("--cpu-frames", type=int, default=3)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    args = ap.parse_args()
    ch = CHAINS[args.config]
    w, h, depth = args.width or ch["width"], args.height or ch["height"], ch["depth"]
    fmt = fmt_of(depth)
    flt = C.CDLL(str(handbrake_b200.LIBHBCU_FILTERS))
    core = C.CDLL(str(handbrake_b200.LIBHBCU))
    core.hbcu_last_error.restype = C.c_char_p
    bind(flt)
    flt.hbcu_use_pinned_buffers(1)
    flt.hb_shim_set_zero_buffers(0)
    flt.hb_shim_set_log_level(-1)
    fb = synth.frame_bytes(fmt, w, h)
    core.hbcu_host_reserve.argtypes = [C.c_size_t, C.c_int]
    n = args.frames
    core.hbcu_host_reserve(fb + 4096, 2 * n + 16)
    gen = synth.interlaced_frame if ch["interlaced"] else synth.progressive_frame
    host = np.stack([gen(fmt, w, h, t) for t in range(4)])
    flags = synth.PIC_FLAG_TOP_FIELD_FIRST if ch["interlaced"] else synth.PIC_FLAG_PROGRESSIVE_FRAME
    cuda = [f"hb_filter_{x}_cuda" for x in ch["filters"]]
    out = {"workload": f"chain_{args.config}", "desc": ch["desc"], "width": w, "height": h, "frames": n, "unit": "input frames/s"}
    for arm, names, sets in (("host", cuda, ch["settings"]),
                             ("device", ["hb_filter_hbcu_upload"] + cuda + ["hb_filter_hbcu_download"], [None] + ch["settings"] + [None])):
        run(flt, names, sets, fmt, w, h, flags, host, min(n, 8))                      # warm-up: pools, first-launch costs
        st = run(flt, names, sets, fmt, w, h, flags, host, n)
        out[arm] = {"value": round(n / st.seconds, 2), "seconds": round(st.seconds, 4), "frames_out": int(st.frames_out),
                    "checksum": int(st.checksum)}
    assert out["host"]["checksum"] == out["device"]["checksum"], "host-hopping and device-resident chains disagree"
    ref_so = REPO / "oracle" / "_ref" / "libhbref.so"
    if args.cpu_frames > 0 and ref_so.exists():
        ref = C.CDLL(str(ref_so), mode=C.RTLD_LOCAL)
        bind(ref)
        ref.hb_shim_set_log_level(-1)
        names = [f"hb_filter_{x}" + ("_mt" if x == "lapsharp" else "") for x in ch["filters"]]
        sets = ch["settings"]
        st = run(ref, names, sets, fmt, w, h, flags, host, args.cpu_frames)
        import os
        out["cpu"] = {"value": round(args.cpu_frames / st.seconds, 3), "kind": "reference", "sample": f"{args.cpu_frames} frames",
                      "cores": os.cpu_count()}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
