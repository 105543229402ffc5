"""tools/time_chain.py -- times one CUDA filter object end to end (host hb_buffer_t in/out) on synthetic frames.
usage: python tools/time_chain.py <filter_symbol> <settings|-> <width> <height> <depth> <frames> [interlaced]"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import handbrake_b200  # noqa: E402
from handbrake_b200 import synth  # noqa: E402
from bench import BenchStats, bind_bench, fmt_of  # noqa: E402

sym, settings, w, h, depth, n = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
interlaced = len(sys.argv) > 7
fmt = fmt_of(depth)
flt = C.CDLL(str(handbrake_b200.LIBHBCU_FILTERS))
core = C.CDLL(str(handbrake_b200.LIBHBCU))
core.hbcu_last_error.restype = C.c_char_p
bind_bench(flt)
flt.hbcu_use_pinned_buffers(1)
fb = synth.frame_bytes(fmt, w, h)
core.hbcu_host_reserve.argtypes = [C.c_size_t, C.c_int]
core.hbcu_host_reserve(fb + 4096, 3 * n + 24)
gen = synth.interlaced_frame if interlaced else synth.progressive_frame
host = np.stack([gen(fmt, w, h, t) for t in range(4)])
proto = C.addressof(C.c_char.in_dll(flt, sym))
s = None if settings == "-" else settings.encode()
st = BenchStats()
for rep in range(2):
    b = flt.hb_bench_open(proto, s, fmt, w, h)
    assert b, core.hbcu_last_error()
    assert flt.hb_bench_run(b, host.ctypes.data, 4, n, C.byref(st)) == 0
    print(f"{sym} {settings} {w}x{h} d{depth}: {st.frames_out} frames out in {st.seconds*1e3:.2f} ms -> "
          f"{n/st.seconds:.1f} input fps, {st.seconds*1e3/n:.3f} ms/input frame (rep {rep})", flush=True)
