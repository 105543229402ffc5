"""ctypes binding of hb_harness.h: run libhb-style filter objects over numpy frames.

`FilterLib(path)` wraps any shared library that exports `hb_filter_*` objects and
the harness (`hb_harness_run_chain`): the product's libhbcu_filters.so, or -- from
tests only -- the reference build oracle/_ref/libhbref.so.
"""
import ctypes as C
import numpy as np

from . import synth


class HarnessIO(C.Structure):
    _fields_ = [
        ("pix_fmt", C.c_int), ("width", C.c_int), ("height", C.c_int),
        ("n_in", C.c_int),
        ("in_", C.c_void_p), ("in_flags", C.c_void_p), ("in_combed", C.c_void_p),
        ("out", C.c_void_p), ("out_capacity", C.c_int),
        ("out_combed", C.c_void_p), ("out_flags", C.c_void_p),
        ("out_start", C.c_void_p), ("out_stop", C.c_void_p), ("out_duration", C.c_void_p),
        ("n_out", C.c_int), ("n_dropped", C.c_int), ("saw_eof", C.c_int),
        ("init_failed", C.c_int), ("vrate_num_out", C.c_int), ("vrate_den_out", C.c_int),
    ]


class FilterResult:
    def __init__(self):
        self.frames = None
        self.combed = None
        self.flags = None
        self.start = None
        self.stop = None
        self.duration = None
        self.saw_eof = False
        self.init_failed = 0
        self.vrate = (0, 0)
        self.n_dropped = 0


class FilterLib:
    def __init__(self, path):
        self.path = str(path)
        self.lib = C.CDLL(self.path, mode=C.RTLD_LOCAL)
        self.lib.hb_harness_run_chain.restype = C.c_int
        self.lib.hb_harness_run_chain.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_char_p), C.POINTER(HarnessIO)]
        self.lib.hb_harness_frame_bytes.restype = C.c_size_t
        self.lib.hb_harness_frame_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
        self.lib.hb_shim_buffers_alive.restype = C.c_long
        self.lib.hb_shim_set_log_level.argtypes = [C.c_int]
        self.lib.hb_shim_set_cpu_count.argtypes = [C.c_int]
        self.lib.hb_shim_set_log_level(-1)

    def filter_object(self, name):
        """address of the exported hb_filter_object_t `name` (e.g. 'hb_filter_nlmeans')"""
        return C.addressof(C.c_char.in_dll(self.lib, name))

    def buffers_alive(self):
        return int(self.lib.hb_shim_buffers_alive())

    def set_cpu_count(self, n):
        self.lib.hb_shim_set_cpu_count(int(n))

    def run(self, filters, settings, frames, pix_fmt, width, height, flags=None, combed=None,
            max_out=None, out_scale=1):
        """filters: list of exported object names; settings: list of 'k=v:k=v' strings (or None).
        frames: (n, frame_bytes) uint8 array.  Returns FilterResult."""
        if isinstance(filters, str):
            filters, settings = [filters], [settings]
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        n_in = frames.shape[0]
        fb = synth.frame_bytes(pix_fmt, width, height)
        assert frames.shape[1] == fb, (frames.shape, fb)
        cap = max_out if max_out is not None else (n_in * 2 * out_scale + 8)
        out = np.zeros((cap, fb), dtype=np.uint8)
        o_combed = np.zeros(cap, dtype=np.uint8)
        o_flags = np.zeros(cap, dtype=np.uint16)
        o_start = np.zeros(cap, dtype=np.int64)
        o_stop = np.zeros(cap, dtype=np.int64)
        o_dur = np.zeros(cap, dtype=np.float64)
        io = HarnessIO()
        io.pix_fmt, io.width, io.height, io.n_in = pix_fmt, width, height, n_in
        io.in_ = frames.ctypes.data
        keep = [frames]
        if flags is not None:
            fl = np.ascontiguousarray(flags, dtype=np.uint16); keep.append(fl)
            io.in_flags = fl.ctypes.data
        if combed is not None:
            cb = np.ascontiguousarray(combed, dtype=np.uint8); keep.append(cb)
            io.in_combed = cb.ctypes.data
        io.out, io.out_capacity = out.ctypes.data, cap
        io.out_combed, io.out_flags = o_combed.ctypes.data, o_flags.ctypes.data
        io.out_start, io.out_stop, io.out_duration = o_start.ctypes.data, o_stop.ctypes.data, o_dur.ctypes.data
        n = len(filters)
        protos = (C.c_void_p * n)(*[self.filter_object(f) for f in filters])
        sets = (C.c_char_p * n)(*[(s.encode() if s else None) for s in settings])
        rc = self.lib.hb_harness_run_chain(n, protos, sets, C.byref(io))
        if rc != 0:
            raise RuntimeError(f"filter chain {filters} failed (HB_FILTER_FAILED)")
        r = FilterResult()
        k = io.n_out
        r.frames, r.combed, r.flags = out[:k], o_combed[:k], o_flags[:k]
        r.start, r.stop, r.duration = o_start[:k], o_stop[:k], o_dur[:k]
        r.saw_eof, r.init_failed = bool(io.saw_eof), io.init_failed
        r.vrate, r.n_dropped = (io.vrate_num_out, io.vrate_den_out), io.n_dropped
        return r
