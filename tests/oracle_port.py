"""ctypes binding of oracle/_ref/liboracle_port.so (TEST INFRASTRUCTURE: the C restatement)."""
import ctypes as C
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
PORT_SO = REPO / "oracle" / "_ref" / "liboracle_port.so"


class PlaneParams(C.Structure):
    _fields_ = [("strength", C.c_double), ("origin_tune", C.c_double),
                ("patch_size", C.c_int), ("range", C.c_int), ("nframes", C.c_int), ("prefilter", C.c_int)]


class OraclePort:
    def __init__(self, path=PORT_SO):
        self.lib = C.CDLL(str(path))
        self.lib.oracle_nlmeans_clip.restype = C.c_int
        self.lib.oracle_nlmeans_clip.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                 C.POINTER(PlaneParams), C.c_void_p]

    def nlmeans_clip(self, clip, width, height, depth, params):
        """params: list of 3 dicts(strength, origin_tune, patch_size, range, nframes)"""
        clip = np.ascontiguousarray(clip, dtype=np.uint8)
        out = np.zeros_like(clip)
        pp = (PlaneParams * 3)()
        for c in range(3):
            d = params[c]
            pp[c] = PlaneParams(d.get("strength", 6), d.get("origin_tune", 1), d.get("patch_size", 7),
                                d.get("range", 3), d.get("nframes", 2), d.get("prefilter", 0))
        rc = self.lib.oracle_nlmeans_clip(clip.ctypes.data, clip.shape[0], width, height, depth, pp, out.ctypes.data)
        assert rc == 0
        return out


class CombParams(C.Structure):
    _fields_ = [("mode", C.c_int), ("spatial_metric", C.c_int), ("motion_threshold", C.c_int),
                ("spatial_threshold", C.c_int), ("filter_mode", C.c_int), ("block_threshold", C.c_int),
                ("block_width", C.c_int), ("block_height", C.c_int)]


COMB_DEFAULTS = dict(mode=3, spatial_metric=2, motion_threshold=3, spatial_threshold=3, filter_mode=2,
                     block_threshold=40, block_width=16, block_height=16)      # comb_detect.c:1118-1125
COMB_KEYS = {"mode": "mode", "spatial-metric": "spatial_metric", "motion-thresh": "motion_threshold",
             "spatial-thresh": "spatial_threshold", "filter-mode": "filter_mode", "block-thresh": "block_threshold",
             "block-width": "block_width", "block-height": "block_height"}


def comb_params(settings):
    d = dict(COMB_DEFAULTS)
    if settings:
        for kv in settings.split(":"):
            k, v = kv.split("=")
            d[COMB_KEYS[k]] = int(v)
    return CombParams(**d)


def _comb_bind(self):
    self.lib.oracle_comb_detect_clip.restype = C.c_int
    self.lib.oracle_comb_detect_clip.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(CombParams), C.c_void_p]
    self.lib.oracle_comb_detect.restype = C.c_int
    self.lib.oracle_comb_detect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                            C.POINTER(CombParams), C.c_int, C.c_void_p, C.c_void_p]


def comb_detect_clip(self, clip, width, height, depth, settings=None):
    _comb_bind(self)
    clip = np.ascontiguousarray(clip, dtype=np.uint8)
    v = np.zeros(clip.shape[0], dtype=np.uint8)
    p = comb_params(settings)
    self.lib.oracle_comb_detect_clip(clip.ctypes.data, clip.shape[0], width, height, depth, C.byref(p), v.ctypes.data)
    return v


def comb_detect_masks(self, prev, cur, nxt, width, height, depth, settings=None, force=0):
    """luma planes (2-D arrays) -> (verdict, raw mask, scored mask)"""
    _comb_bind(self)
    p = comb_params(settings)
    a, b, c = (np.ascontiguousarray(x) for x in (prev, cur, nxt))
    m = np.zeros((height, width), np.uint8)
    f = np.zeros((height, width), np.uint8)
    v = self.lib.oracle_comb_detect(a.ctypes.data, b.ctypes.data, c.ctypes.data, width, height, depth, C.byref(p), force,
                                    m.ctypes.data, f.ctypes.data)
    return v, m, f


OraclePort.comb_detect_clip = comb_detect_clip
OraclePort.comb_detect_masks = comb_detect_masks


def decomb_clip(self, clip, width, height, depth, mode, parity=-1, flags=None, combed=None, postproc=1):
    """-> (out frames, source index per output)"""
    clip = np.ascontiguousarray(clip, dtype=np.uint8)
    n = clip.shape[0]
    out = np.zeros((2 * n, clip.shape[1]), np.uint8)
    src = np.zeros(2 * n, np.int32)
    fl = np.ascontiguousarray(flags, np.uint16) if flags is not None else None
    cb = np.ascontiguousarray(combed, np.uint8) if combed is not None else None
    self.lib.oracle_decomb_clip_pp.restype = C.c_int
    self.lib.oracle_decomb_clip_pp.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                               C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    k = self.lib.oracle_decomb_clip_pp(clip.ctypes.data, n, fl.ctypes.data if fl is not None else None,
                                       cb.ctypes.data if cb is not None else None, width, height, depth, mode, parity,
                                       postproc, out.ctypes.data, src.ctypes.data)
    return out[:k], src[:k]


OraclePort.decomb_clip = decomb_clip


def lapsharp_frame(self, frame, fmt_planes, depth, strengths, kernels):
    """frame: packed planar; fmt_planes: [(w,h)]*3.  Builds libhb-style strided planes (64-byte stride,
    zero stride region, 16-bit mirror applied when bps == 2 -- for 8-bit content libhb's mirror is a no-op),
    runs the port per plane, returns the packed result."""
    bps = 2 if depth > 8 else 1
    dt = np.uint16 if bps == 2 else np.uint8
    self.lib.oracle_lapsharp_plane.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
    a = frame.view(dt)
    out, off = [], 0
    for c, (w, h) in enumerate(fmt_planes):
        stride_el = ((w * bps + 63) // 64 * 64) // bps
        src = np.zeros((h + 1, stride_el), dt)          # one spare row: the mirror writes the start of the next row
        src[:h, :w] = a[off:off + w * h].reshape(h, w)
        if bps == 2:
            margin = stride_el - w
            mf, mb = margin // 2, margin - margin // 2
            flat = src.reshape(-1)
            for yy in range(h):
                pos = yy * stride_el + w
                for ii in range(mb):
                    flat[pos + ii] = flat[pos - ii - 1]
                pos = (yy + 1) * stride_el - 1
                for ii in range(mf):
                    flat[pos - ii] = flat[pos + ii + 1]
        dst = np.zeros((h, stride_el), dt)
        self.lib.oracle_lapsharp_plane(src.ctypes.data, dst.ctypes.data, w, h, stride_el * bps, stride_el * bps, depth,
                                       kernels[c], float(strengths[c]))
        out.append(dst[:, :w].reshape(-1))
        off += w * h
    return np.concatenate(out).view(np.uint8)


OraclePort.lapsharp_frame = lapsharp_frame


def unsharp_clip(self, clip, w, h, depth, strength, size, smooth):
    """libhb/unsharp.c (smooth=0) / chroma_smooth.c (smooth=1) on packed yuv420p frames; strength/size are the per-plane
    values after the filter's cascade and defaults"""
    self.lib.oracle_unsharp_clip.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int),
                                             C.c_int, C.c_void_p]
    clip = np.ascontiguousarray(clip, dtype=np.uint8)
    out = np.zeros_like(clip)
    st = (C.c_double * 3)(*strength)
    sz = (C.c_int * 3)(*size)
    self.lib.oracle_unsharp_clip(clip.ctypes.data, clip.shape[0], w, h, depth, st, sz, int(smooth), out.ctypes.data)
    return out


OraclePort.unsharp_clip = unsharp_clip


def hqdn3d_clip(self, clip, w, h, depth, strengths):
    """libhb/denoise.c on packed yuv420p frames; strengths = [y-spatial, y-temporal, cb-spatial, cb-temporal, cr-spatial,
    cr-temporal] after the filter's defaults"""
    self.lib.oracle_hqdn3d_clip.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p]
    clip = np.ascontiguousarray(clip, dtype=np.uint8)
    out = np.zeros_like(clip)
    st = (C.c_double * 6)(*strengths)
    self.lib.oracle_hqdn3d_clip(clip.ctypes.data, clip.shape[0], w, h, depth, st, out.ctypes.data)
    return out


OraclePort.hqdn3d_clip = hqdn3d_clip
