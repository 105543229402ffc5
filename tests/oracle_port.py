"""ctypes binding of oracle/_ref/liboracle_port.so (TEST INFRASTRUCTURE: the C restatement)."""
import ctypes as C
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
PORT_SO = REPO / "oracle" / "_ref" / "liboracle_port.so"


class PlaneParams(C.Structure):
    _fields_ = [("strength", C.c_double), ("origin_tune", C.c_double),
                ("patch_size", C.c_int), ("range", C.c_int), ("nframes", C.c_int)]


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
                                d.get("range", 3), d.get("nframes", 2))
        rc = self.lib.oracle_nlmeans_clip(clip.ctypes.data, clip.shape[0], width, height, depth, pp, out.ctypes.data)
        assert rc == 0
        return out
