import os, sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import handbrake_b200
from handbrake_b200 import synth
flt = handbrake_b200.filters()
F = synth.PIX_FMT_YUV420P
w, h = 200, 120
for n in (7, 9):
    clip = synth.progressive_clip(F, w, h, n, seed=31)
    for s, m in (("y-strength=6:y-prefilter=1:y-frame-count=2", "devices=0,0:block=2"), ("y-strength=6:y-frame-count=2", "devices=0,0:block=2"),
                 ("y-strength=6:y-prefilter=1:y-frame-count=2", "threads=4")):
        ref = flt.run("hb_filter_nlmeans_cuda", s, clip, F, w, h).frames
        bad = {}
        for it in range(40):
            many = flt.run("hb_filter_nlmeans_cuda", s + ":" + m, clip, F, w, h)
            d = (ref != many.frames)
            key = tuple(int(i) for i in np.argwhere(d.any(axis=1)).ravel())
            bad[key] = bad.get(key, 0) + 1
        print(f"n={n} {s} + {m}: differing-frame sets over 40 runs: {bad}", flush=True)
