import os, sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import handbrake_b200
from handbrake_b200 import synth
flt = handbrake_b200.filters()
F = synth.PIX_FMT_YUV420P
w, h, n = 200, 120, 9
clip = synth.progressive_clip(F, w, h, n, seed=31)
for pre in (1, 2, 4, 16, 256):
  for nf in (2, 3):
    for blk in (2, 3, 4):
        s = f"y-strength=6:y-prefilter={pre}:y-frame-count={nf}"
        one = flt.run("hb_filter_nlmeans_cuda", s, clip, F, w, h)
        many = flt.run("hb_filter_nlmeans_cuda", s + f":devices=0,0:block={blk}", clip, F, w, h)
        d = (one.frames != many.frames)
        per = [int(x.sum()) for x in d]
        # per plane
        ysz = w*h
        print(f"pre={pre} nf={nf} block={blk} differing bytes per frame {per}  luma-only={[int(x[:ysz].sum()) for x in d]}", flush=True)
