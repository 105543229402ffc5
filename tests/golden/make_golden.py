"""Generates tests/golden/nlmeans_golden.json from the reference itself.

The reference tree holds no golden vectors for this path, so the
goldens are per-frame SHA-256 digests of the output of the UNMODIFIED reference
filter (oracle/_ref/libhbref.so, built by oracle/Makefile from /root/reference)
on frames from handbrake_b200.synth.  Run in the build container:
    python tests/golden/make_golden.py
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
from handbrake_b200 import synth  # noqa: E402
from handbrake_b200.hblib import FilterLib  # noqa: E402

FMT = {8: synth.PIX_FMT_YUV420P, 10: synth.PIX_FMT_YUV420P10}

CASES = {
    # BASELINE.json configs[0]: 640x360 8-bit, NLMeans light, 10 frames
    "config1_640x360_light": dict(width=640, height=360, depth=8, frames=10, seed=12345, settings="y-strength=3"),
    "medium_320x180": dict(width=320, height=180, depth=8, frames=6, seed=12345, settings="y-strength=6"),
    "strong_10bit_256x144": dict(width=256, height=144, depth=10, frames=5, seed=99, settings="y-strength=10"),
    "tiny_light_96x64": dict(width=96, height=64, depth=8, frames=4, seed=1, settings="y-strength=3"),
    "tiny_tuned_10bit_64x48": dict(width=64, height=48, depth=10, frames=4, seed=2,
                                   settings="y-strength=6:y-origin-tune=0.5:y-patch-size=5:y-range=5:y-frame-count=3:cb-strength=4:cb-range=3"),
}


def main():
    ref = FilterLib(REPO / "oracle" / "_ref" / "libhbref.so")
    out = {}
    for name, c in CASES.items():
        clip = synth.progressive_clip(FMT[c["depth"]], c["width"], c["height"], c["frames"], seed=c["seed"])
        r = ref.run("hb_filter_nlmeans", c["settings"] + ":threads=2", clip, FMT[c["depth"]], c["width"], c["height"])
        assert r.frames.shape[0] == c["frames"]
        out[name] = dict(c, sha256=[hashlib.sha256(f.tobytes()).hexdigest() for f in r.frames])
        print(name, out[name]["sha256"][0][:16])
    (Path(__file__).parent / "nlmeans_golden.json").write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
