#!/usr/bin/env python
"""tools/eedi2_shard_probe.py -- SURVEY.md 8e option A' measured: frame-sharded decomb EEDI2 where
every block re-runs the filter over K extra leading frames (their outputs dropped) so that the carried edge-mask state
(templates/eedi2_template.c:132: only the top half of the mask is cleared per field; decomb_template.c:391-397) has
K frames = 2K fields (bob) to converge before the first owned field.  Reports, per K, the bytes and frames that differ
from the unsharded stream.  K = 0 is option B (state zeroed per block).  Runs the PRODUCT's CUDA objects; needs a GPU."""
import json
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
import handbrake_b200  # noqa: E402
from handbrake_b200 import sharding, synth  # noqa: E402


def main():
    flt = handbrake_b200.filters()
    out = []
    for (w, h, depth, n, mode, name) in ((720, 480, 8, 24, 31, "eedi2 bob"), (1920, 1080, 10, 16, 31, "eedi2 bob"), (720, 480, 8, 24, 15, "eedi2 (one picture per frame)")):
        fmt = synth.PIX_FMT_YUV420P if depth == 8 else synth.PIX_FMT_YUV420P10
        clip = synth.interlaced_clip(fmt, w, h, n, seed=9)
        flags = np.full(clip.shape[0], synth.PIC_FLAG_TOP_FIELD_FIRST, np.uint16)
        n = clip.shape[0]
        k_out = 2 if mode & 16 else 1
        whole = flt.run("hb_filter_decomb_cuda", f"mode={mode}", clip, fmt, w, h, flags=flags).frames
        assert whole.shape[0] == n * k_out
        for world, block in ((2, 4), (8, 2)):
            for K in (1, 2, 3):      # decomb itself needs one previous frame: K = 1 is the minimum (2 fields of mask warm-up)
                blocks = sharding.plan_blocks(n, world, block, halo_before=K, halo_after=1)
                win = {}

                def clip_of(a, b):
                    win["r"] = (a, b)
                    return clip[a:b]

                def run(fr):
                    a, b = win["r"]
                    return flt.run("hb_filter_decomb_cuda", f"mode={mode}", fr, fmt, w, h, flags=flags[a:b]).frames
                parts = {}
                for rank in range(world):
                    parts.update(sharding.run_rank(blocks, rank, clip_of, run, outputs_per_frame=k_out))
                got = np.concatenate([parts[b.index] for b in blocks])
                d = got != whole
                out.append({"clip": f"{w}x{h} {depth}-bit, {n} frames, decomb mode {mode} ({name})", "world": world, "block_frames": block,
                            "warm_up_frames": K, "bytes_differing": int(d.sum()), "pictures_differing": int(d.any(axis=1).sum()),
                            "pictures": int(whole.shape[0]), "max_abs": int(np.abs(got.astype(np.int32) - whole.astype(np.int32)).max())})
                print(json.dumps(out[-1]), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
