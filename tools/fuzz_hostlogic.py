"""Longer randomised comparisons of the product's host-side filter code (oracle/_ref/libhostlogic.so) with the compiled reference
(oracle/_ref/libhbref.so) than tests/test_hostlogic.py freezes: random settings per filter, random flag / tag streams through chains
(pictures, timestamps, flags, durations, frame rate), random pulldown streams.  CPU only.
usage: python tools/fuzz_hostlogic.py settings|chains|pulldown|multi FIRST_SEED LAST_SEED
  multi: the multi-device dealing of hb_filter_nlmeans_cuda / lapsharp / unsharp (random device lists, block sizes, windows,
         prefilters, clip lengths incl. shorter than a block or the window) against the reference's single stream"""
import sys
from pathlib import Path

REPO = str(Path(__file__).resolve().parent.parent)
MODE = sys.argv.pop(1)
if MODE == "multi":
    sys.path.insert(0, REPO); sys.path.insert(0, REPO + '/tests')
    import numpy as np
    from handbrake_b200 import synth
    from handbrake_b200.hblib import FilterLib
    ref = FilterLib(REPO + '/oracle/_ref/libhbref.so')
    hl = FilterLib(REPO + '/oracle/_ref/libhostlogic.so')
    FMT = {8: synth.PIX_FMT_YUV420P, 10: synth.PIX_FMT_YUV420P10}
    a, b = int(sys.argv[1]), int(sys.argv[2])
    bad = 0
    for seed in range(a, b):
        rng = np.random.default_rng(seed)
        depth = 8 if rng.random() < 0.7 else 10
        fmt = FMT[depth]
        w, h = 40, 24
        n = int(rng.integers(1, 30))
        clip = synth.progressive_clip(fmt, w, h, n, seed=seed)
        parts = [f"y-strength={rng.choice([3, 6, 10])}", "y-patch-size=3", f"y-range={int(rng.choice([1, 3]))}"]
        if rng.random() < 0.7: parts.append(f"y-frame-count={int(rng.integers(1, 6))}")
        if rng.random() < 0.3: parts.append(f"cb-frame-count={int(rng.integers(1, 4))}")
        pre = rng.random() < 0.3
        if pre: parts.append(f"y-prefilter={int(rng.choice([1, 2, 16, 257]))}")
        base = ":".join(parts)
        ndev = int(rng.integers(2, 7))
        multi = f"devices={','.join('0' * ndev)}:block={int(rng.integers(1, 9))}:threads={int(rng.integers(1, 6))}"
        try:
            r = ref.run("hb_filter_nlmeans", base + ":threads=1", clip, fmt, w, h)
            g = hl.run("hb_filter_nlmeans_cuda", base + ":" + multi, clip, fmt, w, h)
            ok = r.frames.shape == g.frames.shape and np.array_equal(r.frames, g.frames) and np.array_equal(r.start, g.start) and g.saw_eof
        except RuntimeError as e:
            ok = False
            print("EXC", e)
        if not ok:
            bad += 1
            print("MISMATCH nlmeans", seed, base, multi, n)
        clip = synth.progressive_clip(fmt, 88, 50, int(rng.integers(1, 40)), seed=seed, noise=15)
        for rname, gname, st in (("hb_filter_lapsharp_mt", "hb_filter_lapsharp_cuda", "y-strength=0.4"), ("hb_filter_unsharp_mt", "hb_filter_unsharp_cuda", "y-strength=0.5:y-size=5")):
            r = ref.run(rname, st, clip, fmt, 88, 50)
            g = hl.run(gname, st + f":devices={','.join('0' * ndev)}", clip, fmt, 88, 50)
            if not (np.array_equal(r.frames, g.frames) and np.array_equal(r.start, g.start)):
                bad += 1
                print("MISMATCH", gname, seed, ndev)
        if hl.buffers_alive() != 0:
            bad += 1
            print("LEAK", seed, hl.buffers_alive())
    print("multi seeds", a, b, "bad", bad)
    sys.exit(1 if bad else 0)

if MODE == "settings":
    sys.path.insert(0, REPO); sys.path.insert(0, REPO + '/tests')
    import numpy as np
    from handbrake_b200 import synth
    from handbrake_b200.hblib import FilterLib
    from test_oracle import decomb_inputs, mixed_interlaced_clip
    ref = FilterLib(REPO + '/oracle/_ref/libhbref.so')
    hl = FilterLib(REPO + '/oracle/_ref/libhostlogic.so')
    FMT = {8: synth.PIX_FMT_YUV420P, 10: synth.PIX_FMT_YUV420P10}
    bad = 0
    def check(tag, rname, gname, s, clip, fmt, w, h, **kw):
        global bad
        try:
            r = ref.run(rname, s, clip, fmt, w, h, **kw)
            g = hl.run(gname, s, clip, fmt, w, h, **kw)
        except RuntimeError as e:
            print("EXC", tag, s, e); bad += 1; return
        ok = (r.init_failed == g.init_failed) and r.frames.shape == g.frames.shape and np.array_equal(r.frames, g.frames) and np.array_equal(r.start, g.start) and list(r.combed) == list(g.combed) and r.vrate == g.vrate
        if not ok:
            bad += 1
            print("MISMATCH", tag, s, r.init_failed, g.init_failed, r.frames.shape, g.frames.shape)
    a, b = int(sys.argv[1]), int(sys.argv[2])
    for seed in range(a, b):
        rng = np.random.default_rng(seed)
        depth = 8 if rng.random() < 0.6 else 10
        fmt = FMT[depth]
        # decomb: any mode 0..63, parity -1/0/1, random flags/tags (non-EEDI2 mostly, EEDI2 sometimes)
        mode = int(rng.integers(0, 64))
        if mode & 8 and rng.random() < 0.6: mode &= ~8
        w, h = (96, 52) if not mode & 8 else (112, 64)
        clip, flags, combed = decomb_inputs(depth, w, h, 5, seed=seed)
        flags = np.array([int(rng.choice([0, 0x08, 0x10, 0x18])) for _ in range(clip.shape[0])], np.uint16)
        combed = rng.integers(0, 3, clip.shape[0]).astype(np.uint8)
        check("decomb", "hb_filter_decomb", "hb_filter_decomb_cuda", f"mode={mode}:parity={int(rng.integers(-1, 2))}", clip, fmt, w, h, flags=flags, combed=combed)
        # comb detect: random settings
        s = f"mode={int(rng.integers(0,4))}:spatial-metric={int(rng.integers(0,3))}:motion-thresh={int(rng.integers(0,8))}:spatial-thresh={int(rng.integers(0,8))}:filter-mode={int(rng.integers(0,3))}:block-thresh={int(rng.integers(1,120))}:block-width={int(rng.integers(4,40))}:block-height={int(rng.integers(4,40))}"
        w, h = 112, 80
        clip = mixed_interlaced_clip(fmt, w, h, 4, seed=seed)
        check("comb", "hb_filter_comb_detect", "hb_filter_comb_detect_cuda", s, clip, fmt, w, h, flags=np.full(clip.shape[0], 8, np.uint16))
        # nlmeans: random per-plane settings on tiny frames
        parts = []
        for c in ("y", "cb", "cr"):
            if c == "y" or rng.random() < 0.5:
                parts.append(f"{c}-strength={rng.choice([0, 1.5, 3, 6, 10, 20])}")
            if rng.random() < 0.4: parts.append(f"{c}-patch-size={int(rng.integers(0, 12))}")
            if rng.random() < 0.4: parts.append(f"{c}-range={int(rng.integers(0, 8))}")
            if rng.random() < 0.4: parts.append(f"{c}-frame-count={int(rng.integers(0, 5))}")
            if rng.random() < 0.3: parts.append(f"{c}-origin-tune={rng.choice([0, 0.005, 0.3, 1, 2.5])}")
            if rng.random() < 0.3: parts.append(f"{c}-prefilter={int(rng.choice([1, 2, 4, 8, 16, 32, 257, 514, 1028, 2049, 1024, 256]))}")
        parts.append("threads=1")
        w, h = 40, 24
        clip = synth.progressive_clip(fmt, w, h, 5, seed=seed)
        check("nlmeans", "hb_filter_nlmeans", "hb_filter_nlmeans_cuda", ":".join(parts), clip, fmt, w, h)
        # lapsharp / unsharp / chroma smooth / hqdn3d
        w, h = 88, 50
        clip = synth.progressive_clip(fmt, w, h, 3, seed=seed, noise=15)
        ks = ["lap", "isolap", "log", "isolog", "bogus"]
        check("lapsharp", "hb_filter_lapsharp_mt", "hb_filter_lapsharp_cuda", f"y-strength={rng.choice([0, 0.2, 1.5, 3])}:y-kernel={rng.choice(ks)}:cb-strength={rng.choice([0, 0.5, 9])}:cr-kernel={rng.choice(ks)}", clip, fmt, w, h)
        check("unsharp", "hb_filter_unsharp_mt", "hb_filter_unsharp_cuda", f"y-strength={rng.choice([-1, 0, 0.25, 1.5, 4])}:y-size={int(rng.integers(0, 20))}:cb-size={int(rng.integers(0, 20))}", clip, fmt, w, h)
        check("chroma", "hb_filter_chroma_smooth_mt", "hb_filter_chroma_smooth_cuda", f"cb-strength={rng.choice([-1, 0, 0.25, 3, 8])}:cb-size={int(rng.integers(0, 20))}:cr-size={int(rng.integers(0, 20))}", clip, fmt, w, h)
        check("hqdn3d", "hb_filter_denoise", "hb_filter_denoise_cuda", f"y-spatial={rng.choice([0, 1, 4, 40, 300])}:cb-temporal={rng.choice([0, 2, 6, 100])}:cr-spatial={rng.choice([0, 3, 9])}", clip, fmt, w, h)
    print("done; bad =", bad)
elif MODE == "chains":
    sys.path.insert(0, REPO); sys.path.insert(0, REPO + '/tests')
    import numpy as np
    from handbrake_b200 import synth
    from handbrake_b200.hblib import FilterLib
    from test_oracle import decomb_inputs
    ref = FilterLib(REPO + '/oracle/_ref/libhbref.so')
    hl = FilterLib(REPO + '/oracle/_ref/libhostlogic.so')
    fmt = synth.PIX_FMT_YUV420P
    bad = 0
    for seed in range(int(sys.argv[1]), int(sys.argv[2])):
        rng = np.random.default_rng(seed)
        w, h = 96, 52
        clip, flags, combed = decomb_inputs(8, w, h, int(rng.integers(1, 7)), seed=seed)
        n = clip.shape[0]
        flags = np.array([int(rng.choice([0, 0x08, 0x10, 0x18, 0x108, 0x100, 0x208])) for _ in range(n)], np.uint16)
        combed = rng.integers(0, 3, n).astype(np.uint8)
        mode = int(rng.integers(0, 64)) & ~8
        chains = [(["hb_filter_decomb"], ["hb_filter_decomb_cuda"], [f"mode={mode}:parity={int(rng.integers(-1,2))}"]),
                  (["hb_filter_comb_detect", "hb_filter_decomb", "hb_filter_nlmeans"], ["hb_filter_comb_detect_cuda", "hb_filter_decomb_cuda", "hb_filter_nlmeans_cuda"], [None, f"mode={mode | 32}", "y-strength=3:y-patch-size=3:y-range=3:threads=1"]),
                  (["hb_filter_detelecine", "hb_filter_decomb"], ["hb_filter_detelecine_cuda", "hb_filter_decomb_cuda"], [None, f"mode={mode}"])]
        for rn, gn, s in chains:
            r = ref.run(rn, s, clip, fmt, w, h, flags=flags, combed=combed)
            g = hl.run(gn, s, clip, fmt, w, h, flags=flags, combed=combed)
            ok = r.frames.shape == g.frames.shape and np.array_equal(r.frames, g.frames) and all(np.array_equal(getattr(r, k), getattr(g, k)) for k in ("start", "stop", "flags", "combed", "duration")) and r.vrate == g.vrate
            if not ok:
                bad += 1
                print("MISMATCH", seed, gn, s, r.frames.shape, g.frames.shape, [k for k in ("start","stop","flags","combed","duration") if not np.array_equal(getattr(r,k), getattr(g,k))], r.vrate, g.vrate)
    print("done; bad =", bad)
else:
    sys.path.insert(0, REPO); sys.path.insert(0, REPO + '/tests')
    import numpy as np
    from handbrake_b200 import synth
    from handbrake_b200.hblib import FilterLib
    ref = FilterLib(REPO + '/oracle/_ref/libhbref.so')
    hl = FilterLib(REPO + '/oracle/_ref/libhostlogic.so')
    w,h=64,48
    fmt = synth.PIX_FMT_YUV420P
    bad = 0
    for seed in range(int(sys.argv[1]), int(sys.argv[2])):
        rng = np.random.default_rng(seed)
        n = 40
        film = [synth.progressive_frame(fmt, w, h, 3*t, seed, int(rng.integers(0, 10))) for t in range(12)]
        frames, flags = [], []
        for i in range(n):
            k = rng.integers(0, 4)
            a, b = film[rng.integers(0, 12)], film[rng.integers(0, 12)]
            if k == 0: f = a
            elif k == 1: f = synth.weave(a, b, fmt, w, h)
            elif k == 2 and frames: f = frames[-1]
            else: f = film[i % 12]
            frames.append(f)
            fl = 0
            if rng.random() < 0.6: fl |= synth.PIC_FLAG_TOP_FIELD_FIRST
            if rng.random() < 0.4: fl |= synth.PIC_FLAG_REPEAT_FIRST_FIELD
            flags.append(fl)
        clip = np.stack(frames); flags = np.array(flags, np.uint16)
        settings = [None, "strict-breaks=0", "strict-breaks=1", "parity=0", "parity=1", "plane=1"][seed % 6]
        r = ref.run("hb_filter_detelecine", settings, clip, fmt, w, h, flags=flags)
        g = hl.run("hb_filter_detelecine_cuda", settings, clip, fmt, w, h, flags=flags)
        ok = r.frames.shape == g.frames.shape and np.array_equal(r.frames, g.frames) and np.array_equal(r.start, g.start)
        if not ok:
            bad += 1
            print("MISMATCH seed", seed, r.frames.shape, g.frames.shape)
    print("done, mismatches:", bad)
