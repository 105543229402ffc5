#!/usr/bin/env python
"""bench.py -- hot-path benchmark (contract in the task statement, section 4).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the NLMeans hot path over one batch of `--batch` synthetic frames per GPU
(weak scaling: every rank filters its own contiguous block of frames, with its own temporal halo).

  value     frames/s, whole job, inputs already resident in HBM (pad + NLMeans kernels, CUDA events)
  e2e       frames/s through hb_filter_nlmeans_cuda.work() with HOST hb_buffer_t frames:
            pinned H2D + kernels + D2H into fresh output buffers, all inside the timed region
  roofline  algorithmic HBM bytes of the NLMeans kernel (3 x frame bytes per output frame,
            SURVEY.md 8d) / its CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline / --impl reference
            the reference's own nlmeans.c + nlmeans_x86.c (oracle/_ref/libhbref.so, compiled
            unmodified from /root/reference) on this box's host cores, same frames, same protocol
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

from handbrake_b200 import synth  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[3]: the configuration the metric ("4K NLMeans frames/sec at 1/2/4/8 B200") is quoted on
    "4k_nlmeans_strong": dict(width=3840, height=2160, depth=8, settings="y-strength=10",
                              desc="3840x2160 yuv420p 8-bit, NLMeans 'strong' (strength 10, patch 7, range 3, frames 2)"),
    "4k_nlmeans_medium": dict(width=3840, height=2160, depth=8, settings="y-strength=6",
                              desc="3840x2160 yuv420p 8-bit, NLMeans 'medium'"),
    "4k_nlmeans_strong_animation": dict(width=3840, height=2160, depth=8,
                                        settings="y-strength=10:y-origin-tune=0.15:y-patch-size=5:y-range=7:y-frame-count=4",
                                        desc="3840x2160 8-bit, NLMeans strong + large search window (patch 5, range 7, frames 4)"),
    # BASELINE.json configs[1]
    "1080p_nlmeans_medium": dict(width=1920, height=1080, depth=8, settings="y-strength=6",
                                 desc="1920x1080 yuv420p 8-bit, NLMeans 'medium'"),
    # BASELINE.json configs[0] (the reference's CPU-runnable case)
    "360p_nlmeans_light": dict(width=640, height=360, depth=8, settings="y-strength=3",
                               desc="640x360 yuv420p 8-bit, NLMeans 'light'"),
    # diagnostic: every plane bypassed (strength 0) -> the transfer pipeline alone
    "4k_copy_only": dict(width=3840, height=2160, depth=8, settings="y-strength=0",
                         desc="3840x2160 yuv420p 8-bit, NLMeans strength 0 (bypass copy): transfer pipeline diagnostic"),
    "4k10_nlmeans_medium": dict(width=3840, height=2160, depth=10, settings="y-strength=6",
                                desc="3840x2160 yuv420p10 NLMeans 'medium'"),
}
NFRAMES = 2   # temporal window of every preset (param.c:408-428)


class BenchStats(C.Structure):
    _fields_ = [("seconds", C.c_double), ("frames_out", C.c_int64), ("bytes_in", C.c_int64),
                ("bytes_out", C.c_int64), ("checksum", C.c_uint64)]


class NlmPlane(C.Structure):
    _fields_ = [("patch_size", C.c_int), ("range", C.c_int), ("nframes", C.c_int), ("bypass", C.c_int),
                ("origin_tune", C.c_double), ("weight_fact", C.c_float), ("diff_max", C.c_int),
                ("exptable", C.c_float * 128), ("prefilter", C.c_int)]


class NlmConfig(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("depth", C.c_int), ("chroma_shift_w", C.c_int),
                ("chroma_shift_h", C.c_int), ("device", C.c_int), ("ring_frames", C.c_int), ("out_slots", C.c_int),
                ("plane", NlmPlane * 3)]


def fmt_of(depth):
    return synth.PIX_FMT_YUV420P if depth == 8 else synth.PIX_FMT_YUV420P10


def bind_bench(lib):
    lib.hb_bench_open.restype = C.c_void_p
    lib.hb_bench_open.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
    lib.hb_bench_run.restype = C.c_int
    lib.hb_bench_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(BenchStats)]
    lib.hb_shim_set_log_level.argtypes = [C.c_int]
    lib.hb_shim_set_log_level(-1)
    lib.hb_shim_set_zero_buffers(0)     # like libhb's buffer pool: recycled, not zeroed
    lib.hb_get_cpu_count.restype = C.c_int


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def measured_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def bind_to_gpu_numa_node(local_rank):
    """one process per GPU: keep the process (and the pinned frame buffers it is about to allocate) on the
    NUMA node the GPU's PCIe link hangs off; best effort, silently skipped when sysfs does not say"""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip()
        bdf = out.lower()
        if bdf.startswith("00000000:"):
            bdf = bdf[4:]
        cpus = Path(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read_text().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        if ids:
            os.sched_setaffinity(0, ids)
            return cpus
    except Exception:
        pass
    try:
        # containers often hide the PCI sysfs tree: `nvidia-smi topo -m` prints the same list in its "CPU Affinity" column
        import re
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
        for line in topo.splitlines():
            line = re.sub(r"\x1b\[[0-9;]*m", "", line)
            tok = line.split()
            if not tok or tok[0] != f"GPU{local_rank}":
                continue
            for t in tok[1:]:
                if re.fullmatch(r"\d+(-\d+)?(,\d+(-\d+)?)*", t) and ("-" in t or "," in t):
                    ids = set()
                    for part in t.split(","):
                        a, _, b = part.partition("-")
                        ids.update(range(int(a), int(b or a) + 1))
                    ids &= os.sched_getaffinity(0)
                    if ids:
                        os.sched_setaffinity(0, ids)
                        return t
    except Exception:
        pass
    return None


def nlm_config(flt, wl, device, ring, out_slots):
    """device configuration from the settings string, built by the filter's own init code
    (hb_nlmeans_cuda_build_config in nlmeans_cuda.c: the numeric contract of nlmeans.c:343-358)"""
    cfg = NlmConfig()
    flt.hb_parse_filter_settings.restype = C.c_void_p
    flt.hb_parse_filter_settings.argtypes = [C.c_char_p]
    flt.hb_nlmeans_cuda_build_config.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(NlmConfig),
                                                 C.c_void_p, C.c_void_p, C.c_void_p]
    d = flt.hb_parse_filter_settings(wl["settings"].encode())
    if flt.hb_nlmeans_cuda_build_config(d, fmt_of(wl["depth"]), wl["width"], wl["height"], C.byref(cfg), None, None, None) != 0:
        raise RuntimeError("bad NLMeans settings")
    cfg.device, cfg.ring_frames, cfg.out_slots = device, ring, out_slots
    return cfg


def run_ours(args, wl, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import handbrake_b200

    handbrake_b200.require_native()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    os.environ["HBCU_DEVICE"] = str(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    W, H, depth = wl["width"], wl["height"], wl["depth"]
    fmt = fmt_of(depth)
    B, K, Wm = args.batch, args.steps, args.warmup
    # the e2e arm keeps one pinned host buffer per input frame of the timed run: bound that to 320 frames per rank
    if (K + Wm) * B > 320:
        B = max(2, 320 // (K + Wm))
    fb = synth.frame_bytes(fmt, W, H)
    n_unique = 4
    # every rank gets its own block of the clip (frame-sharded job): seeds differ per rank
    host = np.stack([synth.progressive_frame(fmt, W, H, rank * 1000 + t) for t in range(n_unique)])

    core = C.CDLL(str(handbrake_b200.LIBHBCU))
    core.hbcu_last_error.restype = C.c_char_p
    core.hbcu_kernel_launches.restype = C.c_uint64
    flt = C.CDLL(str(handbrake_b200.LIBHBCU_FILTERS))
    bind_bench(flt)
    flt.hbcu_use_pinned_buffers(1)

    def ck(rc):
        if rc != 0:
            raise RuntimeError(core.hbcu_last_error().decode())

    # ---------------- value: inputs resident in HBM ----------------
    nin = B + NFRAMES - 1                     # frames a step reads (block + temporal halo)
    cfg = nlm_config(flt, wl, local_rank, ring=nin + 4, out_slots=4)
    h = C.c_void_p()
    ck(core.hbcu_nlmeans_create(C.byref(h), C.byref(cfg)))
    bps = 2 if depth > 8 else 1
    dims = synth.plane_dims(W, H)
    dev_frames = []
    for i in range(nin):                      # nin distinct device buffers: a step's inputs exceed L2
        t = torch.from_numpy(host[i % n_unique]).cuda()
        dev_frames.append(t)
    input_mb = nin * fb / 1e6

    def plane_ptrs(t):
        base, off, ptrs, strides = t.data_ptr(), 0, [], []
        for (w, hh) in dims:
            ptrs.append(base + off); strides.append(w * bps); off += w * hh * bps
        return (C.c_void_p * 3)(*ptrs), (C.c_int * 3)(*strides)

    ptrs = [plane_ptrs(t) for t in dev_frames]
    idx = [0]

    def step_device():
        base = idx[0]
        for i in range(nin):
            ck(core.hbcu_nlmeans_upload_device(h, C.c_int64(base + i), ptrs[i][0], ptrs[i][1]))
        for i in range(B):
            ck(core.hbcu_nlmeans_filter_device(h, C.c_int64(base + i), NFRAMES, None, None))
        idx[0] = base + nin

    for _ in range(Wm):
        step_device()
    ck(core.hbcu_nlmeans_sync(h))
    barrier()
    clk = ClockSampler(local_rank)
    clk.start()
    launches0 = core.hbcu_kernel_launches()
    ck(core.hbcu_nlmeans_mark(h, 0))
    for _ in range(K):
        step_device()
    ck(core.hbcu_nlmeans_mark(h, 1))
    ms = C.c_float()
    ck(core.hbcu_nlmeans_elapsed_ms(h, C.byref(ms)))
    ck(core.hbcu_nlmeans_sync(h))
    launches = int(core.hbcu_kernel_launches() - launches0)
    barrier()
    kms, kcalls = C.c_float(), C.c_int()
    ck(core.hbcu_nlmeans_kernel_ms(h, C.byref(kms), C.byref(kcalls)))
    dev_ms = max_over_ranks(float(ms.value))
    value = world * K * B / (dev_ms / 1e3)

    # ---------------- ordered gather to the muxer rank over NCCL p2p (north-star; only for N > 1) ----------------
    gather = None
    if world > 1:
        core.hbcu_nlmeans_filter_into.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        core.hbcu_nlmeans_stream_wait.argtypes = [C.c_void_p, C.c_void_p]
        block = torch.empty((B, fb), dtype=torch.uint8, device="cuda")            # this rank's finished block
        out_ptrs = [plane_ptrs(block[i]) for i in range(B)]
        if rank == 0:
            recv = [torch.empty((B, fb), dtype=torch.uint8, device="cuda") for _ in range(world - 1)]
            host_out = torch.empty((world, B, fb), dtype=torch.uint8).pin_memory()

        def step_gather():
            base = idx[0]
            for i in range(nin):
                ck(core.hbcu_nlmeans_upload_device(h, C.c_int64(base + i), ptrs[i][0], ptrs[i][1]))
            for i in range(B):
                ck(core.hbcu_nlmeans_filter_into(h, base + i, NFRAMES, out_ptrs[i][0], out_ptrs[i][1]))
            idx[0] = base + nin
            ck(core.hbcu_nlmeans_stream_wait(h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            if rank == 0:
                host_out[0].copy_(block, non_blocking=True)
                for src in range(1, world):          # stream order = frame order: block of rank 1, then 2, ...
                    dist.recv(recv[src - 1], src=src)
                    host_out[src].copy_(recv[src - 1], non_blocking=True)
            else:
                dist.send(block, dst=0)

        for _ in range(Wm):
            step_gather()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            step_gather()
        barrier()
        g_s = max_over_ranks(time.perf_counter() - t0)
        gather = {"value": round(world * K * B / g_s, 2), "unit": "frames/s",
                  "via": "ncclSend/ncclRecv (torch.distributed p2p over NVLink) of finished frames to rank 0, D2H on rank 0 only",
                  "nvlink_bytes_per_step": int((world - 1) * B * fb), "d2h_bytes_per_step_rank0": int(world * B * fb)}
        del block
    core.hbcu_nlmeans_destroy(h)
    del dev_frames
    torch.cuda.empty_cache()

    # ---------------- e2e: host hb_buffer_t frames through the filter object ----------------
    proto = C.addressof(C.c_char.in_dll(flt, "hb_filter_nlmeans_cuda"))
    settings = (wl["settings"] + f":threads={args.inflight}").encode()
    warm = flt.hb_bench_open(proto, settings, fmt, W, H)
    timed = flt.hb_bench_open(proto, settings, fmt, W, H)
    if not warm or not timed:
        raise RuntimeError("hb_filter_nlmeans_cuda.init failed: " + core.hbcu_last_error().decode())
    st = BenchStats()
    # steady state of a running libhb pipeline: every frame buffer comes recycled from the pool
    # (fifo.c:70-135); cudaHostAlloc itself costs milliseconds and must not be on the clock
    core.hbcu_host_reserve.argtypes = [C.c_size_t, C.c_int]
    core.hbcu_host_reserve(fb + 4096, K * B + 24)
    if flt.hb_bench_run(warm, host.ctypes.data, n_unique, max(Wm, 1) * B, C.byref(st)) != 0:
        raise RuntimeError("warm-up stream failed")
    barrier()
    if flt.hb_bench_run(timed, host.ctypes.data, n_unique, K * B, C.byref(st)) != 0:
        raise RuntimeError("timed stream failed")
    barrier()
    clocks = clk.stop()
    e2e_s = max_over_ranks(st.seconds)
    assert st.frames_out == K * B, (st.frames_out, K * B)
    e2e = world * K * B / e2e_s

    peak, peak_src = measured_peaks()
    alg_bytes_per_frame = (NFRAMES + 1) * fb
    # consecutive frames are launched on two alternating compute streams, so two launches are normally in flight:
    # the per-launch figure is the average event-pair duration divided by the measured concurrency
    launch_ms = kms.value / max(kcalls.value, 1)
    in_flight = max(1.0, kms.value / float(ms.value))
    kern_ms_per_frame = launch_ms / in_flight
    achieved = alg_bytes_per_frame / (kern_ms_per_frame / 1e3) / 1e9 if kcalls.value else None

    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this workload
    # (not measured live: a run under ncu is never a bench value)
    traffic, traffic_src = None, None
    cap = REPO / "profiles" / "r01g_nlmeans_fused_ncu.json"
    if args.workload == "4k_nlmeans_strong" and cap.exists():
        traffic = int(json.loads(cap.read_text())["dram_bytes_total"])
        traffic_src = "profiles/r01g_nlmeans_fused_ncu.json (dram__bytes_read.sum + dram__bytes_write.sum, one launch = one frame)"
    out = {
        "metric": "4K NLMeans frames/sec" if W == 3840 else "NLMeans frames/sec",
        "value": round(value, 2), "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": round(dev_ms / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8" if depth == 8 else "u16", "data": "synthetic",
        "config": {"workload": args.workload, "desc": wl["desc"], "frames_per_step_per_gpu": B,
                   "sharding": f"frame blocks x{world}, {NFRAMES - 1}-frame temporal halo per block, no data-path collective",
                   "cpu_affinity": numa_cpus,
                   "l2": f"step inputs {input_mb:.0f} MB in distinct buffers > 126 MB L2" if input_mb > 126 else f"step inputs {input_mb:.0f} MB (fits L2)"},
        "e2e": {"value": round(e2e, 2), "unit": "frames/s", "h2d_bytes_per_step": int(world * st.bytes_in // K),
                "d2h_bytes_per_step": int(world * st.bytes_out // K), "seconds": round(e2e_s, 4), "checksum": int(st.checksum)},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1) if achieved else None, "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4) if achieved else None, "traffic": traffic,
                     "traffic_source": traffic_src,
                     "kernel": ("nlmeans_fast8_kernel" if depth == 8 else "nlmeans_fast16_kernel") + " (all tiles of Y, U, V of one frame in one launch)",
                     "kernel_ms_per_frame": round(kern_ms_per_frame, 4), "launch_ms_avg": round(launch_ms, 4),
                     "launches_in_flight": round(in_flight, 2), "algorithmic_bytes_per_frame": alg_bytes_per_frame,
                     "peak_source": peak_src,
                     "note": "NLMeans is instruction-issue bound on B200, not HBM bound (DESIGN.md): frac is the honest HBM fraction, not the kernel's quality"},
    }
    if gather is not None:
        out["gather"] = gather
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_reference(args, wl, steps=1, warmup=0)["cpu_baseline"]
    if world > 1:
        dist.destroy_process_group()
    return out


def cpu_reference(args, wl, steps, warmup):
    """times the unmodified reference filter (oracle/_ref/libhbref.so) on the host cores"""
    so = REPO / "oracle" / "_ref" / "libhbref.so"
    if not so.exists():
        raise SystemExit("oracle/_ref/libhbref.so missing: run __graft_entry__.build() where /root/reference exists")
    ref = C.CDLL(str(so))
    bind_bench(ref)
    W, H, depth = wl["width"], wl["height"], wl["depth"]
    fmt = fmt_of(depth)
    ncpu = ref.hb_get_cpu_count()
    threads = ncpu // 2 if ncpu >= 32 else (ncpu // 4) * 3 if ncpu >= 16 else ncpu     # nlmeans.c:362-373
    n_unique = 4
    host = np.stack([synth.progressive_frame(fmt, W, H, t) for t in range(n_unique)])
    # one step = one full taskset cycle (threads frames in parallel) + the serial EOF flush of the look-ahead frames
    frames_per_step = args.ref_frames if args.ref_frames else threads + NFRAMES
    proto = C.addressof(C.c_char.in_dll(ref, "hb_filter_nlmeans"))
    times = []
    st = BenchStats()
    for s in range(warmup + steps):
        b = ref.hb_bench_open(proto, wl["settings"].encode(), fmt, W, H)
        if ref.hb_bench_run(b, host.ctypes.data, n_unique, frames_per_step, C.byref(st)) != 0:
            raise RuntimeError("reference stream failed")
        assert st.frames_out == frames_per_step
        if s >= warmup:
            times.append(st.seconds)
    total = sum(times)
    fps = steps * frames_per_step / total
    sample = f"{frames_per_step} frames of {wl['desc']} per step ({threads} frames in one parallel taskset cycle + {NFRAMES} in the serial EOF flush)"
    return {"value": fps, "ms_per_step": 1e3 * total / steps, "frames_per_step": frames_per_step,
            "cpu_baseline": {"value": round(fps, 4), "unit": "frames/s", "cores": threads, "kind": "reference",
                             "sample": sample, "host_logical_cpus": ncpu,
                             "what": "HandBrake libhb nlmeans.c + nlmeans_x86.c (SSE2) compiled unmodified, gcc -O3 -msse2, its own thread heuristic"}}


def run_reference(args, wl, rank, world):
    if rank != 0:
        return None
    r = cpu_reference(args, wl, args.steps, args.warmup)
    return {
        "impl": "reference",
        "metric": "4K NLMeans frames/sec" if wl["width"] == 3840 else "NLMeans frames/sec",
        "value": round(r["value"], 4), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(r["ms_per_step"], 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8" if wl["depth"] == 8 else "u16", "data": "synthetic",
        "config": {"workload": args.workload, "desc": wl["desc"], "frames_per_step": r["frames_per_step"]},
        "cpu_baseline": r["cpu_baseline"],
        "e2e": {"value": round(r["value"], 4), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="4k_nlmeans_strong", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=32, help="frames per step per GPU")
    ap.add_argument("--ref-frames", type=int, default=0, help="frames per reference step (default: threads + 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=6, help="frames in flight in the e2e arm (the filter's `threads` setting)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                     # timing rule: at least three warm-up steps
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        out = run_reference(args, wl, rank, world)
    else:
        out = run_ours(args, wl, rank, world, local_rank)
    if rank == 0 and out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
