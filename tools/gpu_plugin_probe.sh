#!/bin/bash
# gpurun --gpus 2: what bounds the one-process multi-device arm?  The child of bench.py under a few settings.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/pluginprobe
mkdir -p $OUT
run() { tag=$1; shift; echo "== $tag: $*" >> $OUT/probe.txt; env "$@" timeout 300 python bench.py --plugin-multi-child $DEVS --plugin-frames 4096 --plugin-warm 512 --block $BLOCK --inflight $INFL 2>> $OUT/probe.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: d[k] for k in ('value', 'devices', 'block', 'gb_s_per_direction_total', 'pinned_numa', 'mempolicy')})" >> $OUT/probe.txt; }
DEVS=0 BLOCK=8 INFL=6 run one_device A=1
DEVS=0,1 BLOCK=8 INFL=6 run two_devices_default A=1
DEVS=0,1 BLOCK=32 INFL=6 run block32 A=1
DEVS=0,1 BLOCK=8 INFL=14 run inflight14 A=1
DEVS=0,1 BLOCK=2 INFL=6 run block2 A=1
DEVS=0,1 BLOCK=8 INFL=6 run nointerleave HBCU_BENCH_NO_INTERLEAVE=1
DEVS=0,1 BLOCK=8 INFL=6 run wc_inputs HBCU_BENCH_WC_INPUT=1
DEVS=0,0 BLOCK=8 INFL=6 run two_handles_one_gpu A=1
cat $OUT/probe.txt
