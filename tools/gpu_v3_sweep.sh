#!/bin/bash
# One gpurun call: parity of every v3 NLMeans kernel shape + a short bench line per shape + ncu of two shapes.
# usage: gpurun --timeout 1500 -- bash tools/gpu_v3_sweep.sh
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/v3sweep
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
SHAPES=${SHAPES:-"off 12,12,0 12,21,1 12,18,1 8,28,1"}
for V in $SHAPES; do
  tag=$(echo $V | tr ',' '_')
  echo "=== shape $V" | tee -a $OUT/summary.txt
  HBCU_NLMEANS_V3=$V timeout 300 python -m pytest tests/test_nlmeans_gpu.py tests/test_golden_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "nlmeans or golden" > $OUT/test_$tag.log 2>&1
  echo "pytest rc=$? $(tail -1 $OUT/test_$tag.log)" | tee -a $OUT/summary.txt
  HBCU_NLMEANS_V3=$V timeout 300 python bench.py --steps 6 --warmup 3 --batch 128 --no-cpu-baseline --no-extra --no-copy-only > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  echo "bench rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('$OUT/bench_$tag.json').read().strip().splitlines()[-1])
    print('value',d['value'],'e2e',d['e2e']['value'],'kernel_ms',d['roofline']['kernel_ms_per_frame'],'launch_ms',d['roofline']['launch_ms_avg'])
except Exception as e: print('ERR',e)
PY
)" | tee -a $OUT/summary.txt
done
NCU_SHAPES=${NCU_SHAPES:-"12,21,1 12,12,0"}
for V in $NCU_SHAPES; do
  tag=$(echo $V | tr ',' '_')
  HBCU_NLMEANS_V3=$V timeout 600 ncu --set full --clock-control none --import-source on -k regex:nlmeans_v3 -s 4 -c 1 -o $OUT/ncu_$tag -f python bench.py --steps 1 --warmup 3 --batch 2 --no-cpu-baseline --no-extra --no-copy-only > $OUT/ncu_$tag.log 2>&1
  echo "ncu $V rc=$?" | tee -a $OUT/summary.txt
done
if [ -x handbrake_b200/lib/tools/microbench ]; then handbrake_b200/lib/tools/microbench > $OUT/microbench.txt 2>&1; fi
cat $OUT/summary.txt
