#!/bin/bash
# One gpurun call (1 GPU): the whole -m gpu suite after the EEDI2 fusion, EEDI2 throughput, wall time of the default bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/checkf
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$? $(tail -1 $OUT/pytest_gpu.log)" | tee $OUT/summary.txt
timeout 600 python tools/bench_filters.py --only 4k10_decomb_eedi2bob --frames 48 --cpu-frames 0 > $OUT/bench_eedi2.jsonl 2> $OUT/bench_eedi2.err
echo "eedi2 bench rc=$?" | tee -a $OUT/summary.txt
T0=$(date +%s)
timeout 1500 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$? wall $(( $(date +%s) - T0 )) s" | tee -a $OUT/summary.txt
python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1
echo "smoke rc=$? $(tail -1 $OUT/smoke.log)" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
