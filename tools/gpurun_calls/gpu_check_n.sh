#!/bin/bash
# One gpurun call (1 GPU): whole -m gpu suite on the current build, decomb / hqdn3d / lapsharp / comb-detect throughput, hqdn3d launch list.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/checkn
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$? $(tail -1 $OUT/pytest_gpu.log)" | tee $OUT/summary.txt
for w in 4k10_decomb_yadif 4k10_decomb_eedi2bob 4k_hqdn3d 4k_lapsharp 4k10_comb_detect; do
  timeout 300 python tools/bench_filters.py --only $w --frames 64 --cpu-frames 0 >> $OUT/bench.jsonl 2>> $OUT/bench.err
done
echo "bench rc=$?" | tee -a $OUT/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 24 --csv --log-file $OUT/hqdn3d_launches.csv \
    python tools/bench_filters.py --only 4k_hqdn3d --frames 4 --cpu-frames 0 > $OUT/ncu_run.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:decomb_field -c 1 -o $OUT/decomb_field python tools/bench_filters.py --only 4k10_decomb_yadif --frames 8 --cpu-frames 0 > $OUT/ncu_run2.log 2>&1
echo "ncu rc=$?" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
