#!/bin/bash
# One gpurun call (1 GPU): dependent-chain latencies + fp64 rates, the lapsharp occupancy / rows-per-thread variants, hqdn3d after the 3-op chain.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/checkk
mkdir -p $OUT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/latency_bench tools/latency_bench.cu > $OUT/latency_build.log 2>&1
timeout 120 /tmp/latency_bench > $OUT/latency.txt 2>&1
echo "latency rc=$?" | tee $OUT/summary.txt
timeout 300 python -m pytest tests/test_hqdn3d_gpu.py -m gpu -x -q > $OUT/pytest_hqdn3d.log 2>&1
echo "pytest hqdn3d rc=$? $(tail -1 $OUT/pytest_hqdn3d.log)" | tee -a $OUT/summary.txt
for v in 0 1 2 3; do
  HBCU_LAP_VARIANT=$v timeout 300 python -m pytest tests/test_lapsharp_gpu.py -m gpu -x -q > $OUT/pytest_lapsharp_$v.log 2>&1
  echo "pytest lapsharp variant $v rc=$? $(tail -1 $OUT/pytest_lapsharp_$v.log)" | tee -a $OUT/summary.txt
  HBCU_LAP_VARIANT=$v timeout 300 python tools/bench_filters.py --only 4k_lapsharp --frames 64 --cpu-frames 0 >> $OUT/bench_lap_$v.jsonl 2>> $OUT/bench.err
done
timeout 300 python tools/bench_filters.py --only 4k_hqdn3d --frames 64 --cpu-frames 0 >> $OUT/bench_hqdn3d.jsonl 2>> $OUT/bench.err
echo "bench rc=$?" | tee -a $OUT/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 24 --csv --log-file $OUT/hqdn3d_launches.csv \
    python tools/bench_filters.py --only 4k_hqdn3d --frames 4 --cpu-frames 0 > $OUT/ncu_run.log 2>&1
echo "ncu hqdn3d rc=$?" | tee -a $OUT/summary.txt
cat $OUT/summary.txt; cat $OUT/latency.txt
