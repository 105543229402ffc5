#!/bin/bash
# One gpurun call (1 GPU): hqdn3d after the round-2 prefetch change (parity + e2e throughput + launch list).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/checki
mkdir -p $OUT
timeout 600 python -m pytest tests/test_hqdn3d_gpu.py tests/test_golden_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 $OUT/pytest.log)" | tee $OUT/summary.txt
timeout 300 python tools/bench_filters.py --only 4k_hqdn3d --frames 64 --cpu-frames 0 >> $OUT/bench_hqdn3d.jsonl 2>> $OUT/bench_hqdn3d.err
echo "hqdn3d bench rc=$?" | tee -a $OUT/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/hqdn3d_launches.csv \
    python tools/bench_filters.py --only 4k_hqdn3d --frames 8 --cpu-frames 0 > $OUT/ncu_run.log 2>&1
echo "ncu rc=$?" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
