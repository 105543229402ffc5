#!/bin/bash
# One gpurun call (1 GPU): the warp-specialised hqdn3d kernels and the 4-rows-per-thread lapsharp kernel (parity + throughput + launch lists).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/checkj
mkdir -p $OUT
timeout 300 python -m pytest tests/test_hqdn3d_gpu.py -m gpu -x -q > $OUT/pytest_hqdn3d.log 2>&1
echo "pytest hqdn3d rc=$? $(tail -1 $OUT/pytest_hqdn3d.log)" | tee $OUT/summary.txt
timeout 300 python -m pytest tests/test_lapsharp_gpu.py -m gpu -x -q > $OUT/pytest_lapsharp.log 2>&1
echo "pytest lapsharp rc=$? $(tail -1 $OUT/pytest_lapsharp.log)" | tee -a $OUT/summary.txt
timeout 600 python -m pytest tests/test_golden_gpu.py tests/test_fullsize_gpu.py tests/test_device_chain_gpu.py -m gpu -q > $OUT/pytest_rest.log 2>&1
echo "pytest rest rc=$? $(tail -1 $OUT/pytest_rest.log)" | tee -a $OUT/summary.txt
for w in 4k_hqdn3d 4k_lapsharp; do
  timeout 300 python tools/bench_filters.py --only $w --frames 64 --cpu-frames 0 >> $OUT/bench.jsonl 2>> $OUT/bench.err
done
HBCU_HQDN3D_V1=1 timeout 300 python tools/bench_filters.py --only 4k_hqdn3d --frames 64 --cpu-frames 0 >> $OUT/bench_v1.jsonl 2>> $OUT/bench.err
echo "bench rc=$?" | tee -a $OUT/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file $OUT/hqdn3d_launches.csv \
    python tools/bench_filters.py --only 4k_hqdn3d --frames 8 --cpu-frames 0 > $OUT/ncu_run.log 2>&1
echo "ncu hqdn3d rc=$?" | tee -a $OUT/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:lapsharp -c 6 --csv --log-file $OUT/lapsharp_ncu.csv \
    python tools/bench_filters.py --only 4k_lapsharp --frames 8 --cpu-frames 0 > $OUT/ncu_run2.log 2>&1
echo "ncu lapsharp rc=$?" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
