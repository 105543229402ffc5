#!/bin/bash
# One gpurun call (1 GPU): multi-device + sharding GPU tests (repeated), the EEDI2 option-A' probe, dram traffic of one
# EEDI2 field graph (ncu), the 10-bit NLMeans bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/checkb
mkdir -p $OUT
for i in 1 2 3; do timeout 300 python -m pytest tests/test_nlmeans_multi_gpu.py -m gpu -q 2>&1 | tail -2; done > $OUT/pytest_multi_x3.txt 2>&1
timeout 600 python -m pytest tests/test_sharding_gpu.py tests/test_wrap_gpu.py -m gpu -q > $OUT/pytest_sharding_wrap.txt 2>&1
echo "sharding+wrap rc=$? $(tail -1 $OUT/pytest_sharding_wrap.txt)" | tee $OUT/summary.txt
timeout 900 python tools/eedi2_shard_probe.py > $OUT/eedi2_shard_probe.jsonl 2> $OUT/eedi2_shard_probe.err
echo "probe rc=$?" | tee -a $OUT/summary.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --clock-control none -c 600 --csv --log-file $OUT/eedi2_traffic.csv python tools/bench_filters.py --only 4k10_decomb_eedi2bob --frames 3 --cpu-frames 0 > $OUT/eedi2_traffic.log 2>&1
echo "eedi2 traffic rc=$?" | tee -a $OUT/summary.txt
timeout 600 python bench.py --workload 4k10_nlmeans_medium --steps 10 --warmup 3 --batch 256 --no-cpu-baseline --no-extra > $OUT/bench_4k10.json 2> $OUT/bench_4k10.err
echo "bench 4k10 rc=$?" | tee -a $OUT/summary.txt
cat $OUT/pytest_multi_x3.txt
