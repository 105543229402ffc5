#!/bin/bash
# One gpurun call (1 GPU): lapsharp bench + ncu, initcheck of the multi-device path, the plugin multi-device child on two
# handles of one GPU, wall time of the default bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/checkd
mkdir -p $OUT
timeout 600 python tools/bench_filters.py --only 4k_lapsharp --frames 64 --cpu-frames 0 > $OUT/bench_lapsharp.jsonl 2> $OUT/bench_lapsharp.err
echo "lapsharp bench rc=$?" | tee $OUT/summary.txt
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum --clock-control none -k regex:lapsharp -c 6 --csv --log-file $OUT/lapsharp_ncu.csv python tools/bench_filters.py --only 4k_lapsharp --frames 4 --cpu-frames 0 > $OUT/lapsharp_ncu.log 2>&1
echo "lapsharp ncu rc=$?" | tee -a $OUT/summary.txt
timeout 600 compute-sanitizer --tool initcheck --print-limit 5 python -m pytest tests/test_nlmeans_multi_gpu.py -q -m gpu -x -k "two_handles" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Uninitialized" | head -8 > $OUT/initcheck_multi.txt
echo "initcheck rc=$?" | tee -a $OUT/summary.txt
timeout 300 python bench.py --plugin-multi-child 0,0 --plugin-frames 2048 --plugin-warm 256 > $OUT/plugin_child_0_0.json 2> $OUT/plugin_child_0_0.err
echo "plugin child rc=$?" | tee -a $OUT/summary.txt
T0=$(date +%s)
timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$? wall $(( $(date +%s) - T0 )) s" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
