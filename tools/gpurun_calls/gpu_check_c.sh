#!/bin/bash
# One gpurun call (1 GPU): whole -m gpu suite, stencil benches, prefilter NLMeans (v3 variant vs the generic kernel),
# compute-sanitizer passes of round 2.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/checkc
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$? $(tail -1 $OUT/pytest_gpu.log)" | tee $OUT/summary.txt
timeout 600 python tools/bench_filters.py --only lapsharp --frames 64 --cpu-frames 0 > $OUT/bench_lapsharp.jsonl 2> $OUT/bench_lapsharp.err
echo "lapsharp bench rc=$?" | tee -a $OUT/summary.txt
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:lapsharp -c 6 --csv --log-file $OUT/lapsharp_ncu.csv python tools/bench_filters.py --only lapsharp --frames 4 --cpu-frames 0 > $OUT/lapsharp_ncu.log 2>&1
echo "lapsharp ncu rc=$?" | tee -a $OUT/summary.txt
timeout 300 python bench.py --workload 4k_nlmeans_medium_prefilter --steps 5 --warmup 3 --batch 64 --no-cpu-baseline --no-extra --no-copy-only > $OUT/bench_prefilter_v3.json 2> $OUT/bench_prefilter_v3.err
HBCU_NLMEANS_IMPL=1 timeout 600 python bench.py --workload 4k_nlmeans_medium_prefilter --steps 3 --warmup 3 --batch 16 --no-cpu-baseline --no-extra --no-copy-only > $OUT/bench_prefilter_generic.json 2> $OUT/bench_prefilter_generic.err
echo "prefilter benches rc=$?" | tee -a $OUT/summary.txt
bash tools/sanitize_r02.sh > $OUT/sanitize_r02.txt 2>&1
echo "sanitize rc=$?" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
