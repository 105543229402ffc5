#!/bin/bash
# One gpurun call (1 GPU): the final build -- whole -m gpu suite, filters bench, default bench line, reference arm, launch list, smoke.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/checkp
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$? $(tail -1 $OUT/pytest_gpu.log)" | tee $OUT/summary.txt
timeout 900 python tools/bench_filters.py --frames 64 --cpu-frames 0 > $OUT/bench_filters.jsonl 2> $OUT/bench_filters.err
echo "bench_filters rc=$?" | tee -a $OUT/summary.txt
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$? $(wc -c < $OUT/bench.json) bytes" | tee -a $OUT/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
echo "reference rc=$?" | tee -a $OUT/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --batch 8 --no-cpu-baseline --no-extra --no-copy-only > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?" | tee -a $OUT/summary.txt
python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1
echo "smoke rc=$? $(tail -1 $OUT/smoke.log)" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
