#!/bin/bash
# One gpurun call (1 GPU): the round's final build -- whole -m gpu suite, the default bench line, smoke.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/final
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$? $(tail -1 $OUT/pytest_gpu.log)" | tee $OUT/summary.txt
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$? $(wc -c < $OUT/bench.json) bytes" | tee -a $OUT/summary.txt
python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1
echo "smoke rc=$? $(tail -1 $OUT/smoke.log)" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
