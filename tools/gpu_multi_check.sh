#!/bin/bash
# One gpurun --gpus N call: multi-device dealing tests on real devices, then bench.py under torchrun at N ranks.
# usage: gpurun --gpus 2 --timeout 1500 -- bash tools/gpu_multi_check.sh 2
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/multi$N
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 600 python -m pytest tests/test_nlmeans_multi_gpu.py tests/test_sharding_gpu.py tests/test_nlmeans_gpu.py tests/test_golden_gpu.py tests/test_detelecine_gpu.py tests/test_wrap_gpu.py -m gpu -x -q > $OUT/pytest_multi.log 2>&1
echo "pytest rc=$? $(tail -1 $OUT/pytest_multi.log)" | tee $OUT/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "bench N=$N rc=$?" | tee -a $OUT/summary.txt
HBCU_BENCH_WC_INPUT=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --no-gather --no-plugin-multi > $OUT/bench_n${N}_wc.json 2> $OUT/bench_n${N}_wc.err
echo "bench N=$N write-combined inputs rc=$?" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
