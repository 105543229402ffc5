#!/bin/bash
# gpurun --gpus N: bench.py under torchrun at N ranks only (the driver's SCALE step), output kept under gpurun_out/scaleN
N=${1:-8}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/scale$N
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "bench N=$N rc=$?" | tee $OUT/summary.txt
tail -c 600 $OUT/bench_n$N.json
