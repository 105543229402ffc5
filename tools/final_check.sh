# round-end sanity on a B200 box: the GPU test suite, smoke(), the default bench line and the reference arm
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 250 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_final.json
python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print('ours', d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])"
timeout 250 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-300
