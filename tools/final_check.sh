timeout 500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3; timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -2; timeout 250 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_r01n.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r01n.json')); print(d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['roofline']['frac'], d['clocks'], d['config']['cpu_affinity'])"; timeout 400 python tools/bench_filters.py --frames 96 --cpu-frames 4 2>/dev/null | tee gpurun_out/bench_filters_r01n.jsonl | python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print(d['workload'], d.get('value'), d.get('roofline',{}).get('frac'), d['e2e']['value'], d['cpu_baseline']['value'])
    except Exception as e: print(l[:200])
"; rm -f gpurun_out/bench_chain_r01n.jsonl; for c in 4k 3 5; do timeout 300 python tools/bench_chain.py --config $c --frames 24 --cpu-frames 2 2>/dev/null | tail -1 >> gpurun_out/bench_chain_r01n.jsonl; done; cat gpurun_out/bench_chain_r01n.jsonl | cut -c1-700
