#!/bin/bash
# follow-up of the r01 sanitizer pass: racecheck of the device-resident EEDI2 chain (graphs on), initcheck details
CS="compute-sanitizer --print-limit 8"
echo "=== racecheck device chain (graphs on) x2"
for i in 1 2; do timeout 600 $CS --tool racecheck python -m pytest tests/test_device_chain_gpu.py -q -m gpu -x -k "eedi2_bob or full_chain" 2>&1 | grep -E "passed|failed|bytes differ|hazard|SUMMARY" | head -6; done
echo "=== initcheck"
timeout 600 $CS --tool initcheck python -m pytest tests/test_device_chain_gpu.py -q -m gpu -x -k "eedi2_bob" 2>&1 | grep -E "Uninitialized|at .*\(|SUMMARY|passed" | head -40
timeout 600 $CS --tool initcheck python -m pytest tests/test_nlmeans_gpu.py tests/test_comb_detect_gpu.py tests/test_unsharp_gpu.py -q -m gpu -x -k "config1 and 6 or prefilter_modes and 1025 or masks_match and 10 or extreme" 2>&1 | grep -E "Uninitialized|at .*\(|SUMMARY|passed" | head -40
echo "=== nlmeans edgeboost tests"
timeout 300 python -m pytest tests/test_nlmeans_gpu.py -q -m gpu -x -k "prefilter or edgeboost" 2>&1 | tail -3
