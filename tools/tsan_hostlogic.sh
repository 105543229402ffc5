#!/bin/bash
# ThreadSanitizer over the product's HOST-side filter code (the per-device submission threads of the multi-device NLMeans
# path in particular) driven through the CPU stand-ins for the device calls -- no GPU needed.
# usage: bash tools/tsan_hostlogic.sh
set -eu
cd "$(dirname "$0")/../oracle"
R=$(python3 hbcu_rename.py)
L=../handbrake_b200/libhb
mkdir -p _ref
gcc -g -O1 -std=gnu99 -fsanitize=thread -fno-omit-frame-pointer -w -D__LIBHB__ -pthread $R -I$L -I../include -o _ref/tsan_hostlogic \
    ../tools/asan_hostlogic_main.c $L/{nlmeans,detelecine,comb_detect,decomb,lapsharp,unsharp,denoise}_cuda.c port/*.c \
    $L/hbcu_device_frames.c $L/hb_runtime.c $L/hb_harness.c -lm -lpthread
TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" ./_ref/tsan_hostlogic 2>&1 | grep -E "WARNING: ThreadSanitizer|SUMMARY|ok|failed" | sort | uniq -c | head -20
