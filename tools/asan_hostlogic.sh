#!/bin/bash
# AddressSanitizer + UBSan + LeakSanitizer over the product's HOST-side filter code (handbrake_b200/libhb/*_cuda.c) driven
# through the CPU stand-ins for the device calls (oracle/port/hostlogic_*.c) -- no GPU needed.
# usage: bash tools/asan_hostlogic.sh
set -eu
cd "$(dirname "$0")/../oracle"
R=$(python3 hbcu_rename.py)
L=../handbrake_b200/libhb
mkdir -p _ref
gcc -g -O1 -std=gnu99 -fsanitize=address,undefined -fno-omit-frame-pointer -w -D__LIBHB__ -pthread $R -I$L -I../include -o _ref/asan_hostlogic \
    ../tools/asan_hostlogic_main.c $L/{nlmeans,detelecine,comb_detect,decomb,lapsharp,unsharp,denoise}_cuda.c port/*.c \
    $L/hbcu_device_frames.c $L/hb_runtime.c $L/hb_harness.c -lm -lpthread
ASAN_OPTIONS=detect_leaks=1 ./_ref/asan_hostlogic 2>&1 | tail -3
