#!/bin/bash
# ncu source-level capture (dense PC sampling) of conv_tc launches matching a kernel regex inside one sampling step
# usage: bash profiles/ncu_one.sh <out-prefix> <kernel regex> [launch-skip] [count]
OUT=$1; RE=$2; SKIP=${3:-300}; CNT=${4:-2}
ncu --set full --warp-sampling-interval 0 --warp-sampling-max-passes 50 --warp-sampling-buffer-size 536870912 --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:$RE" -s $SKIP -c $CNT -o ${OUT} \
    python bench.py --steps 1 --warmup 1 --profile-mode > ${OUT}.out 2>&1 || true
ls -la ${OUT}*.ncu-rep
