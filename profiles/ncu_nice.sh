#!/bin/bash
# ncu --set full of the three NICE coupling convs (conv1 / conv2 / conv3 of consecutive couplings) inside one sampling step
# usage (on the GPU box, from the repo root): bash profiles/ncu_nice.sh <out-prefix> [launch-skip] [count]
OUT=${1:-gpurun_out/r02_nice}
SKIP=${2:-1000}
CNT=${3:-9}
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:conv_tc_kernel<\(int\)[0-9]+, \(int\)3, \(int\)0' -s $SKIP -c $CNT -o ${OUT} \
    python bench.py --steps 1 --warmup 1 --profile-mode > ${OUT}.out 2>&1 || true
ls -la ${OUT}*.ncu-rep
