#!/bin/bash
# ncu --set full of the tcgen05 conv engine inside one sampling step (B = 64, T = 16, 128x128, fp32 = bf16x3):
#   NICE conv2 on CTA pairs      conv_tc_kernel<256,3,0,0,2>  (the dominant kernel: roofline.traffic comes from this capture)
#   decoder fused convs          conv_tc_kernel<*,3,1,*,*>    (ConvTranspose pairs with fused statistics, conv2 with fused residual, halo mode)
# usage (on the GPU box, from the repo root): bash profiles/ncu_convtc.sh <out-prefix>
set -e
OUT=${1:-gpurun_out/r02_convtc}
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
    -k 'regex:conv_tc_kernel<\(int\)256, \(int\)3, \(bool\)0, \(bool\)0, \(int\)2>' -s 100 -c 3 -o ${OUT}_conv2 \
    python bench.py --steps 1 --warmup 1 --profile-mode > ${OUT}_conv2.out 2>&1 || true
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
    -k 'regex:conv_tc_kernel<\(int\)[0-9]+, \(int\)3, \(bool\)1' -c 11 -o ${OUT}_dec \
    python bench.py --steps 1 --warmup 1 --profile-mode > ${OUT}_dec.out 2>&1 || true
ls -la ${OUT}*.ncu-rep
