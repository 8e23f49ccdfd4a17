#!/bin/bash
# ncu --set full of the tcgen05 conv engine inside one sampling step (B = 64, T = 16, 128x128, fp32 = bf16x3):
#   NICE conv2 on CTA pairs      conv_tc_kernel<256,3,0,0,2>  (the dominant kernel: roofline.traffic comes from this capture)
#   decoder fused convs          conv_tc_kernel<*,3,1|2,*,*>  (ConvTranspose pairs with fused statistics, conv2 with fused residual, halo mode)
#   final conv + fused SPADE     out_conv_kernel
# and the launch list of the whole step (gpu__time_duration.sum per launch).  The raw pages are exported on the box (the reports of the
# decoder pass exceed what gpurun copies back) -- usage (GPU box, repo root): bash profiles/ncu_convtc.sh <out-prefix>
OUT=${1:-gpurun_out/r02_convtc}
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ${OUT}_launches.csv \
    python bench.py --steps 1 --warmup 1 --profile-mode > ${OUT}_launches.out 2>&1 || true
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
    -k 'regex:conv_tc_kernel<\(int\)256, \(int\)3, \(int\)0, \(bool\)0, \(int\)2>' -s 100 -c 2 -o ${OUT}_conv2 \
    python bench.py --steps 1 --warmup 1 --profile-mode > ${OUT}_conv2.out 2>&1 || true
ncu -i ${OUT}_conv2.ncu-rep --page raw --csv > ${OUT}_conv2_raw.csv 2>/dev/null || true
ncu --set full --clock-control none --profile-from-start off --kernel-name-base demangled \
    -k 'regex:conv_tc_kernel<\(int\)[0-9]+, \(int\)3, \(int\)[12]|out_conv_kernel' -c 12 -o /tmp/ipk_dec \
    python bench.py --steps 1 --warmup 1 --profile-mode > ${OUT}_dec.out 2>&1 || true
ncu -i /tmp/ipk_dec.ncu-rep --page raw --csv > ${OUT}_dec_raw.csv 2>/dev/null || true
ls -la ${OUT}* /tmp/ipk_dec.ncu-rep
