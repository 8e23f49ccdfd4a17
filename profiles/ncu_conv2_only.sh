ncu --set full --clock-control none --profile-from-start off --kernel-name-base demangled -k 'regex:conv_tc_kernel<\(int\)256, \(int\)3, \(int\)0, \(bool\)0, \(int\)2>' -s 100 -c 2 -o /tmp/c2 python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/r02_c2nv.out 2>&1
ncu -i /tmp/c2.ncu-rep --page raw --csv > gpurun_out/r02_c2nv_raw.csv 2>/dev/null
ls -la gpurun_out/r02_c2nv_raw.csv
