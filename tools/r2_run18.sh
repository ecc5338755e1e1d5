#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_ln_f16_tcgen05_kernel<\(bool\)1" -s 14 -c 1 -f -o gpurun_out/prof_gemm_ln_fc1_v2 python tools/ncu_target.py 296 2 296 > gpurun_out/ncu_ln_fc1_v2.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_gemm_ln_fc1_v2.ncu-rep
