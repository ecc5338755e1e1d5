#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_tcgen05 -s 2 -c 1 -f -o gpurun_out/prof_gemm_v3 python tools/gemm_probe_one.py 1152 384 128 5 > gpurun_out/ncu_gemm_v3.log 2>&1
tail -2 gpurun_out/ncu_gemm_v3.log
