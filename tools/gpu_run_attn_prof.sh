#!/bin/bash
# source-level ncu capture of the tcgen05 attention kernel (one launch, after warm-up)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 14 -c 1 -f -o gpurun_out/prof_attn_v5 python tools/ncu_target.py 32 2 > gpurun_out/ncu_attn_v5.log 2>&1
tail -2 gpurun_out/ncu_attn_v5.log
