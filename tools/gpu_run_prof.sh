#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + one --set full capture of each dominant kernel
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
# launch list of the default bench command (numbers printed by a run under ncu are never bench values)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_bench.log 2>&1; echo "ncu bench launches exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 25 -c 1 -f -o gpurun_out/prof_attention python tools/ncu_target.py 64 2 > gpurun_out/ncu_att.log 2>&1; echo "ncu attention exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_tcgen05 -s 100 -c 4 -f -o gpurun_out/prof_gemm python tools/ncu_target.py 64 2 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lanczos -s 1 -c 1 -f -o gpurun_out/prof_eigsh python tools/ncu_target.py 256 2 > gpurun_out/ncu_eigsh.log 2>&1; echo "ncu eigsh exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
