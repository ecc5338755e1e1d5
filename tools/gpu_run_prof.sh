#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_configs_gpu.py -m gpu -x -q -s --timeout=600 --timeout-method=thread > gpurun_out/pytest_configs.log 2>&1; echo "configs exit $?" >> gpurun_out/summary.txt
# launch list of one pipeline pass (64 images = 2 ViT sub-batches of 32 + affinity + eigsh), after 2 warm-up passes
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 342 -c 171 --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py 64 3 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 25 -c 1 -f -o gpurun_out/prof_attention python tools/ncu_target.py 64 2 > gpurun_out/ncu_att.log 2>&1; echo "ncu attention exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_tcgen05 -s 100 -c 4 -f -o gpurun_out/prof_gemm python tools/ncu_target.py 64 2 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lanczos -s 1 -c 1 -f -o gpurun_out/prof_eigsh python tools/ncu_target.py 256 2 > gpurun_out/ncu_eigsh.log 2>&1; echo "ncu eigsh exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/pytest_configs.log
