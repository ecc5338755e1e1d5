#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref exit $?" >> gpurun_out/summary.txt
# launch list: 167 launches per pass (2 ViT sub-batches of 32 + affinity + eigsh); skip the 2 warm-up passes
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 334 -c 167 --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py 64 3 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?" >> gpurun_out/summary.txt
for k in gemm_f16_tcgen05 attention_f16 lanczos_laplacian affinity_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 30 -c 2 -f -o gpurun_out/prof_$k python tools/ncu_target.py 64 2 > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k exit $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/bench.json | head -c 1500
