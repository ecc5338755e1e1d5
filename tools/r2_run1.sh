#!/bin/bash
# Round 2, GPU call 1: everything new through the test suite, the cta_group::2 GEMM path, stage probes, first bench line.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1; nproc > gpurun_out/nproc.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
DSS_GEMM_2CTA=1 timeout 300 python -m pytest tests/test_ops_gpu.py -k gemm -q -p no:cacheprovider --timeout=120 > gpurun_out/pytest_2cta.log 2>&1; echo "pytest_2cta exit $?" >> gpurun_out/summary.txt
timeout 200 python tools/gemm_epi_probe.py > gpurun_out/gemm_epi_1cta.log 2>&1; echo "gemm1 exit $?" >> gpurun_out/summary.txt
DSS_GEMM_2CTA=1 timeout 200 python tools/gemm_epi_probe.py > gpurun_out/gemm_epi_2cta.log 2>&1; echo "gemm2 exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/perf_probe.py > gpurun_out/perf_probe.log 2>&1; echo "perf_probe exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/smoke.log; tail -15 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_2cta.log; cat gpurun_out/gemm_epi_1cta.log gpurun_out/gemm_epi_2cta.log; tail -8 gpurun_out/perf_probe.log
