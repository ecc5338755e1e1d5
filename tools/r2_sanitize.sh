#!/bin/bash
# compute-sanitizer memcheck over the small-shape GPU tests (every kernel class at least once)
mkdir -p gpurun_out
S="compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 20"
timeout 700 $S python -m pytest tests/test_ops_gpu.py tests/test_golden_gpu.py -q -m gpu -x --timeout 600 -k "not 901 and not 577" > gpurun_out/sanitize_ops.log 2>&1
echo "sanitize ops rc $?"; tail -4 gpurun_out/sanitize_ops.log | cut -c1-200
timeout 700 $S python -m pytest tests/test_round2_gpu.py tests/test_spectral_gpu.py -q -m gpu -x --timeout 600 -k "(fused_layernorm and (128 or 300 or 77)) or affinity_symmetric or rw_colour or segmentations or kmeans or odd_grid or two_cliques or (structured and 196)" > gpurun_out/sanitize_r2.log 2>&1
echo "sanitize r2 rc $?"; tail -4 gpurun_out/sanitize_r2.log | cut -c1-200
grep -c "ERROR SUMMARY" gpurun_out/sanitize_ops.log gpurun_out/sanitize_r2.log; grep "ERROR SUMMARY\|Invalid\|out of bounds\|misaligned" gpurun_out/sanitize_*.log | head -10
