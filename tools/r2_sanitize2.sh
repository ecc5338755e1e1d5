#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 20 python -m pytest tests/test_vit_gpu.py tests/test_branches_gpu.py -q -m gpu -x --timeout 800 > gpurun_out/sanitize_vit.log 2>&1
echo "memcheck vit/branches rc $?"; tail -3 gpurun_out/sanitize_vit.log | cut -c1-200
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 77 --print-limit 20 python -m pytest tests/test_ops_gpu.py tests/test_round2_gpu.py -q -m gpu -x --timeout 800 -k "(gemm and 128) or epilogues or (attention and 64) or (fused_layernorm and (128 or 300)) or segmentations or two_cliques" > gpurun_out/sanitize_sync.log 2>&1
echo "synccheck rc $?"; tail -3 gpurun_out/sanitize_sync.log | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 77 --print-limit 20 python -m pytest tests/test_ops_gpu.py tests/test_round2_gpu.py tests/test_spectral_gpu.py -q -m gpu -x --timeout 800 -k "(gemm and 128-128) or (attention and 64) or (fused_layernorm and 128) or segmentations or two_cliques or (layernorm and 384) or im2col" > gpurun_out/sanitize_race.log 2>&1
echo "racecheck rc $?"; tail -3 gpurun_out/sanitize_race.log | cut -c1-200
grep "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/sanitize_*.log
