#!/bin/bash
mkdir -p gpurun_out
timeout 500 python tools/io_probe.py 2048 2>&1 | grep --line-buffered -E "PROBE|Error|error|Traceback" | tee gpurun_out/io_probe3.txt
timeout 420 python -m pytest tests/test_round2_gpu.py -q -m gpu -k "extract_all_host" -s --timeout 200 > gpurun_out/pytest_extract_all.log 2>&1
echo "pytest extract_all rc $?"
grep -n "extract_all\|passed\|failed\|Timeout\|Error" gpurun_out/pytest_extract_all.log | head -20
