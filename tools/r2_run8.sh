#!/bin/bash
# GPU call 11: host pipeline probe + the two extract_all tests
mkdir -p gpurun_out
python tools/io_probe.py 2048 2>&1 | grep -E "PROBE|Error|error|Traceback" > gpurun_out/io_probe.txt
timeout 900 python -m pytest tests/test_round2_gpu.py -q -m gpu -k "extract_all" -s 2>&1 | tail -30 > gpurun_out/pytest_extract_all.log
cat gpurun_out/io_probe.txt
grep -n "extract_all:\|passed\|failed" gpurun_out/pytest_extract_all.log
