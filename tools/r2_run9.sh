#!/bin/bash
mkdir -p gpurun_out
timeout 500 python tools/io_probe.py 2048 2>&1 | grep --line-buffered -E "PROBE|Error|error|Traceback" | tee gpurun_out/io_probe2.txt
echo "probe rc ${PIPESTATUS[0]}"; df -h /tmp /dev/shm | head; mount | grep -E " / | /tmp " | head -3
