#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/io_probe.py 2048 2>&1 | grep --line-buffered -E "PROBE|Error|error|Traceback" > gpurun_out/io_probe4.txt
grep -c PROBE gpurun_out/io_probe4.txt
