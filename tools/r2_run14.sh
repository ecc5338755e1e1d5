#!/bin/bash
mkdir -p gpurun_out
echo "cpu.max: $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"
python -c "
import sys; sys.path.insert(0,'tests')
from conftest import load_pkg
iop=load_pkg('io_pipeline'); print('available_cpus', iop.available_cpus(), 'decode threads', iop.default_workers())"
timeout 400 python tools/io_probe.py 2048 2>&1 | grep --line-buffered -E "PROBE|Error|error|Traceback" > gpurun_out/io_probe5.txt
grep "PROBE trace\|PROBE default\|out_on\|decode_only" gpurun_out/io_probe5.txt | cut -c1-420
timeout 420 python -m pytest tests/test_round2_gpu.py -q -m gpu -k "extract_all_host" -s --timeout 200 > gpurun_out/pytest_extract_all.log 2>&1
echo "pytest extract_all rc $?"
grep -n "extract_all\|passed\|failed\|Timeout\|Error" gpurun_out/pytest_extract_all.log | head -20
cat /sys/fs/cgroup/cpu.stat | grep thrott
