#!/bin/bash
# decode load schedule + bucket sort: parity tests first, then timings
set -x
timeout 600 python -m pytest tests/test_yolo_gpu.py tests/test_fullsize_gpu.py tests/test_golden_gpu.py -x -q 2>&1 | tail -5
timeout 200 python tools/perf_probe.py cfg4 2>&1 | grep -i "cfg4\|candidates"
timeout 200 python tools/perf_probe.py cfg2 2>&1 | grep -i "cfg2\|candidates"
