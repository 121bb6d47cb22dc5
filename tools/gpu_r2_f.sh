#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2f}
( time timeout 900 python -m pytest tests/test_gpu_horizon.py -m gpu -q -s --timeout 600 ) > gpurun_out/${TAG}_horizon.log 2>&1
grep -a "HORIZON\|passed\|failed\|Error\|stated" gpurun_out/${TAG}_horizon.log | head -40
