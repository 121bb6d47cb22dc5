#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2c}
( time timeout 1200 python -m pytest tests/test_gpu_slab.py tests/test_gpu_parity.py -m gpu -q -x --timeout 200 ) > gpurun_out/${TAG}_tests.log 2>&1
tail -12 gpurun_out/${TAG}_tests.log
