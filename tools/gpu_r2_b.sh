#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2b}
( time timeout 1200 python -m pytest tests/test_gpu_slab.py -m gpu -q --timeout 200 ) > gpurun_out/${TAG}_slab.log 2>&1
tail -40 gpurun_out/${TAG}_slab.log
