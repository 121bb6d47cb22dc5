#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2q}
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/${TAG}_tests.log 2>&1
tail -6 gpurun_out/${TAG}_tests.log
grep -a "^E  \|FAILED" gpurun_out/${TAG}_tests.log | head
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_ffma2 tools/ubench_ffma2.cu 2> gpurun_out/${TAG}_ubench_build.log && ./tools/ubench_ffma2 > gpurun_out/${TAG}_ubench_ffma2.txt 2>&1
cat gpurun_out/${TAG}_ubench_ffma2.txt | tail -5
