#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2h}
( time timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 ) > gpurun_out/${TAG}_tests.log 2>&1
tail -5 gpurun_out/${TAG}_tests.log
bash tools/gpu_r2_g.sh ${TAG}
