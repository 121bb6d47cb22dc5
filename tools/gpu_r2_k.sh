#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2k}
for MODE in 1 3 4; do
echo "== sweep mode $MODE"
SPH_SWEEP_MODE=$MODE bash tools/gpu_r2_g.sh ${TAG}_m$MODE 2>&1 | grep -v "scene built"
done
SPH_SWEEP_MODE=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_dp_soil|k_cspm_f|k_mui_soil3" -s 2 -c 3 -o gpurun_out/${TAG}_soil_direct -f \
    python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
