#!/bin/bash
# ncu --set full of the soil sweeps (per-particle and per-cell-segment forms)
mkdir -p gpurun_out
TAG=${1:-r2j}
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_cell_sweep|k_dp_soil|k_cspm_f" -s 4 -c 6 -o gpurun_out/${TAG}_soil_cell -f \
    python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/${TAG}_ncu_cell.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_cell.log
ls -la gpurun_out/*.ncu-rep
