#!/bin/bash
# round 2 evidence: launch list of the bench command, ncu --set full of the sweeps of one step, the soil and C5 sweeps
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 200 --csv --log-file gpurun_out/${TAG}_launches_c4.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 --stirred-steps 0 > gpurun_out/${TAG}_ncu_launch.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_launch.log
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_tile_fluid|k_tile_mask|k_wall_gather|k_reorder|k_tile_prep|k_wc_finish" -s 12 -c 9 -o gpurun_out/${TAG}_c4_sweeps -f \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 --stirred-steps 0 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_full.log
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_dp_soil|k_corr_nlist|k_rk_stage|k_advect_pos_xsph" -s 6 -c 5 -o gpurun_out/${TAG}_soil -f \
    python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/${TAG}_ncu_soil.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_soil.log
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_tile_density|k_tile_mask" -s 4 -c 2 -o gpurun_out/${TAG}_c5 -f \
    python bench.py --workload c5 --size 1e7 --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${TAG}_ncu_c5.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_c5.log
ls -la gpurun_out/${TAG}_*.ncu-rep
