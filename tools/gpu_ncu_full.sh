#!/bin/bash
# one full ncu capture: $1 tag, $2 kernel regex, $3 launches to skip, $4 count
mkdir -p gpurun_out
TAG=${1:-f}; KPAT=${2:-k_tile_fluid}; SKIP=${3:-2}; CNT=${4:-1}
timeout 800 ncu --set full --clock-control none --import-source on -k "regex:${KPAT}" -s ${SKIP} -c ${CNT} -o gpurun_out/${TAG} -f \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out/${TAG}.ncu-rep
