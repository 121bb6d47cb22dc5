#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list, one full ncu capture of the dominant kernel.
mkdir -p gpurun_out
TAG=${1:-r1}
KPAT=${2:-k_tile_fluid}
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_tests.log 2>&1
tail -5 gpurun_out/${TAG}_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_raw.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${KPAT} -s 2 -c 1 -o gpurun_out/${TAG}_${KPAT} -f \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out
