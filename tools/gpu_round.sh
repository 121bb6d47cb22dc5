#!/bin/bash
# One gpurun call: bench line, ncu launch list, one full ncu capture of the dominant kernels, GPU parity tests.
mkdir -p gpurun_out
TAG=${1:-r1}
KPAT=${2:-'k_tile_(fluid|mask)'}
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_raw.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:${KPAT}" -s 2 -c 2 -o gpurun_out/${TAG}_tile -f \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
( time timeout ${TEST_TIMEOUT:-600} python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_tests.log 2>&1
tail -5 gpurun_out/${TAG}_tests.log
ls -la gpurun_out
