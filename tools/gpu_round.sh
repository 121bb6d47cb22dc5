#!/bin/bash
# One gpurun call: GPU parity tests, bench line (+ reference arm), ncu launch list, full ncu capture of one step's sweeps.
mkdir -p gpurun_out
TAG=${1:-r1}
( time timeout ${TEST_TIMEOUT:-600} python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_tests.log 2>&1
tail -4 gpurun_out/${TAG}_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
cat gpurun_out/${TAG}_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_raw.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:k_tile_fluid|k_tile_mask|k_wall_gather" -s 5 -c 5 -o gpurun_out/${TAG}_sweeps -f \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out | tail -12
