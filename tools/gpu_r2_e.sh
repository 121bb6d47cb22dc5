#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2e}
( time timeout 600 python -m pytest tests/test_gpu_c5.py -m gpu -q -x --timeout 300 ) > gpurun_out/${TAG}_tests.log 2>&1
tail -8 gpurun_out/${TAG}_tests.log
for SZ in 1e6 1e7 1e8; do
python bench.py --workload c5 --size $SZ --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_c5_$SZ.json 2> gpurun_out/${TAG}_c5_$SZ.err
tail -2 gpurun_out/${TAG}_c5_$SZ.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_c5_$SZ.json").read().strip().splitlines()[-1])
    print("$SZ", "ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["config"]["mean_neighbours"], d["config"]["flagged_cells"])
    print(json.dumps(d["roofline"]["kernel_ms_per_sweep"])); print(d["roofline"]["grid_build"])
except Exception as e: print("fail", e)
PY
done
