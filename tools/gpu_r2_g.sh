#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2g}
for SOIL in dp mui; do
python bench.py --workload c3 --soil $SOIL --steps 10 --warmup 3 --no-cpu ${EXTRA} > gpurun_out/${TAG}_soil_$SOIL.json 2> gpurun_out/${TAG}_soil_$SOIL.err
tail -2 gpurun_out/${TAG}_soil_$SOIL.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_soil_$SOIL.json").read().strip().splitlines()[-1])
    print("$SOIL", "ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
    print(json.dumps(d["roofline"]["kernel_ms"])); print(json.dumps(d["roofline"]["kernel_share_of_step"]))
except Exception as e: print("fail", e)
PY
done
