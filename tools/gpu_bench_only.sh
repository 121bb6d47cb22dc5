#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-b}
python bench.py --steps 20 --warmup 3 --no-cpu $2 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"])
print("e2e", d["e2e"])
print(json.dumps(d["roofline"]["kernel_ms"]))
PY
