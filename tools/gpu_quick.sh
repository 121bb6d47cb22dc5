#!/bin/bash
# Quick GPU iteration: parity tests, then a short bench line without the CPU-baseline leg.
mkdir -p gpurun_out
TAG=${1:-q}
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_tests.log 2>&1
tail -15 gpurun_out/${TAG}_tests.log
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
print(json.dumps(d["roofline"]["kernel_ms"]))
PY
if [ -n "$2" ]; then
python bench.py --steps 20 --warmup 3 --no-cpu $2 > gpurun_out/${TAG}_bench_b.json 2> gpurun_out/${TAG}_bench_b.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_b.json").read().strip().splitlines()[-1])
print("B: ms_per_step", d["ms_per_step"], "value", d["value"])
print(json.dumps(d["roofline"]["kernel_ms"]))
PY
fi
