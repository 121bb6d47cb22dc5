#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2d}
( time timeout 1200 python -m pytest tests -m gpu -q -x --timeout 300 ) > gpurun_out/${TAG}_tests.log 2>&1
tail -6 gpurun_out/${TAG}_tests.log
python bench.py --steps 20 --warmup 3 --no-cpu ${BENCH_EXTRA} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "serial e2e", d["e2e"]["serial_value"])
print(json.dumps(d["roofline"]["kernel_ms"]))
print("stirred", d.get("stirred"))
print("flagged", d["config"].get("flagged_cells"))
PY
