#!/bin/bash
# round 2, call A: regression of the whole GPU suite after the device-side particle count, the native slab step on
# one device (LocalSlabGroup), a short single-GPU bench line.
mkdir -p gpurun_out
TAG=${1:-r2a}
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
( time timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_slab.py ) > gpurun_out/${TAG}_tests.log 2>&1
tail -6 gpurun_out/${TAG}_tests.log
( time timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -q -x --timeout 300 ) > gpurun_out/${TAG}_slab.log 2>&1
tail -30 gpurun_out/${TAG}_slab.log
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
print(json.dumps(d["roofline"]["kernel_ms"]))
PY
