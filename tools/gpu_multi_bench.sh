#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}; TAG=${2:-mb}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --e2e-steps 2 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
tail -4 gpurun_out/${TAG}_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
    print(d["config"].get("rank_ms_per_step"), d["config"].get("owned_particles"), d["config"].get("columns"))
    print(d["roofline"]["kernel_ms"] if d.get("roofline") else None)
except Exception as e:
    print("no bench line", e)
PY
