#!/bin/bash
# multi-GPU check: slab parity tests + one bench line at N GPUs
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-m}
( time timeout 600 python -m pytest tests/test_gpu_slab.py -x -q ) > gpurun_out/${TAG}_slab_tests.log 2>&1
tail -6 gpurun_out/${TAG}_slab_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
tail -2 gpurun_out/${TAG}_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
    print(d["config"].get("rank_ms_per_step"), d["config"].get("rank_host_wait_ms_per_step"), d["config"].get("owned_particles"))
except Exception as e:
    print("no bench line", e)
PY
