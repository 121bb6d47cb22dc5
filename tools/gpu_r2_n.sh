#!/bin/bash
# multi-GPU call: bench at N GPUs (native slab step over CUDA IPC), optionally the multi-process slab tests
N=${1:-2}
TAG=${2:-r2n$N}
EXTRA=${3:-}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 3 $EXTRA ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${TAG}_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
    c=d["config"]
    print("N", d["n_gpus"], "ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
    print("parity", c.get("slab_parity_cases"), "transport", c.get("transport")[:40])
    print("owned", c.get("owned_particles"))
    print("rank_ms", c.get("rank_ms_per_step"))
    for t in c.get("rank_timeline_ms_per_step") or []:
        print({k: v for k, v in t.items() if k != "kernel_ms"})
    print("limiter", c.get("limiter"))
    for r, t in enumerate(c.get("rank_timeline_ms_per_step") or []):
        if r in (0, 1, len(c["rank_timeline_ms_per_step"]) // 2, len(c["rank_timeline_ms_per_step"]) - 1):
            print("rank", r, "kernels", t.get("kernel_ms"))
except Exception as e:
    print("no bench line:", e)
PY
if [ -n "$TESTS" ]; then
( time timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -q -x --timeout 300 -k "$TESTS" ) > gpurun_out/${TAG}_tests.log 2>&1
tail -15 gpurun_out/${TAG}_tests.log
fi
