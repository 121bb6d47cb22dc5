#!/bin/bash
# final multi-GPU measurements of round 2: C4 (native slab step), C5 slab-partitioned; N=2 also the torch.distributed transport
N=${1:-2}
C5SIZES=${2:-1e7}
mkdir -p gpurun_out
run() {  # tag, extra args
  local TAG=$1; shift
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 20 --warmup 3 "$@" ) > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
  grep -v "^W\|^\*\*\*" gpurun_out/${TAG}.err | tail -4
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${TAG}.json").read().strip().splitlines() if l.startswith("{")][-1])
    c=d["config"]
    print("${TAG}: N", d["n_gpus"], "ms_per_step", round(d["ms_per_step"],4), "value", "%.4g" % d["value"], "e2e", (d.get("e2e") or {}).get("value"))
    for k in ("slab_parity", "slab_parity_cases", "transport", "owned_particles", "rank_ms_per_step", "limiter", "mean_neighbours"):
        if k in c: print("  ", k, str(c[k])[:300])
    if "slab_parity" in d: print("   slab_parity", d["slab_parity"])
except Exception as e:
    print("${TAG}: no bench line:", e)
PY
}
nvidia-smi topo -m > gpurun_out/final_n${N}_topo.txt 2>&1
run final_n${N}_c4
for SZ in $C5SIZES; do run final_n${N}_c5_$SZ --workload c5 --size $SZ --steps 10; done
if [ "$N" = "2" ]; then run final_n2_c4_dist --transport dist --no-parity; fi
