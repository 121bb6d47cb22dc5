#!/bin/bash
# A/B of the tile-path switches on one box: TISPHI_FLUID_PIPE x TISPHI_WALL_DYN, C4 at N GPUs
N=${1:-8}
mkdir -p gpurun_out
for v in "1 1" "0 1" "0 0" "1 0"; do
  set -- $v
  export TISPHI_FLUID_PIPE=$1 TISPHI_WALL_DYN=$2
  T=ab2_p$1_d$2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --no-parity > gpurun_out/$T.json 2> gpurun_out/$T.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/$T.json") if l.startswith("{")][-1]); c=d["config"]
    print("pipe $1 dyn $2:", round(d["ms_per_step"],4), c.get("rank_ms_per_step"), "fluid", [t["kernel_ms"].get("tile_fluid") for t in c.get("rank_timeline_ms_per_step", [])][:6], "wall", [t["kernel_ms"].get("tile_wall") for t in c.get("rank_timeline_ms_per_step", [])][:3])
except Exception as e: print("$T fail", e)
PY
done
