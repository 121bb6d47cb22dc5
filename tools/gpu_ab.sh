#!/bin/bash
# A/B on one box: the current library against another build (TISPHI_B200_LIB), C4 at N GPUs, alternating
N=${1:-8}
mkdir -p gpurun_out
for rep in 1 2; do
for which in new old; do
  if [ $which = old ]; then export TISPHI_B200_LIB=$PWD/tisphi_b200/libtisphi_b200_old.so; else unset TISPHI_B200_LIB; fi
  if [ "$N" = "1" ]; then
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 1 --stirred-steps 0 > gpurun_out/ab_${which}_$rep.json 2> gpurun_out/ab_${which}_$rep.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --no-parity > gpurun_out/ab_${which}_$rep.json 2> gpurun_out/ab_${which}_$rep.err
  fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/ab_${which}_$rep.json") if l.startswith("{")][-1]); c=d["config"]
    print("$which $rep", round(d["ms_per_step"],4), c.get("rank_ms_per_step"), [t["kernel_ms"].get("tile_fluid") for t in c.get("rank_timeline_ms_per_step", [])] or d["roofline"]["kernel_ms"].get("tile_fluid"))
except Exception as e: print("$which $rep fail", e)
PY
done; done
