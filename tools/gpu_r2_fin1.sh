#!/bin/bash
# final single-GPU measurements of round 2: GPU tests, the default bench (with cpu_baseline), the reference arm, soil, C5
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 300 ) > gpurun_out/final_tests.log 2>&1
tail -4 gpurun_out/final_tests.log
( time python bench.py ) > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err
tail -2 gpurun_out/final_n1.err
( time python bench.py --impl reference ) > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
tail -2 gpurun_out/final_ref.err
for SOIL in dp mui; do python bench.py --workload c3 --soil $SOIL > gpurun_out/final_soil_$SOIL.json 2> gpurun_out/final_soil_$SOIL.err; tail -1 gpurun_out/final_soil_$SOIL.err; done
for SZ in 1e6 1e7 1e8; do python bench.py --workload c5 --size $SZ > gpurun_out/final_c5_$SZ.json 2> gpurun_out/final_c5_$SZ.err; tail -1 gpurun_out/final_c5_$SZ.err; done
python - <<'PY'
import json
for t in ["n1", "ref", "soil_dp", "soil_mui", "c5_1e6", "c5_1e7", "c5_1e8"]:
    try:
        d = json.loads([l for l in open(f"gpurun_out/final_{t}.json").read().splitlines() if l.startswith("{")][-1])
        print(t, "ms", d.get("ms_per_step"), "value", d.get("value"), "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"),
              "frac", (d.get("roofline") or {}).get("frac"), "launches", d.get("gpu_launches"))
    except Exception as e:
        print(t, "fail", e)
PY
