#!/usr/bin/env python
"""C5 (synthetic uniform box): particles/s of grid build + neighbour count + density sum at one or more sizes.
usage: bench_c5.py [N ...]   (default 1e6 1e7)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tisphi_b200.c5 import UniformBox

for n in [int(float(a)) for a in sys.argv[1:]] or [1_000_000, 10_000_000]:
    box = UniformBox(n)
    for _ in range(3):
        box.sweep()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 10
    e0.record(box.engine.stream)
    for _ in range(k):
        box.sweep()
    e1.record(box.engine.stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / k
    cnt = box.count.cpu().numpy()
    print(json.dumps({"workload": "C5 uniform box: grid build + neighbour count + density sum", "N": box.n, "ms_per_sweep": ms,
                      "particles_per_s": box.n / (ms * 1e-3), "algorithmic_GBps": box.algorithmic_bytes() / (ms * 1e-3) / 1e9,
                      "mean_neighbours": float(cnt.mean()), "max_neighbours": int(cnt.max()), "dtype": "f32"}), flush=True)
    del box
    torch.cuda.empty_cache()
