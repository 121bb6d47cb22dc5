#!/usr/bin/env python
"""Writes data/scenes/*.json: the BASELINE.json configurations as tiSPHi scene files.

C1-C3 are the scene dictionaries stored in the golden fixtures (tests/golden/*.npz, written by oracle/gen_golden.py
from the reference's own scene files with the keys changed as SURVEY 8 lists); C4 is this repo's 3D dambreak
(tisphi_b200/scenes.py).  Engine-only option keys (precision, ...) are stripped so that the files follow the
reference's schema and run_simulation.py --scene_file works on them unchanged.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import Golden          # noqa: E402
from tisphi_b200 import scenes      # noqa: E402

ENGINE_KEYS = ("precision", "wcFresh", "fastSweeps", "neighbourLists", "slabCapacity")


def clean(scene):
    scene = json.loads(json.dumps(scene))
    for k in ENGINE_KEYS:
        scene["Configuration"].pop(k, None)
    return scene


OUT = {
    "test1_db_water.json": lambda: Golden("c1_test1_wc_lf").scene,                 # C1
    "test2_cc_sand_muI.json": lambda: Golden("c2_test2_mui_lf").scene,             # C2
    "test2_cc_sand_dp_rk4_cspm.json": lambda: Golden("c3_test2_dp_rk4_cspm").scene,  # C3
    "test4_in_ver_small.json": lambda: Golden("dp2d_indenter_lf").scene,            # shipped test4 shrunken: static rigid indenter (8 f2)
    "c4_db3d_water_13M.json": lambda: scenes.dambreak3d(scale=1.0),                # C4
    "c4_db3d_water_small.json": lambda: scenes.dambreak3d(scale=0.25),             # C4 coarsened x4 (CPU-baseline sample)
}

if __name__ == "__main__":
    d = os.path.join(ROOT, "data", "scenes")
    os.makedirs(d, exist_ok=True)
    for name, make in OUT.items():
        with open(os.path.join(d, name), "w") as f:
            json.dump(clean(make()), f, indent=4)
            f.write("\n")
        print("wrote", os.path.join("data", "scenes", name))
