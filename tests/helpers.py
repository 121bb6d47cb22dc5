"""Shared helpers for the parity tests (fixtures are produced by oracle/gen_golden.py from the reference sources)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.meta = json.loads(str(self.z["meta"]))
        self.scene = self.meta["scene"]
        self.steps = self.meta["steps"]

    def end(self, step, field):
        return self.z[f"s{step}/end/{field}"]

    def grid(self, step, field):
        return self.z[f"s{step}/grid/{field}"]


def relmax(a, b):
    """max-norm relative difference: max|a-b| / max|b| (0 if both are identically zero)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    num = np.max(np.abs(a - b)) if a.size else 0.0
    return 0.0 if num == 0.0 else (num / den if den > 0 else np.inf)


def sym6_to_9(s6):
    """(n,6) xx,yy,zz,xy,yz,zx -> (n,9) row-major full tensor."""
    xx, yy, zz, xy, yz, zx = [s6[:, k] for k in range(6)]
    return np.stack([xx, xy, zx, xy, yy, yz, zx, yz, zz], axis=1)
