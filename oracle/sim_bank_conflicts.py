#!/usr/bin/env python
"""TEST / ANALYSIS INFRASTRUCTURE (lives under oracle/ because it runs the oracle; nothing in the product imports it).

CPU model of the shared-memory gather traffic of k_tile_fluid (the fluid pass is bound by
LDS.128 gathers, 47 % of whose wavefronts are bank-conflict replays) and of a conflict-aware neighbour order.

An LDS.128 of a warp is served per quarter-warp (8 lanes x 16 B = the 32 banks once).  Two lanes of a quarter that
read DIFFERENT 16-byte entries of the same bank group (entry index mod 8) cost an extra wavefront; identical entries
are broadcast.  Wavefronts of one instruction = sum over the four quarters of max over bank groups of the number of
distinct entries in that group.

The model builds the neighbour lists of a 3D WCSPH state (oracle, float64 predicate), lays the particles out in the
2x2x4-cell footprint tiles exactly as sweeps_tile.cu does (16 runs of 6 cells, tile index = run offset + position in
the run), and replays
  (a) the kernel's order: stencil cells x-major / z fastest, j ascending, two neighbours per cell and half round;
  (b) a conflict-aware order for the recorded lists: in round t, slot s, lane l prefers a remaining neighbour whose
      entry lies in bank group (l + 4 t + s) mod 8 (eight lanes of a quarter -> eight different groups).
usage: sim_bank_conflicts.py [scale] [steps]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path = [ROOT] + [q for q in sys.path if os.path.abspath(q or '.') != os.path.dirname(os.path.abspath(__file__))]
from oracle import oracle as orc          # noqa: E402
from tisphi_b200 import scenes            # noqa: E402

BX, BY, ZB, SENT = 2, 2, 4, 3072


def neighbour_lists(o):
    """per particle: list of (stencil cell 0..26, global index j) in the kernel's order."""
    n = o.n
    gn = [int(v) for v in o.D["grid_num"]]
    cell_end = o.cell_end
    start = lambda g: int(cell_end[g - 1]) if g > 0 else 0
    x, gid = o.x, o.grid_ids
    sup2 = o.P.support ** 2
    out = [None] * n
    for i in range(n):
        g = int(gid[i])
        cx, r = divmod(g, gn[1] * gn[2])
        cy, cz = divmod(r, gn[2])
        lst = []
        cc = 0
        for ox in (-1, 0, 1):
            for oy in (-1, 0, 1):
                for oz in (-1, 0, 1):
                    nx, ny, nz = cx + ox, cy + oy, cz + oz
                    if 0 <= nx < gn[0] and 0 <= ny < gn[1] and 0 <= nz < gn[2]:
                        gg = (nx * gn[1] + ny) * gn[2] + nz
                        j0, j1 = start(gg), int(cell_end[gg])
                        if j1 > j0:
                            d = x[j0:j1] - x[i]
                            r2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
                            for k in np.nonzero(r2 < sup2)[0]:
                                if j0 + k != i:
                                    lst.append((cc, j0 + int(k)))
                    cc += 1
        out[i] = lst
    return out


def tile_index_maps(o):
    """for every footprint (b0, b1, seg): dict global particle index -> tile index, as tile_setup lays the runs out."""
    gn = [int(v) for v in o.D["grid_num"]]
    cell_end = o.cell_end
    start = lambda g: int(cell_end[g - 1]) if g > 0 else 0
    maps = {}
    nb0, nb1, nseg = (gn[0] + BX - 1) // BX, (gn[1] + BY - 1) // BY, (gn[2] + ZB - 1) // ZB
    for b0 in range(nb0):
        for b1 in range(nb1):
            for seg in range(nseg):
                f0 = seg * ZB
                f_lo, f_hi = max(f0 - 1, 0), min(f0 + ZB, gn[2] - 1)
                roff, m = 0, {}
                for rx in range(BX + 2):
                    for ry in range(BY + 2):
                        n0, n1 = b0 * BX + rx - 1, b1 * BY + ry - 1
                        if 0 <= n0 < gn[0] and 0 <= n1 < gn[1]:
                            gb = (n0 * gn[1] + n1) * gn[2]
                            S, E = start(gb + f_lo), int(cell_end[gb + f_hi])
                            for j in range(S, E):
                                m[j] = roff + (j - S)
                            roff += E - S
                maps[(b0, b1, seg)] = m
    return maps, (nb0, nb1, nseg)


def wavefronts(idx):
    """idx: (32,) tile indices read by one LDS.128 -> wavefronts (4 = conflict free)."""
    w = 0
    for q in range(4):
        ent = set(int(v) for v in idx[8 * q:8 * q + 8])
        groups = {}
        for e in ent:
            groups[e & 7] = groups.get(e & 7, 0) + 1
        w += max(groups.values())
    return w


def kernel_schedule(lst, tmap):
    """slots of one lane in the kernel's order: two per stencil cell and half round, padded with the sentinel."""
    slots, k = [], 0
    while k < len(lst):
        cc = lst[k][0]
        take = [tmap[lst[k][1]]]
        if k + 1 < len(lst) and lst[k + 1][0] == cc:
            take.append(tmap[lst[k + 1][1]])
        k += len(take)
        slots += take + [SENT] * (2 - len(take))
    if len(slots) % 4:
        slots += [SENT] * (4 - len(slots) % 4)
    return slots


def aware_schedule(lst, tmap, lane):
    """conflict-aware order of the same neighbours: slot (t, s) of lane l prefers bank group (l + 4 t + s) mod 8."""
    buckets = [[] for _ in range(8)]
    for _, j in lst:
        buckets[tmap[j] & 7].append(tmap[j])
    total = len(lst)
    slots = []
    pos = 0
    while total:
        want = (lane + pos) & 7
        pick = None
        for d in range(8):                                    # nearest non-empty group, wanted one first
            b = buckets[(want + d) & 7]
            if b:
                pick = b.pop()
                break
        slots.append(pick)
        total -= 1
        pos += 1
    if len(slots) % 4:
        slots += [SENT] * (4 - len(slots) % 4)
    return slots


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    scene = scenes.dambreak3d(scale=scale, precision="f64")
    o = orc.Oracle.from_scene(scene, serial=0)
    for _ in range(steps):
        o.step()
    o.grid_build()
    lists = neighbour_lists(o)
    maps, (nb0, nb1, nseg) = tile_index_maps(o)
    gn = [int(v) for v in o.D["grid_num"]]
    cell_end, typ = o.cell_end, o.mat_type
    start = lambda g: int(cell_end[g - 1]) if g > 0 else 0
    tot = {"kernel": [0, 0, 0], "aware": [0, 0, 0]}                # wavefronts, instructions, real pairs
    cells = 0
    for g in range(len(cell_end)):
        is_, ie = start(g), int(cell_end[g])
        nc = ie - is_
        if nc == 0 or nc > 32 or not np.any(typ[is_:ie] == 1):
            continue
        cx, r = divmod(g, gn[1] * gn[2])
        cy, cz = divmod(r, gn[2])
        tmap = maps[(cx // BX, cy // BY, cz // ZB)]
        cells += 1
        for name, sched in (("kernel", lambda l, k: kernel_schedule(l, tmap)), ("aware", lambda l, k: aware_schedule(l, tmap, k))):
            lanes = []
            for k in range(32):
                if k < nc and typ[is_ + k] == 1:
                    lanes.append(sched(lists[is_ + k], k))
                else:
                    lanes.append([])
            rounds = max(len(s) for s in lanes) // 4
            for s in lanes:
                s += [SENT] * (4 * rounds - len(s))
            arr = np.array(lanes, dtype=np.int64)                  # (32, 4 * rounds)
            for c in range(arr.shape[1]):
                tot[name][0] += wavefronts(arr[:, c])
                tot[name][1] += 1
            tot[name][2] += sum(len(lists[is_ + k]) for k in range(nc) if typ[is_ + k] == 1)
    print(f"scene scale {scale}, {steps} steps: N = {o.n}, cells with flow particles = {cells}")
    for name, (w, ins, pairs) in tot.items():
        print(f"  {name:7s}: {w / ins:5.2f} wavefronts per LDS.128 (4 = conflict free), conflict share {100 * (1 - 4 * ins / w):4.1f} %, "
              f"slots per real pair {32 * ins / max(pairs, 1):5.2f} (lane-slots incl. idle lanes and padding)")


if __name__ == "__main__":
    main()
