"""world_size-2/3/4 gloo runs of the slab-decomposition protocol (tisphi_b200/parallel.py) on CPU.

The driver is the product's; the engine behind it is the CPU oracle (tests/slab_oracle.py).  After every step the
owned ranges of all ranks, concatenated in rank order, must equal a single-process oracle run BIT FOR BIT: same cell
ids, same sorted order (id0), same float64 fields -- the property DESIGN.md claims for the multi-GPU path."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import Golden, ROOT

CHECK = {1: ["x", "v", "density", "pressure", "d_vel", "d_density", "v_tmp", "m_V"],
         2: ["x", "v", "density", "pressure", "d_vel", "stress", "strain_equ"],
         3: ["x", "v", "density", "stress", "d_stress", "strain_equ", "strain_equ_p", "d_vel"]}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, nsteps, serial):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["OMP_NUM_THREADS"] = "2"
    from oracle import oracle as orc
    from slab_oracle import OracleSlabEngine
    from tisphi_b200.parallel import SlabDriver, column_weights, partition_columns
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        g = Golden(name)
        P, D = orc.make_params(g.scene, serial=serial)
        _, x, v, rho, typ = orc.build_particles(g.scene)
        obj, dyn = orc.build_particles.last_obj_dyn
        ref = orc.Oracle.from_scene(g.scene, serial=serial) if rank == 0 else None
        n_cols = int(D["grid_num"][0])
        w, cx = column_weights(x[:, 0], typ, float(D["vstart"][0]), D["grid_size"], n_cols)
        cols = partition_columns(w, world)
        a, b = cols[rank]
        mine = np.nonzero((cx >= a) & (cx < b))[0]
        stress = None
        if P.solver == 3:                      # init_stress needs the global column top (base:249-260)
            full = orc.Oracle.from_scene(g.scene, serial=serial)
            stress = full.stress[mine].copy()
        eng = OracleSlabEngine(P, x[mine], v[mine], rho[mine], typ[mine], mine.astype(np.int32), stress, obj[mine], dyn[mine])
        drv = SlabDriver(eng, (a, b), rank, world, n_cols, check=True)
        fields = CHECK[P.solver] + ["id0", "grid_ids", "mat_type"]
        for s in range(nsteps):
            drv.step()
            if rank == 0:
                assert ref.step() == 0
            mine_now = {f: np.ascontiguousarray(getattr(eng.o, f)[drv.own_first:drv.own_first + drv.own_count]) for f in fields}
            gathered = [None] * world
            dist.all_gather_object(gathered, mine_now)
            if rank == 0:
                for f in fields:
                    got = np.concatenate([gd[f] for gd in gathered])
                    want = getattr(ref, f)
                    assert got.shape == want.shape, (name, s, f, got.shape, want.shape)
                    assert np.array_equal(got, want), f"{name}: step {s + 1}: {f} differs from the single-process run"
        if rank == 0:
            assert drv.exchanges > nsteps
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,world,nsteps,serial", [
    ("wc2d_small_lf", 2, 6, 0),
    ("wc2d_small_lf", 3, 4, 1),
    ("wc2d_small_rk4_cspm", 2, 3, 0),
    ("wc3d_tiny_lf", 2, 3, 0),
    ("dp2d_small_lf", 2, 3, 0),
    ("mui2d_small_lf", 2, 3, 0),
    ("c1_test1_wc_lf", 4, 2, 0),              # four slabs of the full test1 scene (92 columns)
    ("c1_test1_wc_lf", 8, 2, 0),              # eight slabs: the width the driver's scaling run goes to
    ("dp2d_indenter_lf", 2, 3, 0),            # a static rigid indenter inside one slab, next to the face
    ("wc2d_collision_lf", 2, 4, 0),           # boundary mode 1: enforced collision, no wall particles
    ("mui2d_dummyrep_lf", 3, 3, 0),           # boundary mode 4: dummy + repulsive wall particles
])
def test_slab_protocol_equals_single_process(name, world, nsteps, serial):
    mp.start_processes(_worker, args=(world, _free_port(), name, nsteps, serial), nprocs=world, join=True,
                       start_method="spawn")


def test_partition_columns_properties():
    from tisphi_b200.parallel import partition_columns
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        for _ in range(20):
            w = rng.random(rng.integers(world, 60)) * (rng.random() < 0.8)
            parts = partition_columns(w, world)
            assert parts[0][0] == 0 and parts[-1][1] == len(w)
            assert all(b > a for a, b in parts)
            assert all(parts[k][1] == parts[k + 1][0] for k in range(world - 1))
    w = np.ones(80)
    assert partition_columns(w, 8) == [(10 * k, 10 * k + 10) for k in range(8)]
