"""Multi-GPU parity: the slab-partitioned engine must reproduce the single-GPU engine BIT FOR BIT -- cell ids, sorted
order and every field -- in both precisions.

  test_native_slab_group_equals_single_gpu   the device-driven slab step (csrc/slab.cu) with 2-4 slabs as independent
                                             contexts of ONE process on ONE device (runs on a single-GPU box)
  test_slab_equals_single_gpu                one process per GPU (needs >= 2 devices): native step over CUDA IPC peer
                                             mappings ("p2p") and the Python SlabDriver over NCCL send / recv ("dist")"""
import os
import socket
import sys

import numpy as np
import pytest

from helpers import Golden, ROOT

pytestmark = pytest.mark.gpu

FIELDS = ["x", "v", "density", "pressure", "d_vel", "d_density", "v_tmp", "m_V"]
IFIELDS = ["id0", "grid_ids", "mat_type"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, prec, nsteps, extra, transport):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import copy
    import torch
    import torch.distributed as dist
    from tisphi_b200.eng.simulation import Simulation, SimConfiger
    from tisphi_b200.parallel import SlabSimulation
    torch.cuda.set_device(rank)
    import datetime
    import traceback
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device(f"cuda:{rank}"), timeout=datetime.timedelta(seconds=90))
    try:
        g = Golden(name)
        scene = copy.deepcopy(g.scene)
        scene["Configuration"]["precision"] = prec
        scene["Configuration"].update(extra)
        slab = SlabSimulation(SimConfiger(config=copy.deepcopy(scene)), f"cuda:{rank}", rank, world, check=True, transport=transport)
        ref = Simulation(SimConfiger(config=copy.deepcopy(scene)), device=f"cuda:{rank}") if rank == 0 else None
        for s in range(nsteps):
            slab.run_steps(1)
            mine = {f: slab.owned(f).detach().cpu().numpy() for f in FIELDS + IFIELDS}
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)
            if rank == 0:
                ref.solver.step()
                for f in FIELDS + IFIELDS:
                    got = np.concatenate([gd[f] for gd in gathered])
                    want = getattr(ref.ps.pt, f).detach().cpu().numpy()
                    assert got.shape == want.shape, (name, prec, s, f, got.shape, want.shape)
                    assert np.array_equal(got, want), f"{name}[{prec}] step {s + 1}: {f} differs from the single-GPU run"
    except BaseException:           # fail fast: a peer blocked in a collective must not wait for the NCCL watchdog
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
    dist.destroy_process_group()


CASES = [
    ("wc2d_small_lf", "f64", {}),
    ("wc2d_small_lf", "f32", {}),
    ("wc3d_tiny_lf", "f32", {}),
    ("wc3d_tiny_lf", "f32", {"fastSweeps": False}),
    ("c1_test1_wc_lf", "f32", {}),
    ("wc2d_small_rk4_cspm", "f64", {}),
]
SOIL_CASES = [
    ("mui2d_small_lf", "f64", {}),           # XSPH + the regularisation sweep on ghosts, carried wall v_tmp (H27)
    ("mui2d_small_lf", "f32", {}),
    ("dp2d_small_rk4_cspm", "f64", {}),      # CSPM_L, four one_steps per step, stress in the messages
    ("dp2d_small_lf", "f32", {}),
    ("dp2d_plate_lf", "f64", {}),            # several soil blocks of different height: init_stress needs the GLOBAL soil top
    ("wc2d_indenter_lf", "f32", {}),         # static rigid block
    ("wc2d_dummyrep_lf", "f64", {}),         # repulsive particles (boundary 4): generic sweeps with the type-dependent force
    ("wc2d_collision_lf", "f32", {}),        # enforced collision (boundary 1): no wall particles at all
]


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("name,prec,extra", CASES + SOIL_CASES)
def test_native_slab_group_equals_single_gpu(name, prec, extra, world):
    import copy
    import torch
    from tisphi_b200.eng.simulation import Simulation, SimConfiger
    from tisphi_b200.parallel import LocalSlabGroup
    if world > 2 and (name, prec) not in (("wc3d_tiny_lf", "f32"), ("c1_test1_wc_lf", "f32"), ("mui2d_small_lf", "f32"),
                                         ("dp2d_small_rk4_cspm", "f64")):
        pytest.skip("3 and 4 slabs are covered on four representative cases")
    g = Golden(name)
    scene = copy.deepcopy(g.scene)
    scene["Configuration"]["precision"] = prec
    scene["Configuration"].update(extra)
    grp = LocalSlabGroup(lambda: SimConfiger(config=copy.deepcopy(scene)), world)
    ref = Simulation(SimConfiger(config=copy.deepcopy(scene)))
    soil = scene["Configuration"]["simulationMethod"] != 1
    fields = FIELDS + (["stress", "strain_equ", "d_stress"] if soil else [])
    for s in range(4):
        grp.run_steps(1)
        ref.solver.run_steps(1)
        torch.cuda.synchronize()
        for f in fields + IFIELDS + (["flag_retmap"] if soil else []):
            got, want = grp.gather(f).cpu().numpy(), getattr(ref.ps.pt, f).detach().cpu().numpy()
            assert got.shape == want.shape, (name, prec, world, s, f, got.shape, want.shape)
            assert np.array_equal(got, want), f"{name}[{prec}] x{world} step {s + 1}: {f} differs from the single-GPU run"
    assert sum(d.own_count for d in grp.drivers) == grp.global_particle_num


@pytest.mark.parametrize("transport", ["p2p", "dist"])
@pytest.mark.parametrize("name,prec,extra", CASES)
def test_slab_equals_single_gpu(name, prec, extra, transport):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(torch.cuda.device_count(), 4) if name == "c1_test1_wc_lf" else 2
    mp.start_processes(_worker, args=(world, _free_port(), name, prec, 4, extra, transport), nprocs=world, join=True, start_method="spawn")
