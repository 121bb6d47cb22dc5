"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): the slab-partitioned engine over NCCL must
reproduce the single-GPU engine BIT FOR BIT -- cell ids, sorted order and every field -- in both precisions."""
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


def _worker(rank, world, port, name, prec, nsteps, extra):
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
        slab = SlabSimulation(SimConfiger(config=copy.deepcopy(scene)), f"cuda:{rank}", rank, world, check=True)
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


@pytest.mark.parametrize("name,prec,extra", [
    ("wc2d_small_lf", "f64", {}),
    ("wc2d_small_lf", "f32", {}),
    ("wc3d_tiny_lf", "f32", {}),
    ("wc3d_tiny_lf", "f32", {"fastSweeps": False}),
    ("c1_test1_wc_lf", "f32", {}),
    ("wc2d_small_rk4_cspm", "f64", {}),
])
def test_slab_equals_single_gpu(name, prec, extra):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(torch.cuda.device_count(), 4) if name == "c1_test1_wc_lf" else 2
    mp.start_processes(_worker, args=(world, _free_port(), name, prec, 4, extra), nprocs=world, join=True, start_method="spawn")
