"""Debug helper: 2-rank slab run of a small fixture scene with a traceback dump if it hangs."""
import faulthandler, os, sys, datetime
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def worker(rank, world, port, name, prec):
    import copy
    import numpy as np
    import torch
    import torch.distributed as dist
    from helpers import Golden
    from tisphi_b200.eng.simulation import Simulation, SimConfiger
    from tisphi_b200.parallel import SlabSimulation
    faulthandler.dump_traceback_later(45, exit=True, file=sys.stderr)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device(f"cuda:{rank}"), timeout=datetime.timedelta(seconds=60))
    g = Golden(name)
    scene = copy.deepcopy(g.scene)
    scene["Configuration"]["precision"] = prec
    slab = SlabSimulation(SimConfiger(config=copy.deepcopy(scene)), f"cuda:{rank}", rank, world, check=True)
    ref = Simulation(SimConfiger(config=copy.deepcopy(scene)), device=f"cuda:{rank}") if rank == 0 else None
    for s in range(4):
        print(f"[{rank}] step {s} begin n={slab.ps.engine.n} own=({slab.driver.own_first},{slab.driver.own_count})", file=sys.stderr, flush=True)
        slab.run_steps(1)
        torch.cuda.synchronize()
        print(f"[{rank}] step {s} done n={slab.ps.engine.n} ghosts={slab.driver.ghost_l},{slab.driver.ghost_r}", file=sys.stderr, flush=True)
        mine = {f: slab.owned(f).detach().cpu().numpy() for f in ("x", "v", "density", "pressure", "id0", "CSPM_f")}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            ref.solver.step()
            for f in mine:
                got = np.concatenate([gd[f] for gd in gathered])
                want = getattr(ref.ps.pt, f).detach().cpu().numpy()
                same = got.shape == want.shape and np.array_equal(got, want)
                err = 0.0 if same or got.shape != want.shape else float(np.max(np.abs(got.astype(np.float64) - want)) / (np.max(np.abs(want)) + 1e-300))
                print(f"step {s} {f}: equal={same} shapes {got.shape} {want.shape} relerr={err:.2e}", file=sys.stderr, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    import torch.multiprocessing as mp
    name, prec = sys.argv[1], sys.argv[2]
    mp.start_processes(worker, args=(2, 29533, name, prec), nprocs=2, join=True, start_method="spawn")
