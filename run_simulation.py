"""Drop-in for the reference's entry point (run_simulation.py:1-33): same flag, same scene JSON, headless.

    python run_simulation.py --scene_file data/scenes/test1_db_water.json [--max_steps N] [--precision f64|f32]

``ti.init(arch=ti.gpu, default_fp=ti.f64)`` (run_simulation.py:23) corresponds to the engine's float64 mode, which is
the default here too; ``--precision f32`` selects the MIXED mode (fp32 sweeps, fp64 positions and densities).
The GGUI window of ``ui_sim`` (needs a display) is replaced by its headless mirror ``tisphi_b200.eng.ui_sim``.
"""
import argparse
import json

from tisphi_b200.eng.simulation import Simulation, SimConfiger
from tisphi_b200.eng.ui_sim import ui_sim

if __name__ == "__main__":
    parser = argparse.ArgumentParser(description='tiSPHi')
    parser.add_argument('--scene_file', default='', help='scene file')
    # headless additions (the reference stops when its window is closed)
    parser.add_argument('--max_steps', type=int, default=None, help='end the run after this many steps')
    parser.add_argument('--precision', default=None, choices=['f64', 'f32'], help='override Configuration.precision')
    parser.add_argument('--out_dir', default=None, help='where sim_<time stamp>/ is created (default: cwd)')
    parser.add_argument('--checkpoint_every', type=int, default=0, help='write checkpoint.<step>.npz every N steps')
    parser.add_argument('--resume', default=None, help='checkpoint file to continue from')
    parser.add_argument('--device', default='cuda:0')
    args = parser.parse_args()
    scene_path = args.scene_file

    cfg = SimConfiger(scene_file_path=scene_path)
    if args.precision is not None:
        cfg.config["Configuration"]["precision"] = args.precision
    scene_name = scene_path.split("/")[-1].split(".")[0]

    print("\n========== SIMULATION ==========")
    case = Simulation(config=cfg, device=args.device)
    result = ui_sim(case=case, max_steps=args.max_steps, out_dir=args.out_dir, checkpoint_every=args.checkpoint_every,
                    resume=args.resume)
    print(json.dumps({k: v for k, v in result.items() if k != "files"} | {"n_files": len(result["files"]), "scene": scene_name}))

    print("\n========== END ==========")
