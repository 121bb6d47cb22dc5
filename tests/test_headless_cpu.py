"""Host logic of the headless driver (tisphi_b200/eng/ui_sim.py): the reference's stop / exit / export rules
(eng/ui_sim.py:211-241 of the reference), file stamps, CSV layout and the VTU writer.  No GPU."""
import json
import os

import numpy as np

from tisphi_b200.eng import ui_sim as U
from tisphi_b200.eng.configer_builder import SimConfiger

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cfg_with(**over):
    scene = json.load(open(os.path.join(ROOT, "data", "scenes", "test1_db_water.json")))
    scene["Configuration"].update(dict(stopEveryStep=0, stopAtStep=0, exitAtStep=0, stopAtTime=0, exitAtTime=0, exportEveryTime=0,
                                       exportEveryRender=0, exportFrame=False, exportVTK=False, exportCSV=False, pauseAtStart=False))
    scene["Configuration"].update(over)
    return SimConfiger(config=scene)


def run_rules(ctl, frames):
    out, step = [], 0
    for k in range(frames + 1):
        export, paused, done = ctl.after_frame(step)
        out.append((step, export, paused, done))
        if paused or done:
            break
        step += ctl.substeps
    return out


def test_exit_at_step_and_export_every_render():
    ctl = U.RunControl(cfg_with(stepsPerRenderUpdate=10, exitAtStep=50, exportEveryRender=2, exportCSV=True), dt=1e-4)
    assert ctl.exports
    log = run_rules(ctl, 100)
    assert [s for s, e, p, d in log if e is not None] == [0, 20, 40]         # step 0, then every 2 renders x 10 steps
    assert all(e[0] == "step" for s, e, p, d in log if e is not None)
    assert log[-1][0] == 50 and log[-1][3] and not log[-1][2]               # ends exactly at exitAtStep


def test_export_every_time_uses_simulated_time():
    ctl = U.RunControl(cfg_with(stepsPerRenderUpdate=5, exitAtTime=0.0101, exportEveryTime=0.002, exportVTK=True), dt=1e-4)
    log = run_rules(ctl, 1000)
    stamps = [U.stamp_of(e) for s, e, p, d in log if e is not None]
    assert stamps[0] == "time.secx1e6.0000000"
    assert stamps[1:] == ["time.secx1e6.%07d" % (2000 * k) for k in range(1, 6)]
    assert abs(log[-1][0] * 1e-4 - 0.0105) < 1e-12 and log[-1][3]           # first frame at or after exitAtTime


def test_stop_rules_pause_the_run():
    assert run_rules(U.RunControl(cfg_with(stepsPerRenderUpdate=10, stopAtStep=30), dt=1e-4), 100)[-1][:3:2] == (30, True)
    assert run_rules(U.RunControl(cfg_with(stepsPerRenderUpdate=10, stopAtTime=0.0025), dt=1e-4), 100)[-1][:3:2] == (30, True)
    assert run_rules(U.RunControl(cfg_with(stepsPerRenderUpdate=10, stopEveryStep=40), dt=1e-4), 100)[-1][:3:2] == (40, True)
    # stopEveryStep smaller than a frame never fires (ui:215)
    assert not any(p for s, e, p, d in run_rules(U.RunControl(cfg_with(stepsPerRenderUpdate=10, stopEveryStep=5), dt=1e-4), 20))


def test_resumed_run_does_not_fire_spent_rules_again():
    """A run that stopped at stopAtStep / stopEveryStep and is resumed from a checkpoint of that step continues; an
    exportEveryTime schedule picks up at the next multiple instead of exporting every frame until it has caught up."""
    def resumed(step, **over):
        ctl = U.RunControl(cfg_with(stepsPerRenderUpdate=10, **over), dt=1e-4)
        ctl.fast_forward(step)
        out = []
        for k in range(1, 13):
            s = step + 10 * k
            export, paused, done = ctl.after_frame(s)
            out.append((s, export, paused, done))
            if paused or done:
                break
        return out
    assert resumed(30, stopAtStep=30, exitAtStep=100)[-1][0] == 100             # not stopped again at 40
    assert resumed(40, stopEveryStep=40)[-1][::2] == (80, True)                  # the NEXT multiple
    assert resumed(30, stopAtTime=0.003, exitAtStep=60)[-1][0] == 60
    log = resumed(50, exportEveryTime=0.002, exportCSV=True, exitAtStep=170)
    assert [s for s, e, p, d in log if e is not None] == [60, 80, 100, 120, 140, 160]


def test_no_exports_without_a_format_or_a_schedule():
    assert not U.RunControl(cfg_with(exportEveryRender=2), dt=1e-4).exports
    assert not U.RunControl(cfg_with(exportCSV=True), dt=1e-4).exports
    assert U.stamp_of(("step", 120)) == "000120"


def test_vtu_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    n = 1234
    x, y, z = rng.normal(size=(3, n))
    data = {"id0": np.arange(n, dtype=np.int64), "density": rng.uniform(900, 1100, n), "vel.x": rng.normal(size=n)}
    f = U.write_vtu(str(tmp_path / "a.vtu"), x, y, z, data)
    back = U.read_vtu(f)
    assert np.array_equal(back["points"], np.stack([x, y, z], axis=1))
    for k, v in data.items():
        assert np.array_equal(back[k], v), k
    assert np.array_equal(back["connectivity"], np.arange(n)) and np.array_equal(back["offsets"], np.arange(1, n + 1))
    assert (back["types"] == 1).all()                                        # VTK_VERTEX
    txt = open(f).read()
    assert txt.startswith('<?xml version="1.0"?>') and f'NumberOfPoints="{n}"' in txt


class _FakePS:
    color_title = 7

    def dump(self):
        n = 5
        pos = {"pos.x": np.arange(n) * 0.1, "pos.y": np.arange(n) * 0.2, "pos.z": np.zeros(n)}
        data = {k: np.arange(n, dtype=np.float64) + j for j, k in enumerate(
            ["id0", "objId", "material", "vel.x", "vel.y", "vel.z", "vel.norm", "density", "stress.xx", "stress.yy", "stress.zz",
             "stress.xy", "stress.yz", "stress.zx", "strain_equ", "pressure"])}
        return pos, data


class _FakeCase:
    ps = _FakePS()


def test_csv_layout_matches_the_reference(tmp_path):
    f = U.export_csv("000010", str(tmp_path), _FakeCase())
    assert os.path.basename(f) == "sim.csv.000010.csv"
    lines = open(f).read().splitlines()
    assert lines[0] == "# id0, objId, material, pos.x, pos.y, pos.z, vel.x, vel.y, vel.z, density, stress.xx, stress.yy, " \
                       "stress.zz, stress.xy, stress.yz, stress.zx, strain_equ"
    rows = np.loadtxt(f, delimiter=",")
    assert rows.shape == (5, 17)
    assert np.allclose(rows[:, 3], np.arange(5) * 0.1) and np.allclose(rows[:, 9], np.arange(5) + 7)   # pos.x, density
    g = U.export_frame("000010", str(tmp_path), _FakeCase())
    z = np.load(g)
    assert str(z["title"]) == "pressure" and np.array_equal(z["value"], np.arange(5) + 15.0)


def test_scene_files_are_the_baseline_configs():
    d = os.path.join(ROOT, "data", "scenes")
    c1 = json.load(open(os.path.join(d, "test1_db_water.json")))["Configuration"]
    assert (c1["simulationMethod"], c1["timeIntegration"], c1["kernel"], c1["boundary"], c1["is2D"]) == (1, 2, 1, 2, True)
    c2 = json.load(open(os.path.join(d, "test2_cc_sand_muI.json")))["Configuration"]
    assert (c2["simulationMethod"], c2["timeIntegration"], c2["xsph"]) == (2, 2, True)
    c3 = json.load(open(os.path.join(d, "test2_cc_sand_dp_rk4_cspm.json")))["Configuration"]
    assert (c3["simulationMethod"], c3["timeIntegration"], c3["kernelCorrection"]) == (3, 4, 1)
    c4 = json.load(open(os.path.join(d, "c4_db3d_water_13M.json")))
    assert c4["Configuration"]["is2D"] is False and c4["Blocks"][0]["size"] == [1.6, 1.0, 0.8]
    for f in os.listdir(d):
        assert "precision" not in json.load(open(os.path.join(d, f)))["Configuration"], f


# ------------------------------------------------------------------------------------------------ viewer helpers (host)
def test_jet_colour_map_is_the_reference_tent_map():
    import torch
    from tisphi_b200.eng.colormap import color_map, ColorMap
    x = torch.tensor([-0.5, 0.0, 0.25, 0.5, 0.75, 1.0, 1.5], dtype=torch.float64)
    rgb = color_map(x).double().numpy()
    assert rgb.shape == (7, 3) and rgb.min() >= 0.0 and rgb.max() <= 1.0
    # colormap.py:20-27 by hand: clamp((w - |clamp(x) - c|) / w * h), jet = (1.5, .37, .37, c) with c = .75 / .5 / .25
    def tent(v, c):
        v = min(1.0, max(0.0, v))
        return min(1.0, max(0.0, (0.37 - abs(v - c)) / 0.37 * 1.5))
    for k, v in enumerate(x.tolist()):
        want = [tent(v, 0.75), tent(v, 0.5), tent(v, 0.25)]
        assert np.allclose(rgb[k], want, atol=1e-6), (v, rgb[k], want)
    assert np.allclose(rgb[3], [0.0, 1.0, 0.0], atol=0.49) and rgb[3][1] == 1.0          # mid-range is green
    assert rgb[1][2] > rgb[1][0] and rgb[5][0] > rgb[5][2]                               # blue end, red end
    asym = ColorMap(1.0, .25, 1, .5)                                                       # bwrR: different widths left / right
    assert abs(float(asym.map(torch.tensor([0.4]))[0]) - (0.25 - 0.1) / 0.25) < 1e-6
    assert abs(float(asym.map(torch.tensor([0.9]))[0]) - (1 - 0.4) / 1) < 1e-6


def test_cell_index_helpers_match_the_oracle_grid():
    from oracle import oracle as orc
    from tisphi_b200.eng.particle_system import pos_to_index, flatten_grid_index
    scene = json.load(open(os.path.join(ROOT, "data", "scenes", "test1_db_water.json")))
    o = orc.Oracle.from_scene(scene, serial=0)
    o.grid_build()
    D = o.D
    idx = pos_to_index(o.x, D["vstart"], D["grid_size"])
    gn = [int(g) for g in D["grid_num"]]
    if len(gn) == 2 or gn[2] == 0:
        gn = [gn[0], gn[1], 1]
    idx[:, 2] = 0                                                                         # 2D scene: z plays no part
    assert np.array_equal(flatten_grid_index(idx, gn), o.grid_ids)


def test_value_range_follows_the_reference_precedence():
    import torch
    from tisphi_b200.eng.particle_system import value_range
    val = torch.tensor([1.0, 5.0, 9.0, 100.0])
    flow = torch.tensor([True, True, True, False])
    assert value_range(val, flow, -1, -1, 0, 0) == (9.0, 1.0)                              # no given range
    assert value_range(val, flow, 20.0, 0.5, 0, 0) == (9.0, 1.0)                           # given range wider: data wins
    assert value_range(val, flow, 6.0, 2.0, 0, 0) == (6.0, 2.0)                            # given range narrower: capped
    assert value_range(val, flow, 20.0, 0.5, 1, 1) == (20.0, 0.5)                          # fixed: always the given values
    assert value_range(val, torch.zeros(4, dtype=torch.bool), -1, -1, 0, 0) == (-float("inf"), float("inf"))
