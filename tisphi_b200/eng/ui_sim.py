"""Headless mirror of the reference's viewer loop ``eng/ui_sim.py``.

The reference drives ``case.solver.step()`` from a Taichi GGUI window (``ui:105``) and applies its stop / exit /
export rules once per rendered frame (``ui:211-240``).  There is no display on a GPU box, so this module keeps the
control flow and the exporters and drops the window:

* a "frame" is ``stepsPerRenderUpdate`` steps enqueued by ONE native call (``solver.run_steps``);
* every rule that would PAUSE the viewer (``pauseAtStart`` aside, which has nobody to press SPACE and is ignored)
  ends a headless run, because nobody can resume it: ``stopAtStep``, ``stopAtTime``, ``stopEveryStep``;
* ``exitAtStep`` / ``exitAtTime`` end the run as in the reference (``ui:236-241``); ``max_steps`` is the headless
  substitute for closing the window;
* exports keep the reference's schedule (``exportEveryTime`` / ``exportEveryRender``, ``ui:224-233``), directory
  name ``sim_<time stamp>`` (``ui:50-55``), file stamps (``ui:306-323``), CSV columns and header (``ui:293-298``) and
  ``_info.txt`` (``ui:176-178, 325-328``).  ``exportVTK`` writes the same point data ``pyevtk.pointsToVTK`` would
  (``ui:300-304``) as a ``.vtu`` file with our own writer (pyevtk is not a dependency); ``exportFrame`` (a screenshot
  in the reference) stores positions + the scalar of ``colorTitle`` as ``.npz``.
* checkpoint / resume (not in the reference): ``save_checkpoint`` / ``load_checkpoint`` store the persistent members.

Paths are joined with ``os.path.join`` (the reference hard-codes Windows separators, ``ui:47,51,295``).
"""
import base64
import os
import struct
from datetime import datetime

import numpy as np

CSV_COLUMNS = ["id0", "objId", "material", "pos.x", "pos.y", "pos.z", "vel.x", "vel.y", "vel.z", "density", "stress.xx",
               "stress.yy", "stress.zz", "stress.xy", "stress.yz", "stress.zx", "strain_equ"]
CSV_HEADER = ", ".join(CSV_COLUMNS)        # ui:298


def get_time_stamp():
    return datetime.today().strftime("%Y_%m_%d_%H%M%S")     # ui:247-248


# ------------------------------------------------------------------------------------------------------ rules
class RunControl:
    """The stop / exit / export decisions of ``ui:211-241`` as a pure state machine (no engine, no files): given the
    step counter after a frame it says whether to export and whether the run ends.  Tested on the CPU."""

    def __init__(self, cfg, dt):
        g = cfg.get_cfg
        self.dt = float(dt)
        self.substeps = max(1, int(g("stepsPerRenderUpdate")))                       # ui:60
        self.stop_at_step, self.exit_at_step = int(g("stopAtStep")), int(g("exitAtStep"))
        self.stop_at_time, self.exit_at_time = float(g("stopAtTime")), float(g("exitAtTime"))
        self.stop_every_step = int(g("stopEveryStep"))
        self.stop_at_step_tmp = self.stop_every_step                                 # ui:61
        self.save_every_time = float(g("exportEveryTime"))
        self.save_every_time_tmp = self.save_every_time                              # ui:62
        self.save_every_render = int(g("exportEveryRender"))
        self.save_frame, self.save_vtk, self.save_csv = bool(g("exportFrame")), bool(g("exportVTK")), bool(g("exportCSV"))
        self.exports = (self.save_every_render > 0 or self.save_every_time > 0) and \
                       (self.save_frame or self.save_vtk or self.save_csv)           # judge_sim_path, ui:49

    def after_frame(self, count_step):
        """-> (export: None | ("time", cur_time) | ("step", count_step), paused: bool, exit: bool)"""
        cur_time = self.dt * count_step                                              # ui:137
        paused = False
        if count_step >= self.stop_at_step and self.stop_at_step > 0:               # ui:212-214
            paused, self.stop_at_step = True, 0
        if self.stop_every_step > 0 and self.stop_every_step >= self.substeps:      # ui:215-218
            if count_step >= self.stop_at_step_tmp:
                paused = True
                self.stop_at_step_tmp += self.stop_every_step
        if cur_time >= self.stop_at_time and self.stop_at_time > 0:                 # ui:219-221
            paused, self.stop_at_time = True, 0
        export = None
        if self.save_every_time > 0 and self.exports:                                # ui:224-229
            if count_step == 0 and not paused:
                export = ("time", cur_time)
            elif cur_time >= self.save_every_time:
                export = ("time", cur_time)
                self.save_every_time += self.save_every_time_tmp
        elif self.save_every_render > 0 and self.exports:                            # ui:230-232
            if (count_step == 0 and not paused) or \
                    (count_step % (self.save_every_render * self.substeps) == 0 and count_step > 0):
                export = ("step", count_step)
        done = (count_step >= self.exit_at_step and self.exit_at_step > 0) or \
               (cur_time >= self.exit_at_time and self.exit_at_time > 0)             # ui:236
        return export, paused, done

    def fast_forward(self, count_step):
        """A run resumed from a checkpoint taken at ``count_step``: the rules that already fired up to (and including) that
        step are spent -- one-shot stops are cleared, the periodic thresholds move past the restored step / time -- so
        that a run that had stopped at one of them continues instead of stopping again at once."""
        cur_time = self.dt * count_step
        if self.stop_at_step > 0 and count_step >= self.stop_at_step:
            self.stop_at_step = 0
        if self.stop_at_time > 0 and cur_time >= self.stop_at_time:
            self.stop_at_time = 0
        if self.stop_every_step > 0:
            while self.stop_at_step_tmp <= count_step:
                self.stop_at_step_tmp += self.stop_every_step
        if self.save_every_time > 0:
            while self.save_every_time <= cur_time:
                self.save_every_time += self.save_every_time_tmp


def stamp_of(export):
    kind, value = export
    if kind == "step":
        return f"{int(value):06d}"                                                   # ui:307
    return f"time.secx1e6.{int(value * 1e6):07d}"                                   # ui:316-317


# ------------------------------------------------------------------------------------------------------ exporters
def export_csv(stamp, simpath, case):
    """ui:293-298: one row per particle, 17 columns, '# '-prefixed header line (numpy.savetxt default)."""
    pos, data = case.ps.dump()
    cols = [data["id0"], data["objId"], data["material"], pos["pos.x"], pos["pos.y"], pos["pos.z"], data["vel.x"],
            data["vel.y"], data["vel.z"], data["density"], data["stress.xx"], data["stress.yy"], data["stress.zz"],
            data["stress.xy"], data["stress.yz"], data["stress.zx"], data["strain_equ"]]
    fname = os.path.join(simpath, "sim.csv.%s.csv" % stamp)
    np.savetxt(fname, np.array(cols).T, delimiter=",", header=CSV_HEADER)
    return fname


def write_vtu(fname, x, y, z, data):
    """Point cloud as a VTK XML UnstructuredGrid (.vtu): points, one VTK_VERTEX cell per point, and every array of
    ``data`` as point data under its own name -- the content ``pyevtk.hl.pointsToVTK`` writes (ui:300-304).  Binary
    payloads are inline base64 with a UInt32 byte-count header (VTK's uncompressed 'binary' format)."""
    n = len(x)

    def enc(a):
        raw = np.ascontiguousarray(a).tobytes()
        return base64.b64encode(struct.pack("<I", len(raw))).decode() + base64.b64encode(raw).decode()

    def arr(name, a, ncomp=1):
        a = np.asarray(a)
        if a.dtype.kind == "f":
            a, ty = a.astype("<f8"), "Float64"
        else:
            a, ty = a.astype("<i8"), "Int64"
        return f'<DataArray type="{ty}" Name="{name}" NumberOfComponents="{ncomp}" format="binary">{enc(a)}</DataArray>\n'

    pts = np.stack([np.asarray(x, dtype="<f8"), np.asarray(y, dtype="<f8"), np.asarray(z, dtype="<f8")], axis=1)
    with open(fname, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian" header_type="UInt32">\n')
        f.write(f'<UnstructuredGrid>\n<Piece NumberOfPoints="{n}" NumberOfCells="{n}">\n<Points>\n')
        f.write(arr("points", pts, 3))
        f.write("</Points>\n<Cells>\n")
        f.write(arr("connectivity", np.arange(n, dtype=np.int64)))
        f.write(arr("offsets", np.arange(1, n + 1, dtype=np.int64)))
        f.write(f'<DataArray type="UInt8" Name="types" format="binary">{enc(np.ones(n, dtype=np.uint8))}</DataArray>\n')
        f.write("</Cells>\n<PointData>\n")
        for name, a in data.items():
            f.write(arr(name, a))
        f.write("</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n")
    return fname


def read_vtu(fname):
    """Inverse of write_vtu for the tests: {'points': (n,3), name: array}."""
    import re
    txt = open(fname).read()
    out = {}
    for m in re.finditer(r'<DataArray type="(\w+)" Name="([^"]+)"(?: NumberOfComponents="(\d+)")? format="binary">([^<]*)</DataArray>', txt):
        ty, name, ncomp, payload = m.group(1), m.group(2), int(m.group(3) or 1), m.group(4)
        nbytes = struct.unpack("<I", base64.b64decode(payload[:8]))[0]
        raw = base64.b64decode(payload[8:])[:nbytes]
        a = np.frombuffer(raw, dtype={"Float64": "<f8", "Int64": "<i8", "UInt8": "u1"}[ty])
        out[name] = a.reshape(-1, ncomp) if ncomp > 1 else a
    return out


def export_vtk(stamp, simpath, case):
    pos, data = case.ps.dump()
    return write_vtu(os.path.join(simpath, "sim.vtk.%s.vtu" % stamp), pos["pos.x"], pos["pos.y"], pos["pos.z"], data)


def export_frame(stamp, simpath, case):
    """Stands in for window.save_image (ui:309, 319): what the viewer would draw -- positions, the scalar selected by
    ``colorTitle`` (solver.assign_value_color), its range (ps.v_maxmin) and the jet colours (ps.set_color)."""
    pos, data = case.ps.dump()
    extra = {}
    if hasattr(case, "solver") and case.ps.color_title > 0:                  # assign_color, ui:286-290
        cfg = case.cfg.get_cfg
        case.solver.assign_value_color()
        case.ps.v_maxmin(cfg("givenMax"), cfg("givenMin"), cfg("fixMax"), cfg("fixMin"))
        case.ps.set_color()
        value = case.ps.pt.val.detach().cpu().numpy()
        extra = {"color": case.ps.pt.color.detach().cpu().numpy(), "vmax": case.ps.vmax[None], "vmin": case.ps.vmin[None]}
    else:
        value = data.get(_COLOR_KEY.get(case.ps.color_title, "vel.norm"), data["vel.norm"])
    fname = os.path.join(simpath, "%s.npz" % stamp)
    np.savez_compressed(fname, x=pos["pos.x"], y=pos["pos.y"], z=pos["pos.z"], value=value,
                        title=choose_color_title(case.ps.color_title), material=data["material"], **extra)
    return fname


# colorTitle codes of the reference (solver_sph_base.py:721-789): the dump() key each one colours by
_COLOR_KEY = {1: "id0", 2: "density", 21: "d_density", 3: "vel.norm", 31: "vel.x", 32: "vel.y", 33: "vel.z", 4: "pos.norm",
              5: "stress.yy", 51: "stress.xx", 52: "stress.yy", 53: "stress.zz", 54: "stress.xy", 6: "strain_equ",
              7: "pressure"}


def choose_color_title(code):
    return _COLOR_KEY.get(int(code), "vel.norm")


def info_text(case, ctl):
    """The strings of the 'Running Info' / 'Simulation Info' panels and _info.txt (ui:136-178)."""
    s, ps = case.solver, case.ps
    str_pt_num = "Total particle number: {ptnum:,}".format(ptnum=ps.particle_num[None])
    str_dt = "dt={dt:.6f}s".format(dt=s.dt[None])
    str_solver = "Solver: " + {1: "Weakly Compressible", 2: "Mohr-Coulomb mu(I)", 3: "Drucker-Prager"}.get(case.solver_type, "None")
    str_ti = "Time integ: " + {1: "1 Symplectic Euler", 2: "2 Leap-Frog", 4: "4 Runge-Kutta"}.get(s.flagTI, "None")
    str_bdy = "Boundary: " + {ps.bdy_collision: "Enforced collision", ps.bdy_dummy: "Dummy particles",
                              ps.bdy_rep: "Repulsive particles", ps.bdy_dummy_rep: "Dummy + repulsive pts"}.get(ps.flag_boundary, "None")
    str_kernel = "Kernel func: " + {0: "Cubic spline", 1: "Wendland C2"}.get(s.flagKernel, "None")
    str_corr = "Kernel corr: " + {1: "CSPM", 2: "MLS"}.get(s.flagKernelCorr, "None")
    str_pos = "Position upd: " + ("XSPH" if s.flagXSPH == 1 else "None")
    str_comment = "Comment: " + str(case.cfg.get_cfg("comment"))
    return ("==== Running Info ====\n%s\n%s\n\n==== Simulation Info ====\n%s\n%s\n%s\n%s\n%s\n%s\n%s\n\n==== Note ====\n"
            "\"time.secx1e6.0349220\" means the frame of 0.349220s\n\n\n\n==== Configure Info ====\n%s"
            % (str_pt_num, str_dt, str_solver, str_ti, str_bdy, str_kernel, str_corr, str_pos, str_comment, case.cfg.config))


def save_info(simpath, text):
    with open(os.path.join(simpath, "_info.txt"), "w") as f:       # ui:325-328
        f.write(text)


# ------------------------------------------------------------------------------------------------------ checkpoints
# every member that travels through the sort (sph_state_fields): v_tmp / density_tmp of WALL particles are carried state
# in mu(I) (the soil loop of a one_step reads the wall velocities extrapolated by the PREVIOUS one_step)
_CKPT_FIELDS = ("x", "v", "density", "pressure", "mat_type", "id0", "v_tmp", "density_tmp")
_CKPT_SOIL = ("strain_equ", "strain_equ_p", "flag_retmap")


def save_checkpoint(path, case, count_step):
    """Persistent members in current order + the step counter.  Everything else is recomputed by the next step()."""
    pt = case.ps.pt
    out = {k: getattr(pt, k).detach().cpu().numpy() for k in _CKPT_FIELDS}
    out["m_V"] = pt.m_V.detach().cpu().numpy()
    out["mass"] = pt.mass.detach().cpu().numpy()
    if case.solver_type != 1:
        out["stress6"] = case.ps.sym6("stress").detach().cpu().numpy()
        for k in _CKPT_SOIL:
            out[k] = getattr(pt, k).detach().cpu().numpy()
    out["count_step"] = np.int64(count_step)
    out["dt"] = np.float64(case.solver.dt[None])
    np.savez(path, **out)
    return path


def load_checkpoint(path, case):
    """Overwrites the state of a freshly built ``case`` (same scene) with a checkpoint; returns the step counter."""
    import torch
    z = np.load(path)
    pt = case.ps.pt
    n = case.ps.particle_num[None]
    if len(z["id0"]) != n:
        raise ValueError(f"checkpoint holds {len(z['id0'])} particles, the scene builds {n}")
    dev = case.ps.engine.device
    for k in _CKPT_FIELDS + ("m_V", "mass"):
        dst = getattr(pt, k)
        dst.copy_(torch.from_numpy(z[k]).to(dev).to(dst.dtype))
    if case.solver_type != 1:
        dst = case.ps.sym6("stress")
        dst.copy_(torch.from_numpy(z["stress6"]).to(dev).to(dst.dtype))
        for k in _CKPT_SOIL:
            dst = getattr(pt, k)
            dst.copy_(torch.from_numpy(z[k]).to(dev).to(dst.dtype))
    return int(z["count_step"])


# ------------------------------------------------------------------------------------------------------ the loop
def ui_sim(case, max_steps=None, out_dir=None, checkpoint_every=0, resume=None, log=print):
    """Runs ``case`` headless under the reference's rules.  Returns a dict with the step count, simulated time, the
    reason the run ended and the exported files."""
    ctl = RunControl(case.cfg, case.solver.dt[None])
    if case.cfg.get_cfg("pauseAtStart"):
        log("pauseAtStart is ignored in a headless run")
    base = out_dir if out_dir is not None else os.getcwd()
    simpath = None
    if ctl.exports or checkpoint_every > 0:
        simpath = os.path.join(base, "sim_" + get_time_stamp())                      # ui:50-54
        os.makedirs(simpath, exist_ok=True)
    log("UI %dD starts to serve! (headless)" % case.ps.dim)                          # ui:88-91
    count_step = load_checkpoint(resume, case) if resume else 0
    if resume:
        ctl.fast_forward(count_step)
    files, info_saved, reason = [], False, "max_steps"

    def do_export(export):
        stamp = stamp_of(export)
        if ctl.save_frame:
            files.append(export_frame(stamp, simpath, case))
        if ctl.save_vtk:
            files.append(export_vtk(stamp, simpath, case))
        if ctl.save_csv:
            files.append(export_csv(stamp, simpath, case))

    # the reference evaluates its rules once before the first step too (count_step == 0 exports the initial state);
    # a resumed run has been through the rules of its restored step already
    first = not resume
    while True:
        if not first:
            nstep = ctl.substeps if max_steps is None else min(ctl.substeps, max_steps - count_step)
            if nstep <= 0:
                break
            case.solver.run_steps(nstep)                                             # ui:102-106
            count_step += nstep
        first = False
        if ctl.exports and not info_saved:
            save_info(simpath, info_text(case, ctl))                                 # ui:176-178
            info_saved = True
        export, paused, done = ctl.after_frame(count_step)
        if export is not None:
            do_export(export)
        if checkpoint_every > 0 and count_step > 0 and count_step % checkpoint_every == 0:
            files.append(save_checkpoint(os.path.join(simpath, "checkpoint.%06d.npz" % count_step), case, count_step))
        if done:
            reason = "exit"
            break
        if paused:
            reason = "stop"
            log("Simulation is stopped!")
            break
        if max_steps is not None and count_step >= max_steps:
            break
    bad = case.ps.engine.L.sph_read_bad_cells(case.ps.engine.h)
    if bad:
        log(f"warning: {bad} particle-steps fell outside the padded grid and were clamped")
    log("Simulator exits!")
    return {"steps": count_step, "time": case.solver.dt[None] * count_step, "reason": reason, "simpath": simpath,
            "files": files, "bad_cells": int(bad)}
