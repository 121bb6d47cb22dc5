"""CPU-only tests: the C-ABI library loads and exports every symbol the header declares; host-side scene logic of
the product agrees with the oracle's independent restatement and with the reference fixtures' metadata."""
import ctypes
import os
import re

import numpy as np
import pytest

from helpers import Golden, ROOT

CASES = ["c1_test1_wc_lf", "c2_test2_mui_lf", "c3_test2_dp_rk4_cspm", "wc3d_tiny_lf", "dp2d_small_lf", "dp2d_indenter_lf"]


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from tisphi_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "tisphi_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(sph_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for sym in declared:
        assert hasattr(L, sym), f"{sym} declared in include/tisphi_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared, "tisphi_b200/_lib.py binds a different symbol set than the header declares"


def test_params_struct_matches_header_size():
    """ctypes mirror and C struct agree (sizeof is checked through sph_arena_bytes behaving sanely)."""
    from tisphi_b200 import _lib
    L = _lib.load()
    P = _lib.SphParams()
    P.dim, P.solver, P.ti, P.precision = 3, 1, 2, 1
    P.gn[0], P.gn[1], P.gn[2] = 10, 10, 10
    a = L.sph_arena_bytes(ctypes.byref(P), 1000)
    P.precision = 0
    b = L.sph_arena_bytes(ctypes.byref(P), 1000)
    assert 0 < a < b < 10_000_000
    assert ctypes.sizeof(_lib.SphParams) == L.sph_params_size() == 14 * 4 + 35 * 8         # + boundary, pad, radius, dstart[3], dend[3]


def test_engine_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from tisphi_b200 import _lib
    with pytest.raises(_lib.SphError):
        _lib.Engine(_lib.SphParams(), 10)


@pytest.mark.parametrize("name", CASES)
def test_scene_builder_matches_oracle_and_fixture(name):
    """Product scene builders (tisphi_b200/eng/particle_func.py) vs the oracle's restatement vs the reference run."""
    from oracle import oracle as orc
    from tisphi_b200.eng import particle_func as pf
    g = Golden(name)
    cfg = g.scene["Configuration"]
    D, x, v, rho, typ = orc.build_particles(g.scene)
    dim, d = D["dim"], D["d"]
    xs = []
    for b in g.scene["Blocks"]:
        pf.chk_block_in_domain(D["domain_start"], D["domain_end"], b["translation"], b["size"], dim)
        xs.append(pf.cube_positions(b["translation"], b["size"], dim, d))
    for lo, hi in pf.calc_dummy_boundary(dim, D["domain_start"], D["domain_end"], D["vstart"], D["vend"]):
        xs.append(pf.cube_positions(lo, hi - lo, dim, d))
    mine = np.concatenate(xs)
    assert mine.shape == x.shape == (g.meta["n"], 3)
    assert np.array_equal(mine, x)
    # creation order + positions agree with the reference run: sort the fixture's first snapshot back by id0
    gx = g.grid(1, "x")
    back = np.empty_like(gx)
    back[g.grid(1, "id0")] = gx
    assert np.array_equal(back, mine)


def test_block_outside_domain_raises_like_reference():
    from tisphi_b200.eng import particle_func as pf
    with pytest.raises(AssertionError, match="Block is not in domain!"):
        pf.chk_block_in_domain([0, 0, 0], [1, 1, 1], [0.5, 0.5, 0], [0.6, 0.2, 0.1], 2)


def test_dt_cfl_uses_taichi_float_modulo():
    """test1 gives dt = 1e-4 with Taichi's floor-mod (9e-5 with C fmod); fixtures hold the reference's dt."""
    from tisphi_b200.eng.solver_sph_base import SPHBase

    class Dummy:
        smoothing_len = 0.03
    s = SPHBase.__new__(SPHBase)
    s.ps = Dummy()
    assert s.calc_dt_CFL(0.2, 60, 1e-5) == Golden("c1_test1_wc_lf").meta["dt"] == 1e-4
    Dummy.smoothing_len = 0.003
    assert s.calc_dt_CFL(0.2, 24.0, 1e-6) == Golden("c2_test2_mui_lf").meta["dt"]
    c = Golden("c3_test2_dp_rk4_cspm").meta
    assert s.calc_dt_CFL(0.2, c["vsound"], 1e-6) == c["dt"]


def test_configer_keyerror_and_sections(tmp_path):
    import json
    from tisphi_b200.eng.configer_builder import SimConfiger
    p = tmp_path / "s.json"
    p.write_text(json.dumps({"Configuration": {"is2D": True}}))
    c = SimConfiger(str(p))
    assert c.get_cfg("is2D") is True
    with pytest.raises(KeyError):
        c.get_cfg("kernel")
    assert c.get_blocks() == [] and c.get_materials() == [] and c.get_bodies() == [] and c.get_motions() == []


def test_bench_replays_long_runs_in_stable_legs():
    """bench.py cuts runs longer than the reference scheme's stable horizon (3D WCSPH diverges ~35 steps after rest)
    into legs from the restored initial state; every requested step is taken exactly once."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for total in (1, 20, 21, 50, 100):
        log = []
        bench.run_in_legs(lambda k: log.append(("run", k)), lambda: log.append(("restore", 0)), total)
        runs = [k for what, k in log if what == "run"]
        assert sum(runs) == total and max(runs) <= bench.STABLE_STEPS
        assert [w for w, _ in log] == ["restore", "run"] * len(runs)          # every leg starts from the restored state


def test_reference_module_paths_resolve_to_the_engine():
    """the reference imports ``eng.simulation`` / ``eng.ui_sim`` (run_simulation.py:3-4); the top-level ``eng``
    package re-exports this repository's classes under those paths."""
    import importlib
    import tisphi_b200.eng as impl
    for mod, names in {"simulation": ["Simulation", "SimConfiger"], "ui_sim": ["ui_sim"], "particle_system": ["ParticleSystem"],
                       "solver_sph_base": ["SPHBase"], "solver_sph_wc": ["WCSPHSolver"], "solver_sph_muI": ["MUISPHSolver"],
                       "solver_sph_dp": ["DPSPHSolver"], "configer_builder": ["SimConfiger"], "particle_func": ["add_cube"]}.items():
        shim = importlib.import_module("eng." + mod)
        real = importlib.import_module("tisphi_b200.eng." + mod)
        for n in names:
            assert getattr(shim, n) is getattr(real, n), (mod, n)


def test_bench_state_hash_is_order_independent_and_bit_sensitive():
    """bench.py::state_hash (the full-size slab parity check): permutations and splits of the particle set leave the
    sum unchanged, one flipped mantissa bit changes it."""
    import importlib.util
    import torch
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    g = torch.Generator().manual_seed(7)
    n = 4096
    x = torch.randn(n, 3, dtype=torch.float64, generator=g)
    v = torch.randn(n, 4, generator=g)
    p = torch.randn(n, generator=g)
    gid = torch.randint(0, 1000, (n,), dtype=torch.int32, generator=g)
    id0 = torch.arange(n, dtype=torch.int32)
    h = int(bench.state_hash(torch, [x, v, p, gid], id0))
    perm = torch.randperm(n, generator=g)
    assert int(bench.state_hash(torch, [x[perm], v[perm], p[perm], gid[perm]], id0[perm])) == h
    parts = [slice(0, 1000), slice(1000, 1001), slice(1001, n)]
    total = sum(bench.state_hash(torch, [x[s], v[s], p[s], gid[s]], id0[s]) for s in parts)
    assert int(total) == h
    v2 = v.clone()
    v2[17, 1] = torch.nextafter(v2[17, 1], torch.tensor(10.0))
    assert int(bench.state_hash(torch, [x, v2, p, gid], id0)) != h
    x2 = x.clone()
    x2[[3, 4]] = x2[[4, 3]]                                     # two particles swap their positions: ids matter
    assert int(bench.state_hash(torch, [x2, v, p, gid], id0)) != h


def test_product_code_never_imports_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke(), bench.py's CPU legs and bench_workloads/
    (their cpu_baseline legs) may import it -- never the package, the eng shim or the run script."""
    import ast
    offenders = []
    files = [os.path.join(ROOT, "run_simulation.py")]
    for top in ("tisphi_b200", "eng"):
        for d, _, names in os.walk(os.path.join(ROOT, top)):
            files += [os.path.join(d, n) for n in names if n.endswith(".py")]
    for f in files:
        for node in ast.walk(ast.parse(open(f).read())):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom) and node.module:
                mods = [node.module]
            if any(m == "oracle" or m.startswith("oracle.") for m in mods):
                offenders.append(os.path.relpath(f, ROOT))
    assert not offenders, offenders
    assert len(files) > 15
