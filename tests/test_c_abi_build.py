"""The header is plain C (not only C++) and a host program without Python links against the library."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def test_header_compiles_as_c99_and_cpp(tmp_path):
    src = tmp_path / "inc.c"
    src.write_text('#include "tisphi_b200.h"\nint main(void) { SphParams p; (void)p; return (int)(sizeof(SphParams) == 0); }\n')
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", INC, "-c", str(src), "-o", str(tmp_path / "a.o")], check=True)
    cpp = tmp_path / "inc.cpp"
    cpp.write_text(src.read_text())
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", INC, "-c", str(cpp), "-o", str(tmp_path / "b.o")], check=True)


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not found")
def test_c_host_example_links_against_the_library(tmp_path):
    import __graft_entry__   # noqa: F401  (puts the repo on sys.path)
    from tisphi_b200 import _build
    _build.build()
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    lib_dir = os.path.join(ROOT, "tisphi_b200")
    exe = tmp_path / "c_abi_minimal"
    r = subprocess.run([nvcc, "-o", str(exe), os.path.join(ROOT, "examples", "c_abi_minimal.c"), "-I", INC, "-L", lib_dir,
                        "-ltisphi_b200", "-Xlinker", f"-rpath={lib_dir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert exe.exists()


@pytest.mark.gpu
def test_c_host_example_runs_and_matches_the_oracle(tmp_path):
    """examples/c_abi_minimal.c EXECUTED on the GPU (no Python, no torch in that process): its 10 steps of 3D WCSPH on
    its own particle set against the float64 CPU oracle on the same particles."""
    import numpy as np
    from oracle import oracle as orc
    from helpers import relmax
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    lib_dir = os.path.join(ROOT, "tisphi_b200")
    exe, out = tmp_path / "c_abi_minimal", tmp_path / "state.bin"
    subprocess.run([nvcc, "-o", str(exe), os.path.join(ROOT, "examples", "c_abi_minimal.c"), "-I", INC, "-L", lib_dir,
                    "-ltisphi_b200", "-Xlinker", f"-rpath={lib_dir}"], check=True, capture_output=True)
    r = subprocess.run([str(exe), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = out.read_bytes()
    n = int(np.frombuffer(raw, dtype=np.int64, count=1)[0])
    off = 8
    id0 = np.frombuffer(raw, dtype=np.int32, count=n, offset=off); off += 4 * n
    x = np.frombuffer(raw, dtype=np.float64, count=3 * n, offset=off).reshape(n, 3); off += 24 * n
    v = np.frombuffer(raw, dtype=np.float32, count=4 * n, offset=off).reshape(n, 4); off += 16 * n
    rho = np.frombuffer(raw, dtype=np.float64, count=n, offset=off)
    # the same particles and constants as the C program builds
    d, nx, ny, nz, layers, size = 0.02, 20, 15, 10, 3, (1.0, 0.6, 0.4)
    h, gs = 1.5 * d, 3 * d
    fx, fz = int(size[0] / d), int(size[2] / d)
    I, J, K = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    xf = np.stack([(I + 0.5) * d, (J + 0.5) * d, (K + 0.5) * d], axis=-1).reshape(-1, 3)
    I, J, K = np.meshgrid(np.arange(fx), np.arange(layers), np.arange(fz), indexing="ij")
    xw = np.stack([(I + 0.5) * d, -(J + 0.5) * d, (K + 0.5) * d], axis=-1).reshape(-1, 3)
    x0 = np.concatenate([xf, xw])
    assert n == len(x0)
    P = orc.OrcParams()
    P.dim, P.kernel, P.kcorr, P.ti, P.xsph, P.solver, P.serial, P.wc_fresh = 3, 1, 0, 2, 0, 1, 0, 0
    for a in range(3):
        P.vstart[a], P.gn[a], P.g[a] = -gs, int(np.ceil((size[a] + 2 * gs) / gs)), 0.0
        P.dstart[a], P.dend[a] = 0.0, size[a]
    P.g[1] = -9.81
    P.h, P.support, P.grid_size, P.m_V0, P.eps = h, 2 * h, gs, d ** 3, 1e-8
    P.rho0, P.visc, P.stiff, P.gamma_, P.vsound, P.boundary, P.radius = 1000.0, 0.01, 5e5, 7.0, 60.0, 2, d / 2
    P.dt = 0.2 * h / 60.0
    o = orc.Oracle(P, x0, np.zeros_like(x0), np.concatenate([np.full(len(xf), 1000.0), np.zeros(len(xw))]),
                   np.concatenate([np.ones(len(xf), np.int32), -np.ones(len(xw), np.int32)]))
    for _ in range(10):
        assert o.step() == 0
    pa, pb = np.argsort(id0), np.argsort(o.id0)
    assert relmax(x[pa], o.x[pb]) < 1e-7 and relmax(rho[pa], o.density[pb]) < 1e-6
    assert relmax(v[pa][:, :3], o.v[pb]) < 1e-4
