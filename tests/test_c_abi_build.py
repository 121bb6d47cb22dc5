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
