import os
import sys

import pytest

# tests/test_gpu_slab.py steps several slab contexts of ONE process in turn: a stream's one-block wait kernel spins until
# another stream's push kernel has run.  With CUDA's default lazy module loading the first launch of a kernel may have to
# synchronise with the context, i.e. with the spinning kernel -- the host would stall until the wait times out.  Eager
# loading (read by the driver when CUDA initialises, so it is set before anything touches CUDA) removes that coupling.
# One process per GPU (the deployment, bench.py) does not need it: there a wait is only ever answered by another process.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
