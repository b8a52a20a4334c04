import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Built on demand."""
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def mg():
    """One libmodsgpu context with the three networks loaded (GPU tests only)."""
    import mods_light_zmq_b200 as M
    ctx = M.ModsGpu(0, load_nets=True)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def synth_pair():
    from mods_light_zmq_b200 import synth
    return synth.image_pair()
