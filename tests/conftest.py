import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


SCENES = {
    "cornell": "scenes/cornell_box/cornell_box_path.json",
    "cornell_dir": "scenes/cornell_box/cornell_box_dir.json",
    "caustics": "scenes/caustics.json",
    "materials": "scenes/material_test/materials.json",
}


def scene_path(name):
    return os.path.join(ROOT, SCENES[name])


def free_port():
    """A TCP port the kernel just handed out on 127.0.0.1 (rendezvous of the two-rank gloo tests): a fixed port derived from the
    pid can collide with a socket another test left in TIME_WAIT."""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.fixture(scope="session")
def root():
    return ROOT


@pytest.fixture(scope="session")
def device():
    """One lmb_ctx for the whole GPU session. Fails (does not skip) when CUDA is unusable: there is no CPU fallback."""
    from lumen_b200 import integrator
    dev = integrator.Device(0)
    yield dev
    dev.close()
