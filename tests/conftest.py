import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden", "ref")
SCENES = os.path.join(ROOT, "scenes", "xml")
REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def have_meshes():
    return os.path.exists(os.path.join(ROOT, "scenes", "meshes", "cbox", "light.obj"))


def have_vader():
    return os.path.exists(os.path.join(ROOT, "scenes", "meshes", "vader", "vader.obj"))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def scenes_dir():
    return SCENES


def load_frame(name, shape=(256, 256, 4)):
    import numpy as np
    return np.fromfile(os.path.join(GOLDEN, name), dtype=np.float32).reshape(shape)


def rel_l2(a, b):
    import numpy as np
    a = a[..., :3].astype(np.float64)
    b = b[..., :3].astype(np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))
