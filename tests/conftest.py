import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    import json
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    data = np.load(path)
    arrays = {k: data[k] for k in data.files if k != "meta"}
    meta = json.loads(bytes(data["meta"]).decode())
    return meta, arrays


def golden_names(prefix):
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith(prefix) and f.endswith(".npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import vq_oracle
    return vq_oracle
