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
    config.addinivalue_line("markers", "single_route: full-size case, far above the small-batch bound: run once")


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


@pytest.fixture(params=["auto", "tensor"])
def assign_route(request):
    """GPU parity tests run twice: with the automatic route choice (small batches are scored exactly on the CUDA cores,
    medtok_b200/csrc/small.cu) and with the tensor-core route forced for every size, so that the small shapes keep
    exercising the tcgen05 kernels' edge cases (ragged tiles, padding codes, duplicated codes ...)."""
    import torch
    if request.param == "tensor" and request.node.get_closest_marker("single_route"):
        pytest.skip("full-size case: both settings take the tensor route")
    if not torch.cuda.is_available():
        yield request.param
        return
    from medtok_b200 import _lib
    _lib.check(_lib.lib().mtvq_debug_set_soft_route(3 if request.param == "tensor" else 4), "route")
    try:
        yield request.param
    finally:
        _lib.lib().mtvq_debug_set_soft_route(4)
