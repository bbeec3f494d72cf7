import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _reproducible_rng(request):
    """Every test starts from global generators seeded by its own name: module initialisations and random draws that do not
    pass an explicit generator are the same on every run (a failure is then a property of the code, not of the draw)."""
    import random
    import zlib
    seed = zlib.crc32(request.node.name.encode()) % (2 ** 31)
    torch.manual_seed(seed)
    np.random.seed(seed % (2 ** 32))
    random.seed(seed)
    yield


def load_golden(name):
    """Committed fixture produced by tests/golden/make_golden.py from the real reference."""
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def golden_names(prefix):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith(".npz"))


def t(a, dtype=None):
    a = np.asarray(a)
    x = torch.from_numpy(np.ascontiguousarray(a)).reshape(a.shape)     # ascontiguousarray promotes 0-d to 1-d
    return x.to(dtype) if dtype is not None else x
