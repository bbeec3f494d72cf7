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


# ------------------------------------------------------------------ observed parity errors (reported next to their bounds)
OBSERVED = {}          # test id -> {"worst_ratio": observed / bound, "observed": max abs error, "bound": its bound, "checks": n}
_CURRENT = {"id": None}


@pytest.fixture(autouse=True)
def _track_current_test(request):
    _CURRENT["id"] = request.node.nodeid
    yield
    _CURRENT["id"] = None


def record_observed(observed: float, bound: float) -> None:
    """Called by the parity helpers: keeps, per test, the check that came closest to its bound."""
    key = _CURRENT["id"] or "?"
    ratio = observed / bound if bound > 0 else (0.0 if observed == 0 else float("inf"))
    cur = OBSERVED.get(key)
    if cur is None:
        OBSERVED[key] = {"worst_ratio": ratio, "observed": observed, "bound": bound, "checks": 1}
    else:
        cur["checks"] += 1
        if ratio > cur["worst_ratio"]:
            cur.update(worst_ratio=ratio, observed=observed, bound=bound)


def pytest_sessionfinish(session, exitstatus):
    """Writes gpurun_out/parity_observed.json: for every parity test the largest observed error as a fraction of its
    tolerance (how much slack each bound has).  Only on GPU runs (the CPU suite records nothing)."""
    if not OBSERVED:
        return
    import json
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "parity_observed.json"), "w") as fh:
        json.dump(OBSERVED, fh, indent=1, sort_keys=True)


def pytest_terminal_summary(terminalreporter):
    if not OBSERVED:
        return
    terminalreporter.write_sep("-", "observed parity error / bound (worst check per test)")
    for key, v in sorted(OBSERVED.items(), key=lambda kv: -kv[1]["worst_ratio"])[:25]:
        terminalreporter.write_line(f"{v['worst_ratio']:8.3f}  obs {v['observed']:.3e}  bound {v['bound']:.3e}  {key}")


def golden_module():
    """tests/golden/make_golden.py as a module (fake datasets and model builders shared with the fixture generator;
    importing it does not touch /root/reference)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
