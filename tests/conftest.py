import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The in-tree CUDA library must exist for every test session (nvcc cross-compiles without a GPU)."""
    from magpy_rv_b200 import build
    if build.needs_build():
        build.build()
    return build.OUT


def load_cases():
    data = np.load(os.path.join(GOLD, "logl_cases.npz"))
    manifest = json.loads(str(data["manifest"]))
    cases = []
    for m in manifest:
        pre = m["name"] + "__"
        c = dict(m)
        for key in ("t", "y", "yerr", "hp", "modpar", "logL", "model_y"):
            c[key] = data[pre + key]
        c["flags"] = data[pre + "flags"] if m["has_flags"] else None
        cases.append(c)
    return cases


def prior_tuples(priors):
    """[(name, prior, [vals])] -> reference-style prior_list via the package's pri_create."""
    from magpy_rv_b200 import parameters as par
    return [par.pri_create(a, b, v) for a, b, v in priors]
