import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    return sorted(f[:-5] for f in os.listdir(GOLDEN) if f.endswith(".json"))


def load_golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        g = json.load(f)
    for c in g["cases"]:
        c["prime"] = int(c["prime"])
        c["points"] = np.array([[int(x) for x in row] for row in c["points"]], dtype=np.uint64)
        c["values"] = np.array([[int(x) for x in row] for row in c["values"]], dtype=np.uint64)
    return g


@pytest.fixture(scope="session")
def oracle():
    from oracle_bind import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from refs import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return Ref()


@pytest.fixture(scope="session")
def ctx():
    import firefly_b200 as ff
    return ff.Context()
