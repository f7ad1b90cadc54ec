import json
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


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "nw_golden_blosum62.json")) as f:
        g = json.load(f)
    lm = {c: i for i, c in enumerate(g["letters"])}
    g["enc"] = {sid: np.array([lm[c] for c in s], dtype=np.uint8) for sid, s in g["seqs"].items()}
    return g


@pytest.fixture(scope="session")
def scoring():
    with open(os.path.join(GOLDEN_DIR, "scoring.json")) as f:
        s = json.load(f)
    s["subst"] = {k: np.array(v, dtype=np.int32) for k, v in s["subst"].items()}
    return s


def case_letters(g, case):
    y = g["enc"][case["y"]]
    x = g["enc"][case["x"]]
    yl, yr = case["y_range"]
    xl, xr = case["x_range"]
    y = y[(yl or 0):(yr if yr is not None else y.size)]
    x = x[(xl or 0):(xr if xr is not None else x.size)]
    return y, x


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (builds oracle/libnworacle.so on first use)."""
    import subprocess
    from oracle import pyoracle
    if not os.path.exists(pyoracle.ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    pyoracle.lib()
    return pyoracle
