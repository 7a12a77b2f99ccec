import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a fresh checkout has no built artefacts (they are git-ignored): build the product library (nvcc cross
    # compiles without a GPU) and the test oracle once
    from channelcoding_b200 import build as _build
    if not os.path.exists(_build.LIB):
        _build.build()
    import oracle
    oracle.build()


@pytest.fixture(scope="session")
def catalogue():
    with open(os.path.join(GOLDEN, "catalogue.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_codes():
    return np.load(os.path.join(GOLDEN, "codes.npz"))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def golden_H(golden_codes, catalogue, name, alt=False):
    e = catalogue[name]
    packed = golden_codes[name + (".H_alt" if alt else ".H")]
    return np.unpackbits(packed, axis=1)[:, :e["n"]]


# variant id of oracle/ref_shim.cc -> (name, alpha, beta, max_iter); see oracle/ccref.py
VARIANT_PARAMS = {
    0: ("MS", 1.0, 0.0, 50), 1: ("NMS", 0.8, 0.0, 50), 2: ("OMS", 1.0, 0.01, 50),
    3: ("SCMS1", 1.0, 0.0, 50), 4: ("SCMS2", 1.0, 0.0, 50), 5: ("2DNMS", 1.0, 1.0, 50),
    6: ("NMS", 0.915, 0.0, 50), 7: ("OMS", 1.0, 0.032, 50), 8: ("2DNMS", 0.968, 907.0 / 125.0, 50),
    9: ("MS", 1.0, 0.0, 1), 10: ("MS", 1.0, 0.0, 5), 11: ("NMS", 0.8, 0.0, 7),
}
