import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def goldens():
    import json
    with open(os.path.join(GOLDEN, "reference_goldens.json")) as fh:
        return json.load(fh)


def load_case(name):
    import tinker_gpu_b200 as tg
    return tg.load_system(os.path.join(GOLDEN, "lf_" + name.lower().replace("-", "_") + ".npz"))


def section(goldens, case, frag):
    for k, v in goldens[case]["sections"].items():
        if frag in k:
            return v
    raise KeyError((case, frag))
