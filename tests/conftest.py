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


def _cuda_devices():
    """Number of usable CUDA devices, without importing torch (libcudart through ctypes)."""
    import ctypes
    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            rt = ctypes.CDLL(name)
        except OSError:
            continue
        n = ctypes.c_int(0)
        return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """Tests marked gpu are skipped (not failed) on a machine without a CUDA device; on the GPU box nothing is skipped."""
    if not any("gpu" in it.keywords for it in items) or _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device on this machine (run with -m gpu on the B200 box)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


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
