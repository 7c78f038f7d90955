"""Host-side checks of the reference-CUDA comparator (oracle/ref_cuda.cu -> oracle/_ref/libref_cuda.so): the library resolves
every symbol it needs (the reference's CUDA translation units + our shim), exports the comparator C ABI, the ctypes mirror of
its system struct has the C layout, and without a GPU it fails with an error message instead of crashing."""
import ctypes as C
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_cuda.so")
pytestmark = pytest.mark.skipif(not os.path.isfile(LIB), reason="oracle/_ref/libref_cuda.so not built (make -C oracle cuda needs /root/reference)")


def test_library_resolves_and_exports_the_comparator_abi():
    lib = C.CDLL(LIB, mode=os.RTLD_NOW)      # RTLD_NOW: every undefined symbol of the reference TUs must be satisfied
    for name in ("refcu_open", "refcu_set_xyz", "refcu_induce", "refcu_energy", "refcu_time", "refcu_last_error"):
        assert hasattr(lib, name), name


def test_ctypes_mirror_has_the_c_layout():
    from oracle.ref_cuda_bridge import _RefcuSystem
    lib = C.CDLL(LIB)
    assert lib.refcu_sizeof_system() == C.sizeof(_RefcuSystem)
    assert lib.refcu_offsetof_dielec() == _RefcuSystem.dielec.offset


def test_without_a_gpu_the_comparator_reports_an_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, "-m", "oracle.ref_cuda_bridge", os.path.join(ROOT, "tests", "golden", "water30.npz")], cwd=ROOT,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CUDA device" in r.stderr
    sys.path.insert(0, ROOT)
    import bench
    out = bench.ref_cuda_sample(ours_induce_ms=1.0)
    assert "unavailable" in out
