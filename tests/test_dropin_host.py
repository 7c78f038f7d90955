"""Host-side checks of the drop-in demonstration (oracle/_ref/libref_dropin.so): the reference's unmodified front-ends
src/amoeba/field.cpp, induce.cpp, emplar.cpp, mpole.cpp and its energy-buffer reductions src/energybuffer.cpp link against integration/apx_adapter.cpp + libapx with no reference CUDA kernel in the
library; every `*_cu` symbol those front-ends call is defined by the adapter; without a GPU the open call reports an error."""
import ctypes as C
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_dropin.so")
pytestmark = pytest.mark.skipif(not os.path.isfile(LIB), reason="oracle/_ref/libref_dropin.so not built (make -C oracle dropin needs /root/reference)")

# the operator boundary of the induce()/dfield()/ufield() front-ends + the energy operators of INTEGRATION.md
CU_SYMBOLS = ["induceMutualPcg1_cu", "sparsePrecondApply_cu", "diagPrecond_cu", "ulspredSaveP1_cu", "ulspredSum_cu", "dfieldEwaldReal_cu",
              "dfieldEwaldRecipSelfP2_cu", "dfieldNonEwald_cu", "ufieldEwaldReal_cu", "ufieldEwaldRecipSelfP1_cu", "ufieldNonEwald_cu",
              "emplar_cu", "empoleEwaldRealSelf_cu", "empoleNonEwald_cu", "empoleChgpenEwaldRecip_cu", "epolarEwaldReal_cu",
              "epolarNonEwald_cu", "epolarEwaldRecipSelf_cu", "epolar0DotProd_cu", "torque_cu", "chkpole_cu", "rotpole_cu", "rpoleToCmp_cu",
              "bsplineFill_cu", "gridMpole_cu", "gridUind_cu", "pmeConv_cu", "fphiMpole_cu", "fphiUind_cu", "fphiUind2_cu", "cmpToFmp_cu",
              "cuindToFuind_cu", "fphiToCphi_cu", "mpoleDataBinding_cu", "epolarDataBinding_cu", "epolarPairwiseExtfield_cu"]
FRONT_ENDS = ["tinker::induce(", "tinker::dfield(", "tinker::ufield(", "tinker::sparsePrecondApply(", "tinker::diagPrecond(",
              "tinker::emplar(", "tinker::mpoleInit(", "tinker::torque(", "tinker::energyReduce(", "tinker::virialReduce("]


def _nm():
    return subprocess.run(["nm", "-DC", "--defined-only", LIB], capture_output=True, text=True, check=True).stdout


def test_library_resolves_with_the_adapter_instead_of_the_reference_kernels():
    C.CDLL(LIB, mode=os.RTLD_NOW)
    out = _nm()
    for s in CU_SYMBOLS:
        assert f"tinker::{s}(" in out, s
    for s in FRONT_ENDS:
        assert s in out, s
    # no kernel of the reference is in this library: its pair / PME kernels carry these names
    for k in ("pcgUdirV2", "dfield_cu1", "ufield_cu1", "emplar_cu1a", "gridPut_cu", "sparsePrecond_cu1", "torque_cu1", "rotpoleNorm"):
        assert k not in subprocess.run(["nm", "-C", LIB], capture_output=True, text=True).stdout, k
    ldd = subprocess.run(["ldd", LIB], capture_output=True, text=True).stdout
    assert "libapx.so" in ldd and "not found" not in ldd.split("libapx.so")[1].splitlines()[0]


def test_without_a_gpu_open_reports_an_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, "-m", "oracle.ref_dropin_bridge", os.path.join(ROOT, "tests", "golden", "water30.npz")], cwd=ROOT,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CUDA device" in r.stderr
