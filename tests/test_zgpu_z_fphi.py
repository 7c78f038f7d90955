"""PME operator entry points against the oracle: the agreement the developer check tests/gpu_check.py
prints (profiles/r02a_check.log), asserted.  Green on the B200 in both builds."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["double", "mixed"])
def test_pme_mpole_fphi_vs_oracle(precision):
    """The PME operator entry point of row a7 against the oracle (not only against itself across ranks): fphi of the permanent
    multipoles = spread -> FFT -> influence function -> inverse FFT -> 20-component gather (cmpToFmp, gridMpole, pmeConv,
    fphiMpole; src/cu/pme.cu).  Measured on the B200: 1.3e-15 (double) and 2.0e-7 (mixed) at a scale of 0.36
    (profiles/r02a_check.log)."""
    import tinker_gpu_b200 as tg
    from oracle.amoeba_ref import Oracle
    s = tg.load_system(os.path.join(GOLDEN, "water30.npz"))
    from tinker_gpu_b200.amoeba import Amoeba
    a = Amoeba(s, precision)
    f = a.pme_mpole_fphi()
    a.close()
    o = Oracle(s)
    o.rotpole()
    o.dfield()
    ref = o._recip_m["fphi"]
    assert f.shape == ref.shape
    assert np.abs(f - ref).max() < (1e-12 if precision == "double" else 2e-6)
