"""One-pass neighbour-row rebuild (rows.cu, APX_ROWS_ONEPASS): the padded-slot search + packing copy, and its overflow
fall-back, must produce exactly the rows of the count + fill path."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["1", "2"])
@pytest.mark.parametrize("blob,precision", [("water30.npz", "double"), ("dhfr2.npz", "mixed")])
def test_onepass_rebuild_equals_two_pass(monkeypatch, mode, blob, precision):
    """mode 1: slots with slack (packing path).  mode 2: zero slack, so any row that grew overflows (fall-back path).
    Build 1 is always count + fill; a rigid shift forces build 2 (same pair set: identical row totals); a jittered
    configuration forces build 3, compared with a fresh two-pass context on the same coordinates."""
    import copy
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba, calc
    s = tg.load_system(os.path.join(GOLDEN, blob))
    tol = 1e-10 if precision == "double" else 2e-6
    monkeypatch.setenv("APX_ROWS_ONEPASS", mode)
    a = Amoeba(s, precision, vdw=True)
    r0 = a.energy(calc.v4)
    st0 = a.stats()
    x1 = s.xyz + np.array([1.7, -1.3, 2.1])
    a.set_positions(x1)
    r1 = a.energy(calc.v4)
    st1 = a.stats()
    assert st1["list_rebuilds"] == st0["list_rebuilds"] + 1
    # a rigid shift keeps every distance; in float coordinates a handful of pairs within 1e-6 A of a cutoff may change sides
    slop = 0 if precision == "double" else 200
    for k in ("nverlet", "nverlet_vdw", "npairs_m"):
        assert abs(st1[k] - st0[k]) <= slop, k
    # (the energies themselves differ at the 1e-5 level: PME on a grid is not translation invariant)
    assert abs(r1["esum"] - r0["esum"]) <= 1e-4 * abs(r0["esum"])
    rng = np.random.default_rng(9)
    x2 = x1 + np.array([-1.4, 1.6, 1.2]) + rng.normal(scale=0.03, size=x1.shape)
    a.set_positions(x2)
    r2 = a.energy(calc.v4)
    st2 = a.stats()
    assert st2["list_rebuilds"] == st1["list_rebuilds"] + 1
    a.close()
    monkeypatch.delenv("APX_ROWS_ONEPASS")
    s2 = copy.copy(s)
    s2.xyz = x2
    b = Amoeba(s2, precision, vdw=True)
    rb = b.energy(calc.v4)
    stb = b.stats()
    assert st2["nverlet"] == stb["nverlet"] and st2["nverlet_vdw"] == stb["nverlet_vdw"] and st2["npairs_m"] == stb["npairs_m"]
    assert abs(r2["esum"] - rb["esum"]) <= tol * abs(rb["esum"])
    assert np.abs(r2["grad"] - rb["grad"]).max() <= 1e2 * tol
    b.close()
