"""Pin the CPU oracle (oracle/amoeba_ref.py) to the reference's own golden vectors
(SURVEY.md §4 / §8c): test/localframe.cpp, test/localframe3.cpp, test/ref/tinkernist.*.
Tolerances are the ones the reference's tests use (its literals carry 4 decimals)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_case, section
from oracle.amoeba_ref import Oracle, V1, DEBYE


def _arr(x):
    return np.array(x)


@pytest.mark.parametrize("case", ["Local-Frame-1", "Local-Frame-2"])
def test_empole(goldens, case):
    ref = section(goldens, case, "empole")
    r = Oracle(load_case(case)).energy(V1)
    if "ref_eng" in ref:
        assert abs(r["em"] - ref["ref_eng"]) < 1e-4
    else:   # PME case pins the real / recip / self split (test/localframe.cpp:124-130)
        assert abs(r["em_real"] - ref["ref_ereal"]) < 1e-4
        assert abs(r["em_recip"] - ref["ref_erecip"]) < 1e-4
        assert abs(r["em_self"] - ref["ref_eself"]) < 1e-4
    g = _arr(ref["ref_grad"])
    assert np.abs(r["grad"] - g)[:18].max() < 5e-4          # do_ij: last NH3 excluded
    assert np.abs(r["grad"] - g)[18:, :2].max() < 5e-4
    assert np.abs(r["virial"] - _arr(ref["ref_v"])).max() < ref["eps_v"]
    nself = 22 if "ref_ereal" in ref else 0
    assert r["nem"] + nself == int(ref["ref_count"])


@pytest.mark.parametrize("case,tag", [("Local-Frame-3", "non-ewald"), ("Local-Frame-4", "pme")])
def test_fields_induce_epolar(goldens, case, tag):
    s = load_case(case)
    o = Oracle(s)
    o.rotpole()
    fd, fp = o.dfield()
    ref = section(goldens, case, "dfield")
    assert np.abs(fd - _arr(ref["ref_dir_field_d"])).max() < 1e-4
    assert np.abs(fp - _arr(ref["ref_dir_field_p"])).max() < 1e-4
    n = s.n
    ud = _arr([[0.1 * (i + 1) + 0.03 * (j + 1) for j in range(3)] for i in range(n)])
    up = _arr([[0.1 * (i + 1) - 0.03 * (j + 1) for j in range(3)] for i in range(n)])
    f1, f2 = o.ufield(ud, up)
    ref = section(goldens, case, "ufield")
    assert np.abs(f1 - _arr(ref["ref_ufield_d"])).max() < 1e-4
    assert np.abs(f2 - _arr(ref["ref_ufield_p"])).max() < 1e-4
    u1, u2 = o.induce()
    ref = section(goldens, case, "induce")
    assert np.abs(u1 * DEBYE - _arr(ref["ref_ud_debye"])).max() < 1e-4
    assert np.abs(u2 * DEBYE - _arr(ref["ref_up_debye"])).max() < 1e-4
    r = o.energy(V1)
    ref = section(goldens, case, "various")
    assert abs(r["ep_dot"] - ref["ref_eng"]) < 1e-4
    assert abs(r["ep_pair"] - ref["ref_eng"]) < 1e-4
    g = _arr(ref["ref_grad"])
    assert np.abs(r["grad"] - g)[:18].max() < 1e-4
    assert np.abs(r["virial"] - _arr(ref["ref_v"])).max() < ref["eps_v"]


@pytest.mark.parametrize("case", ["Local-Frame3-1", "Local-Frame3-2"])
def test_fused_total(goldens, case):
    ref = section(goldens, case, "emplar")
    r = Oracle(load_case(case)).energy(V1)
    assert abs(r["esum"] - ref["ref_eng"]) < 1e-4
    assert np.abs(r["grad"] - _arr(ref["ref_g"])).max() < 1e-4
    assert np.abs(r["virial"] - _arr(ref["ref_v"])).max() < 1e-3


def test_numerical_gradient():
    """Analytic gradient == central difference of the oracle's own energy (testgrad-style)."""
    s = load_case("Local-Frame3-2")
    s.poleps = 1e-12
    o = Oracle(s)
    r = o.energy(V1)
    x0 = o.xyz.copy()
    h = 1e-5
    for (i, c) in [(0, 0), (5, 1), (9, 2), (14, 0)]:
        e = []
        for sg in (+1, -1):
            x = x0.copy()
            x[i, c] += sg * h
            o2 = Oracle(s)
            o2.set_xyz(x)
            e.append(o2.energy(V1)["esum"])
        num = (e[0] - e[1]) / (2 * h)
        assert abs(num - r["grad"][i, c]) < 2e-5


def test_tinkernist_frames():
    """2684-atom water box, two MD frames of the reference: every atom's direct and induced dipole."""
    import tinker_gpu_b200 as tg
    s = tg.load_system(os.path.join(GOLDEN, "water30.npz"))
    fr = np.load(os.path.join(GOLDEN, "tinkernist_frames.npz"))
    o = Oracle(s)
    o.set_xyz(fr["arc"][0])
    u1, _ = o.induce()
    assert np.abs(o.udir * DEBYE - fr["udir"][0]).max() < 1e-5
    assert np.abs(u1 * DEBYE - fr["uind"][0]).max() < 1e-5     # reference test tolerance is 1e-3


def test_oracle_triclinic_gradient_is_consistent():
    """No electrostatics golden of the reference lives in a non-orthogonal cell, so the oracle's triclinic branch is held
    to its own energy: analytic gradient = central difference (fixture made from test/file/local_frame/local_frame2.xyz
    with the cell of test/localframe2.cpp)."""
    import os
    import numpy as np
    import tinker_gpu_b200 as tg
    from conftest import GOLDEN
    from oracle.amoeba_ref import Oracle, V0, V1
    s = tg.load_system(os.path.join(GOLDEN, "lf_triclinic.npz"))
    assert abs(s.lvec[0][1]) > 1 and abs(s.lvec[1][2]) > 1          # really triclinic
    o = Oracle(s)
    r = o.energy(V1)
    x0 = o.xyz.copy()
    h = 1e-5
    for i, c in ((0, 0), (3, 1), (10, 2), (17, 0)):
        xp, xm = x0.copy(), x0.copy()
        xp[i, c] += h
        xm[i, c] -= h
        o.set_xyz(xp)
        ep = o.energy(V0)["esum"]
        o.set_xyz(xm)
        em = o.energy(V0)["esum"]
        assert abs((ep - em) / (2 * h) - r["grad"][i, c]) < 2e-6 * max(1.0, abs(r["grad"][i, c]))
