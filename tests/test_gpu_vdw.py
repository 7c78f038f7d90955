"""Buffered 14-7 vdW term on the GPU (ehal.cu, through the C ABI) against the reference's golden vectors
(test/nacl.cpp, test/localframe2.cpp) and against the float64 oracle (oracle/vdw_ref.py) on the water box
and on dhfr2; additivity inside energy(); decomposed ranks."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _rms(a):
    return float(np.sqrt((np.asarray(a) ** 2).mean()))


def _vload(tag):
    import tinker_gpu_b200 as tg
    return tg.load_system(os.path.join(GOLDEN, "vdw_" + tag + ".npz"))


@pytest.fixture(scope="module")
def vgold():
    with open(os.path.join(GOLDEN, "vdw_goldens.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("precision", ["double", "mixed"])
@pytest.mark.parametrize("tag", ["nacl_no_switch", "nacl_near_cut", "nacl_near_off", "nacl_evcorr_vlambda__10"])
def test_nacl_goldens(vgold, tag, precision):
    """COMPARE_CODE_BLOCK1 of test/nacl.cpp:9-34: every calc version against the same literals, eps 1e-3."""
    from tinker_gpu_b200.amoeba import Amoeba, calc
    g = vgold[tag]
    a = Amoeba(_vload(tag), precision, vdw=True)
    eps = 1.0e-3
    for vers in (calc.v0, calc.v1, calc.v3, calc.v4, calc.v5, calc.v6):
        r = a.evdw(vers)
        if vers & calc.energy:
            assert abs(r["ev"] - g["ref_eng"]) < eps
        if vers & calc.grad:
            assert np.abs(r["grad"] - np.array(g["ref_grad"])).max() < eps
        if vers & calc.virial:
            assert np.abs(r["virial"] - np.array(g["ref_v"])).max() < eps
        if vers & calc.analyz:
            assert r["nev"] == int(g["ref_count"])
    a.close()


@pytest.mark.parametrize("precision", ["double", "mixed"])
@pytest.mark.parametrize("tag", ["local_frame2_1", "local_frame2_2"])
def test_local_frame2_goldens(vgold, tag, precision):
    """Triclinic and monoclinic cells (test/localframe2.cpp:46-98) + gradient/virial against the oracle."""
    from tinker_gpu_b200.amoeba import Amoeba, calc
    from oracle.vdw_ref import VdwOracle
    g = vgold[tag]
    s = _vload(tag)
    a = Amoeba(s, precision, vdw=True)
    r3 = a.evdw(calc.v3)
    # double build: the reference's tolerance.  Mixed build: float coordinates of ~25 A carry 2e-6 A, which the
    # 7th power of r/radmin turns into ~5e-6 of each close-contact energy -- 1e-3 of these 200 kcal/mol
    assert abs(r3["ev"] - g["ref_eng"]) < (1.0e-4 if precision == "double" else 1.0e-3)
    assert r3["nev"] == int(g["ref_count"])
    r = a.evdw(calc.v1)
    o = VdwOracle(s).ehal()
    tol = 1e-9 if precision == "double" else 2e-5
    assert abs(r["ev"] - o["ev"]) < tol * max(1.0, abs(o["ev"]))
    assert np.abs(r["grad"] - o["grad"]).max() < tol * max(1.0, np.abs(o["grad"]).max())
    assert np.abs(r["virial"] - o["virial"]).max() < tol * max(1.0, np.abs(o["virial"]).max())
    a.close()


@pytest.mark.parametrize("name,precision", [("water30", "double"), ("water30", "mixed"), ("dhfr2", "mixed"), ("dhfr2", "double")])
def test_boxes_vs_oracle(name, precision):
    """2684-atom water box and dhfr2 (8.3 M pairs inside 12 A): energy 1e-6 relative, forces, virial, count; then the
    atoms move (no rebuild, then a rebuild) and the comparison is repeated."""
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba, calc
    from oracle.vdw_ref import VdwOracle
    s = tg.load_system(os.path.join(GOLDEN, name + ".npz"))
    a = Amoeba(s, precision, vdw=True)
    o = VdwOracle(s)
    x0 = np.array(s.xyz)
    rng = np.random.default_rng(4)
    frames = [x0] if name == "dhfr2" and precision == "double" else [x0, x0 + rng.normal(scale=0.03, size=x0.shape),
                                                                         x0 + np.array([1.7, -0.4, 2.9])]
    te, tg_, tv = (1e-10, 1e-8, 1e-9) if precision == "double" else (1e-6, 5e-5, 2e-5)
    rebuilds = []
    for x in frames:
        a.set_positions(x)
        o.set_xyz(x)
        r = a.evdw(calc.v1)
        q = o.ehal()
        assert abs(r["ev"] - q["ev"]) < te * abs(q["ev"])
        assert _rms(r["grad"] - q["grad"]) < tg_
        assert np.abs(r["virial"] - q["virial"]).max() < tv * np.abs(q["virial"]).max()
        # pairs within float round-off of the cutoff (where the taper has brought the energy to zero) may fall either side
        assert abs(a.evdw(calc.v3)["nev"] - q["nev"]) <= (0 if precision == "double" else 4)
        rebuilds.append(a.stats()["list_rebuilds"])
    if len(frames) == 3:
        assert rebuilds[1] == rebuilds[0] and rebuilds[2] == rebuilds[1] + 1
    a.close()


def test_energy_includes_vdw_and_overlaps_the_solver():
    """energy() with the term attached = electrostatics + vdW (esum, gradient, virial), same PCG iterations."""
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba, calc
    s = tg.load_system(os.path.join(GOLDEN, "dhfr2.npz"))
    a0 = Amoeba(s, "mixed")
    r0 = a0.energy(calc.v1)
    a0.close()
    a = Amoeba(s, "mixed", vdw=True)
    rv = a.evdw(calc.v1)
    r = a.energy(calc.v1)
    assert r["ev"] == rv["ev"]                                   # fixed-point sums: bit-identical
    assert abs(r["esum"] - (r0["esum"] + rv["ev"])) < 1e-7 * abs(r0["esum"])
    assert _rms(r["grad"] - (r0["grad"] + rv["grad"])) < 2e-5
    assert np.abs(r["virial"] - (r0["virial"] + rv["virial"])).max() < 1e-6 * np.abs(rv["virial"]).max()
    assert r["pcg_iterations"] == r0["pcg_iterations"]
    st = a.stats()
    assert st["ms_ehal"] > 0 and st["nverlet_vdw"] > 0
    a.close()


def test_vdw_on_ranks():
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba, calc
    from tinker_gpu_b200.distributed import run_local_ranks
    s = tg.load_system(os.path.join(GOLDEN, "water30.npz"))
    a = Amoeba(s, "mixed", vdw=True)
    ref = a.energy(calc.v1)
    refv = a.evdw(calc.v3)
    a.close()

    def job(am, rank):
        return am.energy(calc.v1), am.evdw(calc.v3)

    for r, rv in run_local_ranks(s, 2, job, "mixed", vdw=True):
        assert abs(r["ev"] - ref["ev"]) < 1e-7 * abs(ref["ev"]) and rv["nev"] == refv["nev"]     # rows are walked in another order
        assert abs(r["esum"] - ref["esum"]) < 3e-7 * abs(ref["esum"])
        assert _rms(r["grad"] - ref["grad"]) < 3e-5
        assert np.abs(r["virial"] - ref["virial"]).max() < 2e-3 * np.abs(ref["virial"]).max()
