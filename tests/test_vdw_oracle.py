"""The float64 vdW oracle (oracle/vdw_ref.py) against the reference's own golden vectors for the buffered
14-7 term: NaCl-1 (test/nacl.cpp:36-176) and Local-Frame2-1/2 (test/localframe2.cpp:46-98), with the
reference's tolerances; plus parameter-assignment checks on the dhfr2 deck (kvdw.f rules)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def vgold():
    with open(os.path.join(GOLDEN, "vdw_goldens.json")) as fh:
        return json.load(fh)


def _load(tag):
    import tinker_gpu_b200 as tg
    return tg.load_system(os.path.join(GOLDEN, "vdw_" + tag + ".npz"))


@pytest.mark.parametrize("tag", ["nacl_no_switch", "nacl_near_cut", "nacl_near_off", "nacl_evcorr_vlambda__10"])
def test_nacl_pair(vgold, tag):
    from oracle.vdw_ref import VdwOracle
    g = vgold[tag]
    r = VdwOracle(_load(tag)).ehal()
    eps = 1.0e-3                                            # test/nacl.cpp:45
    assert abs(r["ev"] - g["ref_eng"]) < eps
    assert r["nev"] == int(g["ref_count"])
    assert np.abs(r["grad"] - np.array(g["ref_grad"])).max() < eps
    assert np.abs(r["virial"] - np.array(g["ref_v"])).max() < eps


@pytest.mark.parametrize("tag", ["local_frame2_1", "local_frame2_2"])
def test_local_frame2_cells(vgold, tag):
    """18 atoms in a triclinic / monoclinic cell, 7 A cutoff: hydrogen reduction, 1-2/1-3 exclusions, images."""
    from oracle.vdw_ref import VdwOracle
    g = vgold[tag]
    o = VdwOracle(_load(tag))
    r = o.ehal()
    assert abs(r["ev"] - g["ref_eng"]) < 1.0e-4             # test/localframe2.cpp:60
    assert r["nev"] == int(g["ref_count"])
    # gradient = central difference of the oracle's own energy
    x0 = o.xyz.copy()
    rng = np.random.default_rng(2)
    for i, c in zip(rng.integers(0, o.n, 6), rng.integers(0, 3, 6)):
        h = 1e-5
        xp, xm = x0.copy(), x0.copy()
        xp[i, c] += h
        xm[i, c] -= h
        o.set_xyz(xp)
        ep = o.ehal(False)["ev"]
        o.set_xyz(xm)
        em = o.ehal(False)["ev"]
        assert abs((ep - em) / (2 * h) - r["grad"][i, c]) < 2e-5 * max(1.0, abs(r["grad"][i, c]))
    o.set_xyz(x0)
    assert np.abs(r["grad"].sum(0)).max() < 1e-9


def test_dhfr2_vdw_parameters():
    """amoebabio09 rules on the dhfr2 deck: CUBIC-MEAN radii of halved diameters, HHG well depths,
    hydrogens reduced along their single bond, 12 A cutoff tapered from 10.8 A."""
    import tinker_gpu_b200 as tg
    s = tg.load_system(os.path.join(GOLDEN, "dhfr2.npz"))
    v = s.vdw
    assert v is not None and v.cutoff == 12.0 and abs(v.taper - 10.8) < 1e-12
    assert v.radmin.shape == v.epsilon.shape and np.allclose(v.radmin, v.radmin.T) and np.allclose(v.epsilon, v.epsilon.T)
    d = np.diag(v.radmin)
    a, b = 3, 11
    ra, rb = d[a] / 2, d[b] / 2
    assert abs(v.radmin[a, b] - 2 * (ra ** 3 + rb ** 3) / (ra ** 2 + rb ** 2)) < 1e-12
    ea, eb = v.epsilon[a, a], v.epsilon[b, b]
    assert abs(v.epsilon[a, b] - 4 * ea * eb / (np.sqrt(ea) + np.sqrt(eb)) ** 2) < 1e-12
    red = v.kred != 0
    assert red.sum() > 10000 and np.all(v.ired[red] != np.arange(s.n)[red]) and np.all(v.ired[~red] == np.arange(s.n)[~red])
    assert np.all(v.vexclude[:, 0] < v.vexclude[:, 1]) and np.all(v.vexclude_scale == 0.0)
    r = tg.replicate(s, (2, 1, 1), keep_bonds=False)
    assert r.vdw.ired[s.n:].min() >= s.n and r.vdw.vexclude.shape[0] == 2 * v.vexclude.shape[0]
