"""The valence-term oracle (oracle/valence_ref.py) and our parameter assignment (valparams.py) against the
reference's own goldens: test/ref/{bond,angle.1,strbnd,urey,opbend,torsion,pitors,tortor}.txt, the files
test/bond.cpp ... test/tortor.cpp compare the reference to (energy 1e-4, gradient 1e-4..4e-3, virial 1e-3..6e-3).
Fixtures: tests/golden/make_valence_golden.py."""
import importlib
import json
import os

import numpy as np
import pytest

from oracle import valence_ref as vr

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")
vp = importlib.import_module("tinker-gpu_b200.valparams")
GOLD = json.load(open(os.path.join(G, "valence_goldens.json")))


def load(blob):
    import tinker_gpu_b200 as tg
    s = tg.load_system(os.path.join(G, blob))
    return s.xyz, s.valence


@pytest.mark.parametrize("term", list(vr.TERMS))
def test_term_matches_reference_golden(term):
    g = GOLD[term]
    xyz, v = load(g["blob"])
    r = vr.valence(xyz, v, terms=[term])
    assert r["count"][term] == g["count"]
    # the goldens are printed with 4 (energy, gradient) and 3 (virial) decimals: half a unit of the last place
    assert abs(r["esum"] - g["energy"]) <= 5.1e-5
    ref_g = np.array(g["grad"])
    assert np.abs(r["grad"][:len(ref_g)] - ref_g).max() <= 5.1e-5
    assert np.abs(r["virial"] - np.array(g["virial"]).reshape(3, 3)).max() <= 5.1e-4


def test_gradient_is_derivative_of_energy():
    xyz, v = load("val_trpcage.npz")
    r = vr.valence(xyz, v)
    rng = np.random.default_rng(5)
    d = rng.normal(size=xyz.shape)
    d /= np.linalg.norm(d)
    h = 1.0e-5
    ep = vr.valence(xyz + h * d, v, grad=False)["esum"]
    em = vr.valence(xyz - h * d, v, grad=False)["esum"]
    assert abs((ep - em) / (2 * h) - float((r["grad"] * d).sum())) < 1.0e-5
    # translation and rotation invariance: no net force, symmetric virial
    assert np.abs(r["grad"].sum(0)).max() < 1e-9
    assert np.abs(r["virial"] - r["virial"].T).max() < 1e-9


def test_term_switches_follow_prmkey():
    io = importlib.import_module("tinker-gpu_b200.tinkerio")
    use = vp.term_switches(io.read_key(None, text="bondterm only\n"))
    assert use["BONDTERM"] and not use["ANGLETERM"] and not use["MULTIPOLETERM"]
    use = vp.term_switches(io.read_key(None, text="torsionterm none\n"))
    assert use["BONDTERM"] and not use["TORSIONTERM"]
    use = vp.term_switches(io.read_key(None, text="bondterm only\nangleterm\n"))
    assert use["BONDTERM"] and use["ANGLETERM"] and not use["UREYTERM"]


def test_dhfr2_lists():
    """Every bonded interaction of the 23 558-atom DHFR deck gets parameters (build_valence raises otherwise);
    water contributes 2 bonds, 1 angle and 1 Urey-Bradley term per molecule."""
    xyz, v = load("dhfr2.npz")
    assert len(xyz) == 23558
    nwat = 7023
    assert v.count("urey") == nwat
    assert v.count("bond") == 16569 and v.count("angle") == 11584
    assert v.count("tortor") == 147 and v.count("pitors") == 292
    assert v.opbtyp == 1 and abs(v.c("cbnd") + 2.55) < 1e-12 and abs(v.c("torsunit") - 0.5) < 1e-12
    assert (v.bk > 0).all() and (v.ak > 0).all()
