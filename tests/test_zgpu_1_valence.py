"""Valence terms on the GPU (evalence.cu, through the C ABI) against the reference's goldens
test/ref/{bond,angle.1,strbnd,urey,opbend,torsion,pitors,tortor}.txt and against the autograd oracle on dhfr2;
additivity inside energy()."""
import copy
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
TERMS = ("bond", "angle", "strbnd", "urey", "opbend", "torsion", "pitors", "tortor")


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLDEN, "valence_goldens.json")) as fh:
        return json.load(fh)


def _only(v, term):
    v = copy.copy(v)
    v.use = np.array([int(t == term) for t in TERMS], np.int32)
    return v


@pytest.mark.parametrize("precision", ["mixed", "double"])
@pytest.mark.parametrize("term", TERMS)
def test_term_goldens(gold, term, precision):
    """test/bond.cpp ... test/tortor.cpp: `xxxterm only`, every calc version against the printed transcript
    (energy 1e-4, gradient 1e-4, virial 1e-3: the reference's double-precision tolerances -- the valence kernel
    runs in double in both builds)."""
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba, calc
    g = gold[term]
    s = tg.load_system(os.path.join(GOLDEN, g["blob"]))
    a = Amoeba(s, precision)
    a.attach_valence(_only(s.valence, term))
    k = TERMS.index(term)
    ref_g = np.array(g["grad"])
    ref_v = np.array(g["virial"]).reshape(3, 3)
    for vers in (calc.v0, calc.v1, calc.v3, calc.v4, calc.v5, calc.v6):
        r = a.evalence(vers)
        assert r.count[k] == g["count"] and sum(r.count) == g["count"]
        if vers & calc.energy:
            assert abs(r.esum - g["energy"]) < 1.0e-4 and abs(r.e[k] - r.esum) < 1e-12
        if vers & calc.grad:
            assert np.abs(a.valence_gradient()[:len(ref_g)] - ref_g).max() < 1.0e-4
        if vers & calc.virial:
            assert np.abs(np.array(list(r.virial)).reshape(3, 3) - ref_v).max() < 1.0e-3
    a.close()


@pytest.mark.parametrize("precision", ["mixed", "double"])
def test_dhfr2_vs_oracle(precision):
    """All 48 k bonded interactions of the DHFR deck: energies to 1e-9 relative per term, forces to 1e-7 (fixed-point
    quantum 2^-32 times at most ~20 contributions per atom), virial to 1e-6."""
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba, calc
    from oracle import valence_ref as vr
    s = tg.load_system(os.path.join(GOLDEN, "dhfr2.npz"))
    o = vr.valence(s.xyz, s.valence)
    a = Amoeba(s, precision)
    a.attach_valence(s.valence)
    r = a.evalence(calc.v1)
    for k, t in enumerate(TERMS):
        assert abs(r.e[k] - o["energy"][t]) <= 1e-9 * max(1.0, abs(o["energy"][t])) + 5e-6, t
        assert r.count[k] == o["count"][t]
    g = a.valence_gradient()
    assert np.abs(g - o["grad"]).max() < 1e-7
    assert np.abs(np.array(list(r.virial)).reshape(3, 3) - o["virial"]).max() < 1e-5
    # a second evaluation starts from cleared accumulators
    r2 = a.evalence(calc.v1)
    assert abs(r2.esum - r.esum) < 1e-9 and np.abs(a.valence_gradient() - g).max() < 1e-12
    a.close()


def test_dhfr2_vs_reference_arithmetic():
    """The same comparison against the reference's OWN functions (include/seq/bond.h ... tortor.h compiled in place into
    oracle/_ref, which travels to the GPU box): the CUDA kernel and dk_bond ... dk_tortor on the 48 013 interactions of dhfr2."""
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba, calc
    from oracle import ref_bridge
    if not ref_bridge.available("valence"):
        pytest.skip("oracle/_ref not built")
    s = tg.load_system(os.path.join(GOLDEN, "dhfr2.npz"))
    ref = ref_bridge.valence(s)
    a = Amoeba(s, "mixed")
    a.attach_valence(s.valence)
    r = a.evalence(calc.v1)
    assert np.abs(np.array(list(r.e)) - ref["energy"]).max() < 1e-5
    assert np.abs(a.valence_gradient() - ref["grad"]).max() < 1e-7
    assert np.abs(np.array(list(r.virial)).reshape(3, 3) - ref["virial"]).max() < 1e-5
    a.close()


def test_energy_includes_valence():
    """energy(vers) = electrostatics + vdW + valence once the terms are attached (src/energy.cpp:180-215, 319-448);
    esum, gradient and virial are the sums of the separately evaluated parts."""
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba, calc
    s = tg.load_system(os.path.join(GOLDEN, "val_trpcage.npz"))
    a = Amoeba(s, "double", vdw=True)
    base = a.energy(calc.v1)
    a.attach_valence(s.valence)
    val = a.evalence(calc.v1)
    gv = a.valence_gradient()
    tot = a.energy(calc.v1)
    assert abs(tot["evalence"] - val.esum) < 1e-9
    assert abs(tot["esum"] - (base["esum"] + val.esum)) < 1e-7 * abs(tot["esum"])
    assert np.abs(tot["grad"] - (base["grad"] + gv)).max() < 1e-6
    assert np.abs(tot["virial"] - (base["virial"] + np.array(list(val.virial)).reshape(3, 3))).max() < 1e-5
    assert list(tot["nval_term"]) == list(val.count)
    a.close()


def test_bad_lists_are_refused():
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba, ApxError
    s = tg.load_system(os.path.join(GOLDEN, "val_water10.npz"))
    a = Amoeba(s, "mixed")
    v = copy.copy(s.valence)
    v.ibnd = v.ibnd.copy()
    v.ibnd[0, 1] = s.n + 5
    with pytest.raises(ApxError):
        a.attach_valence(v)
    with pytest.raises(ApxError):
        a.evalence()
    a.close()
