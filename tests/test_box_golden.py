"""The reference's box tests (test/box.cpp Box-1 orthogonal, Box-2 monoclinic, Box-3 triclinic): lattice vectors, volume and
minimum-image literals against our lattice builder (params.lattice, the boxLattice / lattice.f rules) and the oracle's
image(); margin 1e-6 as in the reference.  The CUDA image arithmetic is exercised by the triclinic parity tests."""
import numpy as np
import pytest

BOXES = {"ortho": (16, 16, 16, 90, 90, 90), "mono": (32, 24, 20, 90, 30, 90), "tri": (32, 24, 20, 75, 60, 45)}
# (box, r, image(r)) -- test/box.cpp:65-100, 140-180, 212-240
IMAGES = [("ortho", (-16, -12, -8), (0, 4, -8)), ("ortho", (-4, 0, 4), (-4, 0, 4)), ("ortho", (8, 12, 16), (-8, -4, 0)),
          ("ortho", (51, -83, 164), (3, -3, 4)),
          ("mono", (0, 0, 0), (0, 0, 0)), ("mono", (-8, -6, 0), (-8, -6, 0)), ("mono", (5, 10, 15), (2.3589838486, 10, -5)),
          ("mono", (-13, -30, 20), (-15.6410161514, -6.0, 0)), ("mono", (-18, -40, 5), (-3.3205080757, 8.0, -5)),
          ("mono", (-18, -16, 5), (-3.3205080757, 8.0, -5)),
          ("tri", (0, 0, 0), (0, 0, 0)), ("tri", (5, 10, 15), (10.02943725, -4.29107082, -2.11199354)),
          ("tri", (-13, -30, 20), (10.94112550, 6.62061742, 2.88800646)), ("tri", (-18, -40, 5), (-16.05887450, -6.05887450, 5)),
          ("tri", (0.91168825, 10.91168824, 5), (-16.05887450, -6.05887450, 5))]


def _cell(name):
    import importlib
    params = importlib.import_module("tinker-gpu_b200.params")
    return params.lattice(*BOXES[name])


def test_lattice_vectors_and_volumes():
    lv, rc = _cell("mono")
    a, b, c, _, be, _ = BOXES["mono"]
    cb, sb = np.cos(np.radians(be)), np.sin(np.radians(be))
    # lvec1 = (a, 0, c cos(beta)), lvec2 = (0, b, 0), lvec3 = (0, 0, c sin(beta))   (test/box.cpp:118-131)
    assert np.allclose(lv, [[a, 0, c * cb], [0, b, 0], [0, 0, c * sb]], atol=1e-6)
    assert abs(np.linalg.det(lv)) == pytest.approx(a * b * c * sb, abs=1e-6)
    lv, _ = _cell("ortho")
    assert abs(np.linalg.det(lv)) == pytest.approx(16 ** 3, abs=1e-6)
    lv, rc = _cell("tri")
    al, be, ga = (np.cos(np.radians(x)) for x in BOXES["tri"][3:])
    vol = 32 * 24 * 20 * np.sqrt(1 - al * al - be * be - ga * ga + 2 * al * be * ga)
    assert abs(np.linalg.det(lv)) == pytest.approx(vol, abs=1e-6)
    assert np.allclose(rc @ lv.T, np.eye(3), atol=1e-12) or np.allclose(rc @ lv, np.eye(3), atol=1e-12)


@pytest.mark.parametrize("box,r,expect", IMAGES)
def test_minimum_image_literals(box, r, expect):
    from oracle.amoeba_ref import Oracle
    lv, rc = _cell(box)
    o = Oracle.__new__(Oracle)
    o.lvec, o.recip = np.asarray(lv, float), np.asarray(rc, float)
    got = o.image(np.array([r], float))[0]
    # the reference compares |components| for imagen2 and components for image; a point exactly on the cell face (-L/2 vs +L/2)
    # is the same image, so faces are compared by magnitude
    on_face = np.isclose(np.abs(got @ o.recip.T), 0.5, atol=1e-9)
    assert np.allclose(np.where(on_face, np.abs(got), got), np.where(on_face, np.abs(expect), expect), atol=1e-6)
