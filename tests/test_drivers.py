"""ANALYZE / TESTGRAD driver shells (tinker-gpu_b200/drivers.py): the formatting reproduces lines of the reference's own
transcripts character by character (test/ref/vdw14.1.txt, produced by src/xanalyze.cpp / src/xtestgrad.cpp), the moments
follow xAnalyzeMoments; on the GPU the drivers run end to end on the 22-atom local-frame deck."""
import io
import os

import numpy as np
import pytest

from conftest import GOLDEN


def test_energy_and_virial_lines_match_reference_transcript():
    from tinker_gpu_b200.drivers import format_energy_breakdown, format_virial
    txt = format_energy_breakdown(-80.5672, [("Van der Waals", -80.5672, 17290)])
    # test/ref/vdw14.1.txt:3-4
    assert " Energy Component Breakdown :           Kcal/mole        Interactions" in txt.splitlines()
    assert " Van der Waals                           -80.5672            17290" in txt.splitlines()
    assert " Total Potential Energy :                -80.5672 Kcal/mole" in txt.splitlines()
    v = [[-648.117, -119.517, 31.637], [-119.517, -491.019, 30.124], [31.637, 30.124, -462.209]]
    lines = format_virial(v, 100, 1000.0).splitlines()
    # test/ref/vdw14.1.txt:6-8
    assert " Internal Virial Tensor :                -648.117     -119.517       31.637" in lines
    assert "                                         -119.517     -491.019       30.124" in lines
    assert "                                           31.637       30.124     -462.209" in lines
    assert any(ln.startswith(" Pressure (Temp 298 K) :") and ln.endswith("Atmospheres") for ln in lines)


def test_testgrad_lines_match_reference_transcript():
    from tinker_gpu_b200.drivers import format_testgrad
    g = np.array([[-2.2181, -0.7271, -1.6957], [-10.1362, -2.4631, 1.2675]])
    txt = format_testgrad(-80.5672, g, None, 4)
    lines = txt.splitlines()
    # test/ref/vdw14.1.txt:12-15
    assert "  Type      Atom              dE/dX       dE/dY       dE/dZ          Norm" in lines
    assert " Anlyt         1            -2.2181     -0.7271     -1.6957        2.8851" in lines
    assert " Anlyt         2           -10.1362     -2.4631      1.2675       10.5079" in lines
    assert " Total Gradient Norm and RMS Gradient per Atom :" in lines
    assert any(ln.startswith(" Anlyt      Total Gradient Norm Value") for ln in lines)
    both = format_testgrad(1.0, g, g + 1e-4, 6).splitlines()
    assert sum(ln.startswith(" Numer") for ln in both) == 4          # 2 rows + norm + rms


def test_moments_of_simple_charge_sets():
    from tinker_gpu_b200.drivers import moments, DEBYE
    xyz = np.array([[0.5, 0.0, 0.0], [-0.5, 0.0, 0.0]])
    rp = np.zeros((2, 10))
    rp[0, 0], rp[1, 0] = 1.0, -1.0
    mo = moments(xyz, [1.0, 1.0], rp, np.zeros((2, 3)))
    assert abs(mo["netchg"]) < 1e-15 and abs(mo["netdpl"] - DEBYE) < 1e-12 and abs(mo["dipole"][0] - DEBYE) < 1e-12
    assert abs(np.trace(mo["quadrupole"])) < 1e-12
    # an induced dipole adds to the permanent one; a linear quadrupole +q, -2q, +q has no dipole
    mo2 = moments(xyz, [1.0, 1.0], rp, np.array([[0.1, 0, 0], [0.1, 0, 0]]))
    assert abs(mo2["dipole"][0] - 1.2 * DEBYE) < 1e-12
    xyz3 = np.array([[1.0, 0, 0], [0.0, 0, 0], [-1.0, 0, 0]])
    rp3 = np.zeros((3, 10))
    rp3[:, 0] = [1.0, -2.0, 1.0]
    mo3 = moments(xyz3, [1.0, 1.0, 1.0], rp3, np.zeros((3, 3)))
    assert mo3["netdpl"] < 1e-12 and abs(mo3["quadrupole"][0, 0] - 2.0 * DEBYE) < 1e-12
    assert abs(mo3["quadrupole"][1, 1] + 1.0 * DEBYE) < 1e-12


@pytest.mark.gpu
def test_drivers_end_to_end():
    """analyze E/M/V and testgrad (analytical + numerical on three atoms) on the 22-atom deck with PME: the printed
    multipole + polarization energies are the reference's goldens, analytical and numerical gradients agree."""
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.drivers import analyze, testgrad
    s = tg.load_system(os.path.join(GOLDEN, "lf_local_frame_4.npz"))
    buf = io.StringIO()
    res = analyze(s, "EMV", "double", out=buf)
    txt = buf.getvalue()
    assert " Polarization" in txt and " Internal Virial Tensor :" in txt        # this deck says `polarizeterm only`
    assert " Total Potential Energy :                -36.5477 Kcal/mole" in txt.splitlines()
    assert " Total Electric Charge :" in txt
    assert abs(res["E"]["ep"] - (-36.5477)) < 1e-4                    # test/localframe.cpp:508
    assert abs(res["M"]["netchg"]) < 1e-9
    buf = io.StringIO()
    tgr = testgrad(s, True, True, 1e-4, 6, "double", out=buf, atoms=[0, 5, 17])
    for i in (0, 5, 17):
        assert np.abs(tgr["anlyt"][i] - tgr["numer"][i]).max() < 5e-5
    assert " Anlyt" in buf.getvalue() and " Numer" in buf.getvalue()
