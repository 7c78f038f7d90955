"""The reference's PolPair tests (test/polpair.cpp, test/ref/polpair.{1,2}.txt): a NaCl pair whose Thole width comes from a
POLPAIR record (`polpair 7 15 0.05` -> the thlval[jpolar_i][jpolar_k] table of the pair kernels), with and without Ewald.
Total energy, gradient and virial against the reference's transcript literals with the reference's own tolerances
(1e-4, 1e-4, 1e-3): the oracle on the CPU, the CUDA path on the GPU.  Fixtures: tests/golden/make_polpair_golden.py."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

CASES = ["polpair_ewald", "polpair_nonewald"]


def _load(name):
    import tinker_gpu_b200 as tg
    g = json.load(open(os.path.join(GOLDEN, "polpair_goldens.json")))[name]
    return tg.load_system(os.path.join(GOLDEN, name + ".npz")), g


@pytest.mark.parametrize("name", CASES)
def test_reader_applies_the_polpair_record(name):
    s, _ = _load(name)
    assert s.n == 2 and s.use_ewald == (name == "polpair_ewald")
    assert s.thlval[s.jpolar[0], s.jpolar[1]] == pytest.approx(0.05) and s.thlval[s.jpolar[1], s.jpolar[0]] == pytest.approx(0.05)
    assert s.thlval[s.jpolar[0], s.jpolar[0]] == pytest.approx(0.39)      # amoeba09 Thole width everywhere else


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_polpair_transcripts(name):
    from oracle.amoeba_ref import Oracle, V0, V1
    s, g = _load(name)
    o = Oracle(s)
    assert abs(o.energy(V0)["esum"] - g["energy"]) < 1e-4
    r = o.energy(V1)
    assert abs(r["esum"] - g["energy"]) < 1e-4
    assert np.abs(r["grad"] - np.array(g["gradient"])).max() < 1e-4
    assert np.abs(r["virial"] - np.array(g["virial"])).max() < 1e-3


def _cuda_check(name, precision):
    """energy(v0 / v1 / v4 / v5 / v6) as test/polpair.cpp:33-55 calls them."""
    from tinker_gpu_b200.amoeba import Amoeba, calc
    s, g = _load(name)
    a = Amoeba(s, precision)
    try:
        assert abs(a.energy(calc.v0)["esum"] - g["energy"]) < 1e-4
        r = a.energy(calc.v1)
        assert abs(r["esum"] - g["energy"]) < 1e-4
        assert np.abs(r["grad"] - np.array(g["gradient"])).max() < 1e-4
        assert np.abs(r["virial"] - np.array(g["virial"])).max() < 1e-3
        r = a.energy(calc.v4)
        assert abs(r["esum"] - g["energy"]) < 1e-4 and np.abs(r["grad"] - np.array(g["gradient"])).max() < 1e-4
        assert np.abs(a.energy(calc.v5)["grad"] - np.array(g["gradient"])).max() < 1e-4
        r = a.energy(calc.v6)
        assert np.abs(r["grad"] - np.array(g["gradient"])).max() < 1e-4 and np.abs(r["virial"] - np.array(g["virial"])).max() < 1e-3
    finally:
        a.close()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["mixed", "double"])
@pytest.mark.parametrize("name", CASES)
def test_cuda_path_reproduces_the_polpair_transcripts(name, precision):
    """In a child process (a two-atom electrostatics context; green on the B200 since round 2, profiles/r02v_tests.log)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import test_polpair_golden as t; t._cuda_check(%r, %r)"
            % (root, os.path.join(root, "tests"), name, precision))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, (r.stderr or r.stdout)[-800:]
