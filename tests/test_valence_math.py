"""The hand-derived valence gradients of the CUDA library (csrc/valmath.cuh, valterms.cuh, valpack.h) run on the CPU:
tests/valmath_host.cpp compiles those very headers with g++ and walks the interactions in a loop.  Compared with the
autograd oracle (oracle/valence_ref.py) in double (tight) and float (the mixed build's pair-math type), and directly
with the reference goldens.  No GPU involved: this checks the math the kernel executes, not the launch plumbing."""
import ctypes as C
import importlib
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import valence_ref as vr

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
G = os.path.join(HERE, "golden")
vp = importlib.import_module("tinker-gpu_b200.valparams")
am = importlib.import_module("tinker-gpu_b200.amoeba")
GOLD = json.load(open(os.path.join(G, "valence_goldens.json")))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("vm") / "valmath_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "tinker-gpu_b200", "csrc"), os.path.join(HERE, "valmath_host.cpp"), "-o", so])
    lib = C.CDLL(so)
    lib.valmath_host_eval.argtypes = [C.c_int, C.POINTER(am._ApxValence), C.POINTER(C.c_double), C.c_int,
                                      C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    return lib


def run(lib, xyz, v, real_bytes, terms=None):
    import copy
    v = copy.copy(v)
    if terms is not None:
        v.use = np.array([int(t in terms) for t in vp.TERMS], np.int32)
    s, keep = am.valence_struct(v, len(xyz))
    xyz = np.ascontiguousarray(xyz, np.float64)
    e8, g, v6 = np.zeros(8), np.zeros((len(xyz), 3)), np.zeros(6)
    dp = C.POINTER(C.c_double)
    lib.valmath_host_eval(real_bytes, C.byref(s), xyz.ctypes.data_as(dp), 1, e8.ctypes.data_as(dp), g.ctypes.data_as(dp),
                          v6.ctypes.data_as(dp))
    vir = np.array([[v6[0], v6[1], v6[2]], [v6[1], v6[3], v6[4]], [v6[2], v6[4], v6[5]]])
    return e8, g, vir


def load(blob):
    import tinker_gpu_b200 as tg
    s = tg.load_system(os.path.join(G, blob))
    return s.xyz, s.valence


@pytest.mark.parametrize("term", list(vr.TERMS))
def test_double_matches_oracle_per_term(host, term):
    xyz, v = load(GOLD[term]["blob"])
    e8, g, vir = run(host, xyz, v, 8, [term])
    r = vr.valence(xyz, v, terms=[term])
    k = vp.TERMS.index(term)
    assert abs(e8[k] - r["esum"]) <= 1e-10 * max(1.0, abs(r["esum"]))
    assert np.abs(e8).sum() == pytest.approx(abs(e8[k]))
    assert np.abs(g - r["grad"]).max() <= 1e-9
    assert np.abs(vir - r["virial"]).max() <= 1e-8


@pytest.mark.parametrize("term", list(vr.TERMS))
def test_float_matches_reference_golden(host, term):
    """The mixed build's arithmetic (float interaction math, double accumulation) against test/ref/*.txt with the
    reference's own tolerances for its mixed-precision build (test/bond.cpp:22-24 ... test/tortor.cpp:24-26)."""
    gold = GOLD[term]
    xyz, v = load(gold["blob"])
    e8, g, vir = run(host, xyz, v, 4, [term])
    assert abs(e8.sum() - gold["energy"]) <= 1.5e-4
    ref_g = np.array(gold["grad"])
    assert np.abs(g[:len(ref_g)] - ref_g).max() <= 4e-3
    assert np.abs(vir - np.array(gold["virial"]).reshape(3, 3)).max() <= 6e-3


def test_dhfr2_all_terms(host):
    """Full-size deck: all 48k interactions, double and float against the oracle; net force vanishes."""
    xyz, v = load("dhfr2.npz")
    r = vr.valence(xyz, v)
    e8, g, vir = run(host, xyz, v, 8)
    assert abs(e8.sum() - r["esum"]) <= 1e-9 * abs(r["esum"])
    assert np.abs(g - r["grad"]).max() <= 1e-8
    assert np.abs(vir - r["virial"]).max() <= 1e-6
    assert np.abs(g.sum(0)).max() < 1e-9
    # float interaction math (what the reference's mixed build does) misses the north-star force tolerance of
    # 1e-5 kcal/mol/A RMS: r - r0 of a stiff bond loses 6e-8 A, times 2k ~ 1000.  Measured here: 4e-5.  The CUDA
    # kernel therefore evaluates the valence terms in double in BOTH builds (48k interactions: the cost is nil).
    e4, g4, vir4 = run(host, xyz, v, 4)
    assert abs(e4.sum() - r["esum"]) <= 1e-6 * abs(r["esum"])
    rms4 = np.sqrt(((g4 - r["grad"]) ** 2).sum(1).mean())
    assert 1e-5 < rms4 <= 1e-4
