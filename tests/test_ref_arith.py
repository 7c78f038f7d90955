"""oracle/_ref: the reference's OWN arithmetic for the path (its include/seq pair and valence headers compiled in place by
oracle/Makefile) against the oracle restatements and against the CPU build of the CUDA library's valence math.  This is
what ties the oracle -- and through the oracle the CUDA path -- to the reference at full dhfr2 size, where the reference
tree holds no golden.  Skipped where neither /root/reference nor a prebuilt oracle/_ref exists."""
import glob
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, REFERENCE, ROOT


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_bridge
    if os.path.isdir(REFERENCE):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    if not (ref_bridge.available("valence") and ref_bridge.available("realspace") and ref_bridge.available("pme")):
        pytest.skip("oracle/_ref not built (no reference tree here)")
    return ref_bridge


def _load(blob):
    import tinker_gpu_b200 as tg
    return tg.load_system(os.path.join(GOLDEN, blob))


@pytest.mark.parametrize("blob", ["val_trpcage.npz", "val_trpcage_angle.npz", "val_water10.npz", "dhfr2.npz"])
def test_valence_oracle_equals_reference_arithmetic(ref, blob):
    """dk_bond ... dk_tortor of include/seq/*.h in double against the autograd oracle: every term energy, the gradient and
    the virial (incl. the reference's own pi-torsion virial expression), up to the 48 013 interactions of dhfr2."""
    from oracle import valence_ref as vr
    s = _load(blob)
    r = ref.valence(s)
    o = vr.valence(s.xyz, s.valence)
    for k, t in enumerate(vr.TERMS):
        assert abs(r["energy"][k] - o["energy"].get(t, 0.0)) <= 1e-11 * max(1.0, abs(o["energy"].get(t, 0.0))), t
    assert np.abs(r["grad"] - o["grad"]).max() < 1e-10
    assert np.abs(r["virial"] - o["virial"]).max() < 1e-8


@pytest.mark.parametrize("blob", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "lf_*.npz"))))
def test_realspace_oracle_equals_reference_arithmetic_small(ref, blob):
    """pair_mpole / pair_polar / pair_dfield / pair_ufield (EWALD pass + NON_EWALD (scale-1) pass, as the reference's loops
    do) against the oracle's tensor-contraction formulation on every local-frame deck: orthogonal, monoclinic, triclinic
    cells, Ewald and non-Ewald, with 1-2...1-5 and group scaling."""
    from oracle.amoeba_ref import Oracle, V1
    s = _load(blob)
    o = Oracle(s)
    o.rotpole()
    if s.use_polar:
        o.induce()
    r = ref.realspace(o, o.uind if s.use_polar else None, o.uinp if s.use_polar else None)
    rs = o._real_space(V1, s.use_mpole, s.use_polar)
    fd, fp = o.dfield(real_only=True)
    tol = 1e-11
    if s.use_mpole:
        assert abs(r["em"] - rs["em"]) <= tol * max(1.0, abs(rs["em"]))
        assert np.abs(r["gm"] - rs["gm"]).max() <= tol * max(1.0, np.abs(rs["gm"]).max())
        assert np.abs(r["tm"] - rs["tm"]).max() <= tol * max(1.0, np.abs(rs["tm"]).max())
        assert np.abs(r["vm"] - rs["vm"]).max() <= tol * max(1.0, np.abs(rs["vm"]).max())      # pairwise virial, empoleewald.cpp:100-107
    assert np.abs(r["fd"] - fd).max() <= tol and np.abs(r["fp"] - fp).max() <= tol
    if s.use_polar:
        ufd, ufp = o.ufield(o.uind, o.uinp, real_only=True)
        assert abs(r["ep"] - rs["ep"]) <= tol * max(1.0, abs(rs["ep"]))
        assert np.abs(r["gp"] - rs["gp"]).max() <= tol * max(1.0, np.abs(rs["gp"]).max())
        assert np.abs(r["tp"] - rs["tp"]).max() <= tol * max(1.0, np.abs(rs["tp"]).max())
        assert np.abs(r["vp"] - rs["vp"]).max() <= tol * max(1.0, np.abs(rs["vp"]).max())      # epolarewald.cpp:163-170
        assert np.abs(r["ufd"] - ufd).max() <= tol and np.abs(r["ufp"] - ufp).max() <= tol


def test_realspace_dhfr2_fixture_equals_reference_arithmetic(ref):
    """Full size: the oracle's real-space energies, gradient, torque and fields over the 1 644 163 pairs of dhfr2 (fixture
    tests/golden/dhfr2_oracle_real.npz, ~10 oracle minutes, made by make_ref_fixtures.py) against the reference's pair
    functions run here in about a second.  With test_gpu_parity.py::test_dhfr2_vs_oracle_fixture holding the CUDA path to
    the same oracle run, dhfr2 parity is pinned to the reference's arithmetic for everything but the PME reciprocal part."""
    from oracle.amoeba_ref import Oracle
    s = _load("dhfr2.npz")
    z = np.load(os.path.join(GOLDEN, "dhfr2_oracle.npz"))
    f = np.load(os.path.join(GOLDEN, "dhfr2_oracle_real.npz"))
    o = Oracle(s)
    o.rotpole()
    r = ref.realspace(o, z["uind"], z["uinp"])
    assert r["npair"] == int(f["npairs"]) == int(z["npairs"])
    assert abs(r["em"] - float(z["em_real"])) <= 1e-11 * abs(float(z["em_real"]))
    assert abs(r["em"] - float(f["em_real"])) <= 1e-11 * abs(float(f["em_real"]))
    assert abs(r["ep"] - float(f["ep_real"])) <= 1e-11 * abs(float(f["ep_real"]))
    g, t = r["gm"] + r["gp"], r["tm"] + r["tp"]
    assert np.abs(g - f["g_real"]).max() <= 1e-10 * np.abs(f["g_real"]).max()
    assert np.abs(t - f["t_real"]).max() <= 1e-10 * np.abs(f["t_real"]).max()
    for a, b in (("fd", "fd_real"), ("fp", "fp_real"), ("ufd", "ufd_real"), ("ufp", "ufp_real")):
        assert np.abs(r[a] - f[b]).max() <= 1e-12, a


def test_cuda_valence_math_equals_reference_arithmetic(ref, tmp_path):
    """The interaction math the CUDA kernel executes (csrc/valmath.cuh, compiled for the CPU by tests/valmath_host.cpp,
    double) against the reference's functions on dhfr2."""
    import ctypes as C
    import importlib
    am = importlib.import_module("tinker-gpu_b200.amoeba")
    so = str(tmp_path / "valmath_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "tinker-gpu_b200", "csrc"), os.path.join(ROOT, "tests", "valmath_host.cpp"), "-o", so])
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    lib.valmath_host_eval.argtypes = [C.c_int, C.POINTER(am._ApxValence), dp, C.c_int, dp, dp, dp]
    s = _load("dhfr2.npz")
    st, keep = am.valence_struct(s.valence, s.n)
    x = np.ascontiguousarray(s.xyz, np.float64)
    e8, g, v6 = np.zeros(8), np.zeros((s.n, 3)), np.zeros(6)
    lib.valmath_host_eval(8, C.byref(st), x.ctypes.data_as(dp), 1, e8.ctypes.data_as(dp), g.ctypes.data_as(dp), v6.ctypes.data_as(dp))
    r = ref.valence(s)
    assert np.abs(e8 - r["energy"]).max() <= 1e-9
    assert np.abs(g - r["grad"]).max() <= 1e-9
    vir = np.array([[v6[0], v6[1], v6[2]], [v6[1], v6[3], v6[4]], [v6[2], v6[4], v6[5]]])
    assert np.abs(vir - r["virial"]).max() <= 1e-7


def test_bspline_tables_equal_reference(ref):
    """Order-5 B-spline weights and derivatives (the PME spreading / gathering stencil) from bsplgen<4> of
    include/seq/bsplgen.h against the oracle's tables, over the whole range of fractional offsets."""
    from oracle.amoeba_ref import Oracle
    w = np.concatenate([np.linspace(0.0, 1.0, 2001)[:-1], np.random.default_rng(1).uniform(0, 1, 500)])
    assert np.abs(ref.bspline5(w) - Oracle.bspline_theta(w, 5)).max() < 1e-15


def test_vdw_pair_terms_equal_reference(ref):
    """pair_hal_v2 (include/seq/pair_hal.h) against the vdW oracle's pair function over every class pair of dhfr2 at
    distances from contact to beyond the taper: energy and dE/dr, tapered region included."""
    from oracle.vdw_ref import VdwOracle
    s = _load("dhfr2.npz")
    o = VdwOracle(s)
    v = s.vdw
    rng = np.random.default_rng(2)
    m = 200000
    a, b = rng.integers(0, v.radmin.shape[0], m), rng.integers(0, v.radmin.shape[0], m)
    r = rng.uniform(1.2, v.cutoff, m)
    r[:2000] = rng.uniform(v.taper, v.cutoff, 2000)
    rv, eps = v.radmin[a, b], v.epsilon[a, b]
    ok = rv > 0
    e0, de0 = o.pair_terms(r[ok], rv[ok], eps[ok])
    e1, de1 = ref.hal(r[ok], rv[ok], eps[ok], v.taper, v.cutoff, v.ghal, v.dhal)
    assert np.abs(e1 - e0).max() <= 1e-12 * max(1.0, np.abs(e0).max())
    assert np.abs(de1 - de0).max() <= 1e-12 * max(1.0, np.abs(de0).max())


@pytest.mark.parametrize("blob", ["lf_local_frame_2.npz", "lf_local_frame3_2.npz", "lf_triclinic.npz", "water30.npz", "dhfr2.npz"])
def test_pme_operators_equal_reference(ref, blob):
    """The reference's host PME translation unit (src/acc/pme.cpp, compiled unmodified into oracle/_ref/libref_pme.so with a
    17-symbol shim) against the oracle, operator by operator, on cubic and triclinic cells up to the 64^3 grid of dhfr2:
    rpoleToCmp, cmpToFmp, gridMpole, pmeConv (grid, reciprocal energy, virial), fphiMpole, fphiToCphi, cuindToFuind, gridUind,
    fphiUind2, fphiUind.  The FFT between them is numpy's on both sides."""
    from oracle.amoeba_ref import Oracle
    if not ref.available("pme"):
        pytest.skip("oracle/_ref/libref_pme.so not built")
    s = _load(blob)
    o = Oracle(s)
    o.rotpole()
    rp = o._ensure_rpole()
    P = ref.RefPME(o)

    def close(a, b, tol=1e-13):
        assert np.abs(np.asarray(a) - np.asarray(b)).max() <= tol * max(1.0, np.abs(np.asarray(b)).max())
    c = o.rpole_to_cmp(rp)
    close(P.rpole_to_cmp(rp), c)
    f = o.cmp_to_fmp(c)
    close(P.cmp_to_fmp(c), f)
    g = o.grid_mpole(f)
    close(P.grid_mpole(f), g)
    q, e, v = o.pme_convolve(g, True)
    q1, e1, v1 = P.convolve(g, True)
    close(q1, q)
    assert abs(e1 - e) <= 1e-13 * abs(e)
    close(v1, v, 1e-12)
    if blob == "dhfr2.npz":      # the full-size oracle run of the fixture used the same reciprocal multipole energy
        z = np.load(os.path.join(GOLDEN, "dhfr2_oracle.npz"))
        assert abs(e1 - float(z["em_recip"])) <= 1e-12 * abs(e1)
    ph = o.fphi_gather(q.real, 20)
    close(P.fphi_mpole(q), ph)
    close(P.fphi_to_cphi(ph), o.fphi_to_cphi(ph))
    rng = np.random.default_rng(0)
    ud, up = rng.normal(size=(s.n, 3)) * 0.05, rng.normal(size=(s.n, 3)) * 0.05
    fud, fup = P.cuind_to_fuind(ud, up)
    close(fud, o.cuind_to_fuind(ud))
    close(fup, o.cuind_to_fuind(up))
    gu = o.grid_uind(fud, fup)
    close(P.grid_uind(fud, fup), gu)
    qu, _, _ = o.pme_convolve(gu)
    a0, b0 = o.fphi_gather(qu.real, 10), o.fphi_gather(qu.imag, 10)
    a1, b1 = P.fphi_uind2(qu)
    close(a1[:, 1:], a0[:, 1:])      # the reference leaves the potential (component 0) at zero, src/acc/pme.cpp:668
    close(b1[:, 1:], b0[:, 1:])
    a2, b2, d2 = P.fphi_uind(qu)
    close(a2[:, 1:], a0[:, 1:])
    close(b2[:, 1:], b0[:, 1:])
    close(d2, o.fphi_gather(qu.real + qu.imag, 20))


def test_dhfr2_energy_and_dipoles_by_reference_operators(ref):
    """oracle/ref_oracle.py: the oracle's PCG loop and energy assembly with EVERY pair sweep and PME operator replaced by the
    reference's compiled code, on the full dhfr2 deck (15 s): converged induced dipoles, multipole and polarization energies
    and the gradient against the oracle fixture the CUDA path is held to (tests/golden/dhfr2_oracle.npz).  This is what pins
    dhfr2 parity -- for which the reference tree holds no golden -- to the reference's own arithmetic."""
    from oracle.amoeba_ref import V1
    from oracle.ref_oracle import RefOracle
    s = _load("dhfr2.npz")
    z = np.load(os.path.join(GOLDEN, "dhfr2_oracle.npz"))
    r = RefOracle(s)
    e = r.energy(V1)
    assert r.niter == int(z["niter"]) == 7
    assert abs(e["em"] - float(z["em"])) <= 1e-12 * abs(float(z["em"]))
    assert abs(e["ep"] - float(z["ep"])) <= 1e-12 * abs(float(z["ep"]))
    assert np.sqrt(((r.uind - z["uind"]) ** 2).mean()) * 4.803206802 <= 1e-12          # Debye
    assert np.sqrt(((r.uinp - z["uinp"]) ** 2).mean()) * 4.803206802 <= 1e-12
    assert np.abs(e["grad"] - z["grad"]).max() <= 1e-10
    assert np.abs(e["virial"] - z["virial"]).max() <= 1e-11 * np.abs(z["virial"]).max()


@pytest.mark.parametrize("blob", ["lf_local_frame_2.npz", "lf_local_frame3_2.npz", "lf_triclinic.npz", "dhfr2.npz"])
def test_local_frames_equal_reference(ref, blob):
    """chkpole + rotpole and torque -> force / torque virial: the reference's host translation units
    src/acc/amoeba/rotpole.cpp and torque.cpp (compiled unmodified) against the oracle, on decks that use all five frame types
    (None, Z-Only, Z-then-X, Bisector, Z-Bisect, 3-Fold) and on dhfr2."""
    from oracle.amoeba_ref import Oracle
    s = _load(blob)
    o = Oracle(s)
    pc, rp = ref.rotpole(s.xyz, s.zaxis, s.pole)
    o.chkpole()
    assert np.array_equal(pc, o.pole)
    assert np.abs(rp - o.rotpole()).max() <= 1e-14
    trq = np.random.default_rng(3).normal(size=(s.n, 3))
    g1, v1 = ref.torque(s.xyz, s.zaxis, trq)
    g0 = np.zeros((s.n, 3))
    v0 = o.torque(trq, g0, True)
    assert np.abs(g1 - g0).max() <= 1e-11 * max(1.0, np.abs(g0).max())
    assert np.abs(v1 - v0).max() <= 1e-11 * max(1.0, np.abs(v0).max())


@pytest.mark.parametrize("blob", ["lf_local_frame_2.npz", "lf_local_frame3_2.npz", "lf_triclinic.npz", "water30.npz", "dhfr2.npz"])
def test_reciprocal_assembly_equals_reference(ref, blob):
    """Per-atom reciprocal energy / gradient / torque / virial assembly: the reference's host translation units
    src/acc/hippo/empole.cpp (empoleChgpenEwaldRecip_acc, AMOEBA branch) and src/acc/amoeba/epolarewald.cpp
    (epolarEwaldRecipSelf_acc incl. the self term and the structure-factor virial), compiled unmodified, against the oracle's
    empole_recip / epolar_recip_self -- the last pieces of the dhfr2 parity chain that were ours alone.  Induced dipoles are
    random (the assembly is linear in them), so no PCG run is needed; the FFT is numpy's on both sides."""
    from oracle.amoeba_ref import V1, Oracle
    if not ref.available("pme"):
        pytest.skip("oracle/_ref/libref_pme.so not built")
    s = _load(blob)
    if not s.use_ewald:
        pytest.skip("no reciprocal space in this deck")
    o = Oracle(s)
    o.rotpole()
    rp = o._ensure_rpole()
    P = ref.RefPME(o)
    rng = np.random.default_rng(11)
    o.uind, o.uinp = rng.normal(size=(s.n, 3)) * 0.05, rng.normal(size=(s.n, 3)) * 0.05

    def close(a, b, tol=1e-12):
        a, b = np.asarray(a), np.asarray(b)
        assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())
    m0 = o.empole_recip(V1)
    m1 = ref.recip_mpole(P, rp)
    assert abs(m1["e"] - m0["e"]) <= 1e-12 * max(1.0, abs(m0["e"]))
    close(m1["g"], m0["g"])
    close(m1["t"], m0["t"])
    close(m1["v"], m0["v"])
    p0 = o.epolar_recip_self(V1)
    p1 = ref.recip_polar(P, o.uind, o.uinp)
    assert abs(p1["e"] - p0["e"]) <= 1e-12 * max(1.0, abs(p0["e"]))
    close(p1["g"], p0["g"])
    close(p1["t"], p0["t"])
    close(p1["v"], p0["v"], 1e-11)
