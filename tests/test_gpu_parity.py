"""Parity of the CUDA path (through the C ABI, tinker-gpu_b200/libapx*.so) with the float64 oracle.

double build : proves the algorithm (tolerances ~1e-9, limited by fixed-point 2^-32 force quanta)
mixed build  : the shipped precision (float pair math, fixed-point / f64 accumulation); tolerances are
               the north-star ones where float allows: energy 1e-6 relative, dipoles 1e-6 D RMS;
               forces are asserted at the north-star 1e-5 kcal/mol/A RMS.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_case, section

pytestmark = pytest.mark.gpu

DEBYE = 4.803206802


def _amoeba(system, precision):
    from tinker_gpu_b200.amoeba import Amoeba
    return Amoeba(system, precision)


def _rms(a):
    return float(np.sqrt((np.asarray(a) ** 2).mean()))


CASES = ["Local-Frame-1", "Local-Frame-2", "Local-Frame-3", "Local-Frame-4", "Local-Frame3-1", "Local-Frame3-2"]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("precision", ["double", "mixed"])
def test_small_systems_vs_oracle(case, precision):
    from tinker_gpu_b200.amoeba import calc
    from oracle.amoeba_ref import Oracle, V1, V3
    s = load_case(case)
    a = _amoeba(s, precision)
    o = Oracle(s)
    tol = dict(e=1e-10, g=1e-7, v=1e-6, f=1e-12, u=1e-12) if precision == "double" else dict(e=2e-6, g=1e-5, v=5e-4, f=2e-6, u=5e-6)
    o.rotpole()
    assert np.abs(a.rpole() - o.rpole).max() < (1e-14 if precision == "double" else 1e-6)
    if s.use_polar:
        fd, fp = a.dfield()
        od, op = o.dfield()
        assert np.abs(fd - od).max() < tol["f"] and np.abs(fp - op).max() < tol["f"]
        n = s.n
        ud = np.array([[0.1 * (i + 1) + 0.03 * (j + 1) for j in range(3)] for i in range(n)])
        up = np.array([[0.1 * (i + 1) - 0.03 * (j + 1) for j in range(3)] for i in range(n)])
        f1, f2 = a.ufield(ud, up)
        o1, o2 = o.ufield(ud, up)
        scale = np.abs(o1).max()
        assert np.abs(f1 - o1).max() < tol["f"] * max(1.0, scale) * 10 and np.abs(f2 - o2).max() < tol["f"] * max(1.0, scale) * 10
        z1, z2 = a.sparsePrecondApply(ud, up)
        p1, p2 = o.precond(ud, up)
        assert np.abs(z1 - p1).max() < tol["f"] * 100 and np.abs(z2 - p2).max() < tol["f"] * 100
        u1, u2 = a.induce()
        v1, v2 = o.induce()
        assert a.stats()["pcg_iterations"] == o.niter
        assert np.abs(u1 - v1).max() * DEBYE < tol["u"] and np.abs(u2 - v2).max() * DEBYE < tol["u"]
    r = a.energy(calc.v1)
    ro = o.energy(V1)
    assert abs(r["esum"] - ro["esum"]) < tol["e"] * abs(ro["esum"])
    assert _rms(r["grad"] - ro["grad"]) < tol["g"]
    assert np.abs(r["virial"] - ro["virial"]).max() < tol["v"] * max(1.0, np.abs(ro["virial"]).max())
    # ANALYZE path: pairwise polarization energy and interaction counts
    r3 = a.energy(calc.v3)
    ro3 = o.energy(V3)
    assert abs(r3["ep"] - ro3["ep"]) < tol["e"] * max(1.0, abs(ro3["ep"]))
    nself = s.n if s.use_ewald else 0
    if s.use_mpole:
        assert r3["nem"] == ro3["nem"] + nself
    if s.use_polar:
        assert r3["nep"] == ro3["nep"] + nself
    a.close()


@pytest.mark.parametrize("case", CASES)
def test_small_systems_vs_reference_goldens(goldens, case):
    """The CUDA path against the reference's OWN literals (test/localframe.cpp, test/localframe3.cpp),
    with the reference's tolerances."""
    from tinker_gpu_b200.amoeba import calc
    s = load_case(case)
    a = _amoeba(s, "mixed")
    r = a.energy(calc.v1)
    if case.startswith("Local-Frame3"):
        ref = section(goldens, case, "emplar")
        # the literal is printed to 4 decimals (+-5e-5) and the mixed build is held to 1e-6 relative (-109.8384: 1.1e-4): the
        # converged dipoles, and with them E_polar, move by that much with the path the preconditioned solver takes
        assert abs(r["esum"] - ref["ref_eng"]) < 5e-5 + 1e-6 * abs(ref["ref_eng"])
        assert np.abs(r["grad"] - np.array(ref["ref_g"])).max() < 2e-4
        assert np.abs(r["virial"] - np.array(ref["ref_v"])).max() < 1e-3
    elif case in ("Local-Frame-1", "Local-Frame-2"):
        ref = section(goldens, case, "empole")
        eref = ref["ref_eng"] if "ref_eng" in ref else ref["ref_ereal"] + ref["ref_erecip"] + ref["ref_eself"]
        assert abs(r["em"] - eref) < 2e-4
        assert np.abs(r["grad"] - np.array(ref["ref_grad"]))[:18].max() < 5e-4
        assert np.abs(r["virial"] - np.array(ref["ref_v"])).max() < 5e-3
    else:
        ref = section(goldens, case, "induce")
        u1, u2 = a.uind()
        assert np.abs(u1 * DEBYE - np.array(ref["ref_ud_debye"])).max() < 1e-4
        assert np.abs(u2 * DEBYE - np.array(ref["ref_up_debye"])).max() < 1e-4
        ref = section(goldens, case, "various")
        assert abs(r["ep"] - ref["ref_eng"]) < 1e-4
        assert np.abs(r["grad"] - np.array(ref["ref_grad"]))[:18].max() < 2e-4
        assert np.abs(r["virial"] - np.array(ref["ref_v"])).max() < 1e-3
    a.close()


@pytest.mark.parametrize("precision", ["double", "mixed"])
def test_water_box_frames(precision):
    """2684-atom AMOEBA water box: the reference's two MD frames (test/ref/tinkernist.*) through the
    CUDA path -- direct and induced dipoles of every atom -- then energy/gradient vs the oracle."""
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import calc
    from oracle.amoeba_ref import Oracle, V1
    s = tg.load_system(os.path.join(GOLDEN, "water30.npz"))
    fr = np.load(os.path.join(GOLDEN, "tinkernist_frames.npz"))
    a = _amoeba(s, precision)
    for k in range(2):
        a.set_positions(fr["arc"][k])
        u1, _ = a.induce()
        d1, _ = a.udir()
        assert np.abs(d1 * DEBYE - fr["udir"][k]).max() < 1e-5
        assert np.abs(u1 * DEBYE - fr["uind"][k]).max() < 2e-5      # reference test tolerance: 1e-3
    a.set_positions(s.xyz)
    r = a.energy(calc.v1)
    o = Oracle(s)
    ro = o.energy(V1)
    u1, _ = a.uind()
    if precision == "double":
        assert abs(r["esum"] - ro["esum"]) < 1e-10 * abs(ro["esum"])
        assert _rms(r["grad"] - ro["grad"]) < 1e-7
        assert _rms(u1 - o.uind) * DEBYE < 1e-10
    else:
        assert abs(r["esum"] - ro["esum"]) < 1e-6 * abs(ro["esum"])       # north star: 1e-6 relative
        assert _rms(u1 - o.uind) * DEBYE < 1e-6                            # north star: 1e-6 D RMS
        assert _rms(r["grad"] - ro["grad"]) < 1e-5                         # north star: 1e-5 kcal/mol/A RMS
    assert r["pcg_iterations"] == o.niter
    a.close()


def test_dhfr2_properties():
    """Full-size configuration (BASELINE.json configs[0]/[1]) through size-independent properties:
    net force and net torque-free translation invariance, rebuild invariance, fixed-point determinism
    of the force sum, and double-vs-mixed agreement."""
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import calc
    s = tg.load_system(os.path.join(GOLDEN, "dhfr2.npz"))
    am = _amoeba(s, "mixed")
    r = am.energy(calc.v4)
    g = r["grad"]
    # PME forces do not conserve momentum exactly; the residual must be tiny against the force scale
    assert np.abs(g.sum(0)).max() < 1e-2 * _rms(g) * np.sqrt(s.n)
    assert 4 <= r["pcg_iterations"] <= 12
    # rigid translation by a non-lattice vector: same energy (new list, new sort, new PME phases)
    am.set_positions(np.array(s.xyz) + np.array([1.2345, -2.5, 0.777]))
    r2 = am.energy(calc.v4)
    assert abs(r2["esum"] - r["esum"]) < 2e-5 * abs(r["esum"])       # PME grid discretisation ~1e-5
    ad = _amoeba(s, "double")
    rd = ad.energy(calc.v4)
    assert abs(rd["esum"] - r["esum"]) < 1e-6 * abs(rd["esum"])
    assert _rms(rd["grad"] - g) < 1e-4
    ud, _ = ad.uind()
    am.set_positions(s.xyz)
    am.energy(calc.v4)
    um, _ = am.uind()
    assert _rms(ud - um) * DEBYE < 1e-6
    am.close()
    ad.close()


@pytest.mark.parametrize("precision", ["double", "mixed"])
def test_dhfr2_vs_oracle_fixture(precision):
    """BASELINE.json configs[0]: dhfr2 energy + gradient + virial + induced dipoles against the float64 oracle run
    on the same input (tests/golden/dhfr2_oracle.npz, made by tests/golden/make_oracle_fixtures.py: 10 CPU-minutes,
    so the GPU box only loads its result).  North-star tolerances: energy 1e-6 relative, dipoles 1e-6 D RMS;
    forces 1e-5 kcal/mol/A RMS for the mixed build, 1e-7 for the double build."""
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import calc
    fx = np.load(os.path.join(GOLDEN, "dhfr2_oracle.npz"))
    s = tg.load_system(os.path.join(GOLDEN, "dhfr2.npz"))
    a = _amoeba(s, precision)
    r = a.energy(calc.v1)
    ud, up = a.uind()
    d1, d2 = a.udir()
    tol = dict(e=1e-9, g=1e-7, u=1e-9, v=1e-7) if precision == "double" else dict(e=1e-6, g=1e-5, u=1e-6, v=2e-5)
    eref = float(fx["em"]) + float(fx["ep"])
    assert abs(r["em"] - float(fx["em"])) < tol["e"] * abs(eref)
    assert abs(r["ep"] - float(fx["ep"])) < tol["e"] * abs(eref)
    assert abs(r["esum"] - eref) < tol["e"] * abs(eref)
    assert _rms(r["grad"] - fx["grad"]) < tol["g"]
    assert _rms(ud - fx["uind"]) * DEBYE < tol["u"] and _rms(up - fx["uinp"]) * DEBYE < tol["u"]
    assert _rms(d1 - fx["udir"]) * DEBYE < tol["u"] and _rms(d2 - fx["udirp"]) * DEBYE < tol["u"]
    assert np.abs(r["virial"] - fx["virial"]).max() < tol["v"] * np.abs(fx["virial"]).max()
    assert r["pcg_iterations"] == int(fx["niter"])
    # the list is cut from float coordinates (4e-6 A at 60 A): of the 1.6 M pairs the one or two within that of 7 A may differ
    assert abs(a.stats()["npairs_m"] - int(fx["npairs"])) <= 3
    a.close()


def test_pme_convolution_native_fft():
    """The fused 64^3 FFT -> influence function -> inverse FFT kernels (fft64.cu) against numpy's FFT
    and against the cuFFT path of the same library, on a random complex grid (dhfr2 grid, 64^3)."""
    import tinker_gpu_b200 as tg
    s = tg.load_system(os.path.join(GOLDEN, "dhfr2.npz"))
    assert tuple(s.nfft) == (64, 64, 64)
    am = _amoeba(s, "mixed")
    rng = np.random.default_rng(7)
    shape = (64, 64, 64)
    g = rng.normal(size=shape) + 1j * rng.normal(size=shape)
    # influence function from the response to a unit impulse through the cuFFT path
    am.set_native_fft(False)
    imp = np.zeros(shape, dtype=np.complex128)
    imp[0, 0, 0] = 1.0
    qfac = np.fft.fftn(am.pme_convolve_grid(imp)).real / imp.size
    ref = np.fft.ifftn(qfac * np.fft.fftn(g)) * g.size          # unnormalised inverse, like cuFFT
    out_cufft = am.pme_convolve_grid(g)
    am.set_native_fft(True)
    out_native = am.pme_convolve_grid(g)
    scale = np.abs(ref).max()
    assert np.abs(out_cufft - ref).max() < 2e-5 * scale
    assert np.abs(out_native - ref).max() < 2e-5 * scale
    assert np.abs(out_native - out_cufft).max() < 5e-6 * scale
    # and the whole energy step must not care which FFT ran
    from tinker_gpu_b200.amoeba import calc
    r1 = am.energy(calc.v4)
    am.set_native_fft(False)
    r0 = am.energy(calc.v4)
    assert abs(r1["esum"] - r0["esum"]) < 2e-7 * abs(r0["esum"])
    assert _rms(r1["grad"] - r0["grad"]) < 2e-5
    assert r1["pcg_iterations"] == r0["pcg_iterations"]
    am.close()


def test_no_cpu_fallback_message():
    from tinker_gpu_b200 import amoeba
    assert os.path.isfile(amoeba.library_path("mixed")), "libapx.so must be built in-tree"


def test_step_graphs_survive_rebuilds():
    """The whole-step CUDA graphs are kept across list rebuilds (buffers did not move) and dropped when they must be:
    a context that has replayed its graphs through drifting positions, a rebuild and a box-filling translation gives
    the same energies and forces as a fresh context at each of those positions."""
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import calc
    s = tg.load_system(os.path.join(GOLDEN, "water30.npz"))
    x0 = np.array(s.xyz)
    rng = np.random.default_rng(9)
    # a rigid drift of 0.14 A per frame (physics unchanged, rebuild once the shift passes buffer/2 = 1 A) plus thermal-size
    # noise; uncorrelated drifts would stretch the bonds and amplify float round-off beyond any useful tolerance
    drift = np.array([0.11, 0.07, 0.05])
    frames = [x0 + drift * j + rng.normal(scale=0.004, size=x0.shape) for j in range(12)] + [x0 + np.array([3.1, -2.2, 7.7])]
    a = _amoeba(s, "mixed")
    got = []
    for x in frames:
        a.set_positions(x)
        got.append(a.energy(calc.v4))
    nreb = a.stats()["list_rebuilds"]
    a.close()
    assert nreb >= 2                                  # the drift passes 1 A at frame 8, then the translation
    for j in (0, 5, 11, 12):
        b = _amoeba(s, "mixed")
        b.set_positions(frames[j])
        ref = b.energy(calc.v4)
        b.close()
        # the two contexts sorted their atoms at different frames: float sums in another order
        assert abs(got[j]["esum"] - ref["esum"]) < 5e-7 * abs(ref["esum"]), j
        assert _rms(got[j]["grad"] - ref["grad"]) < 3e-5, j
        assert got[j]["pcg_iterations"] == ref["pcg_iterations"]


@pytest.mark.parametrize("precision", ["double", "mixed"])
def test_triclinic_cell(precision):
    """18 atoms, all five local-frame types, in a triclinic cell (28 x 22 x 18 A, 105/110/80 degrees) with PME: the general
    minimum image, the triclinic reciprocal vectors of the PME transforms and the frac->Cartesian field transforms
    against the oracle (whose own gradient is checked against finite differences in tests/test_oracle_golden.py).
    The reference has no electrostatics golden in a non-orthogonal cell (test/localframe2.cpp covers vdW only)."""
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import calc
    from oracle.amoeba_ref import Oracle, V1
    s = tg.load_system(os.path.join(GOLDEN, "lf_triclinic.npz"))
    a = _amoeba(s, precision)
    r = a.energy(calc.v1)
    u1, u2 = a.uind()
    o = Oracle(s)
    ro = o.energy(V1)
    tol = dict(e=1e-10, g=1e-7, v=1e-6, u=1e-10) if precision == "double" else dict(e=2e-6, g=1e-5, v=5e-4, u=5e-6)
    assert abs(r["esum"] - ro["esum"]) < tol["e"] * abs(ro["esum"])
    assert abs(r["em"] - ro["em"]) < tol["e"] * abs(ro["esum"]) and abs(r["ep"] - ro["ep"]) < tol["e"] * abs(ro["esum"])
    # forces of this 18-atom deck reach 186 kcal/mol/A: one float ulp of the largest component is 1.5e-5, so the mixed build is
    # held to the north-star 1e-5 or 1e-7 of the largest force, whichever is larger (measured: 1.25e-5)
    gtol = tol["g"] if precision == "double" else max(tol["g"], 1e-7 * float(np.abs(ro["grad"]).max()))
    assert _rms(r["grad"] - ro["grad"]) < gtol
    assert np.abs(r["virial"] - ro["virial"]).max() < tol["v"] * max(1.0, np.abs(ro["virial"]).max())
    assert np.abs(u1 - o.uind).max() * DEBYE < tol["u"] and np.abs(u2 - o.uinp).max() * DEBYE < tol["u"]
    assert r["pcg_iterations"] == o.niter
    a.close()


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_device_pointer_entry_points(dtype):
    """The _dev entry points (include/apx.h, csrc/devio.cu) -- what the adapter hands the reference's device arrays to --
    against the host-pointer entry points on the same context: same operators, no host staging, and the accumulate contract
    of the gradient / energy hand-back (SURVEY 8b "Ownership")."""
    import torch
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import calc
    s = tg.load_system(os.path.join(GOLDEN, "water30.npz"))
    a = _amoeba(s, "mixed")
    td = getattr(torch, dtype)
    tol = 2e-7 if dtype == "float32" else 1e-12
    dev = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=td, device="cuda")      # noqa: E731
    new = lambda: torch.zeros(s.n, 3, dtype=td, device="cuda")      # noqa: E731
    rel = lambda t, ref: float(np.abs(t.double().cpu().numpy() - ref).max() / np.abs(ref).max())      # noqa: E731
    # positions from the reference's separate x / y / z device arrays
    xyz = np.array(s.xyz) + 0.01
    a.set_positions_dev(dev(xyz[:, 0]), dev(xyz[:, 1]), dev(xyz[:, 2]))
    r1 = a.energy(calc.v0)
    a.set_positions(xyz if dtype == "float64" else xyz.astype(np.float32).astype(np.float64))
    r0 = a.energy(calc.v0)
    assert abs(r1["esum"] - r0["esum"]) < 2e-8 * abs(r0["esum"])      # two evaluations: float atomics on the PME grid
    a.set_positions(s.xyz)
    # operators
    f0, f1 = a.dfield()
    d0, d1 = new(), new()
    a.dfield_dev(d0, d1)
    assert rel(d0, f0) < 5e-6 and rel(d1, f1) < 5e-6      # recomputed: float atomics on the PME grid are not bit-reproducible
    rng = np.random.default_rng(3)
    ud, up = rng.normal(size=(s.n, 3)) * 0.05, rng.normal(size=(s.n, 3)) * 0.05
    if dtype == "float32":
        ud, up = ud.astype(np.float32).astype(np.float64), up.astype(np.float32).astype(np.float64)
    g0, g1 = a.ufield(ud, up)
    a.ufield_dev(dev(ud), dev(up), d0, d1)
    assert rel(d0, g0) < 5e-6 and rel(d1, g1) < 5e-6
    z0, z1 = a.sparsePrecondApply(ud, up)
    a.precond_dev(dev(ud), dev(up), d0, d1)
    assert rel(d0, z0) < tol and rel(d1, z1) < tol
    u0, u1 = new(), new()
    a.induce_dev(u0, u1, d0, d1)
    h0, h1 = a.uind()
    k0, k1 = a.udir()
    assert rel(u0, h0) < tol and rel(u1, h1) < tol and rel(d0, k0) < tol and rel(d1, k1) < tol      # copies of one solve
    # hand-back: added to what the caller's accumulators already hold
    r = a.energy(calc.v1)
    g = r["grad"]
    gx = [torch.full((s.n,), 7 << 32, dtype=torch.int64, device="cuda") for _ in range(3)]      # 7.0 in 2^32 fixed point
    a.add_gradient_dev(*gx)
    back = np.stack([(t.cpu().numpy().astype(np.float64) / 2.0 ** 32) for t in gx], axis=1) - 7.0
    assert np.abs(back - g).max() < 1e-9
    gf = [torch.full((s.n,), 7.0, dtype=td, device="cuda") for _ in range(3)]
    a.add_gradient_dev(*gf)
    back = np.stack([t.double().cpu().numpy() for t in gf], axis=1) - 7.0
    assert np.abs(back - g).max() < (2e-5 if dtype == "float32" else 1e-9)
    eb = torch.full((4,), 5 << 32, dtype=torch.int64, device="cuda")
    a.add_scalars_dev(eb, [r["esum"]])
    assert abs(eb[0].item() / 2.0 ** 32 - 5.0 - r["esum"]) < 1e-8 * abs(r["esum"]) and eb[1].item() == 5 << 32
    a.close()


def test_fixed_point_spreading_is_deterministic_and_agrees():
    """Deterministic PME spreading (apx_set_pme_fixed_point; north star: "fixed-point atomics for spreading onto the PME
    grid"): contributions summed as 2^32 fixed-point integers.  (1) the operators that consist of one spread -> FFT -> gather
    round trip plus the fixed-order row kernels -- ufield() -- must return the SAME BITS on every call, on a fresh context
    too; (2) energy, gradient and induced dipoles must agree with the float-reduction path well inside the north-star
    tolerances (the rounding step is 2.3e-10 per contribution)."""
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import calc
    s = tg.load_system(os.path.join(GOLDEN, "water30.npz"))
    a = _amoeba(s, "mixed")
    r0 = a.energy(calc.v1)
    u0, p0 = a.uind()
    a.set_pme_fixed_point(True)
    r1 = a.energy(calc.v1)
    u1, p1 = a.uind()
    assert abs(r1["esum"] - r0["esum"]) < 1e-7 * abs(r0["esum"])
    assert _rms(r1["grad"] - r0["grad"]) < 2e-6
    assert _rms(u1 - u0) * DEBYE < 1e-7
    assert r1["pcg_iterations"] == r0["pcg_iterations"]
    # ufield(): spread -> FFT -> gather plus the fixed-order row kernels.  (dfield() also runs the exclusion pass, whose few
    # float corrections per atom are added atomically: reproducible only to the last bit or two, so it is not asserted here.)
    fields = [a.ufield(u0, p0) for _ in range(3)]
    a.close()
    b = _amoeba(s, "mixed")
    b.set_pme_fixed_point(True)
    b.induce()
    fields.append(b.ufield(u0, p0))
    b.close()
    for f in fields[1:]:
        assert np.array_equal(f[0], fields[0][0]) and np.array_equal(f[1], fields[0][1])

