"""Integrator oracle (oracle/md_ref.py): what the RESPA splitting must satisfy independent of any force field, and the
reference's own kinetic-energy golden (test/kinetic.cpp)."""
import numpy as np

from oracle import md_ref


def _harmonic(k):
    return lambda x: k * x


def test_respa_with_no_fast_force_is_velocity_verlet():
    rng = np.random.default_rng(3)
    x, v, m = rng.normal(size=(5, 3)), rng.normal(size=(5, 3)), rng.uniform(1, 16, 5)
    zero = lambda x: np.zeros_like(x)
    a = md_ref.Integrator(x, v, m, zero, _harmonic(30.0), 0.001, nrespa=1)
    b = md_ref.Integrator(x, v, m, zero, _harmonic(30.0), 0.001, nrespa=4)
    for _ in range(20):
        a.step()
        b.step()
    assert np.abs(a.x - b.x).max() < 1e-12 and np.abs(a.v - b.v).max() < 1e-12


def test_respa_conserves_energy_and_is_time_reversible():
    rng = np.random.default_rng(4)
    x, v, m = rng.normal(size=(6, 3)), rng.normal(size=(6, 3)) * 3, rng.uniform(1, 16, 6)
    kf, ks = 400.0, 5.0
    it = md_ref.Integrator(x, v, m, _harmonic(kf), _harmonic(ks), 0.002, nrespa=8)

    def etot(o):
        return 0.5 * (kf + ks) * float((o.x ** 2).sum()) + md_ref.kinetic(o.v, m, 1)[0]
    e0 = etot(it)
    for _ in range(200):
        it.step()
    assert abs(etot(it) - e0) < 2e-3 * abs(e0)
    it.v *= -1
    for _ in range(200):
        it.step()
    assert np.abs(it.x - x).max() < 1e-9


def test_kinetic_energy_units():
    """One atom of mass 12 at 10 A/ps: 0.5 * 12 * 100 / 418.4 kcal/mol; T = 2 Ek / (nfree R)."""
    ek, t = md_ref.kinetic(np.array([[10.0, 0, 0]]), np.array([12.0]), 3)
    assert abs(ek - 600.0 / 418.4) < 1e-12
    assert abs(t - 2 * ek / (3 * 1.9872042586e-3)) < 1e-9


def test_bussi_scale_limits():
    # at the target temperature with the mean draws (r = 0, s = nfree - 1) the scale is 1 up to O(1/nfree)
    s = md_ref.bussi_scale(298.0, 0.002, 0.2, 298.0, 70671, 0.0, 70670.0)
    assert abs(s - 1.0) < 1e-6
    # a cold system is heated, a hot one cooled
    assert md_ref.bussi_scale(100.0, 0.002, 0.2, 298.0, 1000, 0.0, 999.0) > 1.0
    assert md_ref.bussi_scale(600.0, 0.002, 0.2, 298.0, 1000, 0.0, 999.0) < 1.0
