"""The staged real-space operator (csrc/staged.cu: 64-atom groups, neighbour records staged in shared memory by cp.async.bulk
copies behind an mbarrier, 16-bit slot rows) against the row operator it replaces (csrc/field.cu) and against the oracle
fixture: same pairs, same math, another summation order -- the fields agree to float rounding, induced dipoles, energies and
forces to the tolerances of the path; also across list rebuilds and with a shared-memory capacity so small that most records
come through the global-memory fall-back."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEBYE = 4.803206802


def _ctx(system, staged, cap=None, **kw):
    from tinker_gpu_b200.amoeba import Amoeba
    old = {k: os.environ.get(k) for k in ("APX_STAGED", "APX_STAGED_CAP")}
    os.environ["APX_STAGED"] = "1" if staged else "0"
    if cap is not None:
        os.environ["APX_STAGED_CAP"] = str(cap)
    try:
        return Amoeba(system, "mixed", device=0, **kw)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("blob,cap", [("water30.npz", None), ("dhfr2.npz", None), ("dhfr2.npz", 8)])
def test_staged_operator_matches_row_operator(blob, cap):
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import calc
    s = tg.load_system(os.path.join(GOLDEN, blob))
    rng = np.random.default_rng(11)
    ud, up = rng.normal(size=(s.n, 3)) * 0.05, rng.normal(size=(s.n, 3)) * 0.05
    out = {}
    for staged in (True, False):
        a = _ctx(s, staged, cap)
        f = a.ufield(ud, up)
        r = a.energy(calc.v1)
        u = a.uind()[0]
        # move every atom by up to 1.5 A: a list rebuild, new groups and slots
        xyz2 = np.array(s.xyz) + np.random.default_rng(5).uniform(-0.02, 0.02, (s.n, 3)) + np.array([1.3, -0.9, 0.7])
        a.set_positions(xyz2)
        r2 = a.energy(calc.v4)
        out[staged] = (f, r, u, r2, a.stats()["list_rebuilds"])
        a.close()
    (f1, r1, u1, q1, nb1), (f0, r0, u0, q0, nb0) = out[True], out[False]
    scale = np.abs(f0[0]).max()
    assert np.abs(f1[0] - f0[0]).max() < 2e-6 * scale and np.abs(f1[1] - f0[1]).max() < 2e-6 * scale
    assert abs(r1["esum"] - r0["esum"]) < 2e-8 * abs(r0["esum"])
    assert r1["pcg_iterations"] == r0["pcg_iterations"]
    assert np.sqrt(((u1 - u0) ** 2).mean()) * DEBYE < 2e-7
    assert np.sqrt(((r1["grad"] - r0["grad"]) ** 2).mean()) < 2e-6
    assert nb1 >= 2 and nb0 >= 2
    assert abs(q1["esum"] - q0["esum"]) < 2e-8 * abs(q0["esum"])
    assert np.sqrt(((q1["grad"] - q0["grad"]) ** 2).mean()) < 2e-6


def test_staged_operator_in_dynamics():
    """20 r-RESPA steps of the water box with the staged operator and with the row operator from the same start: the
    trajectories stay together (float summation order is the only difference) and lists are rebuilt on the way."""
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.drivers import maxwell_velocities
    s = tg.load_system(os.path.join(GOLDEN, "water30.npz"))
    vel = maxwell_velocities(s.mass, 298.0, 3)
    res = {}
    for staged in (True, False):
        a = _ctx(s, staged, vdw=True, valence=True)
        a.md_init(s.mass, vel, dt=0.002, nrespa=4, thermostat=None)
        m = a.md_steps(20)
        res[staged] = (m.epot, m.ekin, a.md_state()[0])
        a.close()
    assert abs(res[True][0] - res[False][0]) < 2e-3 and abs(res[True][1] - res[False][1]) < 2e-3
    assert np.abs(res[True][2] - res[False][2]).max() < 1e-5
