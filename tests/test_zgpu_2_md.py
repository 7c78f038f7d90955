"""Device integrator (csrc/md.cu) through the C ABI: trajectories against the numpy integrator oracle driven by the
force oracles (small system), energy conservation and the Bussi thermostat on the water box, the DYNAMIC driver."""
import io
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _water10():
    import tinker_gpu_b200 as tg
    return tg.load_system(os.path.join(GOLDEN, "val_water10.npz"))


@pytest.mark.parametrize("nrespa", [1, 4])
def test_trajectory_matches_oracle(nrespa):
    """30 atoms, 6 steps of 1 fs: positions and velocities of the double build against oracle/md_ref.py stepping with
    oracle forces (valence fast; electrostatics + vdW slow).  Differences come from the induced-dipole convergence
    (polar-eps) only."""
    from tinker_gpu_b200.amoeba import Amoeba
    from oracle import md_ref, valence_ref
    from oracle.amoeba_ref import Oracle, V4
    from oracle.vdw_ref import VdwOracle
    s = _water10()
    s.poleps = 1e-10
    rng = np.random.default_rng(11)
    vel = rng.normal(size=(s.n, 3)) * 3.0
    eo, vo = Oracle(s), VdwOracle(s)

    def slow(x):
        eo.set_xyz(x)
        vo.set_xyz(x)
        return eo.energy(V4)["grad"] + vo.ehal()["grad"]

    def fast(x):
        return valence_ref.valence(x, s.valence)["grad"]
    ref = md_ref.Integrator(s.xyz, vel, s.mass, fast, slow, 0.001, nrespa)
    a = Amoeba(s, "double", vdw=True, valence=True)
    a.md_init(s.mass, vel, dt=0.001, nrespa=nrespa)
    for _ in range(6):
        ref.step()
    r = a.md_steps(6)
    x, v = a.md_state()
    assert r.total_steps == 6
    assert np.abs(x - ref.x).max() < 2e-7
    assert np.abs(v - ref.v).max() < 2e-4
    ek, temp = md_ref.kinetic(ref.v, s.mass, 3 * s.n - 3)
    assert abs(r.ekin - ek) < 1e-5 * ek and abs(r.temp - temp) < 1e-5 * temp
    a.close()


def test_nve_energy_conservation_water_box():
    """2684-atom AMOEBA water box, RESPA 2 fs / 4 inner steps from 298 K Maxwell velocities: total energy drift over
    100 steps below 0.02 % of the kinetic energy per step window (polar-eps 1e-6)."""
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba
    from tinker_gpu_b200.drivers import maxwell_velocities
    s = tg.load_system(os.path.join(GOLDEN, "water30.npz"))
    vel = maxwell_velocities(s.mass, 298.0, 7)
    a = Amoeba(s, "mixed", vdw=True, valence=True)
    a.md_init(s.mass, vel, dt=0.002, nrespa=4)
    e = []
    for _ in range(10):
        r = a.md_steps(10)
        e.append(r.epot + r.ekin)
    e = np.array(e)
    assert r.temp > 100 and r.temp < 500
    assert np.abs(e - e[0]).max() < 0.02 * r.ekin
    a.close()


def test_bussi_thermostat_drives_temperature():
    """Starting at rest... at 100 K, tau 0.05 ps: after 0.4 ps the kinetic temperature is within 15 % of 298 K and the
    scale factor of every step stayed positive and near 1."""
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba
    from tinker_gpu_b200.drivers import maxwell_velocities
    s = tg.load_system(os.path.join(GOLDEN, "water30.npz"))
    vel = maxwell_velocities(s.mass, 100.0, 3)
    a = Amoeba(s, "mixed", vdw=True, valence=True)
    a.md_init(s.mass, vel, dt=0.002, nrespa=4, thermostat="BUSSI", kelvin=298.0, tautemp=0.05, seed=42)
    temps = []
    for _ in range(20):
        r = a.md_steps(10)
        temps.append(r.temp)
        assert 0.8 < r.last_scale < 1.3
    assert temps[0] < 290
    assert abs(np.mean(temps[-5:]) - 298.0) < 0.15 * 298.0
    a.close()


def test_dynamic_driver_writes_archive_and_restart(tmp_path):
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.drivers import dynamic
    from tinker_gpu_b200.tinkerio import read_dyn
    s = _water10()
    out = io.StringIO()
    base = str(tmp_path / "w10")
    res = dynamic(s, 20, dt_fs=1.0, dtsave_ps=0.01, mode=2, kelvin=298.0, integrator="RESPA", basename=base, out=out)
    text = out.getvalue()
    assert text.count("Instantaneous Values for Frame Saved at") == 2
    assert "Performance:  ns/day" in text and res["nrespa"] == 2
    d = read_dyn(base + ".dyn")
    assert np.abs(d["xyz"] - res["xyz"]).max() < 1e-12 and np.abs(d["vel"] - res["vel"]).max() < 1e-12
    frames = open(base + ".arc").read().splitlines()
    assert len(frames) == 2 * (s.n + 2)
    # restart continues from the .dyn file
    res2 = dynamic(s, 10, dt_fs=1.0, dtsave_ps=0.01, mode=2, kelvin=298.0, integrator="RESPA", basename=base, out=io.StringIO())
    assert np.abs(res2["xyz"] - res["xyz"]).max() > 0
