"""The reference's NVE-Verlet-ArBox test (test/verlet.cpp:38-84): 216 argon atoms, buffered 14-7 vdW only, 20 velocity-Verlet
steps of 1 fs from the restart file; the potential and kinetic energy after every step against the reference's literals with
its margin (1e-4 kcal/mol).  Pins the integrator oracle (oracle/md_ref.py) and the vdW oracle (oracle/vdw_ref.py) -- the two
things the GPU integrator and the GPU vdW term are held to -- to a trajectory of the reference.
Fixtures: tests/golden/make_arbox_verlet_golden.py."""
import os

import numpy as np

from conftest import GOLDEN


def test_oracles_reproduce_the_nve_verlet_arbox_trajectory():
    import tinker_gpu_b200 as tg
    from oracle.md_ref import Integrator, kinetic
    from oracle.vdw_ref import VdwOracle
    s = tg.load_system(os.path.join(GOLDEN, "arbox.npz"))
    z = np.load(os.path.join(GOLDEN, "arbox_verlet.npz"))
    assert s.n == 216 and s.vdw is not None
    o = VdwOracle(s)
    last = {}

    def slow(x):
        o.set_xyz(x)
        r = o.ehal()
        last["e"] = r["ev"]
        return r["grad"]

    it = Integrator(z["xyz"], z["vel"], s.mass, lambda x: np.zeros_like(x), slow, float(z["dt_ps"]), 1)
    eps = float(z["eps"])
    for i in range(int(z["nsteps_checked"])):
        it.step()
        ek = kinetic(it.v, s.mass, 3 * s.n)
        ek = ek[0] if isinstance(ek, tuple) else ek
        assert abs(last["e"] - z["arbox_pot"][i]) < eps, (i, last["e"], z["arbox_pot"][i])
        assert abs(float(ek) - z["arbox_kin"][i]) < eps, (i, float(ek), z["arbox_kin"][i])
    # energy conservation over the run, as the literals themselves show (pot + kin constant to 1e-5)
    assert abs((last["e"] + float(ek)) - (z["arbox_pot"][0] + z["arbox_kin"][0])) < 1e-4
