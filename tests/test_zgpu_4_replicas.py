"""Parity of the LARGE BASELINE configurations (configs[2..4]) through exact replicas.

The reference holds no golden at these sizes and the float64 oracle cannot run them, but an UN-jittered replica of a periodic
cell on the commensurate PME grid IS the cell: every image atom has the same environment, so on the n-cell box
    E = n_cells x E_cell,   forces, induced dipoles and the virial per cell repeat,   the PCG iteration count is the cell's,
and the cell is pinned to the float64 oracle (tests/golden/water30_oracle_eps{5,8}.npz from make_water30_fixtures.py,
dhfr2_oracle.npz from make_oracle_fixtures.py).  Held to the north-star tolerances: energy 1e-6 relative, forces 1e-5
kcal/mol/A RMS, dipoles 1e-6 D RMS.

  configs[2]  water30 x3x3x4 =    96 624 atoms, PME 108x108x144, polar-eps 1e-8
  configs[3]  water30 x8x8x6 = 1 030 656 atoms, PME 288x288x216, polar-eps 1e-5
  configs[4]  dhfr2   x3x3x2 =   424 044 atoms, PME 192x192x128 (the grid commensurate with the cell's 64^3)
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEBYE = 4.803206802

CASES = {
    "water96k": ("water30.npz", (3, 3, 4), 1e-8, "water30_oracle_eps8.npz"),
    "water1m": ("water30.npz", (8, 8, 6), 1e-5, "water30_oracle_eps5.npz"),
    "dhfr424k": ("dhfr2.npz", (3, 3, 2), 1e-5, "dhfr2_oracle.npz"),
}


def _rms(a):
    return float(np.sqrt((np.asarray(a) ** 2).mean()))


@pytest.mark.parametrize("name", list(CASES))
def test_replicated_box_reproduces_its_cell(name):
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba, calc
    blob, reps, eps, fixture = CASES[name]
    cell = tg.load_system(os.path.join(GOLDEN, blob))
    fx = np.load(os.path.join(GOLDEN, fixture))
    m = reps[0] * reps[1] * reps[2]
    s = tg.replicate(cell, reps, jitter=0.0, keep_bonds=False)
    s.nfft = tuple(int(g * r) for g, r in zip(cell.nfft, reps))      # commensurate with the cell's grid
    s.poleps = eps
    s.vdw = s.valence = None
    n0 = cell.n
    assert s.n == m * n0
    a = Amoeba(s, "mixed", device=0)
    r = a.energy(calc.v1)
    ud, up = a.uind()
    st = a.stats()
    a.close()
    e_cell = float(fx["em"]) + float(fx["ep"])
    g = r["grad"].reshape(m, n0, 3)
    u = ud.reshape(m, n0, 3)
    res = dict(n=s.n, cells=m, esum_rel=abs(r["esum"] - m * e_cell) / abs(m * e_cell), em_rel=abs(r["em"] - m * float(fx["em"])) / abs(m * e_cell),
               ep_rel=abs(r["ep"] - m * float(fx["ep"])) / abs(m * e_cell), grad_rms=_rms(g - fx["grad"][None]),
               uind_rms_debye=_rms(u - fx["uind"][None]) * DEBYE, uinp_rms_debye=_rms(up.reshape(m, n0, 3) - fx["uinp"][None]) * DEBYE,
               virial_rel=float(np.abs(r["virial"] - m * fx["virial"]).max() / np.abs(m * fx["virial"]).max()),
               image_spread_grad=float(np.abs(g - g.mean(0)[None]).max()), iters=int(r["pcg_iterations"]), iters_cell=int(fx["niter"]),
               pairs=int(st["npairs_m"]), pairs_cell=int(fx["npairs"]))
    print(name, res)
    assert res["esum_rel"] < 1e-6 and res["em_rel"] < 1e-6 and res["ep_rel"] < 1e-6
    assert res["grad_rms"] < 1e-5
    assert res["uind_rms_debye"] < 1e-6 and res["uinp_rms_debye"] < 1e-6
    assert res["virial_rel"] < 2e-5
    # the list is cut from float coordinates: of the ~3e7 pairs per A of separation around 7 A (1 M atoms) the few inside the
    # coordinates' rounding (4e-6 A at 60 A, 1.5e-5 A at 240 A) may fall on the other side in another image
    assert abs(res["pairs"] - m * res["pairs_cell"]) <= 2e-5 * m * res["pairs_cell"]
    assert res["iters"] == res["iters_cell"]
