"""Spatial decomposition (dist.cu) against the single-GPU path on the same inputs.

The ranks are host threads sharing cuda:0 (in-process transport), so slab ownership, the halo
exchange of dipoles, the slab-decomposed PME FFT and the cross-rank reductions all execute on a
one-GPU box.  The same code runs over NCCL with one process per GPU (bench.py --gpus N)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEBYE = 4.803206802


def _rms(a):
    return float(np.sqrt((np.asarray(a) ** 2).mean()))


def _single(system, precision, vers):
    from tinker_gpu_b200.amoeba import Amoeba
    a = Amoeba(system, precision)
    r = a.energy(vers)
    r["uind"] = a.uind()[0]
    r["fields"] = a.dfield()
    a.close()
    return r


@pytest.mark.parametrize("world,precision", [(2, "mixed"), (4, "mixed"), (3, "double")])
def test_water_box_ranks_match_single_gpu(world, precision):
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import calc
    from tinker_gpu_b200.distributed import run_local_ranks
    s = tg.load_system(os.path.join(GOLDEN, "water30.npz"))
    if s.nfft[1] % world or s.nfft[2] % world:
        pytest.skip("grid not divisible")
    ref = _single(s, precision, calc.v1)

    def job(a, rank):
        r = a.energy(calc.v1)
        r["uind"] = a.uind()[0]
        r["fields"] = a.dfield()
        r["info"] = a.dist_info()
        return r

    outs = run_local_ranks(s, world, job, precision)
    owned = sum(o["info"]["a1"] - o["info"]["a0"] for o in outs)
    assert owned == s.n
    tol = dict(e=1e-9, g=1e-7, u=1e-9, v=1e-6, f=1e-9) if precision == "double" else dict(e=2e-7, g=2e-5, u=2e-7, v=2e-3, f=2e-6)
    for o in outs:      # every rank returns the complete result
        assert abs(o["esum"] - ref["esum"]) < tol["e"] * abs(ref["esum"])
        assert abs(o["em"] - ref["em"]) < tol["e"] * abs(ref["esum"])
        assert _rms(o["grad"] - ref["grad"]) < tol["g"]
        assert _rms(o["uind"] - ref["uind"]) * DEBYE < tol["u"]
        assert np.abs(o["virial"] - ref["virial"]).max() < tol["v"] * max(1.0, np.abs(ref["virial"]).max())
        assert np.abs(o["fields"][0] - ref["fields"][0]).max() < tol["f"]
        assert np.abs(o["fields"][1] - ref["fields"][1]).max() < tol["f"]
        assert o["pcg_iterations"] == ref["pcg_iterations"]


def test_dhfr2_four_ranks_and_moving_atoms():
    """dhfr2 on 4 ranks: energy/gradient/dipoles equal the single-GPU ones, also after the atoms moved
    without a list rebuild (stencils reaching into the PME halo planes) and after a rebuild."""
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba, calc
    from tinker_gpu_b200.distributed import run_local_ranks
    s = tg.load_system(os.path.join(GOLDEN, "dhfr2.npz"))
    x0 = np.array(s.xyz)
    # a smooth displacement field of up to 0.95 A (< buffer/2: same list) plus thermal-size noise.  Uncorrelated
    # 1 A kicks would squeeze bonded atoms to ~0.3 A, where the excluded-pair cancellation loses all float
    # digits and the single-GPU and decomposed summation orders differ by more than any tolerance.
    L = float(np.asarray(s.lvec)[0, 0])
    ph = 2.0 * np.pi * x0 / L
    d = 0.5 * np.stack([np.sin(ph[:, 1]), np.sin(ph[:, 2]), np.sin(ph[:, 0])], axis=1)
    d += np.random.default_rng(7).normal(scale=0.02, size=x0.shape)
    assert np.linalg.norm(d, axis=1).max() < 0.95
    x1 = x0 + d
    x2 = x0 + np.array([0.0, 0.0, 3.3])                                   # rebuild, atoms change slabs
    a = Amoeba(s, "mixed")
    refs = []
    for x in (x0, x1, x2):
        a.set_positions(x)
        r = a.energy(calc.v4)
        r["uind"] = a.uind()[0]
        refs.append(r)
    a.close()

    def job(am, rank):
        res = []
        for x in (x0, x1, x2):
            am.set_positions(x)
            r = am.energy(calc.v4)
            r["uind"] = am.uind()[0]
            r["rebuilds"] = am.stats()["list_rebuilds"]
            res.append(r)
        return res

    outs = run_local_ranks(s, 4, job, "mixed")
    for res in outs:
        assert res[1]["rebuilds"] == res[0]["rebuilds"] and res[2]["rebuilds"] == res[1]["rebuilds"] + 1
        for r, ref in zip(res, refs):
            assert abs(r["esum"] - ref["esum"]) < 3e-7 * abs(ref["esum"])
            assert _rms(r["grad"] - ref["grad"]) < 3e-5
            assert _rms(r["uind"] - ref["uind"]) * DEBYE < 3e-7
            assert r["pcg_iterations"] == ref["pcg_iterations"]


def test_operators_collective():
    """ufield / preconditioner / PME potentials with prescribed dipoles on 2 ranks."""
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba
    from tinker_gpu_b200.distributed import run_local_ranks
    s = tg.load_system(os.path.join(GOLDEN, "water30.npz"))
    n = s.n
    rng = np.random.default_rng(3)
    ud, up = rng.normal(size=(n, 3)) * 0.05, rng.normal(size=(n, 3)) * 0.05
    a = Amoeba(s, "double")
    ref = (a.ufield(ud, up), a.sparsePrecondApply(ud, up), a.pme_uind_fphi(ud, up), a.pme_mpole_fphi())
    a.close()

    def job(am, rank):
        return (am.ufield(ud, up), am.sparsePrecondApply(ud, up), am.pme_uind_fphi(ud, up), am.pme_mpole_fphi())

    for o in run_local_ranks(s, 2, job, "double"):
        for k in range(3):
            assert np.abs(o[k][0] - ref[k][0]).max() < 1e-10
            assert np.abs(o[k][1] - ref[k][1]).max() < 1e-10
        assert np.abs(o[3] - ref[3]).max() < 1e-10


@pytest.mark.parametrize("world,precision", [(2, "mixed"), (3, "double")])
def test_direct_transport_between_processes(world, precision):
    """The peer-memory transport of real multi-GPU runs (dist.cu DirectComm: CUDA IPC registered buffers, one kernel per
    exchange, flag-based all-reduce; no NCCL) with one PROCESS per rank.  On a one-GPU box the processes share the device and
    their exchange kernels alternate by time slice; the result must match the single-GPU path like the in-process ranks do."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "direct_check.py"), "--world", str(world), "--blob", "water30",
                        "--precision", precision, "--timeout", "400"], capture_output=True, text=True, timeout=450)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")]
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert len(lines) == 3 and all(ln.endswith("OK") for ln in lines), r.stdout[-2000:]
