"""The stored-tensor real-space operator (csrc/tlist.cu: one real4 {B1, sqrt|B2| R} per directed pair, written by the first
operator application of an induce() and streamed by every later one) against the row operator that recomputes the pair
geometry each time (csrc/field.cu, APX_TLIST=0): same pairs and the same B1/B2 arithmetic -- the fields agree to float
rounding, iteration counts are identical, induced dipoles, energies and forces agree to the tolerances of the path; across
list rebuilds and position changes (the tensors must follow the positions), with the per-pair Thole table (PolPair deck), and
in the double build."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
DEBYE = 4.803206802


def _ctx(system, tlist, precision="mixed", mode=None, **kw):
    from tinker_gpu_b200.amoeba import Amoeba
    keys = ("APX_TLIST", "APX_TL_MODE")
    old = {k: os.environ.get(k) for k in keys}
    os.environ["APX_TLIST"] = "1" if tlist else "0"
    if mode is not None:
        os.environ["APX_TL_MODE"] = str(mode)
    try:
        return Amoeba(system, precision, device=0, **kw)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("blob,precision", [("water30.npz", "mixed"), ("dhfr2.npz", "mixed"), ("water30.npz", "double")])
def test_tlist_operator_matches_row_operator(blob, precision):
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import calc
    s = tg.load_system(os.path.join(GOLDEN, blob))
    rng = np.random.default_rng(11)
    ud, up = rng.normal(size=(s.n, 3)) * 0.05, rng.normal(size=(s.n, 3)) * 0.05
    out = {}
    for tl in (True, False):
        a = _ctx(s, tl, precision)
        f = a.ufield(ud, up)
        r = a.energy(calc.v1)
        u = a.uind()[0]
        # small move, no rebuild: the tensors of the old positions must not survive
        xyz1 = np.array(s.xyz) + np.random.default_rng(3).uniform(-0.05, 0.05, (s.n, 3))
        a.set_positions(xyz1)
        f1 = a.ufield(ud, up)
        r1 = a.energy(calc.v4)
        # move every atom by up to 1.5 A: a list rebuild
        xyz2 = np.array(s.xyz) + np.random.default_rng(5).uniform(-0.02, 0.02, (s.n, 3)) + np.array([1.3, -0.9, 0.7])
        a.set_positions(xyz2)
        r2 = a.energy(calc.v4)
        r3 = a.energy(calc.v4)      # replayed graphs at unchanged positions
        out[tl] = (f, r, u, f1, r1, r2, r3, a.stats()["list_rebuilds"])
        a.close()
    (f, r, u, f1, r1, r2, r3, nb), (g, q, v, g1, q1, q2, q3, nb0) = out[True], out[False]
    ftol = 2e-6 if precision == "mixed" else 1e-12
    etol = 2e-8 if precision == "mixed" else 1e-12
    scale = np.abs(g[0]).max()
    assert np.abs(f[0] - g[0]).max() < ftol * scale and np.abs(f[1] - g[1]).max() < ftol * scale
    assert np.abs(f1[0] - g1[0]).max() < ftol * scale and np.abs(f1[1] - g1[1]).max() < ftol * scale
    assert np.abs(f1[0] - f[0]).max() > 100 * ftol * scale      # the move did change the field
    for x, y in ((r, q), (r1, q1), (r2, q2), (r3, q3)):
        assert abs(x["esum"] - y["esum"]) < etol * abs(y["esum"])
        assert x["pcg_iterations"] == y["pcg_iterations"]
    assert np.sqrt(((u - v) ** 2).mean()) * DEBYE < (2e-7 if precision == "mixed" else 1e-12)
    assert np.sqrt(((r["grad"] - q["grad"]) ** 2).mean()) < (2e-6 if precision == "mixed" else 1e-10)
    assert np.sqrt(((r2["grad"] - q2["grad"]) ** 2).mean()) < (2e-6 if precision == "mixed" else 1e-10)
    assert nb >= 2 and nb0 >= 2


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_tlist_kernel_variants_agree(mode):
    """APX_TL_MODE: 8 or 16 lanes per atom, with and without the forced memory-level parallelism -- the same sums in another
    order."""
    import tinker_gpu_b200 as tg
    s = tg.load_system(os.path.join(GOLDEN, "water30.npz"))
    rng = np.random.default_rng(2)
    ud, up = rng.normal(size=(s.n, 3)) * 0.05, rng.normal(size=(s.n, 3)) * 0.05
    a = _ctx(s, False)
    g = a.ufield(ud, up)
    a.close()
    import subprocess
    import sys
    import json
    # the mode is latched per process (static): run the variant in a child
    code = ("import os,sys,json,numpy as np;sys.path.insert(0,%r);import tinker_gpu_b200 as tg;from tinker_gpu_b200.amoeba import Amoeba;"
            "s=tg.load_system(%r);rng=np.random.default_rng(2);ud,up=rng.normal(size=(s.n,3))*0.05,rng.normal(size=(s.n,3))*0.05;"
            "a=Amoeba(s,'mixed',device=0);f=a.ufield(ud,up);np.save(sys.argv[1],np.stack(f));a.close()"
            % (ROOT, os.path.join(GOLDEN, "water30.npz")))
    out = "/tmp/tl_mode_%d.npy" % mode
    env = dict(os.environ, APX_TLIST="1", APX_TL_MODE=str(mode))
    r = subprocess.run([sys.executable, "-c", code, out], env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-800:]
    f = np.load(out)
    scale = np.abs(g[0]).max()
    assert np.abs(f[0] - g[0]).max() < 2e-6 * scale and np.abs(f[1] - g[1]).max() < 2e-6 * scale


def test_tlist_with_thole_table():
    """PolPair deck (polpair record: per-pair Thole widths from the thlval table): tensors built by the TABLE variant."""
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import calc
    path = os.path.join(GOLDEN, "polpair_ewald.npz")
    if not os.path.isfile(path):
        pytest.skip("no polpair blob")
    s = tg.load_system(path)
    res = {}
    for tl in (True, False):
        a = _ctx(s, tl)
        res[tl] = a.energy(calc.v1)
        a.close()
    assert abs(res[True]["esum"] - res[False]["esum"]) < 1e-6 * abs(res[False]["esum"])
    assert res[True]["pcg_iterations"] == res[False]["pcg_iterations"]


def test_deferred_solver_batch_and_retry():
    """energy() enqueues the solver's first batch of iterations and the energy epilogue without waiting for the convergence flag
    (csrc/pcg.cu, csrc/mplar.cu: energy_once).  APX_PCG_FIRST_BATCH=3 makes that batch too short on purpose: the evaluation must
    notice, repeat itself with a solver that waits for its batches, and return the same energies, forces and iteration count."""
    import subprocess
    import sys
    import json
    code = ("import os,sys,json,numpy as np;sys.path.insert(0,%r);import tinker_gpu_b200 as tg;from tinker_gpu_b200.amoeba import Amoeba,calc;"
            "s=tg.load_system(%r);a=Amoeba(s,'mixed',device=0,vdw=False);out=[];\n"
            "for j in range(4):\n"
            "    a.set_positions(np.array(s.xyz)+0.01*j)\n"
            "    r=a.energy(calc.v1)\n"
            "    out.append(dict(esum=r['esum'],it=r['pcg_iterations'],g=float(np.abs(r['grad']).sum()),v=float(np.abs(r['virial']).sum())))\n"
            "st=a.stats();a.close();print(json.dumps(dict(out=out,retries=st['energy_retries'])))"
            % (ROOT, os.path.join(GOLDEN, "water30.npz")))
    res = {}
    for forced in ("0", "3"):
        env = dict(os.environ, APX_PCG_FIRST_BATCH=forced, APX_LOOP="0")
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-800:]
        res[forced] = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["0"]["retries"] == 0
    assert res["3"]["retries"] >= 2      # every evaluation after the first (eager) one
    for x, y in zip(res["0"]["out"], res["3"]["out"]):
        assert x["it"] == y["it"]
        assert abs(x["esum"] - y["esum"]) < 2e-8 * abs(x["esum"])
        assert abs(x["g"] - y["g"]) < 1e-6 * x["g"] and abs(x["v"] - y["v"]) < 1e-6 * x["v"]


def test_while_node_loop_matches_batches():
    """APX_LOOP=1: the PCG iterations as the body of a conditional WHILE graph node that the preconditioner kernel leaves from
    the device -- same iteration counts and results as the default iteration batches."""
    import subprocess
    import sys
    import json
    code = ("import os,sys,json,numpy as np;sys.path.insert(0,%r);import tinker_gpu_b200 as tg;from tinker_gpu_b200.amoeba import Amoeba,calc;"
            "s=tg.load_system(%r);a=Amoeba(s,'mixed',device=0,vdw=False);out=[];\n"
            "for j in range(4):\n"
            "    a.set_positions(np.array(s.xyz)+0.01*j)\n"
            "    r=a.energy(calc.v1)\n"
            "    out.append(dict(esum=r['esum'],it=r['pcg_iterations'],g=float(np.abs(r['grad']).sum())))\n"
            "a.close();print(json.dumps(out))"
            % (ROOT, os.path.join(GOLDEN, "water30.npz")))
    res = {}
    for loop in ("0", "1"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, APX_LOOP=loop), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-800:]
        res[loop] = json.loads(r.stdout.strip().splitlines()[-1])
    for x, y in zip(res["0"], res["1"]):
        assert x["it"] == y["it"]
        assert abs(x["esum"] - y["esum"]) < 2e-8 * abs(x["esum"]) and abs(x["g"] - y["g"]) < 1e-6 * x["g"]
