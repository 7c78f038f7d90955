"""Induced-dipole predictors: ulspredSave / ulspredSum of the reference
(src/amoeba/induce.cpp:27-69, src/cu/upredict.cu:34-207) and the predicted start of the PCG solver
(src/cu/amoeba/pcg.cu:26-66).

CPU part: the oracle's ring bookkeeping and coefficients.  GPU part: a short trajectory of the
2684-atom water box solved step by step with the CUDA path and with the oracle, ring fills and
predicted starts included -- same iteration counts, same dipoles."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

DEBYE = 4.803206802


def _water(polpred):
    import tinker_gpu_b200 as tg
    s = tg.load_system(os.path.join(GOLDEN, "water30.npz"))
    s.polpred = polpred
    return s


def test_oracle_ring_and_coefficients():
    """Slot k of the ring holds the solution of age (nualt-1-k) mod maxualt.  GEAR's coefficients are the
    6-point polynomial extrapolation (exact for degree <= 5); ASPC's sum to one (a constant history is
    reproduced)."""
    from oracle.amoeba_ref import Oracle
    for kind, m in (("GEAR", 6), ("ASPC", 16)):
        o = Oracle(_water(kind))
        assert o.maxualt == m
        n = o.n
        rng = np.random.default_rng(5)
        coef = rng.normal(size=(6, n, 3))

        def poly(t):
            return sum(coef[d] * (0.1 * t) ** d for d in range(6))

        steps = 2 * m + 5          # the ring wraps and nualt is folded back into (maxualt, 2 maxualt]
        for t in range(steps):
            o.ulspred_save(poly(t), -poly(t))
            assert o.nualt <= 2 * m
        ud, up = o.ulspred_sum()
        assert np.allclose(ud, -up)
        if kind == "GEAR":
            assert np.abs(ud - poly(steps)).max() < 1e-9 * np.abs(poly(steps)).max()
        assert abs(sum(Oracle.ASPC) - 1.0) < 1e-5 and abs(sum(Oracle.GEAR) - 1.0) < 1e-12
        # a constant history is a fixed point of both predictors
        o2 = Oracle(_water(kind))
        for _ in range(m):
            o2.ulspred_save(coef[0], coef[1])
        ud, up = o2.ulspred_sum()
        assert np.abs(ud - coef[0]).max() < 5e-5 and np.abs(up - coef[1]).max() < 5e-5


def test_keyword_parsing(tmp_path):
    """polar-predict keyword: bare -> ASPC, explicit GEAR (tinker/source/predict.f:48-60)."""
    from tinker_gpu_b200.tinkerio import read_key
    k = read_key(None, text="polar-predict\n")
    assert k.has("POLAR-PREDICT") and not (k.get("POLAR-PREDICT") or "").split()
    k = read_key(None, text="polar-predict gear\n")
    assert k.get("POLAR-PREDICT").split()[0].upper() == "GEAR"


@pytest.mark.gpu
@pytest.mark.parametrize("kind,precision", [("GEAR", "double"), ("GEAR", "mixed"), ("ASPC", "mixed")])
def test_predicted_start_matches_oracle(kind, precision):
    from tinker_gpu_b200.amoeba import Amoeba
    from oracle.amoeba_ref import Oracle
    s = _water(kind)
    m = 6 if kind == "GEAR" else 16
    x0 = np.array(s.xyz)
    rng = np.random.default_rng(11)
    vel = rng.normal(scale=0.004, size=x0.shape)      # ~thermal displacement per 1 fs step, in Angstrom
    acc = rng.normal(scale=0.0002, size=x0.shape)
    a = Amoeba(s, precision)
    o = Oracle(s)
    assert a.upred_count() == (0, m)
    iters_pred, iters_cold = [], []
    nstep = m + 3
    for t in range(nstep):
        x = x0 + vel * t + 0.5 * acc * t * t
        a.set_positions(x)
        o.set_xyz(x)
        u1, u2 = a.induce()
        v1, v2 = o.induce()
        it = a.stats()["pcg_iterations"]
        assert it == o.niter, (t, it, o.niter)
        # mixed build: float noise of the ring entries is amplified by the extrapolation (sum |c| = 63 for GEAR),
        # both solvers then stop anywhere inside polar-eps (1e-5 D RMS residual) of the exact dipoles
        tol_max, tol_rms = (1e-9, 1e-10) if precision == "double" else (1e-5, 1e-6)
        for u, v in ((u1, v1), (u2, v2)):
            assert np.abs(u - v).max() * DEBYE < tol_max
            assert np.sqrt(((u - v) ** 2).mean()) * DEBYE < tol_rms
        (iters_pred if t >= m else iters_cold).append(it)
        assert a.upred_count()[0] == o.nualt
    # the predicted start must pay off: fewer iterations than the direct guess needed
    assert max(iters_pred) < min(iters_cold), (iters_pred, iters_cold)
    # emptying the ring brings the direct guess back
    a.upred_set("NONE")
    a.induce()
    assert abs(a.stats()["pcg_iterations"] - iters_cold[-1]) <= 1
    a.close()


@pytest.mark.gpu
def test_predictor_on_ranks():
    """Decomposed path: the history ring holds every atom on every rank, so the predicted start is the
    same as on one GPU, also across a list rebuild that moves atoms between slabs."""
    from tinker_gpu_b200.amoeba import Amoeba
    from tinker_gpu_b200.distributed import run_local_ranks
    s = _water("GEAR")
    x0 = np.array(s.xyz)
    vel = np.random.default_rng(11).normal(scale=0.004, size=x0.shape)
    frames = [x0 + vel * t for t in range(7)] + [x0 + vel * 7 + np.array([0.0, 0.0, 2.9])]

    def run(am):
        out = []
        for x in frames:
            am.set_positions(x)
            u1, _ = am.induce()
            out.append((u1, am.stats()["pcg_iterations"]))
        return out

    a = Amoeba(s, "double")
    ref = run(a)
    a.close()
    for res in run_local_ranks(s, 2, lambda am, rank: run(am), "double"):
        for (u, it), (v, jt) in zip(res, ref):
            assert it == jt
            assert np.abs(u - v).max() * DEBYE < 1e-9
