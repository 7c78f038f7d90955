#!/usr/bin/env python
"""Developer check on a GPU box: every operator of the CUDA path against the float64 oracle,
printing max-abs differences (not a test; tests/test_gpu_parity.py asserts the same things)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tinker_gpu_b200 as tg  # noqa: E402
from tinker_gpu_b200.amoeba import Amoeba, calc  # noqa: E402
from oracle.amoeba_ref import Oracle, V1, V3  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def rms(a):
    return float(np.sqrt((np.asarray(a) ** 2).mean()))


def check(name, path, precision, big=False):
    s = tg.load_system(path)
    t0 = time.time()
    a = Amoeba(s, precision)
    o = Oracle(s)
    o.rotpole()
    print(f"== {name} [{precision}] n={s.n} ewald={s.use_ewald} mpole={s.use_mpole} polar={s.use_polar} create {time.time()-t0:.2f}s stats={a.stats()}")
    print("  rpole     ", np.abs(a.rpole() - o.rpole).max())
    if s.use_ewald:
        o.dfield()
        print("  fphi mpole", np.abs(a.pme_mpole_fphi() - o._recip_m["fphi"]).max(), "scale", np.abs(o._recip_m["fphi"]).max())
    if s.use_polar:
        fd, fp = a.dfield()
        od, op = o.dfield()
        print("  dfield    ", np.abs(fd - od).max(), np.abs(fp - op).max(), "scale", np.abs(od).max())
        rng = np.random.default_rng(0)
        ud, up = rng.normal(size=(s.n, 3)) * 0.05, rng.normal(size=(s.n, 3)) * 0.05
        f1, f2 = a.ufield(ud, up)
        o1, o2 = o.ufield(ud, up)
        print("  ufield    ", np.abs(f1 - o1).max(), np.abs(f2 - o2).max(), "scale", np.abs(o1).max())
        z1, z2 = a.sparsePrecondApply(ud, up)
        p1, p2 = o.precond(ud, up)
        print("  precond   ", np.abs(z1 - p1).max(), np.abs(z2 - p2).max(), "scale", np.abs(p1).max())
        t0 = time.time()
        u1, u2 = a.induce()
        t1 = time.time() - t0
        v1, v2 = o.induce()
        st = a.stats()
        print(f"  induce     iters gpu {st['pcg_iterations']} oracle {o.niter}  rms diff D {rms(u1 - v1) * 4.803206802:.3e} max {np.abs(u1 - v1).max() * 4.8032:.3e}"
              f"  ms_induce {st['ms_induce']:.3f} wall {t1*1e3:.1f}ms")
    t0 = time.time()
    r = a.energy(calc.v1)
    tg_ = time.time() - t0
    t0 = time.time()
    ro = o.energy(V1)
    print(f"  energy     em {r['em']:.6f} vs {ro['em']:.6f}   ep {r['ep']:.6f} vs {ro['ep']:.6f}  rel {abs(r['esum'] - ro['esum']) / max(1e-30, abs(ro['esum'])):.2e}")
    g = r["grad"]
    print(f"  grad       rms diff {rms(g - ro['grad']):.3e} max {np.abs(g - ro['grad']).max():.3e} scale rms {rms(ro['grad']):.3f}")
    print(f"  virial     max diff {np.abs(r['virial'] - ro['virial']).max():.3e} scale {np.abs(ro['virial']).max():.2f}")
    print(f"  time       gpu energy(v1) {tg_*1e3:.1f} ms (ms_energy {a.stats()['ms_energy']:.3f}), oracle {time.time() - t0:.1f} s")
    r3 = a.energy(calc.v3)
    ro3 = o.energy(V3)
    print(f"  analyz     em {r3['em']:.6f}/{ro3['em']:.6f} ep(pair) {r3['ep']:.6f}/{ro3['ep']:.6f} nem {r3['nem']} nep {r3['nep']}")
    a.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["small"]
    precs = ["double", "mixed"]
    if "small" in which:
        for c in ["lf_local_frame_1", "lf_local_frame_2", "lf_local_frame_3", "lf_local_frame_4", "lf_local_frame3_1",
                  "lf_local_frame3_2"]:
            for p in precs:
                try:
                    check(c, os.path.join(G, c + ".npz"), p)
                except Exception as e:  # keep going: this is a diagnostic script
                    print("  FAILED:", repr(e))
    if "water" in which:
        for p in precs:
            check("water30", os.path.join(G, "water30.npz"), p)
    if "dhfr" in which:
        for p in precs:
            check("dhfr2", os.path.join(G, "dhfr2.npz"), p)
