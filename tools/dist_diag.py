"""GPU diagnostic: where do the decomposed (local-rank) results differ from the single-GPU ones?"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tinker_gpu_b200 as tg                                    # noqa: E402
from tinker_gpu_b200.amoeba import Amoeba, calc                 # noqa: E402
from tinker_gpu_b200.distributed import run_local_ranks         # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "dhfr2"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 4
prec = sys.argv[3] if len(sys.argv) > 3 else "mixed"
s = tg.load_system(os.path.join(ROOT, "tests", "golden", name + ".npz"))
x0 = np.array(s.xyz)
rng = np.random.default_rng(7)
d = rng.normal(scale=0.3, size=x0.shape)
d *= np.minimum(1.0, 0.95 / np.linalg.norm(d, axis=1))[:, None]
L0 = float(np.asarray(s.lvec)[0, 0])
ph = 2.0 * np.pi * x0 / L0
ds = 0.5 * np.stack([np.sin(ph[:, 1]), np.sin(ph[:, 2]), np.sin(ph[:, 0])], axis=1) + rng.normal(scale=0.02, size=x0.shape)
cases = {"x0": x0, "x1_kicks": x0 + d, "x1_smooth": x0 + ds, "x2_shift": x0 + np.array([0.0, 0.0, 3.3])}
a = Amoeba(s, prec)
refs = {}
for k, x in cases.items():
    a.set_positions(x)
    for vers, tag in ((calc.v4, "v4"),):
        r = a.energy(vers)
        r["uind"] = a.uind()[0]
        r["em_only"] = a.empole(calc.v4)
        r["ep_only"] = a.epolar(calc.v4)
        refs[k] = r
a.close()


def job(am, rank):
    out = {}
    for k, x in cases.items():
        am.set_positions(x)
        r = am.energy(calc.v4)
        r["uind"] = am.uind()[0]
        r["em_only"] = am.empole(calc.v4)
        r["ep_only"] = am.epolar(calc.v4)
        r["info"] = am.dist_info()
        out[k] = r
    return out


outs = run_local_ranks(s, world, job, prec)
L = float(np.asarray(s.lvec)[2, 2])
for k in cases:
    ref = refs[k]
    o = outs[0][k]
    for tag, g, gr in (("total", o["grad"], ref["grad"]), ("empole", o["em_only"]["grad"], ref["em_only"]["grad"]),
                       ("epolar", o["ep_only"]["grad"], ref["ep_only"]["grad"])):
        dg = np.abs(g - gr).max(axis=1)
        bad = np.argsort(-dg)[:12]
        print(f"[{name} w={world} {prec}] {k} {tag}: E {o['esum']:.6f} vs {ref['esum']:.6f}  grad rms diff {np.sqrt(((g-gr)**2).mean()):.3e} "
              f"max {dg.max():.3e} n(>1e-3)={(dg > 1e-3).sum()} sum(g-gr)={np.abs((g-gr).sum(0)).max():.3e}")
        if dg.max() > 1e-3:
            for i in bad:
                z = cases[k][i, 2] % L
                print(f"    atom {i:6d} z/L*world={z / L * world:7.3f} frame={list(np.asarray(s.zaxis)[i])} polarity={s.polarity[i]:.3f} dg={dg[i]:.3e}"
                      f" g={g[i]} ref={gr[i]}")
    du = np.abs(o["uind"] - ref["uind"]).max()
    print(f"    uind max diff {du:.3e}; iters {o['pcg_iterations']} vs {ref['pcg_iterations']}; info {[oo[k]['info'] for oo in outs]}")
