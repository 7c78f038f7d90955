"""GPU diagnostic: per-step time, PCG iterations and list rebuilds of the e2e loop of bench.py (rigid drift)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tinker_gpu_b200 as tg
from tinker_gpu_b200.amoeba import Amoeba, calc
s = tg.load_system(os.path.join(ROOT, "tests", "golden", "dhfr2.npz"))
a = Amoeba(s, "mixed")
x0 = np.array(s.xyz)
vel = np.array([0.06, 0.04, 0.035])
rng = np.random.default_rng(1)
frames = [x0 + vel * j + rng.normal(scale=0.002, size=x0.shape) for j in range(24)]
for _ in range(5):
    a.lib.apx_energy(a.ctx, calc.v4, None)
for j, x in enumerate(frames):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a.set_positions(x)
    t1 = time.perf_counter()
    a.lib.apx_energy(a.ctx, calc.v4, None)
    t2 = time.perf_counter()
    g = a.gradient()
    t3 = time.perf_counter()
    st = a.stats()
    print(f"frame {j:2d} set_pos {1e3*(t1-t0):6.3f} energy {1e3*(t2-t1):6.3f} grad {1e3*(t3-t2):6.3f} ms  iters {st['pcg_iterations']} rebuilds {st['list_rebuilds']} "
          f"ms_list {st['ms_list']:.3f} ms_induce {st['ms_induce']:.3f} ms_energy {st['ms_energy']:.3f}")
