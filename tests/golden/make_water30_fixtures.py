#!/usr/bin/env python
"""Float64 oracle on the 2684-atom water cell (test/file/tinkernist/water30.xyz, blob water30.npz) at the two polar-eps values
the replicated BASELINE configurations use: 1e-5 (configs[3], the ~1 M-atom box = cell x8x8x6) and 1e-8 (configs[2], the
~96 k-atom box = cell x3x3x4).  An UN-jittered replica on the commensurate PME grid is the same periodic system, so
E = n_cells x E_cell, forces / dipoles repeat and the PCG iteration count is the cell's: tests/test_zgpu_4_replicas.py.
Writes tests/golden/water30_oracle_eps{5,8}.npz (CPU, a few minutes)."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import tinker_gpu_b200 as tg  # noqa: E402
from oracle.amoeba_ref import Oracle, V1  # noqa: E402

for tag, eps in (("eps5", 1e-5), ("eps8", 1e-8)):
    s = tg.load_system(os.path.join(HERE, "water30.npz"))
    s.poleps = eps
    o = Oracle(s)
    t0 = time.time()
    r = o.energy(V1)
    print(f"water30 oracle energy(V1) poleps {eps:g}: {time.time() - t0:.0f} s  iters {o.niter}  em {r['em']:.8f}  ep {r['ep']:.8f}", flush=True)
    np.savez_compressed(os.path.join(HERE, f"water30_oracle_{tag}.npz"), em=r["em"], ep=r["ep"], grad=r["grad"], virial=r["virial"],
                        uind=o.uind, uinp=o.uinp, niter=o.niter, npairs=o.pairs(s.ewald_cutoff)[0].shape[0], poleps=eps,
                        nfft=np.array(s.nfft))
