#!/usr/bin/env python
"""Run the float64 oracle HERE (CPU, minutes) on the full-size configurations and store its results as
fixtures, so the GPU box only has to run the CUDA path:  tests/golden/dhfr2_oracle.npz"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import tinker_gpu_b200 as tg  # noqa: E402
from oracle.amoeba_ref import Oracle, V1  # noqa: E402

s = tg.load_system(os.path.join(HERE, "dhfr2.npz"))
o = Oracle(s)
t0 = time.time()
r = o.energy(V1)
print("dhfr2 oracle energy(V1)", time.time() - t0, "s  iters", o.niter, "em", r["em"], "ep", r["ep"])
np.savez_compressed(os.path.join(HERE, "dhfr2_oracle.npz"), em=r["em"], ep=r["ep"], ep_pair=r["ep_pair"], em_real=r["em_real"],
                    em_recip=r["em_recip"], em_self=r["em_self"], grad=r["grad"], virial=r["virial"], uind=o.uind, uinp=o.uinp,
                    udir=o.udir, udirp=o.udirp, niter=o.niter, npairs=o.pairs(s.ewald_cutoff)[0].shape[0])
