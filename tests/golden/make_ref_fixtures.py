#!/usr/bin/env python
"""Oracle real-space quantities of the full-size dhfr2 deck, stored so that tests can hold them against oracle/_ref
(the reference's own pair arithmetic compiled in place) in a second instead of the ~10 oracle minutes they take here:

    python tests/golden/make_ref_fixtures.py        ->  tests/golden/dhfr2_oracle_real.npz

em_real / ep_real (pairwise polarization energy), real-space gradient and torque (multipole + polarization), the
real-space permanent field (d scaling) and the real-space mutual field of the converged dipoles of dhfr2_oracle.npz."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import tinker_gpu_b200 as tg  # noqa: E402
from oracle.amoeba_ref import Oracle, V4  # noqa: E402

s = tg.load_system(os.path.join(HERE, "dhfr2.npz"))
z = np.load(os.path.join(HERE, "dhfr2_oracle.npz"))
o = Oracle(s)
o.rotpole()
o.uind, o.uinp = z["uind"], z["uinp"]
t0 = time.time()
rs = o._real_space(V4, True, True)
fd, fp = o.dfield(real_only=True)
ufd, ufp = o.ufield(o.uind, o.uinp, real_only=True)
print("dhfr2 real space", time.time() - t0, "s", rs["em"], rs["ep"])
np.savez_compressed(os.path.join(HERE, "dhfr2_oracle_real.npz"), em_real=rs["em"], ep_real=rs["ep"], g_real=rs["gm"] + rs["gp"],
                    t_real=rs["tm"] + rs["tp"], fd_real=fd, fp_real=fp, ufd_real=ufd, ufp_real=ufp,
                    npairs=o.pairs(s.ewald_cutoff)[0].shape[0])
