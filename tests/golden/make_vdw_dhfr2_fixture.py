"""tests/golden/dhfr2_vdw_oracle.npz: the float64 vdW oracle (oracle/vdw_ref.py, pinned to the reference's NaCl / Local-Frame2
goldens and to pair_hal_v2 in tests/test_ref_arith.py) on dhfr2 -- pair sum WITHOUT the long-range correction (a host
constant the kernels do not compute), gradient on the real atoms, pair virial.  Used by the reference-CUDA comparator check
(oracle/ref_cuda_bridge.py --vdw).  Run from the repository root: python tests/golden/make_vdw_dhfr2_fixture.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import tinker_gpu_b200 as tg  # noqa: E402
from oracle.vdw_ref import VdwOracle  # noqa: E402

s = tg.load_system(os.path.join(ROOT, "tests", "golden", "dhfr2.npz"))
o = VdwOracle(s)
r = o.ehal()
v = s.vdw
vol = float(abs(np.linalg.det(s.lvec)))
ev = r["ev"] - (v.elrc_vol / vol if v.elrc_vol else 0.0)
vir = r["virial"] - (np.eye(3) * (v.vlrc_vol / vol) if v.vlrc_vol else 0.0)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "dhfr2_vdw_oracle.npz"), ev_pairs=ev, grad=r["grad"].astype(np.float64),
                    virial_pairs=vir, npairs=r["npairs"], nev=r["nev"])
print("ev_pairs", ev, "npairs", r["npairs"], "|g|max", np.abs(r["grad"]).max())
