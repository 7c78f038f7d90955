#!/usr/bin/env python
"""Where the force error of the mixed build comes from (GPU box; no oracle run -- fixtures and the double build are the
references).  Child processes evaluate energy(v4) with APX_DIAG_SKIP (1: no real-space pair kernels, 2: no reciprocal force
kernels) on the mixed build, on its precise-math variant (csrc/build/libapx_precise.so: `make -C tinker-gpu_b200/csrc diag`)
and on the double build; the parent prints RMS differences per component.
usage: python tools/diag_precision.py [water30 dhfr2]"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
OUT = os.path.join(ROOT, "gpurun_out")


def child(blob, precision, out):
    sys.path.insert(0, ROOT)
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba, calc
    s = tg.load_system(os.path.join(G, blob + ".npz"))
    a = Amoeba(s, precision, device=0)
    r = a.energy(calc.v4)
    ud, _ = a.uind()
    np.savez(out, grad=r["grad"], uind=ud, esum=r["esum"])
    a.close()


def run(blob, precision, skip=0, lib=None):
    out = os.path.join(OUT, f"diag_{blob}_{precision}_{skip}_{'p' if lib else 'b'}.npz")
    env = dict(os.environ, APX_DIAG_SKIP=str(skip))
    if lib:
        env["APX_LIBRARY_MIXED"] = lib
    subprocess.run([sys.executable, __file__, "--child", blob, precision, out], env=env, check=True, timeout=300)
    return np.load(out)


def rms(a):
    return float(np.sqrt((a ** 2).mean()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(*sys.argv[2:5])
        sys.exit(0)
    os.makedirs(OUT, exist_ok=True)
    precise = os.path.join(ROOT, "tinker-gpu_b200", "csrc", "build", "libapx_precise.so")
    for blob in (sys.argv[1:] or ["water30", "dhfr2"]):
        fx = np.load(os.path.join(G, "water30_oracle_eps5.npz" if blob == "water30" else "dhfr2_oracle.npz"))
        res = {"system": blob}
        d = {k: run(blob, "double", k) for k in (0, 1, 2)}
        m = {k: run(blob, "mixed", k) for k in (0, 1, 2)}
        res["double_vs_fixture"] = rms(d[0]["grad"] - fx["grad"])
        res["mixed_vs_fixture"] = rms(m[0]["grad"] - fx["grad"])
        res["mixed_vs_double"] = rms(m[0]["grad"] - d[0]["grad"])
        res["recip_only_mixed_vs_double"] = rms(m[1]["grad"] - d[1]["grad"])
        res["real_only_mixed_vs_double"] = rms(m[2]["grad"] - d[2]["grad"])
        res["uind_mixed_vs_double_debye"] = rms(m[0]["uind"] - d[0]["uind"]) * 4.803206802
        if os.path.isfile(precise):
            p = {k: run(blob, "mixed", k, precise) for k in (0, 2)}
            res["precise_math_vs_fixture"] = rms(p[0]["grad"] - fx["grad"])
            res["real_only_precise_vs_double"] = rms(p[2]["grad"] - d[2]["grad"])
        print(json.dumps(res))
