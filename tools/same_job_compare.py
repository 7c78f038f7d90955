"""Ours and the reference's CUDA build (oracle/_ref/libref_cuda.so) on dhfr2 in ONE job on ONE GPU under the same conditions:
calls back to back, no L2 flush, device-event time per call (ours: the library's own ms_induce / ms_energy events around
induce() and around the whole energy(energy+grad) call up to the reduced scalars; reference: oracle/ref_cuda.cu refcu_time).
No torch import, so it fits in a few seconds of GPU time.  Prints one JSON line."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def main():
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba, EnergyResult, calc
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    s = tg.load_system(os.path.join(GOLDEN, "dhfr2.npz"))
    a = Amoeba(s, "mixed", device=0)
    r = EnergyResult()
    ind, ene, its = [], [], []
    for k in range(5 + reps):
        a._chk(a.lib.apx_energy(a.ctx, int(calc.v4), C.byref(r)))
        st = a.stats()
        if k >= 5:
            ind.append(st["ms_induce"]), ene.append(st["ms_energy"]), its.append(r.pcg_iterations)
    out = {"ours": {"induce_ms": {"median": float(np.median(ind)), "min": float(min(ind)), "max": float(max(ind))},
                    "energy_ms": {"median": float(np.median(ene)), "min": float(min(ene)), "max": float(max(ene))},
                    "pcg_iterations": float(np.mean(its)), "esum": r.esum, "reps": reps, "vers": "energy+grad (calc.v4)"}}
    a.close()
    c = subprocess.run([sys.executable, "-m", "oracle.ref_cuda_bridge", os.path.join(GOLDEN, "dhfr2.npz"), "--fixture",
                        os.path.join(GOLDEN, "dhfr2_oracle.npz"), "--reps", str(reps), "--warmup", "5"], cwd=ROOT, capture_output=True, text=True)
    if c.returncode == 0:
        out["ref_cuda"] = json.loads(c.stdout.strip().splitlines()[-1])
        out["induce_speedup"] = out["ref_cuda"]["induce_ms"]["median"] / out["ours"]["induce_ms"]["median"]
        out["energy_speedup"] = out["ref_cuda"]["energy_ms"]["median"] / out["ours"]["energy_ms"]["median"]
    else:
        out["ref_cuda"] = {"failed": (c.stderr or c.stdout)[-300:]}
    out["conditions"] = "same job, same GPU, calls back to back (warm L2), device events per call, 5 warm-up calls each"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
