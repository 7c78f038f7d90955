"""Ours and the reference's CUDA build (oracle/_ref/libref_cuda.so) on dhfr2 in ONE job on ONE GPU under the same conditions:
calls back to back, no L2 flush, device-event time per call (ours: the library's own ms_induce / ms_energy events around
induce() and around the whole energy(energy+grad) call up to the reduced scalars; reference: oracle/ref_cuda.cu refcu_time).
No torch import, so it fits in a few seconds of GPU time.  Prints one JSON line.
   python tools/same_job_compare.py [reps] [workload]      workload: dhfr2 (default) | water96k | dhfr424k | water1m (bench.py WORKLOADS;
   the replicated box is written to a temporary blob for the comparator child; parity of the comparator is checked for dhfr2 only)"""
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
    workload = sys.argv[2] if len(sys.argv) > 2 else "dhfr2"
    blob = os.path.join(GOLDEN, "dhfr2.npz")
    if workload == "dhfr2":
        s = tg.load_system(blob)
    else:
        import tempfile
        sys.argv = sys.argv[:1]
        import bench
        s = bench.make_system(workload)
        s.vdw = s.valence = None
        blob = os.path.join(tempfile.mkdtemp(), workload + ".npz")
        tg.save_system(blob, s)
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
    cmd = [sys.executable, "-m", "oracle.ref_cuda_bridge", blob, "--reps", str(reps), "--warmup", "5"]
    if workload == "dhfr2":
        cmd += ["--fixture", os.path.join(GOLDEN, "dhfr2_oracle.npz")]
    c = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
    if c.returncode == 0:
        out["ref_cuda"] = json.loads(c.stdout.strip().splitlines()[0])
        out["esum_rel_ours_vs_ref_cuda"] = abs(out["ref_cuda"]["esum"] - out["ours"]["esum"]) / abs(out["ours"]["esum"])
        out["induce_speedup"] = out["ref_cuda"]["induce_ms"]["median"] / out["ours"]["induce_ms"]["median"]
        out["energy_speedup"] = out["ref_cuda"]["energy_ms"]["median"] / out["ours"]["energy_ms"]["median"]
    else:
        out["ref_cuda"] = {"failed": (c.stderr or c.stdout)[-300:]}
    out["workload"], out["atoms"] = workload, int(s.n)
    out["conditions"] = "same job, same GPU, calls back to back (warm L2), device events per call, 5 warm-up calls each"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
