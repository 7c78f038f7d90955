"""Parity of the DIRECT transport (dist.cu: one kernel per exchange writing into the peers' registered buffers over CUDA IPC,
no NCCL) against the single-GPU path on the same inputs.  The ranks are separate processes; with fewer GPUs than ranks
they share devices (rank % n_gpus), which is how the one-GPU test box exercises the multi-process path: the exchange kernels
of the two processes then alternate by time slice, slow but exact.  usage:
   python tools/direct_check.py [--world 2] [--blob water30] [--rep 1x1x1] [--precision mixed]
Prints one RESULT line per frame (rank 0) and exits non-zero on a mismatch."""
import argparse
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DEBYE = 4.803206802


def frames_of(s):
    x0 = np.array(s.xyz)
    return [x0, x0 + np.random.default_rng(3).normal(scale=0.02, size=x0.shape), x0 + np.array([0.0, 0.0, 3.3])]


def child(args):
    import torch
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba, calc
    rank, world = args.rank, args.world
    dev = rank % max(1, torch.cuda.device_count())
    s = tg.load_system(os.path.join(ROOT, "tests", "golden", args.blob + ".npz"))
    if args.rep != "1x1x1":
        s = tg.replicate(s, tuple(int(v) for v in args.rep.split("x")), keep_bonds=False)
    frames = frames_of(s)
    ref = []
    if rank == 0:
        a = Amoeba(s, args.precision, device=dev)
        for x in frames:
            a.set_positions(x)
            r = a.energy(calc.v1)
            r["uind"] = a.uind()[0]
            ref.append(r)
        a.close()
    am = Amoeba(s, args.precision, device=dev, dist=(rank, world, "direct", bytes.fromhex(args.job)))
    ok = True
    for j, x in enumerate(frames):
        am.set_positions(x)
        r = am.energy(calc.v1)
        u = am.uind()[0]
        info = am.dist_info()
        if rank == 0:
            q = ref[j]
            de = abs(r["esum"] - q["esum"]) / abs(q["esum"])
            dg = float(np.sqrt(((r["grad"] - q["grad"]) ** 2).mean()))
            du = float(np.sqrt(((u - q["uind"]) ** 2).mean())) * DEBYE
            dv = float(np.abs(r["virial"] - q["virial"]).max() / max(1.0, np.abs(q["virial"]).max()))
            tol = (1e-10, 1e-8, 1e-9, 1e-8) if args.precision == "double" else (3e-7, 3e-5, 3e-7, 2e-3)
            good = de < tol[0] and dg < tol[1] and du < tol[2] and dv < tol[3] and r["pcg_iterations"] == q["pcg_iterations"]
            ok = ok and good
            print(f"RESULT direct {args.blob} {args.rep} world={world} frame={j} n={s.n} dE/E={de:.2e} grad_rms={dg:.2e} "
                  f"uind_rms_D={du:.2e} vir={dv:.2e} iters={r['pcg_iterations']}/{q['pcg_iterations']} info={info} "
                  f"{'OK' if good else 'MISMATCH'}", flush=True)
    am.close()
    sys.exit(0 if ok else 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=2)
    ap.add_argument("--blob", default="water30")
    ap.add_argument("--rep", default="1x1x1")
    ap.add_argument("--precision", default="mixed")
    ap.add_argument("--rank", type=int, default=-1)
    ap.add_argument("--job", default="")
    ap.add_argument("--timeout", type=float, default=600.0)
    args = ap.parse_args()
    if args.rank >= 0:
        child(args)
        return
    job = os.urandom(16).hex()
    procs = []
    for r in range(args.world):
        cmd = [sys.executable, os.path.abspath(__file__), "--world", str(args.world), "--blob", args.blob, "--rep", args.rep,
               "--precision", args.precision, "--rank", str(r), "--job", job]
        procs.append(subprocess.Popen(cmd))
    rc = 0
    try:
        for p in procs:
            rc = max(rc, abs(p.wait(timeout=args.timeout)))
    except subprocess.TimeoutExpired:
        rc = 124
        print("direct_check: timeout", flush=True)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        import glob
        for f in glob.glob("/dev/shm/apx_" + job + "_*"):
            try:
                os.remove(f)
            except OSError:
                pass
    sys.exit(rc)


if __name__ == "__main__":
    main()
