"""Parity of the NCCL-transport decomposition on REAL GPUs (one process per GPU, torchrun) against the
single-GPU path computed by rank 0 on the same inputs.  usage:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/nccl_check.py [blob]
Prints one RESULT line (rank 0) and exits non-zero on a mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tinker_gpu_b200 as tg                                    # noqa: E402
from tinker_gpu_b200.amoeba import Amoeba, calc                 # noqa: E402
from tinker_gpu_b200.distributed import nccl_context            # noqa: E402

DEBYE = 4.803206802
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rank, world = dist.get_rank(), dist.get_world_size()
name = sys.argv[1] if len(sys.argv) > 1 else "water30"
s = tg.load_system(os.path.join(ROOT, "tests", "golden", name + ".npz"))
if len(sys.argv) > 2:
    s = tg.replicate(s, tuple(int(v) for v in sys.argv[2].split("x")), keep_bonds=False)
x0 = np.array(s.xyz)
frames = [x0, x0 + np.random.default_rng(3).normal(scale=0.02, size=x0.shape), x0 + np.array([0.0, 0.0, 3.3])]
ref = []
if rank == 0:
    a = Amoeba(s, "mixed", device=lr)
    for x in frames:
        a.set_positions(x)
        r = a.energy(calc.v1)
        r["uind"] = a.uind()[0]
        ref.append(r)
    a.close()
am = nccl_context(s, "mixed")
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
        good = de < 3e-7 and dg < 3e-5 and du < 3e-7 and dv < 2e-3 and r["pcg_iterations"] == q["pcg_iterations"]
        ok = ok and good
        print(f"RESULT {name} world={world} frame={j} n={s.n} dE/E={de:.2e} grad_rms={dg:.2e} uind_rms_D={du:.2e} vir={dv:.2e} "
              f"iters={r['pcg_iterations']}/{q['pcg_iterations']} info={info} {'OK' if good else 'MISMATCH'}", flush=True)
am.close()
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag) else 1)
