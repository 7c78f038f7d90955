"""Host-side logic of the spatial decomposition (dist.cu: plan_halo through the C ABI, no GPU):
the halo plan of every rank is consistent with every other rank's (what r sends to s is what s expects
from r, in the same order) and covers every neighbour within the list range."""
import os
import subprocess
import sys

import numpy as np

from conftest import GOLDEN, ROOT


def _slabs(system, world):
    """Sorted order + ownership exactly as nblist.cu/dist.cu derive them (slab of w3, then anything)."""
    lz = system.lvec[2][2]
    f3 = (np.array(system.xyz)[:, 2] / lz) % 1.0
    w3 = (f3 + 0.5) % 1.0
    slab = np.minimum(world - 1, (w3 * world).astype(int))
    order = np.argsort(slab, kind="stable")
    bounds = np.searchsorted(slab[order], np.arange(world + 1))
    return order, w3[order].astype(np.float32), bounds.astype(np.int32)


def test_plan_covers_all_neighbours_and_pairs_up():
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import dist_plan
    s = tg.load_system(os.path.join(GOLDEN, "water30.npz"))
    world = 4
    order, w3s, bounds = _slabs(s, world)
    rng_ = (s.ewald_cutoff + s.list_buffer)
    range_frac = rng_ / s.lvec[2][2]
    plans = [dist_plan(w3s, bounds, world, r, range_frac) for r in range(world)]
    for r in range(world):
        si, so, ri, ro = plans[r]
        for p in range(world):
            # what r sends to p == what p receives from r
            assert np.array_equal(si[so[p]:so[p + 1]], plans[p][2][plans[p][3][r]:plans[p][3][r + 1]])
        assert so[r] == so[r + 1] and ro[r] == ro[r + 1]           # nothing to self
        own = np.arange(bounds[r], bounds[r + 1])
        assert not np.isin(ri, own).any()
    # coverage: every atom within the list range of an owned atom is owned or in the halo
    x = np.array(s.xyz)[order]
    L = np.diag(np.array(s.lvec))
    r = 1
    own = np.arange(bounds[r], bounds[r + 1])
    have = np.zeros(s.n, bool)
    have[own] = True
    have[plans[r][2]] = True
    d = x[own][:, None, :] - x[None, :, :]
    d -= L * np.round(d / L)
    near = (np.einsum("ikc,ikc->ik", d, d) <= rng_ ** 2).any(0)
    assert have[near].all()


_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import tinker_gpu_b200 as tg
from tinker_gpu_b200.amoeba import dist_plan
from test_dist_plan import _slabs
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
s = tg.load_system(os.path.join({golden!r}, "water30.npz"))
order, w3s, bounds = _slabs(s, world)
si, so, ri, ro = dist_plan(w3s, bounds, world, rank, (s.ewald_cutoff + s.list_buffer) / s.lvec[2][2])
# halo exchange of a per-atom vector over gloo, driven by the plan: afterwards every halo atom holds its owner's value
v = torch.full((s.n,), -1.0, dtype=torch.float64)
v[bounds[rank]:bounds[rank + 1]] = torch.arange(bounds[rank], bounds[rank + 1], dtype=torch.float64) * 2.0 + 1.0
peer = 1 - rank
send = v[torch.from_numpy(si[so[peer]:so[peer + 1]].astype(np.int64))].contiguous()
recv = torch.empty(int(ro[peer + 1] - ro[peer]), dtype=torch.float64)
ops = [dist.P2POp(dist.isend, send, peer), dist.P2POp(dist.irecv, recv, peer)]
for w in dist.batch_isend_irecv(ops):
    w.wait()
idx = ri[ro[peer]:ro[peer + 1]].astype(np.int64)
v[torch.from_numpy(idx)] = recv
ok = bool((v[torch.from_numpy(idx)] == torch.from_numpy(idx).double() * 2.0 + 1.0).all()) and len(idx) > 0
# energy-like partial sums reduce to the whole
part = torch.tensor([float(bounds[rank + 1] - bounds[rank])], dtype=torch.float64)
dist.all_reduce(part)
t = torch.tensor([1.0 if ok and int(part) == s.n else 0.0])
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("RESULT", int(t), len(idx))
dist.destroy_process_group()
"""


def test_halo_exchange_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT, golden=GOLDEN))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29531", str(script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT")][0].split()
    assert line[1] == "1" and int(line[2]) > 100


_RDV_CHILD = r"""
import ctypes, os, sys
import numpy as np
lib = ctypes.CDLL(sys.argv[1])
job = bytes.fromhex(sys.argv[2]); rank = int(sys.argv[3]); world = int(sys.argv[4]); rounds = int(sys.argv[5])
mine = np.arange(64, dtype=np.uint8) + 10 * rank
out = np.zeros((world, 64), dtype=np.uint8)
lib.apx_rendezvous_selftest.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
rc = lib.apx_rendezvous_selftest(job, rank, world, mine.ctypes.data, 64, rounds, out.ctypes.data)
ok = rc == 0 and all(np.array_equal(out[r], (np.arange(64) + 10 * r + rounds - 1).astype(np.uint8)) for r in range(world))
sys.exit(0 if ok else 1)
"""


def test_direct_transport_rendezvous_between_processes():
    """The start-up exchange of transport "direct" (dist.cu FileRendezvous: the CUDA IPC handles travel through /dev/shm when
    there is no NCCL): three PROCESSES, five rounds; every rank must end with every rank's blob of the last round, and a rank's
    files of earlier rounds must be gone (removed once the following round has been read)."""
    import glob
    lib = os.path.join(ROOT, "tinker-gpu_b200", "libapx.so")
    job = os.urandom(16).hex()
    world, rounds = 3, 5
    procs = [subprocess.Popen([sys.executable, "-c", _RDV_CHILD, lib, job, str(r), str(world), str(rounds)]) for r in range(world)]
    try:
        assert [p.wait(timeout=120) for p in procs] == [0] * world
        left = sorted(os.path.basename(f) for f in glob.glob("/dev/shm/apx_" + job + "_*"))
        assert left == [f"apx_{job}_{rounds}_{r}" for r in range(world)]      # only the last round's files remain
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        for f in glob.glob("/dev/shm/apx_" + job + "_*"):
            os.remove(f)
