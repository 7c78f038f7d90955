"""Launch helpers for the spatially decomposed path (libapx, dist.cu).

`run_local_ranks(system, world, fn)` runs `fn(amoeba, rank)` on `world` ranks that are host threads
of this process sharing one GPU (in-process transport): the whole decomposition -- slab ownership,
halo exchange, slab FFT, reductions -- executes exactly as over NCCL, which makes it testable against
the single-GPU path on a one-GPU box.  `nccl_context(system, ...)` builds the rank of a real
one-process-per-GPU job from the torch.distributed environment."""
from __future__ import annotations

import threading

from .amoeba import Amoeba, LocalHub, nccl_unique_id


def run_local_ranks(system, world, fn, precision="mixed", device=0, vdw=False):
    hub = LocalHub(world, precision)
    out = [None] * world
    err = [None] * world

    def work(rank):
        a = None
        try:
            a = Amoeba(system, precision, device=device, dist=(rank, world, "local", hub), vdw=vdw)
            out[rank] = fn(a, rank)
        except BaseException as e:      # noqa: BLE001 -- reported to the caller below
            err[rank] = e
        finally:
            if a is not None:
                try:
                    a.synchronize()
                except Exception:
                    pass

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    hub.close()
    for e in err:
        if e is not None:
            raise e
    return out


def nccl_context(system, precision="mixed", vdw=False):
    """One rank of a torchrun job (RANK/WORLD_SIZE/LOCAL_RANK set, process group initialised)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.cuda.current_device()
    if world == 1:
        return Amoeba(system, precision, device=dev, vdw=vdw)
    ident = [nccl_unique_id(precision) if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    return Amoeba(system, precision, device=dev, dist=(rank, world, "nccl", ident[0]), vdw=vdw)
