#!/usr/bin/env python
"""Warm-cache kernel timeline of one hot-path step via torch.profiler (CUPTI sees the kernels
libapx launches in-process).  Prints per-kernel totals of ONE energy() call and the idle gaps.
usage: python tools/trace_step.py [--system dhfr2.npz] [--steps 3] [--out gpurun_out/trace.txt]"""
import argparse
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import tinker_gpu_b200 as tg  # noqa: E402
from tinker_gpu_b200.amoeba import Amoeba, calc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--system", default=os.path.join(ROOT, "tests", "golden", "dhfr2.npz"))
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "trace.txt"))
    ap.add_argument("--workload", default=None, help="bench.py workload name instead of --system")
    args = ap.parse_args()
    if args.workload:
        import bench
        s = bench.make_system(args.workload)
    else:
        s = tg.load_system(args.system)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        # one process per GPU under torchrun: the spatially decomposed path over NCCL; rank 0 reports its own timeline
        import torch.distributed as dist
        from tinker_gpu_b200.distributed import nccl_context
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
        a = nccl_context(s, "mixed")
    else:
        a = Amoeba(s, "mixed", device=lr)
    for _ in range(5):
        a.lib.apx_energy(a.ctx, calc.v4, None)
    a.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.steps):
            a.lib.apx_energy(a.ctx, calc.v4, None)
        a.synchronize()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # split into steps at k_chkpole (first kernel of energy())
    starts = [i for i, e in enumerate(evs) if "k_chkpole" in e.name]
    lines = []
    if len(starts) >= 2:
        seg = evs[starts[-2]:starts[-1]]
    else:
        seg = evs
    t0 = seg[0].time_range.start
    t1 = max(e.time_range.end for e in seg)
    tot = collections.OrderedDict()
    busy = 0.0
    last_end = t0
    gap = 0.0
    for e in seg:
        name = re.sub(r"\(.*", "", e.name.replace("(anonymous namespace)::", "").replace("void ", ""))[:60]
        d = e.time_range.end - e.time_range.start
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += d
        if e.time_range.start > last_end:
            gap += e.time_range.start - last_end
        last_end = max(last_end, e.time_range.end)
    span = t1 - t0
    lines.append(f"# one energy() step: {len(seg)} GPU activities, span {span:.1f} us, idle gaps {gap:.1f} us ({100*gap/span:.1f}%)")
    s_all = sum(t[1] for t in tot.values())
    lines.append(f"{'kernel':62s} {'n':>4s} {'us':>9s} {'us/launch':>9s} {'share of span':>8s}")
    for k, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k:62s} {t[0]:4d} {t[1]:9.1f} {t[1]/t[0]:9.2f} {100*t[1]/span:7.1f}%")
    lines.append(f"{'sum of kernel durations (streams overlap)':62s} {'':4s} {s_all:9.1f}")
    # timeline of the last iteration-ish: first 60 activities with start offsets
    lines.append("# timeline (start us, dur us, name)")
    for e in seg[: min(len(seg), 700)]:
        lines.append(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:7.1f}  {e.name[:70]}")
    if rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")
        print("\n".join(lines[:45]))
    a.close()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
