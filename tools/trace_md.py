#!/usr/bin/env python
"""Warm timeline of dhfr2 MD steps (torch.profiler / CUPTI): per-step spans, and the full kernel timeline of one ordinary step and
of one step with a list rebuild.  usage: python tools/trace_md.py [--steps 24] [--out gpurun_out/trace_md.txt]"""
import argparse
import collections
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from tinker_gpu_b200.amoeba import Amoeba, MdReport  # noqa: E402
from tinker_gpu_b200.drivers import maxwell_velocities  # noqa: E402


def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("void ", "")
    return name.split("(")[0][:60]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "trace_md.txt"))
    args = ap.parse_args()
    torch.cuda.set_device(0)
    s = bench.make_system("dhfr2")
    a = Amoeba(s, "mixed", device=0, vdw=True, valence=True)
    nfree = 3 * s.n - 3
    vel = maxwell_velocities(s.mass, bench.MD_KELVIN, bench.MD_SEED, nfree)
    a.md_init(s.mass, vel, dt=bench.MD_DT_PS, nrespa=bench.MD_NRESPA, thermostat="BUSSI", kelvin=bench.MD_KELVIN, tautemp=bench.MD_TAU,
              nfree=nfree, seed=bench.MD_SEED)
    rep = MdReport()
    a.lib.apx_md_steps(a.ctx, 8, C.byref(rep))
    a.synchronize()
    from torch.profiler import profile, ProfilerActivity
    reb = []
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            a.lib.apx_md_steps(a.ctx, 1, C.byref(rep))
            reb.append(rep.list_rebuilds)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    starts = [i for i, e in enumerate(evs) if "k_md_zero1" in e.name] + [len(evs)]
    out = []
    segs = [evs[starts[j]:starts[j + 1]] for j in range(len(starts) - 1)]
    out.append("# step  rebuild  span_us  kernel_sum_us  n_activities")
    for j, seg in enumerate(segs):
        t0, t1 = seg[0].time_range.start, seg[-1].time_range.end
        out.append(f"{j:4d} {reb[j] if j < len(reb) else -1:4d} {t1 - t0:10.1f} {sum(e.time_range.end - e.time_range.start for e in seg):10.1f} {len(seg):5d}")
    pick = [j for j in range(len(segs)) if j < len(reb) and reb[j] > 0][:1] + [j for j in range(2, len(segs)) if j < len(reb) and reb[j] == 0 and reb[j - 1] == 0][:1]
    for j in pick:
        seg = segs[j]
        t0 = seg[0].time_range.start
        out.append(f"\n# ---- step {j} (rebuild={reb[j]}): per-kernel totals")
        tot = collections.OrderedDict()
        for e in seg:
            k = short(e.name)
            n, us = tot.get(k, (0, 0.0))
            tot[k] = (n + 1, us + e.time_range.end - e.time_range.start)
        for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            out.append(f"{k:62s} {n:4d} {us:9.1f}")
        out.append("# timeline (start us, dur us, gap before us, name)")
        prev = t0
        for e in seg:
            out.append(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f} {e.time_range.start - prev:8.1f}  {short(e.name)}")
            prev = max(prev, e.time_range.end)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    open(args.out, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:args.steps + 2]))
    a.close()


if __name__ == "__main__":
    main()
