#!/usr/bin/env python
"""Benchmark of the AMOEBA polarizable-electrostatics hot path (BASELINE.json metric) on B200.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N ...            # reference arm: CPU oracle on host cores

Default (--mode dynamics, BASELINE.json configs[1]): one "step" = one 2 fs r-RESPA MD step of dhfr2 on the device integrator
(csrc/md.cu): 4 valence evaluations + kick/drift, neighbour-list check, induce() + electrostatics + vdW, Bussi thermostat;
`value` = ns/day with the state resident in HBM, `md.batch` the same without L2 flushes, `e2e` the plugin call with host buffers.

--mode energy (the round-1 metric, default for the large boxes): one "step" = one pass of the hot path over the system:
energy(energy+grad) restricted to the electrostatic terms = mpoleInit + induce() (PCG) +
fused real-space multipole/polarization + reciprocal space + torque + reductions
(SURVEY.md §3.1 "HOT").  The metric carries both numbers BASELINE.json names:

  value          ns/day-equivalent of the hot path alone at 2 fs per outer RESPA step
                 (one electrostatics evaluation per step; vdW / valence / integrator are NOT
                 in this repo yet -- SURVEY.md §8f -- so this is NOT a full-MD ns/day)
  ms_per_induce  mean device time of one induce() call

N > 1: dhfr2 is latency bound on one GPU, so ranks run independent replicas (SURVEY §8e
"replicas only"); value is the aggregate over replicas, time is the max over ranks.

--workload selects the other BASELINE.json configurations (synthetic boxes built at run time by
replicating the committed cells): water96k (configs[2], polar-eps 1e-8), water1m
(configs[3]) and dhfr424k (configs[4]).  For those, N > 1 runs ONE system spatially decomposed over
the N GPUs (dist.cu: z-slabs, halo exchange of dipoles per CG iteration, slab FFT with all-to-all
transposes over NCCL) -- "scaling": "strong"; --replicas forces independent replicas instead.

`value` is measured with positions resident in HBM; `e2e` goes through the public host API
(set_positions from host memory -> energy -> gradient back to host) inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
FS_PER_STEP = 2.0
METRIC = "ns/day (electrostatics hot path only, 2 fs/step) & ms/induce() AMOEBA DHFR 23.5k atoms"
WORKLOAD = "dhfr2 AMOEBA DHFR 23558 atoms, PME 64^3 order 5, ewald-cutoff 7.0, polar-eps 1e-5, energy+gradient"
# name -> (cell blob, replication, polar-eps override, jitter, description)   
WORKLOADS = {
    "dhfr2": ("dhfr2.npz", None, None, 0.0, WORKLOAD),
    "water96k": ("water30.npz", (3, 3, 4), 1e-8, 0.05,
                 "synthetic AMOEBA water box 96624 atoms (water30 cell x3x3x4, jitter 0.05 A), PME 108x108x144, polar-eps 1e-8"),
    "water1m": ("water30.npz", (8, 8, 6), None, 0.05,
                "synthetic AMOEBA water box 1030656 atoms (water30 cell x8x8x6, jitter 0.05 A), PME 288x288x216, polar-eps 1e-5"),
    "dhfr424k": ("dhfr2.npz", (3, 3, 2), None, 0.05,
                 "replicated dhfr2 cells 424044 atoms (x3x3x2, jitter 0.05 A), dense PME 240x240x150, polar-eps 1e-5"),
}


def make_system(name):
    import tinker_gpu_b200 as tg
    blob, reps, eps, jitter, _ = WORKLOADS[name]
    s = tg.load_system(os.path.join(GOLDEN, blob))
    if reps is not None:
        s = tg.replicate(s, reps, jitter=jitter, keep_bonds=False)
    if eps is not None:
        s.poleps = eps
    return s


def ns_per_day(ms_per_step, replicas=1):
    return replicas * FS_PER_STEP * 1e-6 * (86400.0e3 / ms_per_step)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([t.strip() for t in ln.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6), ("sw_power_cap", 7)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def load_ncu_traffic(kernel_prefix, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this same command (profiles/*<workload>_ncu_full.json; cold-cache replays)."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for fn in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if fn.endswith(f"{workload}_ncu_full.json"):
            d = json.load(open(os.path.join(pdir, fn)))
            for k, v in d.items():
                if k.startswith(kernel_prefix):
                    best = (v["dram_bytes"], fn)
    return best


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("sm_max_mhz", 1965.0), "measured"
    return 6650.0, 1965.0, "fallback"


# ------------------------------------------------------------------------------------------------
# host threads of the CPU legs: the pair sweeps of the reference's operators are cut into that many slices (oracle/ref_oracle.py);
# the reference's own host build runs them serially (OpenACC pragmas without an accelerator), so this is an upper bound of it
CPU_THREADS = max(1, min(64, os.cpu_count() or 1))


def cpu_oracle_sample(system, full=False):
    """Time the CPU path on one core.  With oracle/_ref present (the reference's own pair functions and PME translation unit
    compiled in place) the WHOLE electrostatics step -- induce() to polar-eps, real-space and reciprocal energy and gradient,
    all pairs -- is run with the reference's operators driven by the oracle's PCG loop (oracle/ref_oracle.py).  Without it, the
    numpy port is timed on a bounded sample: a complete induce() and the real-space energy/gradient of a slice of the pair
    list, extrapolated.  Returns (ms_per_step, ms_induce, description, kind)."""
    from oracle import ref_bridge
    from oracle.amoeba_ref import Oracle, V4
    if all(ref_bridge.available(k) for k in ("realspace", "pme")) and not full:
        from oracle.ref_oracle import RefOracle
        o = RefOracle(system, threads=CPU_THREADS)
        o.pairs(system.ewald_cutoff)                 # neighbour search (scipy cKDTree) outside the timed region, like the GPU list
        o.pairs(system.usolve_cutoff)
        t0 = time.perf_counter()
        o.rotpole()
        o.induce()
        t_ind = time.perf_counter() - t0
        t0 = time.perf_counter()
        o._real_space(V4, True, True)
        o.empole_recip(V4)
        o.epolar_recip_self(V4)
        t_rest = time.perf_counter() - t0
        desc = (f"reference operators, pair sweeps on {CPU_THREADS} host threads (PME operators and the numpy FFT on one) (oracle/_ref: include/seq pair_mpole/pair_polar/pair_dfield/pair_ufield and "
                f"src/acc/pme.cpp compiled in place, g++ -O2, double) driven by the oracle's PCG loop, numpy FFT: full induce() "
                f"({o.niter} iterations) + real-space and reciprocal energy/gradient over all {o.pairs(system.ewald_cutoff)[0].shape[0]} pairs")
        return 1e3 * (t_ind + t_rest), 1e3 * t_ind, desc, "reference"
    o = Oracle(system)
    t0 = time.perf_counter()
    o.rotpole()
    o.induce()
    t_ind = time.perf_counter() - t0
    i, k, R, r = o.pairs(system.ewald_cutoff)
    npair = i.shape[0]
    take = npair if full else min(npair, 60000)
    saved = o._pairs[float(system.ewald_cutoff)]
    o._pairs[float(system.ewald_cutoff)] = (i[:take], k[:take], R[:take], r[:take])
    t0 = time.perf_counter()
    o._real_space(V4, True, True)
    t_real = (time.perf_counter() - t0) * (npair / take)
    o._pairs[float(system.ewald_cutoff)] = saved
    t0 = time.perf_counter()
    o.empole_recip(V4)
    o.epolar_recip_self(V4)
    t_rec = time.perf_counter() - t0
    ms_step = 1e3 * (t_ind + t_real + t_rec)
    desc = (f"oracle/amoeba_ref.py (numpy f64 port) on dhfr2: full induce() ({o.niter} iterations) + reciprocal energy/force + "
            f"real-space energy/gradient on {take} of {npair} pairs scaled to all pairs")
    return ms_step, 1e3 * t_ind, desc, "port"


def roofline_block(uf_flops, uf_bytes, uf_ms, fp32_peak, hbm_peak, peak_src, sm_clk, npairs, traffic):
    """The dominant kernel is the real-space CG operator.  Round 2 turned it from a pair kernel that recomputed the pair geometry
    in every application (FP32-bound by SURVEY 8(d)'s assignment; 6-13 % of that roofline, issue-bound) into a stored-tensor
    sparse matrix-vector product (csrc/tlist.cu): per directed pair it STREAMS 16 B of tensor + 4 B of index and gathers a
    32-byte dipole pair -- a bandwidth-bound kernel, so the roofline is the measured HBM copy bandwidth.  `achieved` =
    ALGORITHMIC bytes (20 B per directed pair, both directions stored, + 72 B per atom of vectors and row offsets; the gathers
    are not counted, they hit L1 / L2) / device time of the launch, timed IN SITU, i.e. beside the PME spread of the other
    stream.  At dhfr2 the whole stream (66 MB) is L2-resident between applications, so there the fraction is of a roofline the
    kernel does not touch; the 1 M-atom box (BASELINE configs[3]; ncu: 3.24 GB of DRAM traffic per launch) is where it is one.
    The FP32 view SURVEY 8(d) prescribes for pair kernels is kept as a side key for comparison with round 1."""
    tf = uf_flops / (uf_ms * 1e-3) / 1e12 if uf_ms > 0 else 0.0
    gbs = uf_bytes / (uf_ms * 1e-3) / 1e9 if uf_ms > 0 else 0.0
    tl = os.environ.get("APX_TLIST", "1") != "0"
    hbm = {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak if hbm_peak else 0.0, "algorithmic_bytes": uf_bytes}
    fp32 = {"achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak if fp32_peak else 0.0, "flop_per_pair": 130,
            "note": "130 flop per pair inside the cutoff (SURVEY 8d, the recomputing pair kernel) / launch time: the work the stored "
                    "tensors avoid is still counted, for comparison with round 1 (0.064)",
            "peak_source": f"148 SMs x 128 FP32 lanes x 2 flop x {sm_clk:.0f} MHz (SM clock sampled during the timed region)"}
    common = {"kernel": "real-space CG operator (csrc/tlist.cu: k_ufield_tl, stored pair tensors; 1 launch per PCG iteration)" if tl
                        else "real-space CG operator (csrc/field.cu: k_ufield_rows_rec, pair geometry recomputed; APX_TLIST=0)",
              "traffic": traffic[0] if traffic else None,
              "traffic_source": ("profiles/" + traffic[1] + " (ncu --set full, cold-cache replay)") if traffic else None,
              "pairs": int(npairs), "directed_pairs": int(2 * npairs), "ms_per_launch": uf_ms,
              "ms_per_launch_source": "CUDA events around the operator launch of the second PCG iteration of every timed step (external "
                                      "event nodes inside the iteration-batch graph): in situ, beside the PME spread of the main stream",
              "peak_source": f"MEASURED_PEAKS.json ({peak_src})"}
    if tl:
        return dict(common, bound="hbm", achieved=gbs, peak=hbm_peak, unit="GB/s", frac=hbm["frac"], algorithmic_bytes=uf_bytes, fp32=fp32)
    return dict(common, bound="fp32", achieved=tf, peak=fp32_peak, unit="TFLOP/s", frac=fp32["frac"], hbm=hbm)


def operator_bytes(n, npairs):
    """Algorithmic bytes of one application of the real-space operator: 16 B tensor + 4 B index per directed pair (stored-tensor
    form; the recomputing form reads 4 B index + a 48 B record it gathers), 32 B in + 32 B out + 8 B of row offsets per atom."""
    if os.environ.get("APX_TLIST", "1") != "0":
        return (16 + 4) * 2 * npairs + (32 + 32 + 8) * n
    return (16 + 16 + 24 + 48) * n + 4 * 2 * npairs


MD_METRIC = "ns/day & ms/induce() AMOEBA DHFR 23.5k atoms (dynamic, 2 fs RESPA, NVT)"
MD_WORKLOAD = ("example/dhfr2 AMOEBA DHFR 23558 atoms (amoebabio09): dynamic 2 fs r-RESPA (4 inner valence steps), NVT Bussi 298 K, "
               "PME 64^3 order 5, ewald-cutoff 7.0, vdw-cutoff 12.0, polar-eps 1e-5")
MD_DT_PS, MD_NRESPA, MD_KELVIN, MD_TAU, MD_SEED = 0.002, 4, 298.0, 0.2, 20261017


def ref_cuda_sample(**kw):
    """_ref_cuda_sample behind a catch-all: nothing in the comparator leg may cost the run its JSON line."""
    try:
        return _ref_cuda_sample(**kw)
    except Exception as e:      # noqa: BLE001
        return {"unavailable": f"comparator leg failed: {type(e).__name__}: {e}"}


def _ref_cuda_sample(ours_induce_ms=None, ours_energy_ms=None, ours_md_step_ms=None, timeout_s=150, system=None, ours_esum=None):
    """The reference's own CUDA kernels (oracle/_ref/libref_cuda.so: its src/cu/**/*.cu compiled unmodified for sm_100 with its
    release flags, oracle/ref_cuda.cu) on dhfr2 on the same GPU, in a CHILD process with a hard time limit, after our own
    measurements are complete: ms per mpoleInit + induce() and per fused energy+gradient+virial step (which contains an
    induce()), checked against the committed float64 oracle fixture.  SURVEY 8(d)'s "1.5x comparator"; the reference
    executable itself cannot be linked in this image (Fortran).  Never raises: a missing library, a failure or a timeout is
    reported in the block.  system: another workload (a replicated box) -- written to a temporary blob for the child; there
    is no oracle fixture at those sizes, so the block carries the relative difference of the reference's E to ours."""
    lib = os.path.join(ROOT, "oracle", "_ref", "libref_cuda.so")
    if not os.path.isfile(lib):
        return {"unavailable": "oracle/_ref/libref_cuda.so not built (make -C oracle cuda needs /root/reference)"}
    cmd = [sys.executable, "-m", "oracle.ref_cuda_bridge", os.path.join(GOLDEN, "dhfr2.npz"),
           "--fixture", os.path.join(GOLDEN, "dhfr2_oracle.npz"), "--reps", "30", "--warmup", "5",
           "--vdw", os.path.join(GOLDEN, "dhfr2_vdw_oracle.npz")]
    if system is not None:
        try:
            import tempfile
            import tinker_gpu_b200 as tg
            keep = (system.vdw, system.valence)
            system.vdw = system.valence = None
            blob = os.path.join(tempfile.mkdtemp(prefix="apx_refcuda_"), "system.npz")
            tg.save_system(blob, system)
            system.vdw, system.valence = keep
        except Exception as e:      # noqa: BLE001
            return {"unavailable": f"could not write the system blob for the comparator: {e}"}
        cmd = [sys.executable, "-m", "oracle.ref_cuda_bridge", blob, "--reps", "10", "--warmup", "3"]
        timeout_s = max(timeout_s, 300)
    rc, stdout, stderr = 0, "", ""
    try:
        r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout_s)
        rc, stdout, stderr = r.returncode, r.stdout or "", r.stderr or ""
    except subprocess.TimeoutExpired as e:
        dec = lambda b: b.decode(errors="replace") if isinstance(b, bytes) else (b or "")      # noqa: E731
        rc, stdout, stderr = -999, dec(e.stdout), f"time limit of {timeout_s} s exceeded"
    except Exception as e:      # noqa: BLE001
        return {"unavailable": f"comparator child could not start: {e}"}
    # line 1 = electrostatics (validated on a B200 in round 1); line 2 = {"vdw": ...} from the reference's ehal.cu, which may be
    # missing if that newer part failed -- the first line stands on its own
    out = None
    for ln in stdout.strip().splitlines():
        try:
            d = json.loads(ln)
        except Exception:      # noqa: BLE001
            continue
        if out is None:
            out = d
        else:
            out.update(d)
    if out is None:
        tail = (stderr or stdout).strip().splitlines()[-1:] or [""]
        return {"unavailable": f"comparator child exit {rc}: {tail[0][:200]}"}
    if rc != 0:
        tail = stderr.strip().splitlines()[-1:] or [""]
        out["vdw"] = {"unavailable": f"child exit {rc} after the electrostatics line: {tail[0][:200]}"}
    if ours_esum is not None and "esum" in out:
        out["esum_rel_vs_ours"] = abs(out["esum"] - ours_esum) / abs(ours_esum)
    par = out.get("parity") or {}
    # the reference computes in mixed precision: float pair math, fixed-point sums (the same tolerances our mixed build is held to)
    out["parity_ok"] = (bool(par["esum_rel"] < 1e-5 and par["uind_rms_debye"] < 1e-4 and par["grad_rms"] < 1e-2) if par
                        else (bool(out["esum_rel_vs_ours"] < 1e-5) if "esum_rel_vs_ours" in out else None))
    out["build"] = "reference src/cu/**/*.cu unmodified, nvcc -O3 --use_fast_math sm_100, mixed precision (oracle/Makefile: cuda)"
    out["timing"] = (f"CUDA events on the reference's stream around each call, {(out.get('induce_ms') or {}).get('reps', '?')} calls after "
                     "warm-up, back to back (warm L2, which favours the reference: ours are timed with the L2 flushed), same GPU, right after "
                     "our own measurements")
    if ours_induce_ms:
        out["ours_induce_ms"] = ours_induce_ms
        out["induce_speedup_vs_ref_cuda"] = out["induce_ms"]["median"] / ours_induce_ms
    if ours_energy_ms:
        out["ours_energy_ms"] = ours_energy_ms
        out["energy_speedup_vs_ref_cuda"] = out["energy_ms"]["median"] / ours_energy_ms
    ehal = (out.get("vdw") or {}).get("ehal_ms")
    if ours_md_step_ms and ehal:
        lo = out["energy_ms"]["median"] + ehal["median"]
        out["md_step_lower_bound"] = {"ms": lo, "ns_per_day_upper_bound": ns_per_day(lo), "ours_ms_per_step": ours_md_step_ms,
                                      "speedup_lower_bound": lo / ours_md_step_ms,
                                      "note": "reference electrostatics step + its ehal kernel, run one after the other as the reference does; its "
                                              "valence terms, integrator and per-step list refresh are NOT included, so its real MD step is longer"}
    return out


def cpu_dynamics_sample(system):
    """One MD step of the CPU oracle on a bounded sample: the electrostatics sample of cpu_oracle_sample (full induce(),
    reciprocal space, a slice of the real-space pairs scaled up) + the full vdW oracle + nrespa evaluations of the
    valence oracle.  The integrator's own cost is negligible beside these.  Returns (ms_step, ms_induce, description)."""
    from oracle import valence_ref
    from oracle.vdw_ref import VdwOracle
    from oracle import ref_bridge
    ms_elec, ms_ind, desc, kind = cpu_oracle_sample(system)
    vo = VdwOracle(system)
    if kind == "reference":
        pr = vo.pairs(vo.reduced())                  # neighbour search outside the timed region, like the GPU's Verlet rows
        t0 = time.perf_counter()
        hp = ref_bridge.hal_pairs(vo, pr, threads=CPU_THREADS)
        ms_vdw = 1e3 * (time.perf_counter() - t0)
        vdw_desc = (f"the reference's pair_hal_v2 over all {hp['npairs']} pairs within 12 A on {CPU_THREADS} host threads (oracle/_ref, "
                    f"{ms_vdw:.0f} ms incl. numpy gather of the pair data)")
    else:
        t0 = time.perf_counter()
        vo.ehal()
        ms_vdw = 1e3 * (time.perf_counter() - t0)
        vdw_desc = f"oracle/vdw_ref.py (numpy) all pairs within 12 A ({ms_vdw:.0f} ms)"
    t0 = time.perf_counter()
    valence_ref.valence(system.xyz, system.valence)
    ms_val = 1e3 * (time.perf_counter() - t0)
    val_desc = f"{MD_NRESPA} x oracle/valence_ref.py ({ms_val:.0f} ms each)"
    if ref_bridge.available("valence"):
        t0 = time.perf_counter()
        ref_bridge.valence(system)
        ms_val = 1e3 * (time.perf_counter() - t0)
        val_desc = f"{MD_NRESPA} x the reference's dk_bond ... dk_tortor (oracle/_ref, {ms_val:.1f} ms each)"
    return (ms_elec + ms_vdw + MD_NRESPA * ms_val, ms_ind,
            desc + "; + " + vdw_desc + " + " + val_desc, kind)


REF_BUDGET_S = 170.0      # wall-clock budget of the CPU arm: W warm-up + K timed samples must end within a few minutes


def timed_cpu_steps(sample, args):
    """W untimed + K timed samples of the CPU path; K is the requested --steps unless the budget runs out first (a CPU MD step
    of dhfr2 takes seconds), in which case the line says so (`steps` = what was timed, config.steps_requested / config.cap)."""
    t_start = time.perf_counter()
    out, warm = [], 0
    est = None
    for _ in range(args.warmup):
        if est is not None and time.perf_counter() - t_start + 2 * est > REF_BUDGET_S / 2:
            break
        t0 = time.perf_counter()
        sample()
        est = time.perf_counter() - t0
        warm += 1
    for _ in range(max(1, args.steps)):
        if out and time.perf_counter() - t_start + 1.2 * est > REF_BUDGET_S:
            break
        t0 = time.perf_counter()
        out.append(sample())
        est = time.perf_counter() - t0
    cap = None if len(out) == max(1, args.steps) and warm == args.warmup else (
        f"time budget {REF_BUDGET_S:.0f} s: {warm} of {args.warmup} warm-up and {len(out)} of {args.steps} timed samples ran "
        f"({est:.1f} s per sample on one core)")
    return out, warm, cap


def run_reference_dynamics(args, rank, world):
    import tinker_gpu_b200 as tg
    if rank != 0:
        return
    system = tg.load_system(os.path.join(GOLDEN, "dhfr2.npz"))
    res, warm, cap = timed_cpu_steps(lambda: cpu_dynamics_sample(system), args)
    steps = len(res)
    ms_step = float(np.mean([r[0] for r in res]))
    ms_ind = [r[1] for r in res]
    desc, kind = res[-1][2], res[-1][3]
    val = ns_per_day(ms_step)
    print(json.dumps({
        "impl": "reference", "metric": MD_METRIC, "value": val, "unit": "ns/day", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": ms_step, "ms_per_induce": float(np.mean(ms_ind)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "reference input deck example/dhfr2 (blob tests/golden/dhfr2.npz)",
        "config": {"workload": MD_WORKLOAD, "steps_requested": args.steps, "warmup_requested": args.warmup, "cap": cap,
                   "note": "CPU arm: the reference executable needs gfortran (absent); its own operators compiled in place (oracle/_ref) -- or, "
                           "without them, the oracle ports -- are timed instead, one core"},
        "cpu_baseline": {"value": val, "unit": "ns/day", "cores": CPU_THREADS if kind == "reference" else 1, "kind": kind, "sample": desc},
        "e2e": {"value": val, "unit": "ns/day", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def run_dynamics(args, rank, world, local_rank):
    """BASELINE.json configs[1]: dhfr2 dynamic, 2 fs RESPA, on the device integrator (csrc/md.cu).  One step = one outer
    RESPA step: 4 valence evaluations + kick/drift, list check, induce + electrostatics + vdW, thermostat."""
    import ctypes as C
    import torch
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba, MdReport, calc
    from tinker_gpu_b200.drivers import maxwell_velocities

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    system = make_system("dhfr2")
    n = system.n
    a = Amoeba(system, "mixed", device=local_rank, vdw=True, valence=True)
    ext = torch.cuda.ExternalStream(a.lib.apx_stream(a.ctx), device=local_rank)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    nfree = 3 * n - 3
    vel = maxwell_velocities(system.mass, MD_KELVIN, MD_SEED + rank, nfree)
    a.md_init(system.mass, vel, dt=MD_DT_PS, nrespa=MD_NRESPA, thermostat="BUSSI", kelvin=MD_KELVIN, tautemp=MD_TAU, nfree=nfree,
              seed=MD_SEED + rank)
    rep = MdReport()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def md(k):
        if a.lib.apx_md_steps(a.ctx, k, C.byref(rep)) != 0:
            raise SystemExit("apx_md_steps failed: " + a.lib.apx_last_error().decode())

    # W >= 3 warm-up steps (the timing rules' minimum): step 1 runs eagerly, step 2 captures the step graphs, step 3 replays them
    W = max(args.warmup, 3)
    md(W)
    a.synchronize()

    # ---- resident leg (value): one MD step per timed region, CUDA events on the library stream, L2 flushed between steps
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    a.stats_reset()
    ms_steps, ms_induce, ms_uf, iters, rebuilt = [], [], [], [], []
    rebuilds = 0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        md(1)
        e1.record(ext)
        e1.synchronize()
        ms_steps.append(e0.elapsed_time(e1))
        st = a.stats()
        ms_induce.append(st["ms_induce"])
        ms_uf.append(st["ms_ufield_real"])
        iters.append(st["pcg_iterations"])
        rebuilds += rep.list_rebuilds
        rebuilt.append(rep.list_rebuilds > 0)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = a.stats()["kernel_launches"]
    ms_step = float(np.mean(ms_steps))
    temp, epot, ekin = rep.temp, rep.epot, rep.ekin

    # ---- production batch: K steps in one C-ABI call, no flushes (what `dynamic` does between saves); device time from the library
    md(args.steps)
    ms_batch = rep.ms_device / max(1, args.steps)

    # ---- e2e leg: THE SAME UNIT OF WORK as `value` -- one MD step -- through the C ABI with the integrator state in HOST buffers:
    #      positions and velocities in from pinned host memory (apx_md_set_state), apx_md_steps(1), positions, velocities and the
    #      step report (energies, temperature) back to the host (apx_md_get_state), every step inside the timed region
    hx = torch.empty((n, 3), dtype=torch.float64).pin_memory()
    hv = torch.empty((n, 3), dtype=torch.float64).pin_memory()
    px, pv = hx.numpy(), hv.numpy()
    dp = C.POINTER(C.c_double)
    cx, cv = px.ctypes.data_as(dp), pv.ctypes.data_as(dp)

    def step_e2e():
        if a.lib.apx_md_set_state(a.ctx, cx, cv, 1) != 0 or a.lib.apx_md_steps(a.ctx, 1, C.byref(rep)) != 0 \
                or a.lib.apx_md_get_state(a.ctx, cx, cv) != 0:
            raise SystemExit("e2e MD step failed: " + a.lib.apx_last_error().decode())

    if a.lib.apx_md_get_state(a.ctx, cx, cv) != 0:
        raise SystemExit("apx_md_get_state failed: " + a.lib.apx_last_error().decode())
    for _ in range(2):
        step_e2e()
    barrier()
    reb0 = a.stats()["list_rebuilds"]
    ms_e2e = []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        step_e2e()
        e1.record(ext)
        e1.synchronize()
        ms_e2e.append(e0.elapsed_time(e1))
    barrier()
    ms_e2e_step = float(np.mean(ms_e2e))
    reb_e2e = a.stats()["list_rebuilds"] - reb0
    x_md = np.array(px)

    # ---- side key: force-evaluation rate of the plugin call energy(vers) with host buffers (round 1's e2e): positions in,
    #      electrostatics + vdW + valence energy and gradient, gradient out.  NOT an MD step (no inner valence evaluations).
    rng = np.random.default_rng(1234 + rank)
    drift = np.array([0.06, 0.04, 0.035])
    nframes = 2 + args.steps
    frames = [x_md + drift * float(j) + rng.normal(scale=0.002, size=x_md.shape) for j in range(nframes)]

    def force_eval(j):
        a.set_positions(frames[j % nframes])
        rc = a.lib.apx_energy(a.ctx, calc.v4, None)
        a.gradient()
        return rc

    for j in range(2):
        force_eval(j)
    ms_fe = []
    for j in range(2, 2 + args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        if force_eval(j) != 0:
            raise SystemExit("apx_energy failed: " + a.lib.apx_last_error().decode())
        e1.record(ext)
        e1.synchronize()
        ms_fe.append(e0.elapsed_time(e1))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_force_eval = float(np.mean(ms_fe))

    if dist is not None:
        t = torch.tensor([ms_step, ms_e2e_step, float(np.mean(ms_induce)), ms_batch], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, ms_e2e_step, ms_ind, ms_batch = (float(v) for v in t.tolist())
    else:
        ms_ind = float(np.mean(ms_induce))

    if rank == 0:
        st = a.stats()
        hbm_peak, sm_max, peak_src = load_peaks()
        npairs = max(1, st["npairs_m"])
        uf_ms = float(np.mean(ms_uf)) if ms_uf else 0.0
        uf_bytes = operator_bytes(n, npairs)
        uf_flops = 130.0 * npairs
        achieved_gbs = uf_bytes / (uf_ms * 1e-3) / 1e9 if uf_ms > 0 else 0.0
        sm_clk = (clocks or {}).get("sm_mhz") or sm_max
        traffic = load_ncu_traffic("k_ufield_tl", "dhfr2")
        fp32_peak = 148 * 128 * 2 * sm_clk * 1e6 / 1e12
        line = {
            "metric": MD_METRIC, "value": ns_per_day(ms_step, world), "unit": "ns/day", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms_step, "ms_per_induce": ms_ind,
            "pcg_iterations": float(np.mean(iters)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 pair math + 2^32 fixed-point accumulation; f64 valence terms, positions and velocities",
            "data": "reference input deck example/dhfr2 parsed by our readers (blob tests/golden/dhfr2.npz); Maxwell velocities at 298 K, seeded",
            "config": {"workload": MD_WORKLOAD, "atoms": int(n), "parallelism": "single GPU" if world == 1 else f"replicas x{world}",
                       "l2": "flushed (256 MB write) between timed steps",
                       "timing": "CUDA events on the library stream around each MD step (one apx_md_steps(1) call), mean of steps, max over ranks",
                       "hot_path_only": False,
                       "terms": "multipole + polarization (PCG) + buffered 14-7 vdW every outer step; 8 valence terms on the inner level"},
            "md": {"temperature_K": temp, "epot": epot, "ekin": ekin, "list_rebuilds_in_timed_steps": int(rebuilds),
                   # medians, and the list-rebuild steps reported separately
                   "ms_per_step_median": float(np.median(ms_steps)), "ms_per_induce_median": float(np.median(ms_induce)),
                   "ms_steps": [round(float(m), 3) for m in ms_steps], "rebuilt": [int(r) for r in rebuilt],
                   "ms_list_rebuild_last": st.get("ms_list"), "solver_batch_misses": st.get("energy_retries"),
                   "ms_per_step_without_rebuild": float(np.mean([m for m, r in zip(ms_steps, rebuilt) if not r])) if not all(rebuilt) else None,
                   "ms_per_step_with_rebuild": float(np.mean([m for m, r in zip(ms_steps, rebuilt) if r])) if any(rebuilt) else None,
                   "batch": {"value": ns_per_day(ms_batch, world), "unit": "ns/day", "ms_per_step": ms_batch,
                             "note": f"{args.steps} steps in ONE apx_md_steps call, no L2 flush: the rate a production run sees"}},
            "e2e": {"value": ns_per_day(ms_e2e_step, world), "unit": "ns/day", "ms_per_step": ms_e2e_step,
                    "h2d_bytes_per_step": int(2 * x_md.nbytes), "d2h_bytes_per_step": int(2 * x_md.nbytes) + 8 * 8,
                    "list_rebuilds": int(reb_e2e),
                    "note": "one MD step per call through the C ABI with the integrator state in host buffers: positions + velocities "
                            "in from pinned host memory (apx_md_set_state), apx_md_steps(1), positions + velocities + the step "
                            "report back (apx_md_get_state); the same unit of work as `value`",
                    "force_eval": {"value": ns_per_day(ms_force_eval, world), "unit": "ns/day-equivalent at one evaluation per 2 fs",
                                   "ms_per_call": ms_force_eval,
                                   "note": "plugin call energy(energy+grad) with host buffers (positions in, gradient out): a force "
                                           "evaluation, not an MD step"}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline_block(uf_flops, uf_bytes, uf_ms, fp32_peak, hbm_peak, peak_src, sm_clk, npairs, traffic),
            "vdw": {"ms_ehal_kernel": st["ms_ehal"], "directed_row_entries": int(st["nverlet_vdw"])},
            "wall_s_timed_region": t_wall,
        }
        if not args.no_cpu:
            ms_cpu, ms_cpu_ind, desc, kind = cpu_dynamics_sample(system)
            line["cpu_baseline"] = {"value": ns_per_day(ms_cpu), "unit": "ns/day", "cores": CPU_THREADS if kind == "reference" else 1, "kind": kind, "sample": desc,
                                    "ms_per_step": ms_cpu, "ms_per_induce": ms_cpu_ind,
                                    "note": "the reference EXECUTABLE cannot be linked here (no Fortran compiler); kind 'reference' = its own "
                                            "pair functions and PME translation unit compiled in place and run serially, as its host build "
                                            "does (OpenACC pragmas ignored by g++); reported, not a target"}
        if not args.no_cpu and not args.no_ref_cuda and world == 1:
            line["ref_cuda"] = ref_cuda_sample(ours_induce_ms=float(np.median(ms_induce)), ours_md_step_ms=float(ms_step))
    a.close()
    del flush
    torch.cuda.empty_cache()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    # ---- the decomposed 1 M-atom box on the same N GPUs (strong scaling), after the dhfr2 legs released the GPUs
    strong = None if args.no_strong else strong_scaling_leg(args, rank, world)
    if rank == 0:
        line["strong_scaling"] = strong if strong is not None else {"unavailable": "skipped (--no-strong)"}
        print(json.dumps(line))


STRONG_WORKLOAD = "water1m"


def run_strong_child(args, rank, world, local_rank):
    """Child of the strong-scaling leg: ONE ~1 M-atom water box (BASELINE.json configs[3]) spatially decomposed over the N GPUs
    of this job (csrc/dist.cu: z-slabs, halo exchange of dipoles per CG iteration, slab FFT with all-to-all transposes), one
    energy+gradient evaluation of the electrostatics path per step; N = 1 is the single-GPU path on the same box."""
    import ctypes as C
    import torch
    if os.environ.get("APX_BENCH_STRONG_FAKE"):
        # tests/test_bench_strong_spawn.py (CPU, gloo, 2 ranks under torchrun): only the plumbing of this leg -- the children's own
        # rendezvous beside the parent job's, the cleaned environment, the STRONG line back to rank 0 -- without a GPU
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t)
        if rank == 0:
            print("STRONG " + json.dumps({"fake": True, "n_gpus": world, "sum_of_ranks": float(t)}), flush=True)
        dist.destroy_process_group()
        return
    from tinker_gpu_b200.amoeba import calc
    from tinker_gpu_b200.distributed import nccl_context
    torch.cuda.set_device(local_rank)
    import torch.distributed as dist
    from tinker_gpu_b200.amoeba import Amoeba
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), rank=rank, world_size=world)
    system = make_system(STRONG_WORKLOAD)
    t0 = time.perf_counter()
    a = nccl_context(system, "mixed") if world > 1 else Amoeba(system, "mixed", device=local_rank)
    t_create = time.perf_counter() - t0
    ext = torch.cuda.ExternalStream(a.lib.apx_stream(a.ctx), device=local_rank)
    a.lib.apx_dist_profile.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(args.warmup, 3)
    for _ in range(W):
        if a.lib.apx_energy(a.ctx, calc.v4, None) != 0:
            raise SystemExit("apx_energy failed: " + a.lib.apx_last_error().decode())
    a.synchronize()
    a.lib.apx_dist_profile(a.ctx, 1, None)
    barrier()
    ms_steps, ms_induce, ms_uf, iters = [], [], [], []
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(ext)
        rc = a.lib.apx_energy(a.ctx, calc.v4, None)
        e1.record(ext)
        e1.synchronize()
        if rc != 0:
            raise SystemExit("apx_energy failed: " + a.lib.apx_last_error().decode())
        ms_steps.append(e0.elapsed_time(e1))
        st = a.stats()
        ms_induce.append(st["ms_induce"])
        ms_uf.append(st["ms_ufield_real"])
        iters.append(st["pcg_iterations"])
    barrier()
    prof = (C.c_double * 8)()
    a.lib.apx_dist_profile(a.ctx, 0, prof)
    vals = [float(np.mean(ms_steps)), float(np.mean(ms_induce)), float(np.mean(ms_uf))] + [float(prof[k]) / args.steps for k in range(4)]
    if world > 1:
        t = torch.tensor(vals, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = [float(v) for v in t.tolist()]
    info = a.dist_info() if world > 1 else None
    if rank == 0:
        st = a.stats()
        it = float(np.mean(iters))
        napp = max(1.0, float(prof[5]) / args.steps)      # PME round trips per step: iterations + r0 + permanent + converged dipoles
        p2p = os.environ.get("APX_DIST_P2P", "3")
        out = {
            "workload": WORKLOADS[STRONG_WORKLOAD][4] + ", energy+gradient of the electrostatics path", "atoms": int(system.n),
            "n_gpus": world, "scaling": "strong", "steps": args.steps, "warmup": W,
            "parallelism": "single GPU" if world == 1 else f"spatial decomposition over {world} GPUs: z-slabs, halo exchange of dipoles per CG "
                                                           "iteration, slab PME FFT with all-to-all transposes",
            "transport": "none" if world == 1 else {"0": "NCCL grouped send/recv", "1": "CUDA-IPC peer windows, copy engines",
                                                    "2": "CUDA-IPC peer windows, fused push/pull kernels (k_xfer)",
                                                    "3": "direct: one kernel per exchange writes strided / per-atom messages into the "
                                                         "peers' CUDA-IPC-registered buffers over NVLink (k_dxchg), flag-based scalar "
                                                         "all-reduce (k_dar); NCCL only for the start-up handshake and the force reduction"
                                                    }.get(p2p, p2p),
            "ms_per_step": vals[0], "ms_per_induce": vals[1], "pcg_iterations": it,
            "timing": "CUDA events on the library stream around each apx_energy call, barrier before every step, mean of steps, max over "
                      "ranks; working set 1.2 GB >> L2, no flush needed",
            "breakdown_ms_per_step": {"real_space_operator_per_launch": vals[2], "halo_exchange": vals[3], "fft_forward_with_transpose": vals[4],
                                      "fft_inverse_with_transpose": vals[5], "scalar_allreduce": vals[6]},
            "per_operator_application_us": None if world == 1 else {
                "real_space_operator": 1e3 * vals[2], "fft_forward_with_transpose": 1e3 * vals[4] / napp,
                "fft_inverse_with_transpose": 1e3 * vals[5] / napp, "halo_exchange": 1e3 * vals[3] / max(1.0, float(prof[4]) / args.steps)},
            "pairs_within_cutoff": int(st["npairs_m"]) if st["npairs_m"] > 0 else None,
            "create_s": t_create,
        }
        if info is not None:
            out["rank0_slab"] = {"atoms_owned": int(info["a1"] - info["a0"]), "halo_atoms": int(info["halo_atoms"]), "planes": int(info["planes"]),
                                 "halo_planes": [int(info["halo_lo"]), int(info["halo_hi"])]}
        print("STRONG " + json.dumps(out), flush=True)
    a.close()
    if world > 1:
        dist.destroy_process_group()


def strong_scaling_leg(args, rank, world):
    """The decomposed 1 M-atom curve north_star names, measured in EVERY bench line: child processes -- one per
    rank, their own rendezvous -- so that a hang or a crash of the decomposed path can cost this block but never the line."""
    env = dict(os.environ)
    env["MASTER_ADDR"] = env.get("MASTER_ADDR", "127.0.0.1")
    env["MASTER_PORT"] = str(int(env.get("MASTER_PORT", "29500")) + 23)
    env["RANK"], env["WORLD_SIZE"] = str(rank), str(world)
    env.setdefault("LOCAL_RANK", "0")
    for k in ("TORCHELASTIC_RUN_ID", "TORCHELASTIC_RESTART_COUNT", "TORCHELASTIC_MAX_RESTARTS", "TORCHELASTIC_USE_AGENT_STORE"):
        env.pop(k, None)
    cmd = [sys.executable, os.path.abspath(__file__), "--strong-child", "--gpus", str(world), "--steps", str(max(3, min(args.steps, 10))),
           "--warmup", "3"]
    limit = 420
    try:
        r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=limit)
    except subprocess.TimeoutExpired:
        return {"unavailable": f"strong-scaling child exceeded {limit} s"}
    except Exception as e:      # noqa: BLE001
        return {"unavailable": f"strong-scaling child could not start: {e}"}
    if rank != 0:
        return None
    for ln in (r.stdout or "").splitlines():
        if ln.startswith("STRONG "):
            try:
                return json.loads(ln[7:])
            except Exception:      # noqa: BLE001
                pass
    tail = ((r.stderr or "") + (r.stdout or "")).strip().splitlines()[-1:] or [""]
    return {"unavailable": f"strong-scaling child exit {r.returncode}: {tail[0][:300]}"}


def run_reference(args, rank, world):
    import tinker_gpu_b200 as tg
    if rank != 0:
        return
    system = tg.load_system(os.path.join(GOLDEN, "dhfr2.npz"))
    res, warm, cap = timed_cpu_steps(lambda: cpu_oracle_sample(system), args)
    steps = len(res)
    ms_step = float(np.mean([r[0] for r in res]))
    ms_ind = [r[1] for r in res]
    desc, kind = res[-1][2], res[-1][3]
    val = ns_per_day(ms_step)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "ns/day", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": ms_step, "ms_per_induce": float(np.mean(ms_ind)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "reference input deck example/dhfr2 (blob tests/golden/dhfr2.npz)",
        "config": {"workload": WORKLOAD, "steps_requested": args.steps, "warmup_requested": args.warmup, "cap": cap, "note": "CPU arm: the reference executable needs gfortran (absent); its operators compiled in place (oracle/_ref) or the oracle port are timed instead"},
        "cpu_baseline": {"value": val, "unit": "ns/day", "cores": CPU_THREADS if kind == "reference" else 1, "kind": kind, "sample": desc},
        "e2e": {"value": val, "unit": "ns/day", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args, rank, world, local_rank):
    import torch
    import tinker_gpu_b200 as tg
    from tinker_gpu_b200.amoeba import Amoeba, calc

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    system = make_system(args.workload)
    decomposed = world > 1 and args.workload != "dhfr2" and not args.replicas
    replicas = 1 if decomposed else world
    if decomposed:
        from tinker_gpu_b200.distributed import nccl_context
        a = nccl_context(system, "mixed", vdw=args.vdw)
    else:
        a = Amoeba(system, "mixed", device=local_rank, vdw=args.vdw)
    ext = torch.cuda.ExternalStream(a.lib.apx_stream(a.ctx), device=local_rank)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    vers = calc.v4
    rng = np.random.default_rng(1234 + (0 if decomposed else rank))     # decomposed: every rank passes the same positions
    xyz0 = np.array(system.xyz)
    # per-step host inputs for the e2e leg: the system drifts rigidly by 0.08 A per step (plus thermal-size noise), so the
    # list check is real AND the list is rebuilt every ~13 steps, as in a dynamics run (buffer/2 = 1 A criterion), while
    # the physics -- and with it the solver's iteration count -- stays that of the reference deck.  (Uncorrelated
    # per-atom drifts would stretch every bond and change the problem being solved.)
    vel = np.array([0.06, 0.04, 0.035])
    nframes = 2 + args.steps
    frames = [xyz0 + vel * float(j) + rng.normal(scale=0.002, size=xyz0.shape) for j in range(nframes)]   # outside the timed region

    def frame(j):
        return frames[j % nframes]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return a.lib.apx_energy(a.ctx, vers, None)

    def step_e2e(j):
        a.set_positions(frame(j))
        r = a.lib.apx_energy(a.ctx, vers, None)
        a.gradient()
        return r

    W = max(args.warmup, 3)      # the timing rules' minimum
    for _ in range(W):
        step_resident()
    a.synchronize()

    # ---- resident leg (value): per-step CUDA events on the library stream, L2 flushed between steps
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    a.stats_reset()
    ms_steps, ms_induce, ms_uf, iters = [], [], [], []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        rc = step_resident()
        e1.record(ext)
        e1.synchronize()
        if rc != 0:
            raise SystemExit("apx_energy failed: " + a.lib.apx_last_error().decode())
        ms_steps.append(e0.elapsed_time(e1))
        st = a.stats()
        ms_induce.append(st["ms_induce"])
        ms_uf.append(st["ms_ufield_real"])
        iters.append(st["pcg_iterations"])
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = a.stats()["kernel_launches"]
    ms_step = float(np.mean(ms_steps))

    # ---- e2e leg: host positions in, gradient out, every step
    for j in range(2):
        step_e2e(j)
    barrier()
    rebuilds0 = a.stats()["list_rebuilds"]
    ms_e2e = []
    for j in range(2, 2 + args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        step_e2e(j)
        e1.record(ext)
        e1.synchronize()
        ms_e2e.append(e0.elapsed_time(e1))
    barrier()
    clocks = sampler.stop() if rank == 0 else None      # sampled over both timed regions (resident + e2e)
    ms_e2e_step = float(np.mean(ms_e2e))
    rebuilds = a.stats()["list_rebuilds"] - rebuilds0

    # max over ranks (replicas): the job advances at the pace of the slowest replica
    if dist is not None:
        t = torch.tensor([ms_step, ms_e2e_step, float(np.mean(ms_induce))], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, ms_e2e_step, ms_ind = (float(v) for v in t.tolist())
    else:
        ms_ind = float(np.mean(ms_induce))

    if rank == 0:
        st = a.stats()
        hbm_peak, sm_max, peak_src = load_peaks()
        n = system.n
        K = int(np.prod(system.nfft))
        npairs = max(1, st["npairs_m"])
        uf_ms = float(np.mean(ms_uf)) if ms_uf else 0.0
        # real-space ufield row kernel: algorithmic bytes = read (pos,pdamp,thole,ud,up) + rmw (field d,p) per atom
        # + one 4-byte neighbor index per directed pair
        uf_bytes = operator_bytes(n, npairs)
        uf_flops = 130.0 * npairs
        achieved_gbs = uf_bytes / (uf_ms * 1e-3) / 1e9 if uf_ms > 0 else 0.0
        sm_clk = (clocks or {}).get("sm_mhz") or sm_max
        traffic = load_ncu_traffic("k_ufield_tl", args.workload)
        fp32_peak = 148 * 128 * 2 * sm_clk * 1e6 / 1e12
        wl_desc = WORKLOADS[args.workload][4]
        metric = METRIC if args.workload == "dhfr2" else METRIC.replace("AMOEBA DHFR 23.5k atoms", wl_desc.split(",")[0])
        par = "single GPU" if world == 1 else (f"spatial decomposition over {world} GPUs (z-slabs, peer-memory halo exchange + slab FFT all-to-all transposes, dist.cu)"
                                                if decomposed else f"replicas x{world}")
        if args.vdw:
            metric = metric.replace("electrostatics hot path only", "electrostatics hot path + buffered 14-7 vdW")
        line = {
            "metric": metric, "value": ns_per_day(ms_step, replicas), "unit": "ns/day", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms_step, "ms_per_induce": ms_ind,
            "pcg_iterations": float(np.mean(iters)), "higher_is_better": True, "scaling": "strong" if decomposed else "weak",
            "vs_baseline": None,
            "dtype": "f32 pair math + 2^32 fixed-point / f64 accumulation",
            "data": ("reference input deck example/dhfr2 parsed by our readers (blob tests/golden/dhfr2.npz)" if args.workload == "dhfr2"
                     else "synthetic: committed reference cell replicated at run time (BASELINE.md section 4), seeded jitter"),
            "config": {"workload": wl_desc + ", energy+gradient" * (args.workload != "dhfr2"), "atoms": int(n), "parallelism": par,
                       "l2": "flushed (256 MB write) between timed steps",
                       "timing": "CUDA events on the library stream around each step, mean of steps, max over ranks",
                       "hot_path_only": True},
            "e2e": {"value": ns_per_day(ms_e2e_step, replicas), "unit": "ns/day", "ms_per_step": ms_e2e_step,
                    "h2d_bytes_per_step": int(xyz0.nbytes), "d2h_bytes_per_step": int(xyz0.nbytes) + 136,
                    "list_rebuilds": rebuilds,
                    "note": "positions drift 0.08 A/step: neighbor-list rebuilds happen inside the timed region"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline_block(uf_flops, uf_bytes, uf_ms, fp32_peak, hbm_peak, peak_src, sm_clk, npairs, traffic),
            "wall_s_timed_region": t_wall,
        }
        if args.vdw:
            line["vdw"] = {"ms_ehal_kernel": st["ms_ehal"], "directed_row_entries": int(st["nverlet_vdw"]),
                           "cutoff": float(system.vdw.cutoff), "note": "ehal runs on its own stream beside induce()"}
        if not args.no_cpu and args.workload == "dhfr2":
            ms_cpu, ms_cpu_ind, desc, kind = cpu_oracle_sample(system)
            line["cpu_baseline"] = {"value": ns_per_day(ms_cpu), "unit": "ns/day", "cores": CPU_THREADS if kind == "reference" else 1, "kind": kind, "sample": desc,
                                    "ms_per_step": ms_cpu, "ms_per_induce": ms_cpu_ind,
                                    "note": "vectorised-numpy port on one core, about two orders of magnitude slower than the reference's "
                                            "compiled host build would be (it cannot be linked here: no Fortran compiler); reported, not a target"}
        if not args.no_cpu and not args.no_ref_cuda and world == 1 and not args.vdw:
            if args.workload == "dhfr2":
                line["ref_cuda"] = ref_cuda_sample(ours_energy_ms=float(ms_step))
            else:
                try:
                    system.xyz = xyz0
                    a.set_positions(xyz0)
                    e_ours = float(a.energy(calc.v0)["esum"])
                except Exception:      # noqa: BLE001
                    e_ours = None
                line["ref_cuda"] = ref_cuda_sample(ours_energy_ms=float(ms_step), system=system, ours_esum=e_ours)
        print(json.dumps(line))
    a.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the reference-CUDA comparator leg (oracle/_ref/libref_cuda.so)")
    ap.add_argument("--workload", default="dhfr2", choices=sorted(WORKLOADS))
    ap.add_argument("--vdw", action="store_true", help="also evaluate the buffered 14-7 vdW term (SURVEY 8f rank 1) in every step")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (decomposed 1 M-atom box in child processes)")
    ap.add_argument("--strong-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--replicas", action="store_true", help="N > 1: independent replicas also for the large workloads")
    ap.add_argument("--mode", default=None, choices=["dynamics", "energy"],
                    help="dynamics: full MD steps on the device integrator (default for dhfr2, BASELINE configs[1]); "
                         "energy: one electrostatics energy+gradient per step (default for the large boxes)")
    args = ap.parse_args()
    if args.mode is None:
        args.mode = "dynamics" if (args.workload == "dhfr2" and not args.vdw) else "energy"
    if args.mode == "dynamics" and args.workload != "dhfr2":
        ap.error("--mode dynamics is built for the dhfr2 workload")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.strong_child:
        run_strong_child(args, rank, world, local_rank)
        return
    if args.impl == "reference":
        (run_reference_dynamics if args.mode == "dynamics" else run_reference)(args, rank, world)
    elif args.mode == "dynamics":
        run_dynamics(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
