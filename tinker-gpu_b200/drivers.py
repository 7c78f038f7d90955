"""Fortran-free ANALYZE and TESTGRAD driver shells over the C ABI.

They produce the reference's transcripts for the terms this repository evaluates (atomic multipoles,
polarization, buffered 14-7 vdW): same prompts' options, same line formats as

    xAnalyzeE / xAnalyzeM / xAnalyzeV      src/xanalyze.cpp:40-137, 139-386, 390-420
    xTestgrad                              src/xtestgrad.cpp:133-147, 207-317

so that the reference's test/ref/*.txt transcripts can be diffed against them term by term.  The numbers
come from the CUDA back end (tinker-gpu_b200/amoeba.py); the formatting functions take plain values and
are tested on the CPU against lines of the reference's own transcripts.

    python -m tinker_gpu_b200.drivers analyze  system.xyz [-k system.key] [EMV]
    python -m tinker_gpu_b200.drivers testgrad system.xyz [-k system.key] [Y|N analytical] [Y|N numerical] [step]
    python -m tinker_gpu_b200.drivers dynamic  system.xyz [-k system.key] nstep dt(fs) dtsave(ps) mode [kelvin]

DYNAMIC (xDynamic, src/xdynamic.cpp:17-196; mdsave, tinker/source/mdsave.f) runs on the device integrator of
csrc/md.cu: INTEGRATOR VERLET or RESPA, ensembles (1) NVE and (2) NVT with the Bussi thermostat; the Beeman default,
the Nose-Hoover / Langevin-piston integrators and the barostats of modes 3-4 are not built and are refused.  It
appends frames to <name>.arc, rewrites <name>.dyn at every save and restarts from an existing <name>.dyn.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np

DEBYE = 4.803206802         # tinker/source/units.f:97
GASCONST = 1.9872042586e-3  # units.f:84
PRESCON = 6.85684112e4      # units.f:98


# ------------------------------------------------------------------------------------------------ formatting
def format_energy_breakdown(esum, rows):
    """rows: [(label, energy, count)].  xAnalyzeE, src/xanalyze.cpp:47-52 and the per-term prints."""
    out = ["", " Total Potential Energy :        %16.4f Kcal/mole" % esum, "",
           " Energy Component Breakdown :           Kcal/mole        Interactions", ""]
    for label, e, cnt in rows:
        out.append(" %-29s %18.4f %16d" % (label, e, cnt))
    return "\n".join(out) + "\n"


def format_virial(vir, n, volume, use_bounds=True, temp=298):
    """xAnalyzeV, src/xanalyze.cpp:397-419."""
    v = np.asarray(vir, float).reshape(9)
    fmt = " %-36s%12.3f %12.3f %12.3f"
    out = ["", fmt % ("Internal Virial Tensor :", v[0], v[1], v[2]), fmt % ("", v[3], v[4], v[5]), fmt % ("", v[6], v[7], v[8])]
    pres = 0.0
    if use_bounds:
        pres_vir = -(v[0] + v[4] + v[8])
        pres = (3 * n * GASCONST * temp + pres_vir) * PRESCON / (3 * volume)
        pres_vir *= PRESCON / (3 * volume)
        out += ["", " Pressure (Temp %3d K) :            %13.3f Atmospheres" % (temp, pres),
                " Pressure From Virial               %13.3f Atmospheres" % pres_vir]
    else:
        out += ["", " Pressure (Temp %3d K) :            %13.3f Atmospheres" % (temp, pres)]
    return "\n".join(out) + "\n"


def moments(xyz, mass, rpole, uind):
    """Net charge, dipole (Debye), traceless quadrupole (Buckingham) about the centre of mass and its principal
    values: xAnalyzeMoments, src/xanalyze.cpp:139-342 (atomic-multipole branch; rpole in MPL_PME order)."""
    x = np.asarray(xyz, float)
    m = np.asarray(mass, float)
    cm = (x * m[:, None]).sum(0) / m.sum() if m.sum() != 0 else np.zeros(3)
    r = x - cm
    c = rpole[:, 0]
    d = rpole[:, 1:4] + uind
    netchg = float(c.sum())
    dpl = (r * c[:, None]).sum(0) + d.sum(0)
    q = np.einsum("i,ia,ib->ab", c, r, r) + np.einsum("ia,ib->ab", r, d) + np.einsum("ia,ib->ba", r, d)
    q = 1.5 * (q - np.eye(3) * np.trace(q) / 3.0)
    # atomic quadrupoles, MPL_PME order xx, xy, xz, yy, yz, zz
    aq = np.zeros((3, 3))
    aq[0, 0], aq[0, 1], aq[0, 2] = rpole[:, 4].sum(), rpole[:, 5].sum(), rpole[:, 6].sum()
    aq[1, 1], aq[1, 2], aq[2, 2] = rpole[:, 7].sum(), rpole[:, 8].sum(), rpole[:, 9].sum()
    aq[1, 0], aq[2, 0], aq[2, 1] = aq[0, 1], aq[0, 2], aq[1, 2]
    q = (q + 3.0 * aq) * DEBYE
    dpl = dpl * DEBYE
    return dict(netchg=netchg, dipole=dpl, netdpl=float(np.linalg.norm(dpl)), quadrupole=q, principal=np.sort(np.linalg.eigvalsh(q)))


def format_moments(mo):
    """xAnalyzeM, src/xanalyze.cpp:346-366."""
    q, d, p = mo["quadrupole"], mo["dipole"], mo["principal"]
    out = ["", " Total Electric Charge :%12s%13.5f Electrons" % ("", mo["netchg"]), "",
           " Dipole Moment Magnitude :%10s%13.3f Debye" % ("", mo["netdpl"]), "",
           " Dipole X,Y,Z-Components :%10s%13.3f%13.3f%13.3f" % ("", d[0], d[1], d[2]), "",
           " Quadrupole Moment Tensor :%9s%13.3f%13.3f%13.3f" % ("", q[0, 0], q[0, 1], q[0, 2]),
           "      (Buckinghams)%17s%13.3f%13.3f%13.3f" % ("", q[1, 0], q[1, 1], q[1, 2]),
           "%36s%13.3f%13.3f%13.3f" % ("", q[2, 0], q[2, 1], q[2, 2]), "",
           " Principal Axes Quadrupole :%8s%13.3f%13.3f%13.3f" % ("", p[0], p[1], p[2])]
    return "\n".join(out) + "\n"


_GRAD_FMT = {
    8: ("\n  Type    Atom %8s dE/dX %9s dE/dY %9s dE/dZ %9s Norm\n", "\n %s%8d %16.8f%16.8f%16.8f%16.8f"),
    6: ("\n  Type      Atom %9s dE/dX %7s dE/dY %7s dE/dZ %9s Norm\n", "\n %s%10d   %14.6f%14.6f%14.6f  %14.6f"),
    4: ("\n  Type      Atom %12s dE/dX %5s dE/dY %5s dE/dZ %8s Norm\n", "\n %s%10d       %12.4f%12.4f%12.4f  %12.4f"),
}


def format_testgrad(energy, anlyt=None, numer=None, digits=4):
    """xTestgrad's transcript, src/xtestgrad.cpp:133-147 (formats) and 246-309 (layout)."""
    header, row = _GRAD_FMT[digits if digits in _GRAD_FMT else 4]
    out = []
    if anlyt is not None:
        out.append("\n Total Potential Energy :%*.*f Kcal/mole\n\n" % (20 + digits, digits, energy))
    out.append(header % ("", "", "", ""))
    n = len(anlyt) if anlyt is not None else len(numer)
    for i in range(n):
        for tag, g in (("Anlyt", anlyt), ("Numer", numer)):
            if g is not None:
                out.append(row % (tag, i + 1, g[i][0], g[i][1], g[i][2], float(np.linalg.norm(g[i]))))
    out.append("\n\n Total Gradient Norm and RMS Gradient per Atom :\n")
    for title, scale in (("Total Gradient Norm Value", 1.0), ("RMS Gradient over All Atoms", 1.0 / np.sqrt(n))):
        for tag, g in (("Anlyt", anlyt), ("Numer", numer)):
            if g is not None:
                out.append("\n %s      %-30s%*.*f" % (tag, title, 13 + digits, digits, float(np.linalg.norm(g)) * scale))
        out.append("\n")
    return "".join(out)


# ------------------------------------------------------------------------------------------------ drivers
def analyze(system, options="E", precision="mixed", device=0, out=sys.stdout):
    """`tinker9 analyze xyz EMV` for the terms in this repository; returns the dict of numbers it printed."""
    from .amoeba import Amoeba, calc
    opts = options.upper()
    a = Amoeba(system, precision, device=device, vdw=system.vdw is not None)
    res = {}
    try:
        if "E" in opts:
            r = a.energy(calc.v3)
            rows = []
            if system.vdw is not None:
                rows.append(("Van der Waals", r["ev"], r["nev"]))
            if system.use_mpole:
                rows.append(("Atomic Multipoles", r["em"], r["nem"]))
            if system.use_polar:
                rows.append(("Polarization", r["ep"], r["nep"]))
            out.write(format_energy_breakdown(r["esum"], rows))
            res["E"] = r
        if "M" in opts:
            a.mpoleInit()
            rp = a.rpole()
            u = a.induce()[0] if system.use_polar else np.zeros((system.n, 3))
            mass = system.mass if getattr(system, "mass", None) is not None else np.ones(system.n)
            mo = moments(system.xyz, mass, rp, u)
            out.write(format_moments(mo))
            res["M"] = mo
        if "V" in opts:
            r = a.energy(calc.v6)
            out.write(format_virial(r["virial"], system.n, system.volume))
            res["V"] = r
    finally:
        a.close()
    return res


def testgrad(system, analytical=True, numerical=False, eps=1.0e-5, digits=4, precision="double", device=0, out=sys.stdout, atoms=None):
    """`tinker9 testgrad xyz Y/N Y/N eps`: analytical gradient and/or central differences of the energy
    (src/xtestgrad.cpp:185-205 moves one coordinate by -eps/2 and +eps/2).  `atoms` limits the numerical part."""
    from .amoeba import Amoeba, calc
    a = Amoeba(system, precision, device=device, vdw=system.vdw is not None)
    try:
        x0 = np.array(system.xyz, float)
        r = a.energy(calc.v4 if analytical else calc.v0)
        ga = r.get("grad") if analytical else None
        gn = None
        if numerical:
            idx = range(system.n) if atoms is None else atoms
            gn = np.zeros((system.n, 3))
            for i in idx:
                for c in range(3):
                    x = x0.copy()
                    x[i, c] -= 0.5 * eps
                    a.set_positions(x)
                    e0 = a.energy(calc.v0)["esum"]
                    x[i, c] += eps
                    a.set_positions(x)
                    e1 = a.energy(calc.v0)["esum"]
                    gn[i, c] = (e1 - e0) / eps
            a.set_positions(x0)
        out.write(format_testgrad(r["esum"], ga, gn, digits))
        return dict(energy=r["esum"], anlyt=ga, numer=gn)
    finally:
        a.close()


# ------------------------------------------------------------------------------------------------ dynamic
EKCAL = 418.4               # units.f:90


def md_options(key):
    """Keywords xDynamic / mdinit.f read: INTEGRATOR, THERMOSTAT, TAU-TEMPERATURE, RESPA-INNER (a count), RANDOMSEED."""
    def word(kw, default):
        v = key.get(kw) if key is not None else None
        return (v.split()[0].upper() if v and v.split() else default)

    def num(kw, default):
        v = key.get(kw) if key is not None else None
        return float(v.split()[0]) if v and v.split() else default
    return dict(integrator=word("INTEGRATOR", "BEEMAN"), thermostat=word("THERMOSTAT", "BUSSI"),
                tautemp=num("TAU-TEMPERATURE", 0.2), nrespa=int(num("RESPA-INNER", 0)),
                seed=int(num("RANDOMSEED", 123456789)))


def respa_inner_steps(dt):
    """mdstuf::nrespa default of mdinit.f:69: max(1, nint(dt / 0.0005 ps)) -- inner steps of about 0.5 fs (dt = 1 fs gives
    2, which test/respa.cpp:39 checks; the dhfr2 benchmark's 2 fs gives 4).  RESPA-INNER overrides it."""
    return max(1, int(np.floor(dt / 0.0005 + 0.5)))


def maxwell_velocities(mass, kelvin, seed, nfree=None):
    """Starting velocities of mdinit.f:411-440: Maxwell-Boltzmann speeds in random directions, centre-of-mass motion
    removed (mdrest.f), then scaled to exactly `kelvin`.  A/ps.  (numpy's generator, not Tinker's `random`: the
    distribution is the reference's, the stream is not.)"""
    rng = np.random.default_rng(seed)
    m = np.asarray(mass, float)
    n = len(m)
    ok = m > 0
    sigma = np.sqrt(0.8314462618 * kelvin / np.where(ok, m, 1.0))      # units::boltzmann in g A^2/ps^2/mol/K
    v = rng.normal(size=(n, 3)) * np.where(ok, sigma, 0.0)[:, None]
    v -= (m[:, None] * v).sum(0) / m.sum()
    v[~ok] = 0.0
    nfree = 3 * n - 3 if nfree is None else nfree
    ek = 0.5 * float((m[:, None] * v * v).sum()) / EKCAL
    t = 2.0 * ek / (nfree * GASCONST)
    if t > 0 and kelvin > 0:
        v *= np.sqrt(kelvin / t)
    return v


def format_md_frame(istep, dt, epot, eksum, box6, isave, arcname, digits=4):
    """The block mdsave.f:86-113, 116-122, 235-236, 268-269 prints at every trajectory save."""
    w = {4: (15, 4), 6: (17, 6), 8: (19, 8)}[4 if digits < 6 else 6 if digits < 8 else 8]
    s = f"\n Instantaneous Values for Frame Saved at{istep:10d} Dynamics Steps\n"
    s += f"\n Current Time{'':8s}{istep * dt:{w[0]}.{w[1]}f} Picosecond\n"
    s += f" Current Potential{'':3s}{epot:{w[0]}.{w[1]}f} Kcal/mole\n"
    s += f" Current Kinetic{'':5s}{eksum:{w[0]}.{w[1]}f} Kcal/mole\n"
    if box6 is not None:
        s += " Lattice Lengths" + " " * 6 + "".join(f"{v:14.6f}" for v in box6[:3]) + "\n"
        s += " Lattice Angles" + " " * 7 + "".join(f"{v:14.6f}" for v in box6[3:]) + "\n"
    s += f" Frame Number{'':13s}{isave:10d}\n"
    s += f" Coordinate File{'':13s}{arcname}\n"
    return s


def format_md_performance(nsday, wall_s, nstep, updates, dt_fs, natoms):
    """src/xdynamic.cpp:181-189."""
    s = "\n"
    s += " %-14s%-9s%18.4f\n" % ("Performance:", "ns/day", nsday)
    s += " %-14s%-9s%18.4f\n" % ("", "Wall Time", wall_s)
    s += " %-14s%-9s%18d\n" % ("", "Steps", nstep)
    s += " %-14s%-9s%18d\n" % ("", "Updates", updates)
    s += " %-14s%-9s%18.4f\n" % ("", "Time Step", dt_fs)
    s += " %-14s%-9s%18d\n" % ("", "Atoms", natoms)
    return s


def dynamic(system, nstep, dt_fs=1.0, dtsave_ps=0.1, mode=1, kelvin=298.0, integrator="RESPA", thermostat="BUSSI",
            tautemp=0.2, nrespa=0, seed=123456789, basename=None, precision="mixed", device=0, out=sys.stdout,
            velocities=None):
    """`tinker9 dynamic xyz nstep dt dtsave mode [kelvin]` (src/xdynamic.cpp:17-196) for INTEGRATOR VERLET / RESPA in the
    NVE (mode 1) and NVT-Bussi (mode 2) ensembles.  Returns dict(ns_day, reports, xyz, vel)."""
    import os
    import time
    from .amoeba import Amoeba
    from .tinkerio import XYZ, append_arc_frame, read_dyn, write_dyn
    integrator = integrator.upper()
    if integrator not in ("VERLET", "RESPA"):
        raise NotImplementedError(f"integrator {integrator}: only VERLET and RESPA are built (csrc/md.cu)")
    if mode not in (1, 2):
        raise NotImplementedError("ensembles (3) NPH and (4) NPT need a barostat, which is not built")
    if mode == 2 and thermostat.upper() != "BUSSI":
        raise NotImplementedError(f"thermostat {thermostat}: only BUSSI is built")
    if system.valence is None or system.mass is None:
        raise ValueError("dynamics needs the valence terms and the atomic masses of the system")
    dt = dt_fs * 0.001
    tautemp = max(tautemp, dt)                                   # xdynamic.cpp:63
    iwrite = max(1, int(round(dtsave_ps / dt)))
    nrespa = (nrespa if nrespa > 0 else respa_inner_steps(dt)) if integrator == "RESPA" else 1
    nfree = 3 * system.n - (3 if mode == 2 else 0)               # mdinit.f:343-366, periodic system without constraints
    lv = np.asarray(system.lvec, float)
    lens = np.linalg.norm(lv, axis=1)
    ang = lambda u, w: float(np.degrees(np.arccos(np.clip(np.dot(u, w) / (np.linalg.norm(u) * np.linalg.norm(w)), -1, 1))))  # noqa: E731
    box6 = [lens[0], lens[1], lens[2], ang(lv[1], lv[2]), ang(lv[0], lv[2]), ang(lv[0], lv[1])]
    xyz0 = np.array(system.xyz, float)
    vel0 = velocities
    dynfile = basename + ".dyn" if basename else None
    if dynfile and os.path.isfile(dynfile):                      # mdinit.f:300-310: restart
        d = read_dyn(dynfile)
        xyz0, vel0 = d["xyz"], d["vel"]
    elif vel0 is None:
        vel0 = maxwell_velocities(system.mass, kelvin if mode == 2 else 0.0, seed, nfree) if mode == 2 else np.zeros_like(xyz0)
    a = Amoeba(system, precision, device=device, vdw=system.vdw is not None, valence=True)
    reports = []
    try:
        a.set_positions(xyz0)
        a.md_init(system.mass, vel0, dt=dt, nrespa=nrespa, thermostat="BUSSI" if mode == 2 else None, kelvin=kelvin,
                  tautemp=tautemp, nfree=nfree, seed=seed)
        t0 = time.perf_counter()
        done = isave = 0
        while done < nstep:
            k = min(iwrite - done % iwrite, nstep - done)
            r = a.md_steps(k)
            done += k
            if done % iwrite == 0:
                isave += 1
                arc = (basename + ".arc") if basename else "(not written)"
                out.write(format_md_frame(done, dt, r.epot, r.ekin, box6 if system.use_ewald else None, isave, arc))
                reports.append(dict(step=done, epot=r.epot, ekin=r.ekin, temp=r.temp, pcg_iterations=r.pcg_iterations))
                if basename:
                    x, v = a.md_state()
                    names = system.names or ["X"] * system.n
                    bonds = system.bonds or [[] for _ in range(system.n)]
                    append_arc_frame(basename + ".arc", XYZ(system.n, system.title, names, x, system.types, bonds,
                                                            tuple(box6) if system.use_ewald else None))
                    g = a.gradient() + a.valence_gradient()
                    minv = np.where(system.mass > 0, 1.0 / np.where(system.mass > 0, system.mass, 1.0), 0.0)
                    write_dyn(dynfile, system.title, box6 if system.use_ewald else None, x, v, -EKCAL * g * minv[:, None])
        a.synchronize()
        wall = time.perf_counter() - t0
        x, v = a.md_state()
    finally:
        a.close()
    nsday = (dt_fs * nstep * 86400.0) / (wall * 1.0e6) if wall > 0 else 0.0      # xdynamic.cpp:179
    out.write(format_md_performance(nsday, wall, nstep, nstep // iwrite, dt_fs, system.n))
    return dict(ns_day=nsday, wall_s=wall, reports=reports, xyz=x, vel=v, nrespa=nrespa)


def main(argv=None):
    ap = argparse.ArgumentParser(prog="tinker_gpu_b200.drivers")
    ap.add_argument("program", choices=["analyze", "testgrad", "dynamic"])
    ap.add_argument("xyz")
    ap.add_argument("-k", "--key", default=None)
    ap.add_argument("rest", nargs="*")
    args = ap.parse_intermixed_args(argv)
    import tinker_gpu_b200 as tg
    s = tg.load_tinker(args.xyz, args.key)
    if args.program == "analyze":
        analyze(s, args.rest[0] if args.rest else "E")
    elif args.program == "dynamic":
        import os
        from .tinkerio import read_key
        rest = args.rest
        if len(rest) < 4:
            ap.error("dynamic needs: nstep dt(fs) dtsave(ps) mode [kelvin]")
        keypath = args.key or (os.path.splitext(args.xyz)[0] + ".key")
        o = md_options(read_key(keypath) if os.path.isfile(keypath) else None)
        dynamic(s, int(rest[0]), float(rest[1]), float(rest[2]), int(rest[3]), float(rest[4]) if len(rest) > 4 else 298.0,
                integrator=o["integrator"], thermostat=o["thermostat"], tautemp=o["tautemp"], nrespa=o["nrespa"], seed=o["seed"],
                basename=os.path.splitext(args.xyz)[0])
    else:
        yes = lambda k, d: (args.rest[k].upper().startswith("Y") if len(args.rest) > k else d)     # noqa: E731
        testgrad(s, yes(0, True), yes(1, False), float(args.rest[2]) if len(args.rest) > 2 else 1.0e-5)


if __name__ == "__main__":
    main()
