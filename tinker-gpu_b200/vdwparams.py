"""Buffered 14-7 van der Waals term: parameter assignment.

Restates what the reference obtains from the Fortran Tinker routines before evdwData() runs
(src/evdw.cpp:62-470): kvdw.f (class radii / well depths, combination rules, reduction factors,
special pairs), cutoffs.f (vdw-cutoff, taper), initprm.f defaults, evcorr.f (long-range correction)
and the 1-2..1-5 exclusion list of evdwData (src/evdw.cpp:196-262).  Only BUFFERED-14-7 is built --
the AMOEBA functional form; the other vdwtyp branches of evdw() belong to other force fields.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

TWOSIX = 1.122462048309372981        # 2**(1/6), tinker/source/math.f


@dataclass
class VdwTerm:
    """What evdwData uploads for Vdw::HAL (src/evdw.cpp:160-169, 380-441)."""
    ired: np.ndarray            # (n,) i32   atom the reduced site hangs off (itself for heavy atoms)   vdw::ired
    kred: np.ndarray            # (n,) f64   reduction factor (0 for heavy atoms)                       vdw::kred
    jvdw: np.ndarray            # (n,) i32   compressed class index                                     vdw::jvdw
    radmin: np.ndarray          # (nj,nj)    pair minimum-energy distances                              vdw::radmin
    epsilon: np.ndarray         # (nj,nj)    pair well depths                                           vdw::epsilon
    vexclude: np.ndarray        # (nv,2) i32 i<k pairs whose scale differs from 1
    vexclude_scale: np.ndarray  # (nv,)
    cutoff: float               # switchOff(Switch::VDW)
    taper: float                # switchCut(Switch::VDW)
    ghal: float = 0.12
    dhal: float = 0.07
    elrc_vol: float = 0.0       # long-range energy correction x box volume (src/evdw.cpp:443-452)
    vlrc_vol: float = 0.0
    list_buffer: float = 2.0


def switch_coeffs(cut, off):
    """c0..c5 of Tinker's multiplicative switch (tinker/source/switch.f:117-125)."""
    if not cut < off:
        return (0.0,) * 6
    d = (off - cut) ** 5
    off2, cut2 = off * off, cut * cut
    return (off * off2 * (off2 - 5.0 * off * cut + 10.0 * cut2) / d, -30.0 * off2 * cut2 / d,
            30.0 * (off2 * cut + off * cut2) / d, -10.0 * (off2 + 4.0 * off * cut + cut2) / d,
            15.0 * (off + cut) / d, -6.0 / d)


def evcorr1(jvdw, radmin, epsilon, cut, off, ghal, dhal, volume):
    """Long-range vdW energy and virial correction by numerical integration from the taper start to
    100 A (tinker/source/evcorr.f:193-372, BUFFERED-14-7 branch, vlambda = 1).  Returns (elrc, vlrc)."""
    c0, c1, c2, c3, c4, c5 = switch_coeffs(cut, off)
    nstep, rng = 2, 100.0
    ndelta = int(nstep * (rng - cut))
    rdelta = (rng - cut) / ndelta
    offset = cut - 0.5 * rdelta
    r = offset + np.arange(1, ndelta + 1) * rdelta
    r2, r3 = r * r, r * r * r
    r6, r7 = r3 * r3, r3 * r3 * r
    taper = c5 * r2 * r3 + c4 * r2 * r2 + c3 * r3 + c2 * r2 + c1 * r + c0
    dtaper = 5.0 * c5 * r2 * r2 + 4.0 * c4 * r3 + 3.0 * c3 * r2 + 2.0 * c2 * r + c1
    inside = r < off
    cls, cnt = np.unique(jvdw, return_counts=True)
    elrc = vlrc = 0.0
    for a in range(len(cls)):
        fi = 4.0 * math.pi * cnt[a]
        for b in range(a, len(cls)):
            fik = fi * cnt[b] * (0.5 if a == b else 1.0)
            rv, eps = radmin[cls[b], cls[a]], epsilon[cls[b], cls[a]]
            rv7 = rv ** 7
            rho = r7 + ghal * rv7
            tau = (dhal + 1.0) / (r + dhal * rv)
            tau7 = tau ** 7
            dtau = tau / (dhal + 1.0)
            gtau = eps * tau7 * r6 * (ghal + 1.0) * (rv7 / rho) ** 2
            e = eps * rv7 * tau7 * ((ghal + 1.0) * rv7 / rho - 2.0)
            de = -7.0 * (dtau * e + gtau)
            de = np.where(inside, de * (1.0 - taper) - e * dtaper, de)
            e = np.where(inside, e * (1.0 - taper), e)
            elrc += fik * float((e * rdelta * r2).sum())
            vlrc += fik * float((de * rdelta * r3).sum())
    return elrc / volume, vlrc / (3.0 * volume)


def build_vdw(n, types, atom_class, i12, i13, i14, i15, key, ff, use_bounds, volume, list_buffer=2.0):
    """Returns a VdwTerm, or None when the force field has no BUFFERED-14-7 term or `vdwterm none`."""
    def kget(kw, default=None):
        v = key.get(kw)
        if v is None and ff.keywords is not None:
            v = ff.keywords.get(kw)
        return default if v is None else v

    def khas(kw):
        return key.has(kw) or (ff.keywords is not None and ff.keywords.has(kw))

    def kword(kw, default):
        v = kget(kw)
        return (v.split() or [default])[0].upper() if v is not None else default

    def kfloat(kw, default):
        v = kget(kw)
        try:
            return float(v.split()[0].replace("D", "E").replace("d", "e"))
        except (AttributeError, IndexError, ValueError):
            return default

    # initprm.f:336-353 defaults, overridden by the .prm header and the key file
    if kword("VDWTYPE", "LENNARD-JONES") != "BUFFERED-14-7":
        return None
    if kword("VDWTERM", "") == "NONE":
        return None
    by_type = kword("VDWINDEX", "CLASS") == "TYPE"
    radrule, radtyp = kword("RADIUSRULE", "ARITHMETIC"), kword("RADIUSTYPE", "R-MIN")
    radsiz, epsrule = kword("RADIUSSIZE", "RADIUS"), kword("EPSILONRULE", "GEOMETRIC")
    ghal, dhal = kfloat("GAMMA-HALGREN", 0.12), kfloat("DELTA-HALGREN", 0.07)
    vscale = [kfloat("VDW-12-SCALE", 0.0), kfloat("VDW-13-SCALE", 0.0), kfloat("VDW-14-SCALE", 1.0), kfloat("VDW-15-SCALE", 1.0)]
    vscale = [1.0 / s if s > 1.0 else s for s in vscale]           # readprm.f accepts the inverse form

    # vdw records: class/type, radius, well depth, optional reduction (readprm.f; key file wins)
    rad, eps, reduct = {}, {}, {}
    for src in ((ff.keywords.lines if ff.keywords is not None else []), key.lines):
        for kw, rest, _ in src:
            if kw == "VDW":
                t = rest.split()
                try:
                    k = int(t[0])
                    rad[k], eps[k] = float(t[1]), float(t[2])
                    reduct[k] = float(t[3]) if len(t) > 3 else 0.0
                except (IndexError, ValueError):
                    continue
    index_of = (lambda i: int(types[i])) if by_type else (lambda i: int(atom_class[int(types[i])]))
    idx = np.array([index_of(i) for i in range(n)])
    used = sorted(set(idx.tolist()))
    if not any(k in rad for k in used):
        return None
    slot = {k: j for j, k in enumerate(used)}
    jvdw = np.array([slot[k] for k in idx], np.int32)
    # kvdw.f:340-352: sigma -> r-min, diameter -> radius, |eps|
    r1 = np.array([rad.get(k, 0.0) for k in used])
    e1 = np.abs(np.array([eps.get(k, 0.0) for k in used]))
    if radtyp == "SIGMA":
        r1 = r1 * TWOSIX
    if radsiz == "DIAMETER":
        r1 = 0.5 * r1
    nj = len(used)
    radmin = np.zeros((nj, nj))
    epsilon = np.zeros((nj, nj))
    se = np.sqrt(e1)
    for a in range(nj):
        for b in range(nj):
            ra, rb, ea, eb = r1[a], r1[b], e1[a], e1[b]
            # kvdw.f:393-404
            if ra == 0.0 and rb == 0.0:
                rd = 0.0
            elif radrule == "ARITHMETIC":
                rd = ra + rb
            elif radrule == "GEOMETRIC":
                rd = 2.0 * math.sqrt(ra) * math.sqrt(rb)
            elif radrule == "CUBIC-MEAN":
                rd = 2.0 * (ra ** 3 + rb ** 3) / (ra ** 2 + rb ** 2)
            else:
                rd = ra + rb
            # kvdw.f:424-440
            if ea == 0.0 and eb == 0.0:
                ep = 0.0
            elif epsrule == "ARITHMETIC":
                ep = 0.5 * (ea + eb)
            elif epsrule == "GEOMETRIC":
                ep = se[a] * se[b]
            elif epsrule == "HARMONIC":
                ep = 2.0 * ea * eb / (ea + eb)
            elif epsrule == "HHG":
                ep = 4.0 * ea * eb / (se[a] + se[b]) ** 2
            elif epsrule == "W-H":
                ep = 2.0 * se[a] * se[b] * (ra * rb) ** 3 / (ra ** 6 + rb ** 6)
            else:
                ep = se[a] * se[b]
            radmin[a, b], epsilon[a, b] = rd, ep
    # special pairs (kvdw.f:556-575): vdwpr / vdwpair  ia ib radius eps
    for src in ((ff.keywords.lines if ff.keywords is not None else []), key.lines):
        for kw, rest, _ in src:
            if kw in ("VDWPR", "VDWPAIR"):
                t = rest.split()
                try:
                    ia, ib, rp, ep = int(t[0]), int(t[1]), float(t[2]), float(t[3])
                except (IndexError, ValueError):
                    continue
                if ia in slot and ib in slot:
                    if radtyp == "SIGMA":
                        rp *= TWOSIX
                    a, b = slot[ia], slot[ib]
                    radmin[a, b] = radmin[b, a] = rp
                    epsilon[a, b] = epsilon[b, a] = abs(ep)
    # reduction factors (kvdw.f:541-552): only atoms with exactly one bond are moved along it
    ired = np.arange(n, dtype=np.int32)
    kred = np.array([reduct.get(int(k), 0.0) for k in idx])
    for i in range(n):
        if len(i12[i]) == 1 and kred[i] != 0.0:
            ired[i] = i12[i][0]
    # exclusions (src/evdw.cpp:196-258): every 1-2..1-5 pair whose scale is not one
    pairs = {}
    for sc, lists in zip(vscale, (i12, i13, i14, i15)):
        if sc == 1.0:
            continue
        for i in range(n):
            for k in lists[i]:
                if k > i:
                    pairs[(i, k)] = sc
    ik = np.array(sorted(pairs), np.int32).reshape(-1, 2)
    sc = np.array([pairs[tuple(p)] for p in ik.tolist()])
    # cutoffs.f:43,63,163-199,241
    off = 9.0 if use_bounds else 1.0e12
    if khas("CUTOFF"):
        off = kfloat("CUTOFF", off)
    off = kfloat("VDW-CUTOFF", off)
    tap = kfloat("VDW-TAPER", kfloat("TAPER", 0.90))
    cut = tap * off if tap < 1.0 else tap
    if khas("TRUNCATE"):
        cut = 1.0e12
    off, cut = min(off, 1.0e12), min(cut, 1.0e12)
    elrc = vlrc = 0.0
    if khas("VDW-CORRECTION") and use_bounds:
        elrc, vlrc = evcorr1(jvdw, radmin, epsilon, cut, off, ghal, dhal, volume)
    return VdwTerm(ired=ired, kred=kred, jvdw=jvdw, radmin=radmin, epsilon=epsilon, vexclude=ik, vexclude_scale=sc,
                   cutoff=off, taper=cut, ghal=ghal, dhal=dhal, elrc_vol=elrc * volume, vlrc_vol=vlrc * volume,
                   list_buffer=list_buffer)


_VDW_ARRAYS = ("ired", "kred", "jvdw", "radmin", "epsilon", "vexclude", "vexclude_scale")
_VDW_SCALARS = ("cutoff", "taper", "ghal", "dhal", "elrc_vol", "vlrc_vol", "list_buffer")


def vdw_to_dict(v: VdwTerm) -> dict:
    d = {"vdw_" + k: getattr(v, k) for k in _VDW_ARRAYS}
    d.update({"vdw__" + k: np.array(getattr(v, k)) for k in _VDW_SCALARS})
    return d


def vdw_from_npz(z) -> VdwTerm | None:
    if "vdw_ired" not in z.files:
        return None
    kw = {k: z["vdw_" + k] for k in _VDW_ARRAYS}
    kw.update({k: float(z["vdw__" + k]) for k in _VDW_SCALARS})
    return VdwTerm(**kw)


def replicate_vdw(v: VdwTerm, n0: int, m: int, volume_ratio: float) -> VdwTerm:
    """Tile a VdwTerm m times (index arrays offset per image), for params.replicate()."""
    shift = np.arange(m) * n0
    ired = (v.ired[None, :] + shift[:, None]).reshape(-1).astype(np.int32)
    ik = (v.vexclude[None, :, :] + shift[:, None, None]).reshape(-1, 2).astype(np.int32) if v.vexclude.size else v.vexclude.copy()
    # the correction is elrc_vol / V with elrc_vol ~ N^2: m^2 in the numerator, m in the volume
    return VdwTerm(ired=ired, kred=np.tile(v.kred, m), jvdw=np.tile(v.jvdw, m), radmin=v.radmin.copy(), epsilon=v.epsilon.copy(),
                   vexclude=ik, vexclude_scale=np.tile(v.vexclude_scale, m), cutoff=v.cutoff, taper=v.taper, ghal=v.ghal, dhal=v.dhal,
                   elrc_vol=v.elrc_vol * m * m, vlrc_vol=v.vlrc_vol * m * m, list_buffer=v.list_buffer)
