"""Valence (bonded) terms of the AMOEBA force fields: topology lists and parameter assignment
(SURVEY.md section 8f rank 3).

Restates what the reference reads from the Fortran Tinker modules before the `e*Data(RcOp)` uploads of
src/bonded/*.cpp run: the term lists of bonds.f / angles.f / torsions.f / bitors.f and the class look-ups
of kbond.f, kangle.f, kstrbnd.f, kurey.f, kopbend.f, ktors.f (+ torphase.f), kpitors.f, ktortor.f, with
the units and anharmonic constants of initprm.f / prmkey.f.  Built: the eight terms the AMOEBA protein /
nucleic-acid / water parameter files use (bond, angle incl. in-plane, stretch-bend, Urey-Bradley,
out-of-plane bend, torsion, pi-orbital torsion, torsion-torsion).  Ring-specific parameter classes
(bond3/4/5, angle3/4/5, torsion4/5), Fourier/linear angles, MMFF94, electronegativity corrections and the
improper / strtor / angtor terms belong to other force fields and are rejected loudly.

Indices are 0-based throughout.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

RADIAN = 57.29577951308232088     # tinker/source/math.f

TERMS = ("bond", "angle", "strbnd", "urey", "opbend", "torsion", "pitors", "tortor")
_TERM_KEYWORD = {"bond": "BONDTERM", "angle": "ANGLETERM", "strbnd": "STRBNDTERM", "urey": "UREYTERM",
                 "opbend": "OPBENDTERM", "torsion": "TORSIONTERM", "pitors": "PITORSTERM", "tortor": "TORTORTERM"}
# every potential-energy switch of tinker/source/prmkey.f (an "xxxTERM ONLY" switches all others off)
ALL_TERM_KEYWORDS = ("BONDTERM", "ANGLETERM", "STRBNDTERM", "UREYTERM", "ANGANGTERM", "OPBENDTERM", "OPDISTTERM",
                     "IMPROPTERM", "IMPTORSTERM", "TORSIONTERM", "PITORSTERM", "STRTORTERM", "ANGTORTERM",
                     "TORTORTERM", "VDWTERM", "REPULSTERM", "DISPERSIONTERM", "CHARGETERM", "CHGDPLTERM",
                     "DIPOLETERM", "MULTIPOLETERM", "POLARIZETERM", "CHGTRNTERM", "CHGFLXTERM", "RXNFIELDTERM",
                     "SOLVATETERM", "METALTERM", "RESTRAINTERM", "EXTRATERM", "VALENCETERM")
# parameter records of terms / special cases that are not built: {keyword: number of leading class fields}.  A record only
# matters when every class it names occurs in the system (ANGLEF is a fall-back of kangle.f: an angle that needs it is
# reported as undefined instead)
_UNSUPPORTED = {"BOND3": 2, "BOND4": 2, "BOND5": 2, "ANGLE3": 3, "ANGLE4": 3, "ANGLE5": 3, "TORSION4": 4, "TORSION5": 4,
                "ELECTNEG": 3, "IMPROPER": 4, "IMPTORS": 4, "STRTORS": 4, "ANGTORS": 4, "ANGANG": 1, "OPDIST": 4}


@dataclass
class ValenceTerms:
    """What ebondData ... etortorData upload (src/bonded/ebond.cpp ... etortor.cpp), with the atom indices of
    every term resolved (the reference keeps indirections through iang / ibnd / ibitor)."""
    # bond (ebond.cpp:30-52)
    ibnd: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))
    bk: np.ndarray = field(default_factory=lambda: np.zeros(0))
    bl: np.ndarray = field(default_factory=lambda: np.zeros(0))
    # angle (eangle.cpp); angtyp 0 = HARMONIC, 1 = IN-PLANE; iang[:,3] = out-of-plane atom or -1
    iang: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.int32))
    ak: np.ndarray = field(default_factory=lambda: np.zeros(0))
    anat: np.ndarray = field(default_factory=lambda: np.zeros(0))
    angtyp: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    # stretch-bend (estrbnd.cpp): atoms a,b,c; force constants for the a-b and c-b stretches; ideal values
    isb: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.int32))
    sbk: np.ndarray = field(default_factory=lambda: np.zeros((0, 2)))
    sb_anat: np.ndarray = field(default_factory=lambda: np.zeros(0))
    sb_bl: np.ndarray = field(default_factory=lambda: np.zeros((0, 2)))
    # Urey-Bradley (eurey.cpp)
    iury: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.int32))
    uk: np.ndarray = field(default_factory=lambda: np.zeros(0))
    ul: np.ndarray = field(default_factory=lambda: np.zeros(0))
    # out-of-plane bend (eopbend.cpp): atoms a, b (centre), c, d (the out-of-plane atom)
    iopb: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.int32))
    opbk: np.ndarray = field(default_factory=lambda: np.zeros(0))
    # torsion (etors.cpp): per fold 1..6 amplitude, phase (deg)
    itors: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.int32))
    tors_v: np.ndarray = field(default_factory=lambda: np.zeros((0, 6)))
    tors_phase: np.ndarray = field(default_factory=lambda: np.zeros((0, 6)))
    # pi-orbital torsion (epitors.cpp)
    ipit: np.ndarray = field(default_factory=lambda: np.zeros((0, 6), np.int32))
    kpit: np.ndarray = field(default_factory=lambda: np.zeros(0))
    # torsion-torsion (etortor.cpp): atoms a..e already in table order, chirality probe atom (or -1), grid id
    itt: np.ndarray = field(default_factory=lambda: np.zeros((0, 5), np.int32))
    tt_chk: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    tt_grid: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    # grids, flattened: grid g occupies [tt_off[g], tt_off[g] + tnx*tny) of tbf/tbx/tby/tbxy (x fastest) and
    # [tt_xoff[g], +tnx) of ttx, [tt_yoff[g], +tny) of tty
    tnx: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    tny: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    tt_off: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    tt_xoff: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    tt_yoff: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    ttx: np.ndarray = field(default_factory=lambda: np.zeros(0))
    tty: np.ndarray = field(default_factory=lambda: np.zeros(0))
    tbf: np.ndarray = field(default_factory=lambda: np.zeros(0))
    tbx: np.ndarray = field(default_factory=lambda: np.zeros(0))
    tby: np.ndarray = field(default_factory=lambda: np.zeros(0))
    tbxy: np.ndarray = field(default_factory=lambda: np.zeros(0))
    # constants: [bndunit, cbnd, qbnd, angunit, cang, qang, pang, sang, stbnunit, ureyunit, cury, qury,
    #             opbunit, copb, qopb, popb, sopb, torsunit, ptorunit, ttorunit]
    consts: np.ndarray = field(default_factory=lambda: np.zeros(20))
    opbtyp: int = 0                      # 0 = W-D-C, 1 = ALLINGER (initprm.f default W-D-C)
    use: np.ndarray = field(default_factory=lambda: np.ones(8, np.int32))   # order of TERMS

    def active(self, name: str) -> bool:
        return bool(self.use[TERMS.index(name)]) and self.count(name) > 0

    def count(self, name: str) -> int:
        return {"bond": len(self.ibnd), "angle": len(self.iang), "strbnd": len(self.isb), "urey": len(self.iury),
                "opbend": len(self.iopb), "torsion": len(self.itors), "pitors": len(self.ipit),
                "tortor": len(self.itt)}[name]

    def c(self, name: str) -> float:
        return float(self.consts[CONST_NAMES.index(name)])


CONST_NAMES = ("bndunit", "cbnd", "qbnd", "angunit", "cang", "qang", "pang", "sang", "stbnunit", "ureyunit", "cury",
               "qury", "opbunit", "copb", "qopb", "popb", "sopb", "torsunit", "ptorunit", "ttorunit")
_CONST_KEYWORD = {"bndunit": "BONDUNIT", "cbnd": "BOND-CUBIC", "qbnd": "BOND-QUARTIC", "angunit": "ANGLEUNIT",
                  "cang": "ANGLE-CUBIC", "qang": "ANGLE-QUARTIC", "pang": "ANGLE-PENTIC", "sang": "ANGLE-SEXTIC",
                  "stbnunit": "STRBNDUNIT", "ureyunit": "UREYUNIT", "cury": "UREY-CUBIC", "qury": "UREY-QUARTIC",
                  "opbunit": "OPBENDUNIT", "copb": "OPBEND-CUBIC", "qopb": "OPBEND-QUARTIC", "popb": "OPBEND-PENTIC",
                  "sopb": "OPBEND-SEXTIC", "torsunit": "TORSIONUNIT", "ptorunit": "PITORSUNIT",
                  "ttorunit": "TORTORUNIT"}
_CONST_DEFAULT = {"bndunit": 1.0, "angunit": 1.0 / RADIAN ** 2, "stbnunit": 1.0 / RADIAN, "ureyunit": 1.0,
                  "opbunit": 1.0 / RADIAN ** 2, "torsunit": 1.0, "ptorunit": 1.0, "ttorunit": 1.0}   # initprm.f:302-332


def _fl(tok):
    return float(tok.replace("D", "E").replace("d", "e"))


def _nums(rest, nint):
    """Leading integers then reals of a parameter line, Fortran list-directed style (missing = 0)."""
    t = rest.replace(",", " ").split()
    ints = [int(x) for x in t[:nint]]
    reals = []
    for x in t[nint:]:
        try:
            reals.append(_fl(x))
        except ValueError:
            break
    return ints, reals


def term_switches(key, ff_keywords=None):
    """use_bond ... use_tortor after prmkey.f: 'xxxTERM NONE' switches one off, 'xxxTERM ONLY' switches every
    other potential off (potoff), a bare keyword switches it on.  Lines are processed in file order, the
    parameter file first (field.f reads it before the keyfile)."""
    use = {k: True for k in ALL_TERM_KEYWORDS}
    for src in ((ff_keywords.lines if ff_keywords is not None else []), key.lines):
        for kw, rest, _ in src:
            if kw in use:
                v = (rest.split() or [""])[0].upper()
                if v == "ONLY":
                    use = {k: False for k in use}
                    use[kw] = True
                elif v == "NONE":
                    use[kw] = False
                else:
                    use[kw] = True
                if kw == "VALENCETERM":      # prmkey.f: a whole-family switch
                    for k in _TERM_KEYWORD.values():
                        use[k] = (v != "NONE")
                    if v == "ONLY":
                        use["VALENCETERM"] = True
    return use


def bond_list(i12):
    """ibnd and bndlist of bonds.f:31-60."""
    ibnd = []
    index = {}
    for i, nb in enumerate(i12):
        for k in nb:
            if i < k:
                index[(i, k)] = len(ibnd)
                ibnd.append((i, k))
    return np.array(ibnd, np.int32).reshape(-1, 2), index


def angle_list(i12):
    """iang of angles.f:40-66: per central atom every pair of neighbours; for a trivalent centre the
    fourth entry is the remaining neighbour."""
    iang = []
    for i, nb in enumerate(i12):
        first = len(iang)
        for j in range(len(nb) - 1):
            for k in range(j + 1, len(nb)):
                iang.append([nb[j], i, nb[k], -1])
        if len(nb) == 3:
            iang[first + 2][3] = nb[0]
            iang[first + 1][3] = nb[1]
            iang[first + 0][3] = nb[2]
    return np.array(iang, np.int32).reshape(-1, 4)


def torsion_list(ibnd, i12):
    """itors of torsions.f:44-62."""
    out = []
    for ib, ic in ibnd:
        for ia in i12[ib]:
            if ia != ic:
                for idd in i12[ic]:
                    if idd != ib and idd != ia:
                        out.append((ia, ib, ic, idd))
    return np.array(out, np.int32).reshape(-1, 4)


def bitorsion_list(iang, i12):
    """ibitor of bitors.f:46-66."""
    out = []
    for ib, ic, idd, _ in iang:
        for ia in i12[ib]:
            if ia != ic and ia != idd:
                for ie in i12[idd]:
                    if ie != ic and ie != ib and ie != ia:
                        out.append((ia, ib, ic, idd, ie))
    return np.array(out, np.int32).reshape(-1, 5)


def _spline_slopes(xs, ys, cyclic):
    """First derivatives at the nodes of the interpolating cubic spline: the `bs` that cspline.f (periodic)
    and nspline.f (natural end conditions) hand back to ktortor.f:194-250."""
    from scipy.interpolate import CubicSpline
    cs = CubicSpline(xs, ys, bc_type="periodic" if cyclic else "natural")
    return cs(xs, 1)


def _tortor_tables(lines):
    """TORTORS records: returns [(classes5, nx, ny, ttx, tty, tbf, tbx, tby, tbxy)], ktortor.f:48-250."""
    out = {}
    order = []
    i = 0
    while i < len(lines):
        kw, rest, raw = lines[i]
        i += 1
        if kw != "TORTORS":
            continue
        ints, _ = _nums(rest, 7)
        cl, nx, ny = tuple(ints[:5]), ints[5], ints[6]
        vals = []
        while len(vals) < 3 * nx * ny and i < len(lines):
            vals += [_fl(t) for t in lines[i][2].split()]
            i += 1
        v = np.array(vals[:3 * nx * ny]).reshape(-1, 3)
        tx, ty, tf = v[:, 0], v[:, 1], v[:, 2]
        srt = np.argsort(360.0 * ty + tx, kind="stable")
        tbf = tf[srt]
        ttx = np.unique(tx)
        tty = np.unique(ty)
        nx, ny = len(ttx), len(tty)
        eps = 1.0e-6
        cyclic = abs(abs(ttx[0] - ttx[-1]) - 360.0) <= eps and abs(abs(tty[0] - tty[-1]) - 360.0) <= eps
        f = tbf.reshape(ny, nx)                       # x fastest
        if cyclic and (np.abs(f[:, 0] - f[:, -1]).max() > eps or np.abs(f[0] - f[-1]).max() > eps):
            raise ValueError("KTORTOR  --  Warning, Unequal Tor-Tor Values")
        bx = np.stack([_spline_slopes(ttx, f[j], cyclic) for j in range(ny)])
        by = np.stack([_spline_slopes(tty, f[:, k], cyclic) for k in range(nx)], axis=1)
        bxy = np.stack([_spline_slopes(tty, bx[:, k], cyclic) for k in range(nx)], axis=1)
        if cl not in out:
            order.append(cl)
        out[cl] = (cl, nx, ny, ttx, tty, f.ravel(), bx.ravel(), by.ravel(), bxy.ravel())
    return [out[c] for c in order]


def build_valence(n, types, atom_class, atomic, i12, key, ff) -> ValenceTerms:
    """Assign all valence parameters of a system (kbond.f ... ktortor.f)."""
    srcs = ((ff.keywords.lines if ff.keywords is not None else []), key.lines)
    present = {int(atom_class[int(t)]) for t in types}
    for src in srcs:
        for kw, rest, _ in src:
            if kw in _UNSUPPORTED:
                try:
                    ks = [abs(int(t)) for t in rest.split()[:_UNSUPPORTED[kw]]]
                except ValueError:
                    continue
                if all(k == 0 or k in present for k in ks):
                    raise NotImplementedError(f"valence record {kw} {rest.strip()} applies to this system but the term is "
                                              "not built (SURVEY.md section 8f rank 3 covers the AMOEBA terms)")

    def kget(kw):
        v = key.get(kw)
        if v is None and ff.keywords is not None:
            v = ff.keywords.get(kw)
        return v

    cls = np.array([atom_class[int(t)] for t in types], np.int64)
    atn = np.array([atomic.get(int(t), 0) for t in types], np.int64)
    v = ValenceTerms()
    for k, name in enumerate(CONST_NAMES):
        s = kget(_CONST_KEYWORD[name])
        v.consts[k] = _fl(s.split()[0]) if s and s.split() else _CONST_DEFAULT.get(name, 0.0)
    s = kget("OPBENDTYPE")
    v.opbtyp = 1 if (s and s.split() and s.split()[0].upper() == "ALLINGER") else 0
    sw = term_switches(key, ff.keywords)
    v.use = np.array([int(sw[_TERM_KEYWORD[t]]) for t in TERMS], np.int32)

    # ---- class tables and atom-specific overrides (negative numbers), in file order
    kb, ka, kap, ksb, ku, kopb, kt, kpt = {}, {}, {}, {}, {}, {}, {}, {}
    spec = {k: [] for k in ("BOND", "ANGLE", "ANGLEP", "STRBND", "UREYBRAD", "OPBEND", "TORSION")}
    for src in srcs:
        for kw, rest, _ in src:
            if kw == "BOND":
                (a, b), r = _nums(rest, 2)
                r += [0.0] * 2
                if min(a, b) < 0:
                    spec[kw].append((abs(a) - 1, abs(b) - 1, r[0], r[1]))
                else:
                    kb[(min(a, b), max(a, b))] = (r[0], r[1])
            elif kw in ("ANGLE", "ANGLEP"):
                (a, b, c), r = _nums(rest, 3)
                r += [0.0] * 4
                if min(a, b, c) < 0:
                    spec[kw].append((abs(a) - 1, abs(b) - 1, abs(c) - 1, r[0], r[1]))
                elif min(a, b, c) > 0:
                    if a > c:
                        a, c = c, a
                    if kw == "ANGLE":
                        an = [r[1], r[2], r[3]]
                        if an[1] == 0.0 and an[2] == 0.0:
                            an[1] = an[2] = an[0]
                        ka[(a, b, c)] = (r[0], an)
                    else:
                        an = [r[1], r[2]]
                        if an[1] == 0.0:
                            an[1] = an[0]
                        kap[(a, b, c)] = (r[0], an)
            elif kw == "STRBND":
                (a, b, c), r = _nums(rest, 3)
                r += [0.0] * 2
                if min(a, b, c) < 0:
                    spec[kw].append((abs(a) - 1, abs(b) - 1, abs(c) - 1, r[0], r[1]))
                elif a <= c:
                    ksb[(a, b, c)] = (r[0], r[1])
                else:                                   # kstrbnd.f stores the key sorted and swaps the constants
                    ksb[(c, b, a)] = (r[1], r[0])
            elif kw == "UREYBRAD":
                (a, b, c), r = _nums(rest, 3)
                r += [0.0] * 2
                if min(a, b, c) < 0:
                    spec[kw].append((abs(a) - 1, abs(b) - 1, abs(c) - 1, r[0], r[1]))
                else:
                    ku[(min(a, c), b, max(a, c))] = (r[0], r[1])
            elif kw == "OPBEND":
                (a, b, c, d), r = _nums(rest, 4)
                r += [0.0]
                if min(a, b, c, d) < 0:
                    spec[kw].append((abs(a) - 1, abs(b) - 1, abs(c) - 1, abs(d) - 1, r[0]))
                else:
                    kopb[(a, b, min(c, d), max(c, d))] = r[0]
            elif kw == "TORSION":
                (a, b, c, d), _ = _nums(rest, 4)
                t = rest.replace(",", " ").split()[4:]
                vt, st = np.zeros(6), np.zeros(6)
                for j in range(0, len(t) - 2, 3):         # amplitude, phase, periodicity triples (torphase.f)
                    amp, ph, fold = _fl(t[j]), _fl(t[j + 1]), int(float(t[j + 2]))
                    while ph < -180.0:
                        ph += 360.0
                    while ph > 180.0:
                        ph -= 360.0
                    if 1 <= fold <= 6:
                        vt[fold - 1], st[fold - 1] = amp, ph
                if min(a, b, c, d) < 0:
                    spec[kw].append((abs(a) - 1, abs(b) - 1, abs(c) - 1, abs(d) - 1, vt, st))
                else:
                    if b < c or (b == c and a <= d):
                        kt[(a, b, c, d)] = (vt, st)
                    else:
                        kt[(d, c, b, a)] = (vt, st)
            elif kw == "PITORS":
                (a, b), r = _nums(rest, 2)
                kpt[(min(a, b), max(a, b))] = (r + [0.0])[0]

    # ---- bonds (kbond.f:211-262)
    ibnd, bindex = bond_list(i12)
    nb = len(ibnd)
    v.ibnd = ibnd
    v.bk, v.bl = np.zeros(nb), np.zeros(nb)
    missing = []
    for i, (a, b) in enumerate(ibnd):
        p = kb.get((min(cls[a], cls[b]), max(cls[a], cls[b])))
        if p is not None:
            v.bk[i], v.bl[i] = p
        elif min(atn[a], atn[b]) != 0 and v.use[0]:
            missing.append(("Bond", a + 1, b + 1))
    for a, b, fc, bd in spec["BOND"]:
        j = bindex.get((min(a, b), max(a, b)))
        if j is not None:
            v.bk[j], v.bl[j] = fc, bd

    # ---- angles (kangle.f:371-470)
    iang = angle_list(i12)
    na = len(iang)
    v.ak, v.anat, v.angtyp = np.zeros(na), np.zeros(na), np.zeros(na, np.int32)
    for i, (a, b, c, d) in enumerate(iang):
        ta, tc = cls[a], cls[c]
        pt = (min(ta, tc), cls[b], max(ta, tc))
        nhyd = sum(1 for k in i12[b] if k != a and k != c and atn[k] == 1)        # 0-based row of ang(:,j)
        done = False
        p = ka.get(pt)
        if p is not None and nhyd < 3 and p[1][nhyd] != 0.0:
            v.ak[i], v.anat[i] = p[0], p[1][nhyd]
            done = True
        if not done and len(i12[b]) == 3:
            p = kap.get(pt)
            if p is not None and nhyd < 2 and p[1][nhyd] != 0.0:
                v.ak[i], v.anat[i], v.angtyp[i] = p[0], p[1][nhyd], 1
                done = True
        if not done and min(atn[a], atn[b], atn[c]) != 0 and v.use[1]:
            missing.append(("Angle", a + 1, b + 1, c + 1))
    aindex = {(int(a), int(b), int(c)): i for i, (a, b, c, _) in enumerate(iang)}
    for a, b, c, fc, an in spec["ANGLE"] + spec["ANGLEP"]:
        j = aindex.get((a, b, c), aindex.get((c, b, a)))
        if j is not None:
            v.ak[j], v.anat[j] = fc, an
    for a, b, c, fc, an in spec["ANGLE"]:
        j = aindex.get((a, b, c), aindex.get((c, b, a)))
        if j is not None:
            v.angtyp[j] = 0
    for a, b, c, fc, an in spec["ANGLEP"]:
        j = aindex.get((min(a, c), b, max(a, c)))
        if j is not None:
            v.angtyp[j] = 1

    # ---- stretch-bend (kstrbnd.f:133-175)
    isb, sbk, sban, sbbl = [], [], [], []
    sb_angle = []
    if ksb:
        for i, (a, b, c, _) in enumerate(iang):
            ta, tc = cls[a], cls[c]
            p = ksb.get((min(ta, tc), cls[b], max(ta, tc)))
            if p is None:
                continue
            isb.append((a, b, c))
            sb_angle.append(i)
            sbk.append(p if ta <= tc else (p[1], p[0]))
            sban.append(v.anat[i])
            sbbl.append((v.bl[bindex[(min(a, b), max(a, b))]], v.bl[bindex[(min(c, b), max(c, b))]]))
    for a, b, c, s1, s2 in spec["STRBND"]:
        for j, (ia, ib, ic) in enumerate(isb):
            if b == ib and ((a == ia and c == ic) or (a == ic and c == ia)):
                sbk[j] = (s1, s2)
                break
    v.isb = np.array(isb, np.int32).reshape(-1, 3)
    v.sbk = np.array(sbk, float).reshape(-1, 2)
    v.sb_anat = np.array(sban, float)
    v.sb_bl = np.array(sbbl, float).reshape(-1, 2)

    # ---- Urey-Bradley (kurey.f:108-135)
    iury, uk, ul = [], [], []
    if ku:
        for a, b, c, _ in iang:
            ta, tc = cls[a], cls[c]
            p = ku.get((min(ta, tc), cls[b], max(ta, tc)))
            if p is not None:
                iury.append((a, b, c))
                uk.append(p[0])
                ul.append(p[1])
    for a, b, c, bb, tt in spec["UREYBRAD"]:
        for j, (ia, ib, ic) in enumerate(iury):
            if b == ib and ((a == ia and c == ic) or (a == ic and c == ia)):
                uk[j], ul[j] = bb, tt
                break
    v.iury = np.array(iury, np.int32).reshape(-1, 3)
    v.uk, v.ul = np.array(uk, float), np.array(ul, float)

    # ---- out-of-plane bend (kopbend.f:121-205); centres without parameters lose their in-plane reference atom
    iopb, opbk = [], []
    if kopb:
        jopb = {k[1] for k in kopb}
        for i, (a, b, c, d) in enumerate(iang):
            if cls[b] in jopb and len(i12[b]) == 3:
                ta, tc, tb, td = cls[a], cls[c], cls[b], cls[d]
                for pt in ((td, tb, min(ta, tc), max(ta, tc)), (td, tb, 0, 0), (0, tb, 0, 0)):
                    if pt in kopb:
                        iopb.append((a, b, c, d))
                        opbk.append(kopb[pt])
                        break
                else:
                    if v.use[4]:
                        missing.append(("Angle-OP", d + 1, b + 1, a + 1, c + 1))
            else:
                iang[i, 3] = b
    for a, b, c, d, f in spec["OPBEND"]:
        for j, (ia, ib, ic, idd) in enumerate(iopb):
            if a == idd and b == ib and ((c == ia and d == ic) or (c == ic and d == ia)):
                opbk[j] = f
                break
    v.iang = iang
    v.iopb = np.array(iopb, np.int32).reshape(-1, 4)
    v.opbk = np.array(opbk, float)

    # ---- torsions (ktors.f:208-330): exact, then one wildcard end, then both
    itors = torsion_list(ibnd, i12)
    nt = len(itors)
    v.itors = itors
    v.tors_v, v.tors_phase = np.zeros((nt, 6)), np.zeros((nt, 6))
    for i, (a, b, c, d) in enumerate(itors):
        ta, tb, tc, td = cls[a], cls[b], cls[c], cls[d]
        pt = (ta, tb, tc, td) if (tb < tc or (tb == tc and ta <= td)) else (td, tc, tb, ta)
        p = kt.get(pt)
        if p is None:
            p = kt.get((pt[0], pt[1], pt[2], 0))
            q = kt.get((0, pt[1], pt[2], pt[3]))
            if p is not None and q is not None:
                # both half-wildcards exist: the reference takes the one that comes first in its table
                order = list(kt.keys())
                p = p if order.index((pt[0], pt[1], pt[2], 0)) < order.index((0, pt[1], pt[2], pt[3])) else q
            elif p is None:
                p = q
        if p is None:
            p = kt.get((0, pt[1], pt[2], 0))
        if p is not None:
            v.tors_v[i], v.tors_phase[i] = p
        elif min(atn[a], atn[b], atn[c], atn[d]) != 0 and v.use[5]:
            missing.append(("Torsion", a + 1, b + 1, c + 1, d + 1))
    for a, b, c, d, vt, st in spec["TORSION"]:
        for j, (ia, ib, ic, idd) in enumerate(itors):
            if (a, b, c, d) == (ia, ib, ic, idd) or (a, b, c, d) == (idd, ic, ib, ia):
                v.tors_v[j], v.tors_phase[j] = vt, st
                break

    # ---- pi-orbital torsions (kpitors.f:75-112)
    ipit, kpit = [], []
    if kpt:
        for a, b in ibnd:
            if len(i12[a]) == 3 and len(i12[b]) == 3:
                p = kpt.get((min(cls[a], cls[b]), max(cls[a], cls[b])))
                if p is not None:
                    ra = [k for k in i12[a] if k != b]
                    rb = [k for k in i12[b] if k != a]
                    ipit.append((ra[0], ra[1], a, b, rb[0], rb[1]))
                    kpit.append(p)
    v.ipit = np.array(ipit, np.int32).reshape(-1, 6)
    v.kpit = np.array(kpit, float)

    # ---- torsion-torsions (ktortor.f:252-300, chirality probe of src/bonded/etortor.cpp:86-132)
    tables = _tortor_tables([ln for src in srcs for ln in src])
    if tables:
        tkey = {t[0]: g for g, t in enumerate(tables)}
        itt, chk, grid = [], [], []
        for a, b, c, d, e in bitorsion_list(iang, i12):
            pt1 = (cls[a], cls[b], cls[c], cls[d], cls[e])
            pt2 = pt1[::-1]
            g = None
            for gi, t in enumerate(tables):           # first table that matches in either direction
                if t[0] == pt1:
                    g, atoms = gi, (a, b, c, d, e)
                    break
                if t[0] == pt2:
                    g, atoms = gi, (e, d, c, b, a)
                    break
            if g is None:
                continue
            ib_, ic_, id_ = atoms[1], atoms[2], atoms[3]
            probe = -1
            if len(i12[ic_]) == 4:
                j, k = [m for m in i12[ic_] if m != ib_ and m != id_]
                if types[j] > types[k]:
                    probe = j
                if types[k] > types[j]:
                    probe = k
                if atn[j] > atn[k]:
                    probe = j
                if atn[k] > atn[j]:
                    probe = k
            itt.append(atoms)
            chk.append(probe)
            grid.append(g)
        v.itt = np.array(itt, np.int32).reshape(-1, 5)
        v.tt_chk = np.array(chk, np.int32)
        v.tt_grid = np.array(grid, np.int32)
        v.tnx = np.array([t[1] for t in tables], np.int32)
        v.tny = np.array([t[2] for t in tables], np.int32)
        sz = v.tnx.astype(np.int64) * v.tny
        v.tt_off = np.concatenate([[0], np.cumsum(sz)[:-1]]).astype(np.int32)
        v.tt_xoff = np.concatenate([[0], np.cumsum(v.tnx)[:-1]]).astype(np.int32)
        v.tt_yoff = np.concatenate([[0], np.cumsum(v.tny)[:-1]]).astype(np.int32)
        v.ttx = np.concatenate([t[3] for t in tables])
        v.tty = np.concatenate([t[4] for t in tables])
        v.tbf = np.concatenate([t[5] for t in tables])
        v.tbx = np.concatenate([t[6] for t in tables])
        v.tby = np.concatenate([t[7] for t in tables])
        v.tbxy = np.concatenate([t[8] for t in tables])

    if missing:
        raise ValueError("Undefined valence parameters: " + "; ".join(" ".join(map(str, m)) for m in missing[:8])
                         + (" ..." if len(missing) > 8 else ""))
    return v


_V_ARRAYS = ("ibnd", "bk", "bl", "iang", "ak", "anat", "angtyp", "isb", "sbk", "sb_anat", "sb_bl", "iury", "uk", "ul",
             "iopb", "opbk", "itors", "tors_v", "tors_phase", "ipit", "kpit", "itt", "tt_chk", "tt_grid", "tnx", "tny",
             "tt_off", "tt_xoff", "tt_yoff", "ttx", "tty", "tbf", "tbx", "tby", "tbxy", "consts", "use")


def valence_to_dict(v: ValenceTerms) -> dict:
    d = {"val_" + k: getattr(v, k) for k in _V_ARRAYS}
    d["val_opbtyp"] = np.array(v.opbtyp)
    return d


def valence_from_npz(z):
    if "val_consts" not in z.files:
        return None
    v = ValenceTerms(**{k: z["val_" + k] for k in _V_ARRAYS})
    v.opbtyp = int(z["val_opbtyp"])
    return v


def replicate_valence(v: ValenceTerms, n0: int, m: int) -> ValenceTerms:
    """Copies of the term lists for m images of an n0-atom cell (params.replicate)."""
    import copy
    out = copy.copy(v)

    def rep_idx(a):
        if len(a) == 0:
            return a
        off = (np.arange(m, dtype=np.int64) * n0)[:, None, None]
        b = a[None].astype(np.int64) + np.where(a[None] >= 0, off, 0)
        return b.reshape(-1, a.shape[1]).astype(np.int32)

    for k in ("ibnd", "iang", "isb", "iury", "iopb", "itors", "ipit", "itt"):
        setattr(out, k, rep_idx(getattr(v, k)))
    chk = v.tt_chk[None].astype(np.int64) + np.where(v.tt_chk[None] >= 0, (np.arange(m) * n0)[:, None], 0)
    out.tt_chk = chk.reshape(-1).astype(np.int32)
    for k in ("bk", "bl", "ak", "anat", "angtyp", "sb_anat", "uk", "ul", "opbk", "kpit", "tt_grid"):
        setattr(out, k, np.tile(getattr(v, k), m))
    for k in ("sbk", "sb_bl", "tors_v", "tors_phase"):
        setattr(out, k, np.tile(getattr(v, k), (m, 1)))
    return out
