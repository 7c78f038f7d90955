"""CPU oracle for the AMOEBA valence terms -- TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's
cpu_baseline leg); the product path never imports it.

Float64 restatement of the eight bonded energy functions the reference evaluates in src/cu/evalence.cu through
include/seq/{bond,angle,strbnd,urey,opbend,torsion,pitors,tortor}.h.  Only the ENERGY of each term is written
down here (from the functional forms in those headers and tinker/source/e*.f); gradients come from reverse-mode
differentiation of that energy (torch autograd, float64) and the internal virial from sum_i r_i (x) dE/dr_i,
which for translation-invariant terms equals the reference's per-term vxx = xab*dedxia + ... sums.  The CUDA
kernels carry hand-derived analytic gradients, so agreement is evidence rather than a shared derivation.

Pinned against the reference's goldens test/ref/{bond,angle.1,angle.2,strbnd,urey,opbend,torsion,pitors,tortor}.txt
(Trp-cage, amoebapro13) in tests/test_valence_oracle.py.
"""
from __future__ import annotations

import numpy as np
import torch

RADIAN = 57.29577951308232088
TERMS = ("bond", "angle", "strbnd", "urey", "opbend", "torsion", "pitors", "tortor")


def _t(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype)


def _idx(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.int64)


def _cross(a, b):
    return torch.linalg.cross(a, b, dim=-1)


def _dot(a, b):
    return (a * b).sum(-1)


def _angle_deg(u, w):
    """Angle between two vector fields in degrees (include/seq/angle.h:60-66)."""
    c = _dot(u, w) / torch.sqrt(_dot(u, u) * _dot(w, w))
    return RADIAN * torch.acos(torch.clamp(c, -1.0, 1.0))


def _dihedral_cs(a, b, c, d):
    """cos and sin of the a-b-c-d dihedral with Tinker's sign convention (include/seq/torsion.h:70-95)."""
    ba, cb, dc = b - a, c - b, d - c
    t, u = _cross(ba, cb), _cross(cb, dc)
    tu = _cross(t, u)
    rtru = torch.sqrt(_dot(t, t) * _dot(u, u))
    rcb = torch.sqrt(_dot(cb, cb))
    return _dot(t, u) / rtru, _dot(cb, tu) / (rcb * rtru)


def e_bond(x, v):
    """include/seq/bond.h:44-53 (HARMONIC with cubic and quartic terms)."""
    i = _idx(v.ibnd)
    d = x[i[:, 0]] - x[i[:, 1]]
    dt = torch.sqrt(_dot(d, d)) - _t(v.bl)
    dt2 = dt * dt
    return v.c("bndunit") * _t(v.bk) * dt2 * (1 + v.c("cbnd") * dt + v.c("qbnd") * dt2)


def _poly6(dt, k, unit, c, q, p, s):
    dt2 = dt * dt
    return unit * k * dt2 * (1 + c * dt + q * dt2 + p * dt2 * dt + s * dt2 * dt2)


def e_angle(x, v):
    """include/seq/angle.h: HARMONIC (sextic polynomial in degrees) and IN-PLANE (centre projected onto the
    plane of its three neighbours, lines 156-190)."""
    i = _idx(v.iang)
    a, b, c = x[i[:, 0]], x[i[:, 1]], x[i[:, 2]]
    ang = _angle_deg(a - b, c - b)
    inpl = torch.as_tensor(np.asarray(v.angtyp) == 1)
    if bool(inpl.any()):
        d = x[torch.clamp(i[:, 3], min=0)]
        ad, bd, cd = a - d, b - d, c - d
        t = _cross(ad, cd)
        delta = -_dot(t, bd) / _dot(t, t)
        p = b + t * delta[:, None]
        ang_p = _angle_deg(a - p, c - p)
        ang = torch.where(inpl, ang_p, ang)
    return _poly6(ang - _t(v.anat), _t(v.ak), v.c("angunit"), v.c("cang"), v.c("qang"), v.c("pang"), v.c("sang"))


def e_strbnd(x, v):
    """include/seq/strbnd.h:60-84: stbnunit (k1 dr_ab + k2 dr_cb) (theta - theta0), theta in degrees."""
    i = _idx(v.isb)
    a, b, c = x[i[:, 0]], x[i[:, 1]], x[i[:, 2]]
    ab, cb = a - b, c - b
    dt = _angle_deg(ab, cb) - _t(v.sb_anat)
    bl = _t(v.sb_bl)
    k = _t(v.sbk)
    dr1 = torch.sqrt(_dot(ab, ab)) - bl[:, 0]
    dr2 = torch.sqrt(_dot(cb, cb)) - bl[:, 1]
    return v.c("stbnunit") * (k[:, 0] * dr1 + k[:, 1] * dr2) * dt


def e_urey(x, v):
    """include/seq/urey.h:37-44."""
    i = _idx(v.iury)
    d = x[i[:, 0]] - x[i[:, 2]]
    dt = torch.sqrt(_dot(d, d)) - _t(v.ul)
    dt2 = dt * dt
    return v.c("ureyunit") * _t(v.uk) * dt2 * (1 + v.c("cury") * dt + v.c("qury") * dt2)


def e_opbend(x, v):
    """include/seq/opbend.h:84-118: Wilson-Decius-Cross or Allinger out-of-plane angle, sextic polynomial."""
    i = _idx(v.iopb)
    a, b, c, d = x[i[:, 0]], x[i[:, 1]], x[i[:, 2]], x[i[:, 3]]
    ab, cb, db = a - b, c - b, d - b
    if v.opbtyp == 0:
        cc = _dot(ab, ab) * _dot(cb, cb) - _dot(ab, cb) ** 2
    else:
        ad, cd = a - d, c - d
        cc = _dot(ad, ad) * _dot(cd, cd) - _dot(ad, cd) ** 2
    ee = _dot(db, _cross(ab, cb))
    rdb2 = torch.clamp(_dot(db, db), min=1.0e-4)
    sine = torch.clamp(torch.abs(ee) / torch.sqrt(cc * rdb2), max=1.0)
    ang = RADIAN * torch.asin(sine)
    return _poly6(ang, _t(v.opbk), v.c("opbunit"), v.c("copb"), v.c("qopb"), v.c("popb"), v.c("sopb"))


def e_torsion(x, v):
    """include/seq/torsion.h:96-130: sum over folds 1..6 of V_n (1 + cos(n phi - phase_n))."""
    i = _idx(v.itors)
    cs, sn = _dihedral_cs(x[i[:, 0]], x[i[:, 1]], x[i[:, 2]], x[i[:, 3]])
    phi = torch.atan2(sn, cs)
    amp = _t(v.tors_v)
    ph = _t(v.tors_phase) / RADIAN
    n = torch.arange(1, 7, dtype=torch.float64)
    return v.c("torsunit") * (amp * (1 + torch.cos(n[None] * phi[:, None] - ph))).sum(1)


def e_pitors(x, v):
    """include/seq/pitors.h:62-118: two-fold torsion about the c-d bond between the normals of the two
    trigonal planes, V (1 - cos 2 phi)."""
    i = _idx(v.ipit)
    a, b, c, d, e, g = (x[i[:, k]] for k in range(6))
    p = _cross(a - d, b - d) + c          # sic: the neighbours of c are taken relative to d, and vice versa
    q = _cross(e - c, g - c) + d
    cs, sn = _dihedral_cs(p, c, d, q)
    cos2 = cs * cs - sn * sn
    return v.c("ptorunit") * _t(v.kpit) * (1 - cos2)


def _bicubic(y, y1, y2, y12, x1l, x1u, x2l, x2u, x1, x2):
    """Bicubic patch through four corner values / slopes / cross-slopes (Numerical Recipes bcuint, the algorithm
    behind include/seq/tortor.h:9-60), written as a Hermite tensor product instead of the 16 tabulated
    coefficients.  Corner order: (l,l), (u,l), (u,u), (l,u)."""
    d1, d2 = x1u - x1l, x2u - x2l
    t, u = (x1 - x1l) / d1, (x2 - x2l) / d2

    def h(s):    # Hermite basis: value at 0, value at 1, slope at 0, slope at 1
        return 2 * s ** 3 - 3 * s ** 2 + 1, -2 * s ** 3 + 3 * s ** 2, s ** 3 - 2 * s ** 2 + s, s ** 3 - s ** 2

    t0, t1, ts0, ts1 = h(t)
    u0, u1, us0, us1 = h(u)
    # corners: 0 -> (t=0,u=0), 1 -> (1,0), 2 -> (1,1), 3 -> (0,1)
    tw = (t0, t1, t1, t0)
    uw = (u0, u0, u1, u1)
    tsw = (ts0, ts1, ts1, ts0)
    usw = (us0, us0, us1, us1)
    out = 0
    for k in range(4):
        out = out + y[:, k] * tw[k] * uw[k] + d1 * y1[:, k] * tsw[k] * uw[k] \
            + d2 * y2[:, k] * tw[k] * usw[k] + d1 * d2 * y12[:, k] * tsw[k] * usw[k]
    return out


def e_tortor(x, v):
    """include/seq/tortor.h:150-262: bicubic interpolation of the (phi1, phi2) grid, both angles negated when the
    central atom is a chiral centre of the opposite hand (chkttor)."""
    i = _idx(v.itt)
    a, b, c, d, e = (x[i[:, k]] for k in range(5))
    c1, s1 = _dihedral_cs(a, b, c, d)
    c2, s2 = _dihedral_cs(b, c, d, e)
    v1 = RADIAN * torch.atan2(s1, c1)
    v2 = RADIAN * torch.atan2(s2, c2)
    chk = _idx(v.tt_chk)
    probe = x[torch.clamp(chk, min=0)]
    cb, dc, ac = c - b, d - c, probe - c
    vol = ac[:, 0] * (-cb[:, 1] * dc[:, 2] + cb[:, 2] * dc[:, 1]) - cb[:, 0] * (dc[:, 1] * ac[:, 2] - dc[:, 2] * ac[:, 1]) \
        + dc[:, 0] * (-ac[:, 1] * cb[:, 2] + ac[:, 2] * cb[:, 1])
    flip = (chk >= 0) & (vol < 0)
    v1 = torch.where(flip, -v1, v1)
    v2 = torch.where(flip, -v2, v2)
    v1 = torch.where(v1 < -180, v1 + 360, torch.where(v1 >= 180, v1 - 360, v1))
    v2 = torch.where(v2 < -180, v2 + 360, torch.where(v2 >= 180, v2 - 360, v2))
    g = _idx(v.tt_grid)
    tnx, tny = _idx(v.tnx)[g], _idx(v.tny)[g]
    off, xoff, yoff = _idx(v.tt_off)[g], _idx(v.tt_xoff)[g], _idx(v.tt_yoff)[g]
    xlo = torch.floor((v1.detach() + 180) * (tnx - 1) / 360).to(torch.int64)
    ylo = torch.floor((v2.detach() + 180) * (tny - 1) / 360).to(torch.int64)
    ttx, tty = _t(v.ttx), _t(v.tty)
    x1l, x1u = ttx[xoff + xlo], ttx[xoff + xlo + 1]
    y1l, y1u = tty[yoff + ylo], tty[yoff + ylo + 1]
    pos1 = off + ylo * tnx + xlo
    pos2 = pos1 + tnx
    corners = torch.stack([pos1, pos1 + 1, pos2 + 1, pos2], dim=1)
    tbf, tbx, tby, tbxy = _t(v.tbf), _t(v.tbx), _t(v.tby), _t(v.tbxy)
    val = _bicubic(tbf[corners], tbx[corners], tby[corners], tbxy[corners], x1l, x1u, y1l, y1u, v1, v2)
    return v.c("ttorunit") * val


_FUNCS = {"bond": e_bond, "angle": e_angle, "strbnd": e_strbnd, "urey": e_urey, "opbend": e_opbend,
          "torsion": e_torsion, "pitors": e_pitors, "tortor": e_tortor}


def _pitors_virial(xyz, v):
    """The reference's pi-torsion virial (include/seq/pitors.h:178-186) treats the two plane normals p and q as
    sites: V = dc (x) (g_d + g_a + g_b) + cp (x) g_p - qd (x) g_q per term, which differs from sum_i r_i (x) g_i
    because |p - c| is quadratic in the bond lengths.  Reproduced here so that the oracle states what the
    reference computes; g_p, g_q are taken by differentiating the same energy with p, q as free sites."""
    i = _idx(v.ipit)
    X = torch.tensor(np.asarray(xyz, np.float64))[i].clone().requires_grad_(True)      # (np, 6, 3) per-term copies
    a, b, c, d, e, g = (X[:, k] for k in range(6))
    p = _cross(a - d, b - d) + c
    q = _cross(e - c, g - c) + d
    p.retain_grad()
    q.retain_grad()
    cs, sn = _dihedral_cs(p, c, d, q)
    en = (v.c("ptorunit") * _t(v.kpit) * (1 - (cs * cs - sn * sn))).sum()
    en.backward()
    G = X.grad
    Xd = X.detach()
    dc = Xd[:, 3] - Xd[:, 2]
    cp = Xd[:, 2] - p.detach()
    qd = q.detach() - Xd[:, 3]
    vterm = G[:, 3] + G[:, 0] + G[:, 1]
    M = dc.T @ vterm + cp.T @ p.grad - qd.T @ q.grad            # M[a][b] = r_a g_b
    return M.numpy()


def _mirror_lower(M):
    """vxx, vyx, vzx, vyy, vzy, vzz -> symmetric 3x3 the way the reference stores it (lower triangle mirrored)."""
    out = np.array(M, dtype=np.float64)
    for a in range(3):
        for b in range(a + 1, 3):
            out[a, b] = out[b, a]
    return out


def valence(xyz, v, terms=None, grad=True):
    """Energies (kcal/mol), interaction counts, gradient (n,3) and internal virial (3,3) of the active valence
    terms.  `terms` restricts the evaluation (default: every term whose switch is on and whose list is not
    empty, as energy_core does, src/energy.cpp:180-215)."""
    names = [t for t in TERMS if v.active(t)] if terms is None else list(terms)
    xyz = np.asarray(xyz, np.float64)
    out = {"energy": {}, "count": {}}
    gsum = np.zeros_like(xyz)
    vir = np.zeros((3, 3))
    esum = 0.0
    for t in names:
        if v.count(t) == 0:
            out["energy"][t], out["count"][t] = 0.0, 0
            continue
        x = torch.tensor(xyz, requires_grad=grad)
        e = _FUNCS[t](x, v).sum()
        out["energy"][t] = float(e.detach())
        out["count"][t] = v.count(t)
        esum += out["energy"][t]
        if grad:
            (g,) = torch.autograd.grad(e, x)
            g = g.numpy()
            gsum += g
            vir += _pitors_virial(xyz, v) if t == "pitors" else xyz.T @ g
    out["esum"] = esum
    if grad:
        out["grad"] = gsum
        out["virial"] = _mirror_lower(vir)
    return out
