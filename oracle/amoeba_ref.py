"""CPU ORACLE (test infrastructure, float64 numpy) for the AMOEBA polarizable-electrostatics path.

THIS IS NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it (task statement ③).  The product
path (tinker-gpu_b200/csrc, libapx.so) never calls into this module.

It restates the reference's algorithm for SURVEY.md §8 rows a2-a14:

  rotpole / chkpole      include/seq/rotpole.h:9-223
  dfield / ufield        include/seq/pair_field.h:124-256,330-420, include/seq/damp.h:8-44,136-151,
                         src/acc/amoeba/fieldewald.cpp:14-31,212-254 (recip + self parts)
  induce (PCG)           src/cu/amoeba/pcg.cu:14-185, src/cu/induce.cu:17-232,
                         src/acc/amoeba/induce.cpp:296-468 (diag / sparse preconditioner)
  PME                    include/seq/bsplgen.h:27-116, src/acc/pme.cpp:11-197 (spread),
                         220-303 (conv), 323-714 (gather), 731-960 (transforms),
                         src/pme.cpp:84-117 + tinker/source/pmestuf.f:134-237 (bsmod)
  empole real/self       include/seq/pair_mpole.h:237-442, src/acc/amoeba/empoleewald.cpp:11-215
  epolar real            include/seq/pair_polar.h:367-742
  recip E/F/torque/vir   src/cu/hippo/empole.cu:202-298, src/cu/epolarrecip.cu:9-511
  torque -> force        src/acc/amoeba/torque.cpp:20-388

The real-space pair physics is written as one generic Cartesian interaction-tensor
contraction (derivatives T^(n) of a radial hierarchy B_n with B_{n+1} = -(1/r) dB_n/dr)
instead of the reference's expanded scalar formulas; Ewald (erfc), Thole (lambda_3..9)
and exclusion scaling only change the B_n that are fed in.  Equality of the two forms is
what tests/test_oracle_golden.py checks against the reference's own golden vectors
(test/localframe.cpp, test/localframe3.cpp, test/ref/tinkernist.*).

PARITY PINNING: pinned against the reference's golden vectors for the 22-/18-atom
local-frame systems and the 2684-atom water box (tolerance 1e-4..1e-3, the precision the
goldens are printed with).  dhfr2 itself has NO golden in the reference tree;
its dhfr2-size values are pinned by outputs of the reference itself run here: the reference's
host arithmetic compiled in place (oracle/_ref, tests/test_ref_arith.py: 1e-11 .. 1e-13) and
the reference's own CUDA build on a B200 (oracle/ref_cuda.cu, tests/test_zgpu_9_refcuda.py:
6.6e-7 energy, 3.9e-7 D dipoles, 4.9e-5 gradient against tests/golden/dhfr2_oracle.npz).
"""
from __future__ import annotations

import itertools
import math

import numpy as np

SQRTPI = math.sqrt(math.pi)
DEBYE = 4.803206802

# calc flags (include/tool/rcman.h:107-133)
ENERGY, GRAD, VIRIAL, ANALYZ = 0x10, 0x20, 0x40, 0x80
V0, V1, V3, V4, V5, V6 = ENERGY, ENERGY + GRAD + VIRIAL, ENERGY + ANALYZ, ENERGY + GRAD, GRAD, GRAD + VIRIAL


def _full_q(p10):
    """(n,10) MPL_PME order -> full symmetric quadrupole (n,3,3)."""
    q = np.empty(p10.shape[:-1] + (3, 3))
    q[..., 0, 0] = p10[..., 4]
    q[..., 1, 1] = p10[..., 5]
    q[..., 2, 2] = p10[..., 6]
    q[..., 0, 1] = q[..., 1, 0] = p10[..., 7]
    q[..., 0, 2] = q[..., 2, 0] = p10[..., 8]
    q[..., 1, 2] = q[..., 2, 1] = p10[..., 9]
    return q


_EYE = np.eye(3)


def _tensors(R, B, order):
    """Derivative tensors T^(m)_{ab..} = d^m T / dR_a dR_b.. for m = 1..order of a radial
    function T(R) whose hierarchy is B[:,m] (B_{m+1} = -(1/R) dB_m/dR)."""
    out = [None] * (order + 1)
    P = R.shape[0]
    if order >= 1:
        out[1] = -R * B[:, 1, None]
    if order >= 2:
        out[2] = np.einsum("pa,pb->pab", R, R) * B[:, 2, None, None] - _EYE[None] * B[:, 1, None, None]
    if order >= 3:
        t = -np.einsum("pa,pb,pc->pabc", R, R, R) * B[:, 3, None, None, None]
        dR = np.einsum("ab,pc->pabc", _EYE, R)
        t += (dR + dR.transpose(0, 1, 3, 2) + dR.transpose(0, 3, 2, 1)) * B[:, 2, None, None, None]
        out[3] = t
    if order >= 4:
        t = np.einsum("pa,pb,pc,pd->pabcd", R, R, R, R) * B[:, 4, None, None, None, None]
        dRR = np.einsum("ab,pc,pd->pabcd", _EYE, R, R)
        s = np.zeros_like(t)
        for (i, j) in itertools.combinations(range(4), 2):
            rest = [k for k in range(4) if k not in (i, j)]
            perm = [0] * 4
            perm[i], perm[j], perm[rest[0]], perm[rest[1]] = 0, 1, 2, 3
            s += dRR.transpose(0, *[q + 1 for q in perm])
        t -= s * B[:, 3, None, None, None, None]
        dd = np.einsum("ab,cd->abcd", _EYE, _EYE)
        dd = dd + dd.transpose(0, 2, 1, 3) + dd.transpose(0, 3, 2, 1)
        t += dd[None] * B[:, 2, None, None, None, None]
        out[4] = t
    if order >= 5:
        t = -np.einsum("pa,pb,pc,pd,pe->pabcde", R, R, R, R, R) * B[:, 5, None, None, None, None, None]
        dRRR = np.einsum("ab,pc,pd,pe->pabcde", _EYE, R, R, R)
        s = np.zeros_like(t)
        for (i, j) in itertools.combinations(range(5), 2):
            rest = [k for k in range(5) if k not in (i, j)]
            perm = [0] * 5
            perm[i], perm[j] = 0, 1
            for q, k in enumerate(rest):
                perm[k] = 2 + q
            s += dRRR.transpose(0, *[q + 1 for q in perm])
        t += s * B[:, 4, None, None, None, None, None]
        ddR = np.einsum("ab,cd,pe->pabcde", _EYE, _EYE, R)
        s = np.zeros_like(t)
        for e in range(5):
            rest = [k for k in range(5) if k != e]
            a0 = rest[0]
            for b0 in rest[1:]:
                cd = [k for k in rest if k not in (a0, b0)]
                perm = [0] * 5
                perm[a0], perm[b0], perm[cd[0]], perm[cd[1]], perm[e] = 0, 1, 2, 3, 4
                s += ddR.transpose(0, *[q + 1 for q in perm])
        t -= s * B[:, 3, None, None, None, None, None]
        out[5] = t
    return out


def pair_interaction(R, ci, di, Qi, ck, dk, Qk, B, do_g=True):
    """Generic multipole-multipole interaction for P pairs; R = r_k - r_i.

    U = D_k D_i T with D_i = c_i - d_i.grad + Q_i:grad grad, D_k = c_k + d_k.grad + Q_k:grad grad.
    Returns U, gk (gradient on k; gradient on i is -gk), torque on i, torque on k."""
    order = 5 if do_g else 4
    T = _tensors(R, B, order)
    T0 = B[:, 0]
    # potential and its R-derivatives generated by site i, evaluated at k
    A0 = ci * T0 - np.einsum("pa,pa->p", di, T[1]) + np.einsum("pab,pab->p", Qi, T[2])
    A1 = ci[:, None] * T[1] - np.einsum("pb,pab->pa", di, T[2]) + np.einsum("pbc,pabc->pa", Qi, T[3])
    A2 = ci[:, None, None] * T[2] - np.einsum("pc,pabc->pab", di, T[3]) + np.einsum("pcd,pabcd->pab", Qi, T[4])
    U = ck * A0 + np.einsum("pa,pa->p", dk, A1) + np.einsum("pab,pab->p", Qk, A2)
    if not do_g:
        return U, None, None, None
    A3 = ci[:, None, None, None] * T[3] - np.einsum("pd,pabcd->pabc", di, T[4]) + np.einsum("pde,pabcde->pabc", Qi, T[5])
    gk = ck[:, None] * A1 + np.einsum("pb,pab->pa", dk, A2) + np.einsum("pbc,pabc->pa", Qk, A3)
    # torque on k: -(d x A1) - 2 eps (Q A2)
    QA = np.einsum("pab,pbc->pac", Qk, A2)
    tk = -np.cross(dk, A1) - 2.0 * np.stack([QA[:, 1, 2] - QA[:, 2, 1], QA[:, 2, 0] - QA[:, 0, 2], QA[:, 0, 1] - QA[:, 1, 0]], 1)
    # potential derivatives generated by k at i, w.r.t. r_i = -d/dR
    C1 = -(ck[:, None] * T[1] + np.einsum("pb,pab->pa", dk, T[2]) + np.einsum("pbc,pabc->pa", Qk, T[3]))
    C2 = ck[:, None, None] * T[2] + np.einsum("pc,pabc->pab", dk, T[3]) + np.einsum("pcd,pabcd->pab", Qk, T[4])
    QC = np.einsum("pab,pbc->pac", Qi, C2)
    ti = -np.cross(di, C1) - 2.0 * np.stack([QC[:, 1, 2] - QC[:, 2, 1], QC[:, 2, 0] - QC[:, 0, 2], QC[:, 0, 1] - QC[:, 1, 0]], 1)
    return U, gk, ti, tk


def field_from(R, c, d, Q, B, sign):
    """Electric field E = -grad phi at the other site generated by (c,d,Q).
    sign=+1: source at i, field at k (derivatives w.r.t. R); sign=-1: source at k, field at i."""
    T = _tensors(R, B, 3)
    if sign > 0:
        A1 = c[:, None] * T[1] - np.einsum("pb,pab->pa", d, T[2]) + np.einsum("pbc,pabc->pa", Q, T[3])
        return -A1
    C1 = -(c[:, None] * T[1] + np.einsum("pb,pab->pa", d, T[2]) + np.einsum("pbc,pabc->pa", Q, T[3]))
    return -C1


def dipole_field(R, uk, ui, B):
    """Field at i from dipole uk and at k from dipole ui (pair_field.h:330-420)."""
    RR = np.einsum("pa,pb->pab", R, R)
    T2 = RR * B[:, 2, None, None] - _EYE[None] * B[:, 1, None, None]
    # phi(k<-i) = -ui.grad T ; E_k = -grad = +T2.ui ;  E_i = +T2.uk (T2 is even in R)
    return np.einsum("pab,pb->pa", T2, uk), np.einsum("pab,pb->pa", T2, ui)


def ewald_bn(r, aewald, order):
    """damp_ewald (include/seq/damp.h:136-151): bn[0..order-1]."""
    bn = np.empty((r.shape[0], order))
    ra = aewald * r
    from scipy.special import erfc
    bn[:, 0] = erfc(ra) / r
    alsq2 = 2.0 * aewald * aewald
    alsq2n = 1.0 / (SQRTPI * aewald)
    exp2a = np.exp(-ra * ra)
    rr2 = 1.0 / (r * r)
    for j in range(1, order):
        alsq2n *= alsq2
        bn[:, j] = ((2 * j - 1) * bn[:, j - 1] + alsq2n * exp2a) * rr2
    return bn


def coulomb_rr(r, order):
    """rr1, rr3, 3 rr5, 15 rr7 ... i.e. the undamped hierarchy."""
    rr = np.empty((r.shape[0], order))
    rr[:, 0] = 1.0 / r
    rr2 = 1.0 / (r * r)
    for j in range(1, order):
        rr[:, j] = (2 * j - 1) * rr[:, j - 1] * rr2
    return rr


def thole_lambda(r, pdi, pdk, pga, order):
    """lambda_1(=1), lambda_3, lambda_5, lambda_7, lambda_9 (damp.h:8-44,117-134)."""
    lam = np.ones((r.shape[0], order))
    damp = pdi * pdk
    with np.errstate(divide="ignore", invalid="ignore"):
        x = np.where(damp == 0, 1.0e16, pga * (r / np.where(damp == 0, 1.0, damp)) ** 3)
    ex = np.exp(-x)
    if order > 1:
        lam[:, 1] = 1 - ex
    if order > 2:
        lam[:, 2] = 1 - ex * (1 + x)
    if order > 3:
        lam[:, 3] = 1 - ex * (1 + x + 0.6 * x * x)
    if order > 4:
        lam[:, 4] = 1 - ex * (1 + x * (1 + x * (18.0 / 35.0 + 9.0 / 35.0 * x)))
    if order > 5:
        lam[:, 5] = 1.0  # never used
    return lam


class Oracle:
    def __init__(self, system, chunk=20000):
        s = self.s = system
        self.n = s.n
        self.chunk = chunk
        self.f = s.electric / s.dielec
        self.xyz = np.array(s.xyz, float)
        self.lvec = np.array(s.lvec, float)
        self.recip = np.array(s.recip, float)
        self.pole = np.array(s.pole, float)
        self.zaxis = np.array(s.zaxis).copy()
        self._pairs = {}
        self._excl = None
        self.rpole = None
        self.niter = 0
        # induced-dipole predictor (ulspredSave / ulspredSum, src/amoeba/induce.cpp:27-69)
        self.polpred = str(getattr(s, "polpred", "NONE") or "NONE").upper()
        self.maxualt = {"ASPC": 16, "GEAR": 6}.get(self.polpred, 0)
        self.nualt = 0
        self.udalt = np.zeros((self.maxualt, self.n, 3))
        self.upalt = np.zeros((self.maxualt, self.n, 3))

    # ------------------------------------------------------------------ geometry
    def set_xyz(self, xyz):
        self.xyz = np.array(xyz, float)
        self._pairs = {}
        self.rpole = None

    def image(self, dr):
        """Minimum image (include/ff/image.h:16-65) via fractional coordinates."""
        fr = dr @ self.recip.T
        fr -= np.floor(fr + 0.5)
        return fr @ self.lvec.T

    def pairs(self, cutoff):
        """All i<k pairs with |image(r_k - r_i)| <= cutoff -> (i, k, R, r)."""
        key = float(cutoff)
        if key in self._pairs:
            return self._pairs[key]
        n = self.n
        L = np.diag(self.lvec)
        big = cutoff > 0.5 * L.min()
        if n <= 3000 or not self.s.orthogonal or big:
            if n > 6000:
                raise NotImplementedError("oracle pair search for large non-orthogonal / small boxes")
            i, k = np.triu_indices(n, 1)
            R = self.image(self.xyz[k] - self.xyz[i])
            r = np.sqrt((R * R).sum(1))
            m = r <= cutoff
            i, k, R, r = i[m], k[m], R[m], r[m]
        else:
            from scipy.spatial import cKDTree
            fr = self.xyz @ self.recip.T
            fr -= np.floor(fr)
            w = fr * L
            w = np.where(w >= L, w - L, w)
            tree = cKDTree(w, boxsize=L)
            pr = tree.query_pairs(cutoff * (1 + 1e-12), output_type="ndarray")
            i = np.minimum(pr[:, 0], pr[:, 1])
            k = np.maximum(pr[:, 0], pr[:, 1])
            o = np.lexsort((k, i))
            i, k = i[o], k[o]
            R = self.image(self.xyz[k] - self.xyz[i])
            r = np.sqrt((R * R).sum(1))
            m = r <= cutoff
            i, k, R, r = i[m], k[m], R[m], r[m]
        self._pairs[key] = (i.astype(np.int64), k.astype(np.int64), R, r)
        return self._pairs[key]

    def _scales(self, i, k):
        """Per-pair (m, d, p, u) scale factors; 1 unless listed in mdpuexclude."""
        n = self.n
        sc = np.ones((i.shape[0], 4))
        ex = self.s.mdpuexclude
        if ex.shape[0]:
            if self._excl is None:
                keys = ex[:, 0].astype(np.int64) * n + ex[:, 1]
                o = np.argsort(keys)
                self._excl = (keys[o], self.s.mdpuexclude_scale[o])
            keys, vals = self._excl
            q = i * n + k
            pos = np.searchsorted(keys, q)
            pos[pos >= keys.shape[0]] = 0
            hit = keys[pos] == q
            sc[hit] = vals[pos[hit]]
        return sc

    # ------------------------------------------------------------------ frames
    def chkpole(self):
        """include/seq/rotpole.h:9-49.  Flips y-related components at chiral Z-then-X sites."""
        z = self.zaxis
        x = self.xyz
        sel = np.where((z[:, 3] == 2) & (z[:, 2] != 0))[0]
        for i in sel:
            k = int(z[i, 2])
            ia, ib, ic, idd = i, z[i, 0], z[i, 1], abs(k) - 1
            ad, bd, cd = x[ia] - x[idd], x[ib] - x[idd], x[ic] - x[idd]
            c1 = bd[1] * cd[2] - bd[2] * cd[1]
            c2 = cd[1] * ad[2] - cd[2] * ad[1]
            c3 = ad[1] * bd[2] - ad[2] * bd[1]
            vol = ad[0] * c1 + bd[0] * c2 + cd[0] * c3
            if (k < 0 and vol > 0) or (k > 0 and vol < 0):
                z[i, 2] = -k
                self.pole[i, 2] = -self.pole[i, 2]
                self.pole[i, 7] = -self.pole[i, 7]
                self.pole[i, 9] = -self.pole[i, 9]

    def rotmat(self):
        """Rotation matrices a[i] with rows x,y,z axes (rotpole.h:132-223)."""
        n = self.n
        x = self.xyz
        z = self.zaxis
        a = np.tile(np.eye(3), (n, 1, 1))
        pol = z[:, 3]

        def unit(v):
            return v / np.sqrt((v * v).sum(1))[:, None]

        has = pol != 0
        idx = np.where(has)[0]
        if idx.size == 0:
            return a
        zz = unit(x[z[idx, 0]] - x[idx])
        xx = np.zeros_like(zz)
        p = pol[idx]
        zo = p == 1
        okay = ~(np.abs(zz[:, 0]) > 0.866)
        xx[zo, 0] = np.where(okay[zo], 1.0, 0.0)
        xx[zo, 1] = np.where(okay[zo], 0.0, 1.0)
        nzo = ~zo
        xx[nzo] = unit(x[z[idx[nzo], 1]] - x[idx[nzo]])
        yy = np.zeros_like(zz)
        yb = (p == 4) | (p == 5)
        if yb.any():
            yy[yb] = unit(x[np.abs(z[idx[yb], 2]) - 1] - x[idx[yb]])
        bis = p == 3
        zz[bis] = unit(zz[bis] + xx[bis])
        zb = p == 4
        xx[zb] = unit(xx[zb] + yy[zb])
        f3 = p == 5
        zz[f3] = unit(zz[f3] + xx[f3] + yy[f3])
        dot = (xx * zz).sum(1)[:, None]
        xx = unit(xx - dot * zz)
        yv = np.cross(zz, xx)
        a[idx, 0] = xx
        a[idx, 1] = yv
        a[idx, 2] = zz
        return a

    def rotpole(self):
        self.chkpole()
        a = self.rotmat()
        p = self.pole
        rp = np.zeros_like(p)
        rp[:, 0] = p[:, 0]
        rp[:, 1:4] = np.einsum("nj,nji->ni", p[:, 1:4], a)
        Ql = _full_q(p)
        Qg = np.einsum("nki,nmj,nmk->nij", a, a, Ql)   # rp[i][j] = sum a[k][i] a[m][j] mp[m][k]
        rp[:, 4], rp[:, 5], rp[:, 6] = Qg[:, 0, 0], Qg[:, 1, 1], Qg[:, 2, 2]
        rp[:, 7], rp[:, 8], rp[:, 9] = Qg[:, 0, 1], Qg[:, 0, 2], Qg[:, 1, 2]
        self.rpole = rp
        return rp

    def _ensure_rpole(self):
        if self.rpole is None:
            self.rotpole()
        return self.rpole

    # ------------------------------------------------------------------ PME
    @staticmethod
    def bspline_theta(w, order):
        """theta[n, order, 4]: value, 1st, 2nd, 3rd derivative (bsplgen.h, LEVEL=4)."""
        n = w.shape[0]
        b = np.zeros((n, order + 1, order + 1))   # b[:, level k, index i] 1-based like bsbuild(k? , i)
        # bsbuild(j, i): j = order level, i = index
        b[:, 2, 2] = w
        b[:, 2, 1] = 1 - w
        b[:, 3, 3] = 0.5 * w * b[:, 2, 2]
        b[:, 3, 2] = 0.5 * ((1 + w) * b[:, 2, 1] + (2 - w) * b[:, 2, 2])
        b[:, 3, 1] = 0.5 * (1 - w) * b[:, 2, 1]
        for i in range(4, order + 1):
            k = i - 1
            den = 1.0 / k
            b[:, i, i] = den * w * b[:, k, k]
            for j in range(1, i - 1):
                b[:, i, i - j] = den * ((w + j) * b[:, k, i - j - 1] + (i - j - w) * b[:, k, i - j])
            b[:, i, 1] = den * (1 - w) * b[:, k, 1]

        def diff(level_row, upto):
            # differentiate row `level_row` in place, treating entries 1..upto
            b[:, level_row, upto] = b[:, level_row, upto - 1]
            for i in range(upto - 1, 1, -1):
                b[:, level_row, i] = b[:, level_row, i - 1] - b[:, level_row, i]
            b[:, level_row, 1] = -b[:, level_row, 1]

        k = order - 1
        diff(k, order)
        k = order - 2
        diff(k, order - 1)
        diff(k, order)
        k = order - 3
        diff(k, order - 2)
        diff(k, order - 1)
        diff(k, order)
        th = np.zeros((n, order, 4))
        for i in range(1, order + 1):
            for j in range(1, 5):
                th[:, i - 1, j - 1] = b[:, order - j + 1, i]
        return th

    @staticmethod
    def bsmod(nfft, order):
        """tinker/source/pmestuf.f:134-237 via src/pme.cpp:84-117."""
        c = np.zeros(order + 1)
        x = 0.0
        c[1] = 1.0 - x
        c[2] = x
        for k in range(3, order + 1):
            den = 1.0 / (k - 1)
            c[k] = x * c[k - 1] * den
            for i in range(1, k - 1):
                c[k - i] = ((x + i) * c[k - i - 1] + (k - i - x) * c[k - i]) * den
            c[1] = (1.0 - x) * c[1] * den
        bsarray = np.zeros(nfft)
        bsarray[1:order + 1] = c[1:order + 1]
        j = np.arange(nfft)
        arg = 2.0 * math.pi / nfft * np.outer(j, j)
        s1 = (bsarray[None, :] * np.cos(arg)).sum(1)
        s2 = (bsarray[None, :] * np.sin(arg)).sum(1)
        mod = s1 ** 2 + s2 ** 2
        eps = 1.0e-7
        if mod[0] < eps:
            mod[0] = 0.5 * mod[1]
        for i in range(1, nfft - 1):
            if mod[i] < eps:
                mod[i] = 0.5 * (mod[i - 1] + mod[i + 1])
        if mod[nfft - 1] < eps:
            mod[nfft - 1] = 0.5 * mod[nfft - 2]
        jcut = 50
        for i in range(1, nfft + 1):
            k = i - 1
            if i > nfft // 2:
                k -= nfft
            if k == 0:
                zeta = 1.0
            else:
                fac = math.pi * k / nfft
                jj = np.arange(1, jcut + 1)
                a1 = fac / (fac + math.pi * jj)
                a2 = fac / (fac - math.pi * jj)
                sum1 = 1.0 + (a1 ** order).sum() + (a2 ** order).sum()
                sum2 = 1.0 + (a1 ** (2 * order)).sum() + (a2 ** (2 * order)).sum()
                zeta = sum2 / sum1
            mod[i - 1] *= zeta * zeta
        return mod

    def pme_setup(self):
        if getattr(self, "_pme", None) is not None and self._pme["xyz_id"] is self.xyz:
            return self._pme
        s = self.s
        n1, n2, n3 = s.nfft
        order = s.bsorder
        fr = self.xyz @ self.recip.T
        w = fr + 0.5 - np.floor(fr + 0.5)
        nf = np.array([n1, n2, n3], float)
        frg = w * nf
        ig = np.floor(frg).astype(np.int64)
        ww = frg - ig
        ig = ig - order + 1
        ig += np.where(ig < 0, np.array([n1, n2, n3]), 0)
        th = [self.bspline_theta(ww[:, d], order) for d in range(3)]
        ar = np.arange(order)
        ix = (ig[:, 0, None] + ar) % n1
        iy = (ig[:, 1, None] + ar) % n2
        iz = (ig[:, 2, None] + ar) % n3
        # flat index [iz][iy][ix] for every atom and every (z,y,x) stencil point
        flat = (iz[:, :, None, None] * n2 + iy[:, None, :, None]) * n1 + ix[:, None, None, :]
        if getattr(self, "_bsmod_cache", None) is None:
            self._bsmod_cache = [self.bsmod(m, order) for m in (n1, n2, n3)]
        # a[i][j]: cart->frac matrices (src/acc/pme.cpp:757-766)
        a = np.stack([n1 * self.recip[0], n2 * self.recip[1], n3 * self.recip[2]], 1)   # a[c][f] = nfft_f * recip_f[c]
        self._pme = dict(th=th, flat=flat.reshape(self.n, -1), bsmod=self._bsmod_cache, a=a, xyz_id=self.xyz)
        return self._pme

    def _ctf6(self, a):
        qi1 = [0, 1, 2, 0, 0, 1]
        qi2 = [0, 1, 2, 1, 2, 2]
        ctf = np.zeros((6, 6))
        for i1 in range(3):
            k = qi1[i1]
            for i2 in range(6):
                i, j = qi1[i2], qi2[i2]
                ctf[i2, i1] = a[i, k] * a[j, k]
        for i1 in range(3, 6):
            k, m = qi1[i1], qi2[i1]
            for i2 in range(6):
                i, j = qi1[i2], qi2[i2]
                ctf[i2, i1] = a[i, k] * a[j, m] + a[j, k] * a[i, m]
        return ctf

    def _ftc6(self, at):
        # `at` is the frac_to_cart matrix a[f][c] = nfft_f * recip_f[c]  (src/acc/pme.cpp:884-892)
        qi1 = [0, 1, 2, 0, 0, 1]
        qi2 = [0, 1, 2, 1, 2, 2]
        ftc = np.zeros((6, 6))
        for i1 in range(3):
            k = qi1[i1]
            for i2 in range(3):
                i = qi1[i2]
                ftc[i2, i1] = at[i, k] * at[i, k]
            for i2 in range(3, 6):
                i, j = qi1[i2], qi2[i2]
                ftc[i2, i1] = 2 * at[i, k] * at[j, k]
        for i1 in range(3, 6):
            k, m = qi1[i1], qi2[i1]
            for i2 in range(3):
                i = qi1[i2]
                ftc[i2, i1] = at[i, k] * at[i, m]
            for i2 in range(3, 6):
                i, j = qi1[i2], qi2[i2]
                ftc[i2, i1] = at[i, k] * at[j, m] + at[i, m] * at[j, k]
        return ftc

    def rpole_to_cmp(self, rp):
        cmp_ = rp.copy()
        cmp_[:, 7:10] *= 2.0
        return cmp_

    def cmp_to_fmp(self, cmp_):
        a = self.pme_setup()["a"]
        ctf = self._ctf6(a)
        fmp = np.zeros_like(cmp_)
        fmp[:, 0] = cmp_[:, 0]
        fmp[:, 1:4] = cmp_[:, 1:4] @ a          # fmp[j] = sum_k a[k][j] cmp[k]
        fmp[:, 4:10] = cmp_[:, 4:10] @ ctf      # fmp[j] = sum_k ctf[k][j] cmp[k]
        return fmp

    def cuind_to_fuind(self, u):
        a = self.pme_setup()["a"]
        return u @ a

    def fphi_to_cphi(self, fphi):
        a = self.pme_setup()["a"]
        at = a.T                                  # at[f][c]
        ftc = self._ftc6(at)
        cphi = np.zeros((fphi.shape[0], 10))
        cphi[:, 0] = fphi[:, 0]
        cphi[:, 1:4] = fphi[:, 1:4] @ at         # cphi[j] = sum_k at[k][j] fphi[k]
        cphi[:, 4:10] = fphi[:, 4:10] @ ftc
        return cphi

    def _stencil(self, lv):
        """Per-atom 125-point weights for derivative orders (l1,l2,l3) along the three axes."""
        p = self.pme_setup()
        th1, th2, th3 = p["th"]
        return lambda l1, l2, l3: (th3[:, :, None, None, l3] * th2[:, None, :, None, l2] * th1[:, None, None, :, l1]).reshape(self.n, -1)

    def grid_mpole(self, fmp):
        p = self.pme_setup()
        n1, n2, n3 = self.s.nfft
        W = self._stencil(None)
        val = (fmp[:, 0, None] * W(0, 0, 0) + fmp[:, 1, None] * W(1, 0, 0) + fmp[:, 2, None] * W(0, 1, 0)
               + fmp[:, 3, None] * W(0, 0, 1) + fmp[:, 4, None] * W(2, 0, 0) + fmp[:, 5, None] * W(0, 2, 0)
               + fmp[:, 6, None] * W(0, 0, 2) + fmp[:, 7, None] * W(1, 1, 0) + fmp[:, 8, None] * W(1, 0, 1)
               + fmp[:, 9, None] * W(0, 1, 1))
        q = np.zeros(n1 * n2 * n3)
        np.add.at(q, p["flat"].ravel(), val.ravel())
        return q.reshape(n3, n2, n1).astype(complex)

    def grid_uind(self, fud, fup):
        p = self.pme_setup()
        n1, n2, n3 = self.s.nfft
        W = self._stencil(None)
        w100, w010, w001 = W(1, 0, 0), W(0, 1, 0), W(0, 0, 1)
        vd = fud[:, 0, None] * w100 + fud[:, 1, None] * w010 + fud[:, 2, None] * w001
        vp = fup[:, 0, None] * w100 + fup[:, 1, None] * w010 + fup[:, 2, None] * w001
        qd = np.zeros(n1 * n2 * n3)
        qp = np.zeros(n1 * n2 * n3)
        np.add.at(qd, p["flat"].ravel(), vd.ravel())
        np.add.at(qp, p["flat"].ravel(), vp.ravel())
        return (qd + 1j * qp).reshape(n3, n2, n1)

    def conv_factors(self):
        """expterm of pmeConv (src/acc/pme.cpp:241-279) on the [k3][k2][k1] grid."""
        if getattr(self, "_conv", None) is not None:
            return self._conv
        n1, n2, n3 = self.s.nfft
        b1, b2, b3 = self.pme_setup()["bsmod"]
        k1 = np.arange(n1)
        k2 = np.arange(n2)
        k3 = np.arange(n3)
        r1 = np.where(k1 < (n1 + 1) // 2, k1, k1 - n1)
        r2 = np.where(k2 < (n2 + 1) // 2, k2, k2 - n2)
        r3 = np.where(k3 < (n3 + 1) // 2, k3, k3 - n3)
        ra, rb, rc = self.recip
        h = (ra[None, None, None, :] * r1[None, None, :, None] + rb[None, None, None, :] * r2[None, :, None, None]
             + rc[None, None, None, :] * r3[:, None, None, None])
        hsq = (h * h).sum(-1)
        pterm = (math.pi / self.s.aewald) ** 2
        term = -pterm * hsq
        denom = hsq * math.pi * self.s.volume * b1[None, None, :] * b2[None, :, None] * b3[:, None, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            expterm = np.where(term > -50, np.exp(term) / denom, 0.0)
        expterm[0, 0, 0] = 0.0
        self._conv = (expterm, h, hsq, term)
        return self._conv

    def pme_convolve(self, qgrid, want_ev=False):
        """fftfront + pmeConv + fftback.  Returns the real-space convolved grid (complex) and,
        if asked, the reciprocal energy and virial of |Q|^2 (pmeConv DO_E/DO_V)."""
        expterm, h, hsq, term = self.conv_factors()
        Q = np.fft.fftn(qgrid)
        e = v = None
        if want_ev:
            struc2 = (Q.real ** 2 + Q.imag ** 2)
            eterm = 0.5 * self.f * expterm * struc2
            e = eterm.sum()
            with np.errstate(divide="ignore", invalid="ignore"):
                vterm = np.where(hsq > 0, (2.0 / hsq) * (1 - term) * eterm, 0.0)
            v = np.einsum("zyx,zyxa,zyxb->ab", vterm, h, h) - np.eye(3) * e
        out = np.fft.ifftn(Q * expterm) * qgrid.size
        return out, e, v

    def fphi_gather(self, grid, nder):
        """fphiGet: potential and derivatives up to total order 3 (20 values) or 2 (10)."""
        p = self.pme_setup()
        W = self._stencil(None)
        g = grid.reshape(-1)[p["flat"]]          # (n,125)
        combos20 = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (2, 0, 0), (0, 2, 0), (0, 0, 2), (1, 1, 0), (1, 0, 1),
                    (0, 1, 1), (3, 0, 0), (0, 3, 0), (0, 0, 3), (2, 1, 0), (2, 0, 1), (1, 2, 0), (0, 2, 1), (1, 0, 2),
                    (0, 1, 2), (1, 1, 1)]
        out = np.zeros((self.n, nder))
        for j, (a, b, c) in enumerate(combos20[:nder]):
            out[:, j] = (g * W(a, b, c)).sum(1)
        return out

    # ------------------------------------------------------------------ fields
    def _pair_params(self, i, k):
        s = self.s
        pga = s.thlval[s.jpolar[i], s.jpolar[k]]
        return s.pdamp[i], s.pdamp[k], pga

    def dfield(self, real_only=False):
        """Permanent field (d and p scalings) at every atom; dfield() of src/amoeba/field.cpp:56-64.
        real_only: the real-space pair sum alone (what dfieldEwaldReal adds, for the check against oracle/_ref)."""
        s = self.s
        n = self.n
        rp = self._ensure_rpole()
        fd = np.zeros((n, 3))
        fp = np.zeros((n, 3))
        if s.use_ewald and not real_only:
            cmp_ = self.rpole_to_cmp(rp)
            fmp = self.cmp_to_fmp(cmp_)
            grid, e, v = self.pme_convolve(self.grid_mpole(fmp), want_ev=True)
            self._recip_m = dict(e=e, v=v, cmp=cmp_, fmp=fmp)
            fphi = self.fphi_gather(grid.real, 20)
            cphi = self.fphi_to_cphi(fphi)
            self._recip_m.update(fphi=fphi, cphi=cphi)
            term = 4.0 / 3.0 * s.aewald ** 3 / SQRTPI
            fd += -cphi[:, 1:4] + term * rp[:, 1:4]
            fp[:] = fd
        i, k, R, r = self.pairs(s.ewald_cutoff)
        sc = self._scales(i, k)
        pdi, pdk, pga = self._pair_params(i, k)
        lam = thole_lambda(r, pdi, pdk, pga, 4)
        rr = coulomb_rr(r, 4)
        bn = ewald_bn(r, s.aewald, 4) if s.use_ewald else rr
        Q = _full_q(rp)
        for col, out in ((1, fd), (2, fp)):
            B = bn - (1 - sc[:, col, None] * lam) * rr
            ek = field_from(R, rp[i, 0], rp[i, 1:4], Q[i], B, +1)
            ei = field_from(R, rp[k, 0], rp[k, 1:4], Q[k], B, -1)
            np.add.at(out, k, ek)
            np.add.at(out, i, ei)
        return fd, fp

    def ufield(self, ud, up, real_only=False):
        """Mutual field of dipoles (ud, up); ufield() of src/amoeba/field.cpp:111-117."""
        s = self.s
        n = self.n
        fd = np.zeros((n, 3))
        fp = np.zeros((n, 3))
        if s.use_ewald and not real_only:
            a = self.pme_setup()["a"]
            grid, _, _ = self.pme_convolve(self.grid_uind(ud @ a, up @ a))
            f1 = self.fphi_gather(grid.real, 10)
            f2 = self.fphi_gather(grid.imag, 10)
            term = 4.0 / 3.0 * s.aewald ** 3 / SQRTPI
            fd += term * ud - f1[:, 1:4] @ a.T
            fp += term * up - f2[:, 1:4] @ a.T
        i, k, R, r = self.pairs(s.ewald_cutoff)
        sc = self._scales(i, k)
        pdi, pdk, pga = self._pair_params(i, k)
        lam = thole_lambda(r, pdi, pdk, pga, 3)
        rr = coulomb_rr(r, 3)
        bn = ewald_bn(r, s.aewald, 3) if s.use_ewald else rr
        B = bn - (1 - sc[:, 3, None] * lam) * rr
        for u, out in ((ud, fd), (up, fp)):
            ei, ek = dipole_field(R, u[k], u[i], B)
            np.add.at(out, i, ei)
            np.add.at(out, k, ek)
        return fd, fp

    def precond(self, rd, rp_):
        """diagPrecond / sparsePrecondApply (src/acc/amoeba/induce.cpp:296-468, precond.cu:13-43)."""
        s = self.s
        pol = s.polarity[:, None]
        if not (s.pcgprec and s.usolve_cutoff > 0):
            return pol * rd, pol * rp_
        zd = s.uaccel * pol * rd
        zp = s.uaccel * pol * rp_
        i, k, R, r = self.pairs(s.usolve_cutoff)
        sc = self._scales(i, k)
        pdi, pdk, pga = self._pair_params(i, k)
        lam = thole_lambda(r, pdi, pdk, pga, 3)
        polik = s.polarity[i] * s.polarity[k]
        rr3 = sc[:, 3] * lam[:, 1] * polik / r ** 3
        rr5 = 3 * sc[:, 3] * lam[:, 2] * polik / r ** 5
        M = np.einsum("pa,pb->pab", R, R) * rr5[:, None, None] - _EYE[None] * rr3[:, None, None]
        for rs, z in ((rd, zd), (rp_, zp)):
            np.add.at(z, i, np.einsum("pab,pb->pa", M, rs[k]))
            np.add.at(z, k, np.einsum("pab,pb->pa", M, rs[i]))
        return zd, zp

    def induce(self, verbose=False):
        """induceMutualPcg1 (src/cu/amoeba/pcg.cu:14-185)."""
        s = self.s
        n = self.n
        pol = s.polarity[:, None]
        pinv = 1.0 / np.maximum(s.polarity, 1e-16)[:, None]
        fd, fp = self.dfield()
        self.udir, self.udirp = pol * fd, pol * fp
        if s.poltyp == "DIRECT":
            self.uind, self.uinp = self.udir.copy(), self.udirp.copy()
            self.niter = 0
            return self.uind, self.uinp
        # pcg.cu:26-31: the predictor replaces the direct guess once its history ring is full
        predict = self.maxualt > 0 and self.nualt >= self.maxualt
        if predict:
            ud, up = self.ulspred_sum()
            fd, fp = self.ufield(ud, up)
            rd = (self.udir - ud) * pinv + fd          # pcgRsd0V2, src/cu/induce.cu:46-58
            rp_ = (self.udirp - up) * pinv + fp
        elif s.pcgguess:
            ud, up = self.udir.copy(), self.udirp.copy()
            rd, rp_ = self.ufield(ud, up)
        else:
            ud, up = np.zeros((n, 3)), np.zeros((n, 3))
            rd, rp_ = fd.copy(), fp.copy()
        zero = (s.polarity == 0)
        rd[zero] = 0
        rp_[zero] = 0
        zd, zp = self.precond(rd, rp_)
        cd, cp = zd.copy(), zp.copy()
        sm, smp = (rd * zd).sum(), (rp_ * zp).sum()
        it = 0
        done = False
        miniter = min(3, n)
        while not done:
            it += 1
            fd, fp = self.ufield(cd, cp)
            vd = pinv * cd - fd
            vp = pinv * cp - fp
            a, ap = (cd * vd).sum(), (cp * vp).sum()
            a = sm / a if a != 0 else 0.0
            ap = smp / ap if ap != 0 else 0.0
            ud += a * cd
            up += ap * cp
            rd -= a * vd
            rp_ -= ap * vp
            rd[zero] = 0
            rp_[zero] = 0
            zd, zp = self.precond(rd, rp_)
            sm1, smp1 = (rd * zd).sum(), (rp_ * zp).sum()
            b = sm1 / sm if sm != 0 else 0.0
            bp = smp1 / smp if smp != 0 else 0.0
            cd = zd + b * cd
            cp = zp + bp * cp
            sm, smp = sm1, smp1
            eps = DEBYE * math.sqrt(max((rd * rd).sum(), (rp_ * rp_).sum()) / n)
            if verbose:
                print(f" {it:8d}       {eps:<16.10f}")
            if eps < s.poleps:
                done = True
            if it < miniter:
                done = False
            if it >= s.politer:
                done = True
            if done:
                ud += s.pcgpeek * pol * rd
                up += s.pcgpeek * pol * rp_
        self.niter = it
        self.uind, self.uinp = ud, up
        self.ulspred_save(ud, up)
        return ud, up

    # ------------------------------------------------------------------ dipole predictors
    ASPC = (62. / 17., -310. / 51., 2170. / 323., -2329. / 400., 1701. / 409., -806. / 323., 1024. / 809., -479. / 883.,
            257. / 1316., -434. / 7429., 191. / 13375., -62. / 22287., 3. / 7217., -3. / 67015., 2. / 646323., -1. / 9694845.)
    GEAR = (6., -15., 20., -15., 6., -1.)

    def ulspred_save(self, ud, up):
        """ulspredSave (src/amoeba/induce.cpp:29-63): ring of the last maxualt solutions."""
        m = self.maxualt
        if m == 0:
            return
        pos = self.nualt % m
        self.udalt[pos] = ud
        self.upalt[pos] = up
        self.nualt += 1
        if self.nualt > 2 * m:
            self.nualt -= m

    def ulspred_sum(self):
        """ulspredSumASPC_cu / ulspredSumGEAR_cu (src/cu/upredict.cu:34-207): slot k holds the solution
        of age (nualt-1-k) mod maxualt and gets the coefficient of that age."""
        m = self.maxualt
        coef = self.ASPC if self.polpred == "ASPC" else self.GEAR
        ud, up = np.zeros((self.n, 3)), np.zeros((self.n, 3))
        for k in range(m):
            c = coef[(self.nualt - 1 - k) % m]
            ud += c * self.udalt[k]
            up += c * self.upalt[k]
        return ud, up

    # ------------------------------------------------------------------ torque -> gradient
    def torque(self, trq, grad, do_v=False):
        """src/acc/amoeba/torque.cpp:20-388.  Adds to grad in place; returns the torque virial."""
        x = self.xyz
        z = self.zaxis
        vir = np.zeros((3, 3))

        def nrm(v):
            return np.sqrt((v * v).sum())

        for i in range(self.n):
            axe = z[i, 3]
            if axe == 0:
                continue
            ia, ib, ic, idd = z[i, 0], i, z[i, 1], abs(z[i, 2]) - 1
            u = x[ia] - x[ib]
            usiz = nrm(u)
            u = u / usiz
            if axe != 1:
                v = x[ic] - x[ib]
                vsiz = nrm(v)
            else:
                foo = not (abs(u[0]) > 0.866)
                v = np.array([1.0 if foo else 0.0, 0.0 if foo else 1.0, 0.0])
                vsiz = 1.0
            v = v / vsiz
            if axe in (4, 5):
                w = x[idd] - x[ib]
            else:
                w = np.cross(u, v)
            wsiz = nrm(w)
            w = w / wsiz
            t = trq[i]
            dphidu, dphidv, dphidw = -t @ u, -t @ v, -t @ w
            frcz = np.zeros(3)
            frcx = np.zeros(3)
            frcy = np.zeros(3)
            if axe in (1, 2, 3):
                uv = np.cross(v, u)
                uv /= nrm(uv)
                uw = np.cross(w, u)
                uw /= nrm(uw)
                uvcos = u @ v
                uvsin = math.sqrt(1 - uvcos * uvcos)
                if axe == 1:
                    du = uv * dphidv / (usiz * uvsin) + uw * dphidw / usiz
                    frcz = du
                elif axe == 2:
                    du = uv * dphidv / (usiz * uvsin) + uw * dphidw / usiz
                    dv = -uv * dphidu / (vsiz * uvsin)
                    frcz, frcx = du, dv
                else:
                    vw = np.cross(w, v)
                    vw /= nrm(vw)
                    du = uv * dphidv / (usiz * uvsin) + 0.5 * uw * dphidw / usiz
                    dv = -uv * dphidu / (vsiz * uvsin) + 0.5 * vw * dphidw / vsiz
                    frcz, frcx = du, dv
            elif axe == 4:
                r = v + w
                s_ = np.cross(u, r)
                r /= nrm(r)
                s_ /= nrm(s_)
                ur = np.cross(r, u)
                ur /= nrm(ur)
                us = np.cross(s_, u)
                us /= nrm(us)
                urcos = u @ r
                ursin = math.sqrt(1 - urcos * urcos)
                vscos = v @ s_
                vssin = math.sqrt(1 - vscos * vscos)
                wscos = w @ s_
                wssin = math.sqrt(1 - wscos * wscos)
                t1 = v - s_ * vscos
                t2 = w - s_ * wscos
                t1 /= nrm(t1)
                t2 /= nrm(t2)
                ut1cos = u @ t1
                ut1sin = math.sqrt(1 - ut1cos * ut1cos)
                ut2cos = u @ t2
                ut2sin = math.sqrt(1 - ut2cos * ut2cos)
                dphidr, dphids = -t @ r, -t @ s_
                du = ur * dphidr / (usiz * ursin) + us * dphids / usiz
                dv = (vssin * s_ - vscos * t1) * dphidu / (vsiz * (ut1sin + ut2sin))
                dw = (wssin * s_ - wscos * t2) * dphidu / (wsiz * (ut1sin + ut2sin))
                frcz, frcx, frcy = du, dv, dw
            elif axe == 5:
                p = u + v + w
                psiz = nrm(p)
                p /= psiz
                wpcos, upcos, vpcos = w @ p, u @ p, v @ p

                def leg(a_, b_, c_, csiz, cpcos):
                    # force on the atom defining c_ from rotation about the other two (a_, b_)
                    r = a_ + b_
                    r /= nrm(r)
                    rccos = r @ c_
                    rcsin = math.sqrt(1 - rccos * rccos)
                    dphidr = -t @ r
                    dl = np.cross(r, c_)
                    dl /= nrm(dl)
                    dphiddel = -t @ dl
                    eps = np.cross(dl, c_)
                    return dl * dphidr / (csiz * rcsin) + eps * dphiddel * cpcos / (csiz * psiz)

                frcy = leg(u, v, w, wsiz, wpcos)
                frcz = leg(v, w, u, usiz, upcos)
                frcx = leg(u, w, v, vsiz, vpcos)
            grad[ia] += frcz
            grad[ib] -= frcz + frcx + frcy
            if axe != 1:
                grad[ic] += frcx
            if axe in (4, 5):
                grad[idd] += frcy
            if do_v:
                iaz = i if ia == -1 else ia
                iax = i if ic == -1 else ic
                iay = i if idd == -1 else idd
                dz, dx, dy = x[iaz] - x[i], x[iax] - x[i], x[iay] - x[i]
                m = np.outer(dx, frcx) + np.outer(dy, frcy) + np.outer(dz, frcz)
                vir += 0.5 * (m + m.T)
        return vir

    # ------------------------------------------------------------------ energies
    def _real_space(self, vers, do_m, do_p):
        """Real-space multipole and polarization energy / gradient / torque / virial over all
        pairs within the cutoff, exclusion scaling folded into the radial hierarchy."""
        s = self.s
        n = self.n
        do_g = bool(vers & GRAD)
        rp = self._ensure_rpole()
        Q = _full_q(rp)
        i_all, k_all, R_all, r_all = self.pairs(s.ewald_cutoff)
        sc_all = self._scales(i_all, k_all)
        em = ep = 0.0
        nem = nep = 0
        gm = np.zeros((n, 3))
        gp = np.zeros((n, 3))
        tm = np.zeros((n, 3))
        tp = np.zeros((n, 3))
        vm = np.zeros((3, 3))
        vp = np.zeros((3, 3))
        f = self.f
        if do_p:
            ud, up = self.uind, self.uinp
        zc = None
        for lo in range(0, i_all.shape[0], self.chunk):
            sl = slice(lo, lo + self.chunk)
            i, k, R, r, sc = i_all[sl], k_all[sl], R_all[sl], r_all[sl], sc_all[sl]
            P = i.shape[0]
            rr = coulomb_rr(r, 6)
            bn = ewald_bn(r, s.aewald, 6) if s.use_ewald else rr
            ci, di, Qi = rp[i, 0], rp[i, 1:4], Q[i]
            ck, dk, Qk = rp[k, 0], rp[k, 1:4], Q[k]
            if do_m:
                B = bn - (1 - sc[:, 0, None]) * rr
                U, gk, ti, tk = pair_interaction(R, ci, di, Qi, ck, dk, Qk, B, do_g)
                em += f * U.sum()
                nem += int((sc[:, 0] != 0).sum())
                if do_g:
                    gk *= f
                    np.add.at(gm, k, gk)
                    np.add.at(gm, i, -gk)
                    np.add.at(tm, i, f * ti)
                    np.add.at(tm, k, f * tk)
                    m = np.einsum("pa,pb->ab", R, gk)
                    vm += 0.5 * (m + m.T)
            if do_p:
                pdi, pdk, pga = self._pair_params(i, k)
                lam = thole_lambda(r, pdi, pdk, pga, 6)
                zc = np.zeros(P)
                zQ = np.zeros((P, 3, 3))
                Bp = bn - (1 - sc[:, 2, None] * lam) * rr
                Bd = bn - (1 - sc[:, 1, None] * lam) * rr
                Bu = bn - (1 - sc[:, 3, None] * lam) * rr
                for B_ in (Bp, Bd, Bu):
                    B_[:, 0] = 0.0
                hf = 0.5 * f
                # permanent(i) - induced(k)  and induced(i) - permanent(k), p-scaled with ud, d-scaled with up
                for (uu, B_, is_e) in ((ud, Bp, True), (up, Bd, False)):
                    U1, g1, ti1, _ = pair_interaction(R, ci, di, Qi, zc, uu[k], zQ, B_, do_g)
                    U2, g2, _, tk2 = pair_interaction(R, zc, uu[i], zQ, ck, dk, Qk, B_, do_g)
                    if is_e:
                        ep += hf * (U1.sum() + U2.sum())
                        nep += int((sc[:, 2] != 0).sum())
                    if do_g:
                        gk = hf * (g1 + g2)
                        np.add.at(gp, k, gk)
                        np.add.at(gp, i, -gk)
                        np.add.at(tp, i, hf * ti1)
                        np.add.at(tp, k, hf * tk2)
                        m = np.einsum("pa,pb->ab", R, gk)
                        vp += 0.5 * (m + m.T)
                if do_g and s.poltyp == "MUTUAL":
                    _, g1, _, _ = pair_interaction(R, zc, ud[i], zQ, zc, up[k], zQ, Bu, True)
                    _, g2, _, _ = pair_interaction(R, zc, up[i], zQ, zc, ud[k], zQ, Bu, True)
                    gk = hf * (g1 + g2)
                    np.add.at(gp, k, gk)
                    np.add.at(gp, i, -gk)
                    m = np.einsum("pa,pb->ab", R, gk)
                    vp += 0.5 * (m + m.T)
        return dict(em=em, ep=ep, nem=nem, nep=nep, gm=gm, gp=gp, tm=tm, tp=tp, vm=vm, vp=vp)

    _D1 = np.array([2, 5, 8, 9, 11, 16, 18, 14, 15, 20]) - 1
    _D2 = np.array([3, 8, 6, 10, 14, 12, 19, 16, 20, 17]) - 1
    _D3 = np.array([4, 9, 10, 7, 15, 17, 13, 20, 18, 19]) - 1

    @staticmethod
    def _trq_cmp_cphi(cmp_, cphi):
        t = np.zeros((cmp_.shape[0], 3))
        c, p = cmp_, cphi
        t[:, 0] = (c[:, 3] * p[:, 2] - c[:, 2] * p[:, 3] + 2 * (c[:, 6] - c[:, 5]) * p[:, 9] + c[:, 8] * p[:, 7]
                   + c[:, 9] * p[:, 5] - c[:, 7] * p[:, 8] - c[:, 9] * p[:, 6])
        t[:, 1] = (c[:, 1] * p[:, 3] - c[:, 3] * p[:, 1] + 2 * (c[:, 4] - c[:, 6]) * p[:, 8] + c[:, 7] * p[:, 9]
                   + c[:, 8] * p[:, 6] - c[:, 8] * p[:, 4] - c[:, 9] * p[:, 7])
        t[:, 2] = (c[:, 2] * p[:, 1] - c[:, 1] * p[:, 2] + 2 * (c[:, 5] - c[:, 4]) * p[:, 7] + c[:, 7] * p[:, 4]
                   + c[:, 9] * p[:, 8] - c[:, 7] * p[:, 5] - c[:, 8] * p[:, 9])
        return t

    @staticmethod
    def _vir_cmp_cphi(c, p):
        v = np.zeros((3, 3))
        vxx = -c[:, 1] * p[:, 1] - 2 * c[:, 4] * p[:, 4] - c[:, 7] * p[:, 7] - c[:, 8] * p[:, 8]
        vxy = (-0.5 * (c[:, 2] * p[:, 1] + c[:, 1] * p[:, 2]) - (c[:, 4] + c[:, 5]) * p[:, 7]
               - 0.5 * c[:, 7] * (p[:, 4] + p[:, 5]) - 0.5 * (c[:, 8] * p[:, 9] + c[:, 9] * p[:, 8]))
        vxz = (-0.5 * (c[:, 3] * p[:, 1] + c[:, 1] * p[:, 3]) - (c[:, 4] + c[:, 6]) * p[:, 8]
               - 0.5 * c[:, 8] * (p[:, 4] + p[:, 6]) - 0.5 * (c[:, 7] * p[:, 9] + c[:, 9] * p[:, 7]))
        vyy = -c[:, 2] * p[:, 2] - 2 * c[:, 5] * p[:, 5] - c[:, 7] * p[:, 7] - c[:, 9] * p[:, 9]
        vyz = (-0.5 * (c[:, 3] * p[:, 2] + c[:, 2] * p[:, 3]) - (c[:, 5] + c[:, 6]) * p[:, 9]
               - 0.5 * c[:, 9] * (p[:, 5] + p[:, 6]) - 0.5 * (c[:, 7] * p[:, 8] + c[:, 8] * p[:, 7]))
        vzz = -c[:, 3] * p[:, 3] - 2 * c[:, 6] * p[:, 6] - c[:, 8] * p[:, 8] - c[:, 9] * p[:, 9]
        v[0, 0], v[1, 1], v[2, 2] = vxx.sum(), vyy.sum(), vzz.sum()
        v[0, 1] = v[1, 0] = vxy.sum()
        v[0, 2] = v[2, 0] = vxz.sum()
        v[1, 2] = v[2, 1] = vyz.sum()
        return v

    def _frac_grad(self, f123):
        """(f1*nfft1, f2*nfft2, f3*nfft3) -> Cartesian h (recipa.x*f1 + recipb.x*f2 + ...)."""
        nf = np.array(self.s.nfft, float)
        return (f123 * nf) @ self.recip

    def empole_recip(self, vers):
        """empoleEwaldRecip (src/cu/hippo/empole.cu:202-336)."""
        rp = self._ensure_rpole()
        cmp_ = self.rpole_to_cmp(rp)
        fmp = self.cmp_to_fmp(cmp_)
        grid, e_conv, v_conv = self.pme_convolve(self.grid_mpole(fmp), want_ev=True)
        fphi = self.fphi_gather(grid.real, 20)
        cphi = self.fphi_to_cphi(fphi)
        self._recip_m = dict(e=e_conv, v=v_conv, cmp=cmp_, fmp=fmp, fphi=fphi, cphi=cphi)
        f = self.f
        out = dict(e=0.5 * f * (fmp * fphi[:, :10]).sum())
        if vers & GRAD:
            f123 = np.stack([(fmp * fphi[:, self._D1]).sum(1), (fmp * fphi[:, self._D2]).sum(1),
                             (fmp * fphi[:, self._D3]).sum(1)], 1)
            out["g"] = f * self._frac_grad(f123)
            out["t"] = f * self._trq_cmp_cphi(cmp_, cphi)
            out["v"] = f * self._vir_cmp_cphi(cmp_, cphi) + v_conv
        return out

    def epolar_recip_self(self, vers):
        """epolarEwaldRecipSelf (src/cu/epolarrecip.cu:414-511); needs empole_recip()/dfield() state."""
        s = self.s
        f = self.f
        rp = self._ensure_rpole()
        m = self._recip_m
        cmp_, fmp, fphi, cphi = m["cmp"], m["fmp"], m["fphi"], m["cphi"]
        ud, up = self.uind, self.uinp
        a = self.pme_setup()["a"]
        fud, fup = ud @ a, up @ a
        out = {}
        aew = s.aewald
        # recip energy (dot of fractional dipoles with the permanent potential gradient) + self energy
        e = 0.5 * f * (fud * fphi[:, 1:4]).sum()
        e += (-2.0 * f * aew ** 3 / 3.0 / SQRTPI) * (rp[:, 1:4] * ud).sum()
        out["e"] = e
        if not (vers & GRAD):
            return out
        grid, _, _ = self.pme_convolve(self.grid_uind(fud, fup))
        fphid = self.fphi_gather(grid.real, 10)
        fphip = self.fphi_gather(grid.imag, 10)
        fphidp = self.fphi_gather(grid.real + grid.imag, 20)
        d1, d2, d3 = self._D1, self._D2, self._D3
        f123 = np.zeros((self.n, 3))
        for c, dd in enumerate((d1, d2, d3)):
            j = dd[1:4]
            acc = ((fud + fup) * fphi[:, j]).sum(1)
            if s.poltyp == "MUTUAL":
                acc += (fud * fphip[:, j]).sum(1) + (fup * fphid[:, j]).sum(1)
            acc += (fmp * fphidp[:, dd]).sum(1)
            f123[:, c] = 0.5 * acc
        out["g"] = f * self._frac_grad(f123)
        fphidp_s = fphidp.copy()
        fphidp_s[:, :] *= 0.5 * f
        cphidp = self.fphi_to_cphi(fphidp_s)
        ubar = 0.5 * (ud + up)
        t = self._trq_cmp_cphi(cmp_, cphidp)
        t += (f * 4.0 / 3.0 * aew ** 3 / SQRTPI) * np.cross(rp[:, 1:4], ubar)
        out["t"] = t
        if vers & VIRIAL:
            v = -m["v"].copy()
            at = a.T
            cphid = (f * fphid[:, 1:4]) @ at
            cphip = (f * fphip[:, 1:4]) @ at
            cphi_f = f * cphi
            usum = ud + up
            v2 = np.zeros((3, 3))
            # permanent multipoles with the averaged induced potential
            v2 += self._vir_cmp_cphi(cmp_, cphidp)
            # induced dipoles with the permanent potential gradient
            mat = -0.5 * np.einsum("na,nb->ab", usum, cphi_f[:, 1:4])
            v2 += 0.5 * (mat + mat.T)
            if s.poltyp == "MUTUAL":
                mat = -0.5 * (np.einsum("na,nb->ab", up, cphid) + np.einsum("na,nb->ab", ud, cphip))
                v2 += 0.5 * (mat + mat.T)
            v += v2
            # structure-factor cross term of (M + up) and (M + ud)   (epolarrecip.cu:476-509)
            cp_ = cmp_.copy()
            cp_[:, 1:4] += up
            cd_ = cmp_.copy()
            cd_[:, 1:4] += ud
            Qp = np.fft.fftn(self.grid_mpole(self.cmp_to_fmp(cp_)))
            Qd = np.fft.fftn(self.grid_mpole(self.cmp_to_fmp(cd_)))
            expterm, h, hsq, term = self.conv_factors()
            struc2 = Qd.real * Qp.real + Qd.imag * Qp.imag
            eterm = 0.5 * f * expterm * struc2
            with np.errstate(divide="ignore", invalid="ignore"):
                vterm = np.where(hsq > 0, (2.0 / hsq) * (1 - term) * eterm, 0.0)
            v += np.einsum("zyx,zyxa,zyxb->ab", vterm, h, h) - np.eye(3) * eterm.sum()
            out["v"] = v
        return out

    def empole_self(self):
        s = self.s
        rp = self._ensure_rpole()
        aew = s.aewald
        fterm = -self.f * aew / SQRTPI
        a2 = 2 * aew * aew
        cii = rp[:, 0] ** 2
        dii = (rp[:, 1:4] ** 2).sum(1)
        qii = 2 * (rp[:, 7] ** 2 + rp[:, 8] ** 2 + rp[:, 9] ** 2) + rp[:, 4] ** 2 + rp[:, 5] ** 2 + rp[:, 6] ** 2
        return fterm * (cii + a2 * (dii / 3 + 2 * a2 * qii / 5)).sum()

    def energy(self, vers=V1, dot_energy=None):
        """The electrostatic part of energy(vers): empole + epolar (or the fused emplar) + torque.
        Returns dict with em, ep, esum, grad (n,3), virial (3,3), component breakdown."""
        s = self.s
        n = self.n
        do_g = bool(vers & GRAD)
        do_v = bool(vers & VIRIAL)
        self.rotpole()
        res = {}
        em = ep = 0.0
        grad = np.zeros((n, 3))
        trq = np.zeros((n, 3))
        vir = np.zeros((3, 3))
        if s.use_polar:
            self.induce()
        rs = self._real_space(vers, s.use_mpole, s.use_polar)
        if s.use_mpole:
            res["em_real"] = rs["em"]
            em += rs["em"]
            res["nem"] = rs["nem"]
            grad += rs["gm"]
            trq += rs["tm"]
            vir += rs["vm"]
            if s.use_ewald:
                res["em_self"] = self.empole_self()
                rc = self.empole_recip(vers)
                res["em_recip"] = rc["e"]
                em += res["em_self"] + rc["e"]
                if do_g:
                    grad += rc["g"]
                    trq += rc["t"]
                    vir += rc["v"]
        if s.use_polar:
            res["nep"] = rs["nep"]
            res["ep_real"] = rs["ep"]
            ep_pair = rs["ep"]
            grad += rs["gp"]
            trq += rs["tp"]
            vir += rs["vp"]
            if s.use_ewald:
                if not s.use_mpole:
                    self.empole_recip(vers)     # provides fmp/fphi/cmp/cphi state (epolarrecip.cu:419-420)
                rc = self.epolar_recip_self(vers)
                res["ep_recip_self"] = rc["e"]
                ep_pair += rc["e"]
                if do_g:
                    grad += rc["g"]
                    trq += rc["t"]
                    if do_v:
                        vir += rc["v"]
            pinv = 1.0 / np.maximum(s.polarity, 1e-16)
            ep_dot = -0.5 * self.f * (pinv[:, None] * self.uind * self.udirp).sum()
            res["ep_pair"], res["ep_dot"] = ep_pair, ep_dot
            use_dot = (not (vers & ANALYZ)) if dot_energy is None else dot_energy
            ep = ep_dot if use_dot else ep_pair
        if do_g:
            vir += self.torque(trq, grad, do_v)
        res.update(em=em, ep=ep, esum=em + ep, grad=grad, virial=vir, trq=trq)
        return res
