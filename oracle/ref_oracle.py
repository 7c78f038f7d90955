"""The oracle's solver and energy assembly driven by the REFERENCE'S OWN operators (oracle/_ref) -- TEST INFRASTRUCTURE ONLY.

RefOracle keeps the control flow of oracle/amoeba_ref.py (the PCG loop of induceMutualPcg1, the sparse preconditioner,
numpy's FFT) and replaces every O(pairs), O(grid) and per-atom assembly operator by the reference's compiled code:
empoleChgpenEwaldRecip_acc / epolarEwaldRecipSelf_acc (src/acc/hippo/empole.cpp, src/acc/amoeba/epolarewald.cpp) for the
reciprocal energy / gradient / torque / virial, pair_dfield / pair_ufield / pair_mpole / pair_polar (include/seq/*.h) for the real-space sweeps and
gridMpole / gridUind / pmeConv / fphiMpole / fphiUind2 ... (src/acc/pme.cpp) for the reciprocal ones, chkpole / rotpole / torque
(src/acc/amoeba/rotpole.cpp, torque.cpp) for the local frames.  Two uses: it pins the
oracle's converged dipoles and energies at dhfr2 size to the reference's arithmetic (tests/test_ref_arith.py), and it is the
CPU baseline bench.py times ("reference arithmetic, one core") -- about 20x faster than the vectorised-numpy operators.
"""
import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import ref_bridge
from .amoeba_ref import SQRTPI, Oracle

_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int)


class RefOracle(Oracle):
    def __init__(self, system, chunk=20000, threads=1):
        super().__init__(system, chunk)
        self.threads = threads      # > 1: the pair sweeps run on that many host threads (bench.py's CPU legs only)
        self._rs = C.CDLL(os.path.join(ref_bridge.HERE, "_ref", "libref_realspace.so"))
        self._rs.ref_field_real.argtypes = [C.c_int, C.c_int, C.c_longlong, _IP, _IP, _DP, _DP, _DP, _DP, _DP, _DP, _DP, C.c_double, C.c_int,
                                            _DP, _DP]
        self._refpme = None
        self._plist = None
        self._pmat = None

    def set_xyz(self, xyz):
        super().set_xyz(xyz)
        self._refpme = None
        self._plist = None
        self._pmat = None

    # ---- sparse preconditioner: BUILD once per set of positions, APPLY per iteration -- the split the reference makes
    # (sparsePrecondBuild / sparsePrecondApply, src/acc/amoeba/induce.cpp:296-468: `minv`, 6 numbers per pair, is filled once per
    # induce()).  Oracle.precond recomputes the pair tensors in every application, which is fine for a checker and made the
    # CPU baseline of bench.py spend two thirds of its induce() there.  Same arithmetic, summed in CSR row order.
    def precond(self, rd, rp_):
        s = self.s
        if not (s.pcgprec and s.usolve_cutoff > 0):
            return super().precond(rd, rp_)
        if getattr(self, "_pmat", None) is None:
            from scipy import sparse
            from .amoeba_ref import thole_lambda
            i, k, R, r = self.pairs(s.usolve_cutoff)
            sc = self._scales(i, k)
            pdi, pdk, pga = self._pair_params(i, k)
            lam = thole_lambda(r, pdi, pdk, pga, 3)
            polik = s.polarity[i] * s.polarity[k]
            rr3 = sc[:, 3] * lam[:, 1] * polik / r ** 3
            rr5 = 3 * sc[:, 3] * lam[:, 2] * polik / r ** 5
            M = np.einsum("pa,pb->pab", R, R) * rr5[:, None, None] - np.eye(3)[None] * rr3[:, None, None]      # symmetric 3x3 per pair
            a = np.arange(3)
            rows = np.concatenate([(3 * i[:, None, None] + a[None, :, None]) + 0 * a[None, None, :],
                                   (3 * k[:, None, None] + a[None, :, None]) + 0 * a[None, None, :]]).ravel()
            cols = np.concatenate([(3 * k[:, None, None] + a[None, None, :]) + 0 * a[None, :, None],
                                   (3 * i[:, None, None] + a[None, None, :]) + 0 * a[None, :, None]]).ravel()
            vals = np.concatenate([M, M]).ravel()
            self._pmat = sparse.csr_matrix((vals, (rows, cols)), shape=(3 * self.n, 3 * self.n))
        pol = s.polarity[:, None]
        zd = s.uaccel * pol * rd + (self._pmat @ np.ascontiguousarray(rd).reshape(-1)).reshape(-1, 3)
        zp = s.uaccel * pol * rp_ + (self._pmat @ np.ascontiguousarray(rp_).reshape(-1)).reshape(-1, 3)
        return zd, zp

    # ---- shared inputs of the pair sweeps
    def _pairs_c(self):
        if self._plist is None:
            i, k, R, r = self.pairs(self.s.ewald_cutoff)
            self._plist = (np.ascontiguousarray(i, np.int32), np.ascontiguousarray(k, np.int32), np.ascontiguousarray(R, np.float64),
                           np.ascontiguousarray(self._scales(i, k), np.float64),
                           np.ascontiguousarray(self._pair_params(i, k)[2], np.float64), np.ascontiguousarray(self.s.pdamp, np.float64))
        return self._plist

    def _rpme(self):
        if self._refpme is None:
            self._refpme = ref_bridge.RefPME(self)
        return self._refpme

    def _field_real(self, mode, ud=None, up=None):
        i, k, R, sc, pga, pd = self._pairs_c()
        rp = np.ascontiguousarray(self._ensure_rpole(), np.float64)
        fd, fp = np.zeros((self.n, 3)), np.zeros((self.n, 3))
        dp = ref_bridge._dp
        u1 = None if ud is None else np.ascontiguousarray(ud, np.float64)
        u2 = None if up is None else np.ascontiguousarray(up, np.float64)

        def sweep(lo, hi, ofd, ofp):
            self._rs.ref_field_real(mode, self.n, hi - lo, i[lo:hi].ctypes.data_as(_IP), k[lo:hi].ctypes.data_as(_IP), dp(R[lo:hi]), dp(sc[lo:hi]),
                                    dp(rp), dp(pd), dp(pga[lo:hi]), None if u1 is None else dp(u1), None if u2 is None else dp(u2),
                                    float(self.s.aewald), int(bool(self.s.use_ewald)), dp(ofd), dp(ofp))

        T = max(1, min(int(self.threads), len(i) // 4096 or 1))
        if T == 1:
            sweep(0, len(i), fd, fp)
            return fd, fp
        # bench.py's CPU legs: the pair list cut into T contiguous slices, one host thread each (ctypes releases the GIL), partial
        # fields summed -- the reference's host build runs this loop serially, so this is MORE than its CPU path can do
        cuts = np.linspace(0, len(i), T + 1).astype(int)
        parts = [(np.zeros((self.n, 3)), np.zeros((self.n, 3))) for _ in range(T)]
        with ThreadPoolExecutor(T) as ex:
            list(ex.map(lambda t: sweep(int(cuts[t]), int(cuts[t + 1]), parts[t][0], parts[t][1]), range(T)))
        for a, b in parts:
            fd += a
            fp += b
        return fd, fp

    # ---- local frames (src/acc/amoeba/rotpole.cpp, torque.cpp)
    def chkpole(self):
        self.pole, _ = ref_bridge.rotpole(self.xyz, self.zaxis, self.pole)

    def rotpole(self):
        self.pole, self.rpole = ref_bridge.rotpole(self.xyz, self.zaxis, self.pole)
        return self.rpole

    def torque(self, trq, grad, do_v=False):
        g, v = ref_bridge.torque(self.xyz, self.zaxis, trq)
        grad += g
        return v

    # ---- operators
    def grid_mpole(self, fmp):
        return self._rpme().grid_mpole(fmp)

    def grid_uind(self, fud, fup):
        return self._rpme().grid_uind(fud, fup)

    def pme_convolve(self, qgrid, want_ev=False):
        return self._rpme().convolve(qgrid, want_ev)

    def fphi_gather(self, grid, nder):
        P = self._rpme()
        if nder == 20:
            return P.fphi_mpole(np.asarray(grid, complex))
        return super().fphi_gather(grid, nder)

    def cmp_to_fmp(self, cmp_):
        return self._rpme().cmp_to_fmp(cmp_)

    def fphi_to_cphi(self, fphi):
        return self._rpme().fphi_to_cphi(fphi)

    def dfield(self, real_only=False):
        s = self.s
        rp = self._ensure_rpole()
        fd, fp = self._field_real(0)
        if s.use_ewald and not real_only:
            cmp_ = self.rpole_to_cmp(rp)
            fmp = self.cmp_to_fmp(cmp_)
            grid, e, v = self.pme_convolve(self.grid_mpole(fmp), want_ev=True)
            self._recip_m = dict(e=e, v=v, cmp=cmp_, fmp=fmp)
            fphi = self._rpme().fphi_mpole()            # the convolved grid is already in place
            cphi = self.fphi_to_cphi(fphi)
            self._recip_m.update(fphi=fphi, cphi=cphi)
            rec = -cphi[:, 1:4] + (4.0 / 3.0 * s.aewald ** 3 / SQRTPI) * rp[:, 1:4]
            fd += rec
            fp += rec
        return fd, fp

    def ufield(self, ud, up, real_only=False):
        s = self.s
        fd, fp = self._field_real(1, ud, up)
        if s.use_ewald and not real_only:
            P = self._rpme()
            a = self.pme_setup()["a"]
            fud, fup = P.cuind_to_fuind(ud, up)
            P.convolve(P.grid_uind(fud, fup))
            f1, f2 = P.fphi_uind2()
            term = 4.0 / 3.0 * s.aewald ** 3 / SQRTPI
            fd += term * ud - f1[:, 1:4] @ a.T
            fp += term * up - f2[:, 1:4] @ a.T
        return fd, fp

    # ---- per-atom reciprocal energy / gradient / torque / virial assembly (src/acc/hippo/empole.cpp:260-388,
    #      src/acc/amoeba/epolarewald.cpp:358-667, compiled unmodified; always evaluated as calc::v1)
    def empole_recip(self, vers):
        return ref_bridge.recip_mpole(self._rpme(), self._ensure_rpole())

    def epolar_recip_self(self, vers):
        return ref_bridge.recip_polar(self._rpme(), self.uind, self.uinp)

    def _real_space(self, vers, do_m, do_p):
        r = ref_bridge.realspace(self, self.uind if do_p else None, self.uinp if do_p else None, threads=self.threads)
        i, k, R, sc, pga, pd = self._pairs_c()
        return dict(em=r["em"] if do_m else 0.0, ep=r["ep"] if do_p else 0.0, nem=int((sc[:, 0] != 0).sum()), nep=int((sc[:, 2] != 0).sum()),
                    gm=r["gm"], gp=r["gp"], tm=r["tm"], tp=r["tp"], vm=r["vm"], vp=r["vp"])
