"""CPU ORACLE (test infrastructure, float64 numpy) for the buffered 14-7 van der Waals term.

THIS IS NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may
import it.  It restates

  ehalReduceXyz / ehalResolveGradient   src/cu/ehal.cu:13-70 (reduced hydrogen sites, force hand-back)
  pair_hal_v2 + switchTaper5            include/seq/pair_hal.h:52-92, include/math/switch.h:23-32
  the pair sweep and its counters       src/cu/ehal_cu1.cc (energy, gradient, virial on reduced sites, nev)
  evdw(): long-range correction         src/evdw.cpp:493-512

PARITY PINNING: pinned against the reference's golden vectors NaCl-1 (test/nacl.cpp:36-176: energy,
gradient, virial, count at three separations and with vdw-correction) and Local-Frame2-1/2
(test/localframe2.cpp:46-98: energy and count in a triclinic and a monoclinic cell).
"""
from __future__ import annotations

import numpy as np


class VdwOracle:
    def __init__(self, system):
        self.s = system
        self.v = system.vdw
        if self.v is None:
            raise ValueError("system has no buffered 14-7 term")
        self.n = system.n
        self.xyz = np.array(system.xyz, float)
        self.lvec = np.array(system.lvec, float)
        self.recip = np.array(system.recip, float)

    def set_xyz(self, xyz):
        self.xyz = np.array(xyz, float)

    def reduced(self):
        """xred = kred (x_i - x_iv) + x_iv   (src/cu/ehal.cu:17-27)."""
        v = self.v
        xiv = self.xyz[v.ired]
        return v.kred[:, None] * (self.xyz - xiv) + xiv

    def image(self, dr):
        fr = dr @ self.recip.T
        fr -= np.floor(fr + 0.5)
        return fr @ self.lvec.T

    def pairs(self, xr):
        """All i<k with minimum-image distance <= off (cell binning when the box allows, else brute force)."""
        off = self.v.cutoff
        n = self.n
        out_i, out_k = [], []
        L = np.diag(self.lvec)
        ortho = np.allclose(self.lvec, np.diag(L)) and off < 1e6
        ncell = np.maximum(1, np.floor(L / off).astype(int)) if ortho else None
        if ortho and ncell.min() >= 3 and n > 2000:
            fr = (xr / L) % 1.0
            ci = np.minimum((fr * ncell).astype(int), ncell - 1)
            key = (ci[:, 0] * ncell[1] + ci[:, 1]) * ncell[2] + ci[:, 2]
            order = np.argsort(key, kind="stable")
            ks = key[order]
            start = np.searchsorted(ks, np.arange(ncell.prod() + 1))
            for c in range(int(ncell.prod())):
                a = order[start[c]:start[c + 1]]
                if a.size == 0:
                    continue
                cz = c % ncell[2]
                cy = (c // ncell[2]) % ncell[1]
                cx = c // (ncell[1] * ncell[2])
                nb = set()
                for dx in (-1, 0, 1):
                    for dy in (-1, 0, 1):
                        for dz in (-1, 0, 1):
                            nb.add((((cx + dx) % ncell[0]) * ncell[1] + (cy + dy) % ncell[1]) * ncell[2] + (cz + dz) % ncell[2])
                b = np.concatenate([order[start[q]:start[q + 1]] for q in sorted(nb)])
                d = xr[b][None, :, :] - xr[a][:, None, :]
                d -= L * np.round(d / L)
                r2 = np.einsum("ikc,ikc->ik", d, d)
                ii, kk = np.nonzero(r2 <= off * off)
                gi, gk = a[ii], b[kk]
                m = gi < gk
                out_i.append(gi[m])
                out_k.append(gk[m])
        else:
            for i0 in range(0, n, 512):
                a = np.arange(i0, min(n, i0 + 512))
                d = self.image((xr[None, :, :] - xr[a][:, None, :]).reshape(-1, 3)).reshape(a.size, n, 3)
                r2 = np.einsum("ikc,ikc->ik", d, d)
                ii, kk = np.nonzero(r2 <= off * off)
                gi = a[ii]
                m = gi < kk
                out_i.append(gi[m])
                out_k.append(kk[m])
        i = np.concatenate(out_i) if out_i else np.zeros(0, int)
        k = np.concatenate(out_k) if out_k else np.zeros(0, int)
        return i, k

    def pair_terms(self, r, rv, eps):
        """Buffered 14-7 energy and dE/dr of every pair, tapered between `taper` and `cutoff` (include/seq/pair_hal.h:49-86)."""
        v = self.v
        ghal, dhal = v.ghal, v.dhal
        rho = r / np.where(rv > 0, rv, 1.0)
        rho6 = rho ** 6
        rho7 = rho6 * rho
        s1 = 1.0 / (rho + dhal) ** 7
        s2 = 1.0 / (rho7 + ghal)
        t1 = (1.0 + dhal) ** 7 * s1
        t2 = (1.0 + ghal) * s2
        e = eps * t1 * (t2 - 2.0)
        dt1 = -7.0 * (rho + dhal) ** 6 * t1 * s1
        dt2 = -7.0 * rho6 * t2 * s2
        de = eps * (dt1 * (t2 - 2.0) + t1 * dt2) / np.where(rv > 0, rv, 1.0)
        cut, off = v.taper, v.cutoff
        sw = r > cut
        if sw.any():
            x = (r[sw] - off) / (cut - off)
            taper = x ** 3 * (6.0 * x * x - 15.0 * x + 10.0)
            dtaper = 30.0 * (x * (1.0 - x)) ** 2 / (cut - off)
            de[sw] = e[sw] * dtaper + de[sw] * taper
            e[sw] = e[sw] * taper
        return e, de

    def ehal(self, want_grad=True):
        """Returns dict(ev, nev, grad (n,3) on the real atoms, virial (3,3))."""
        v = self.v
        n = self.n
        xr = self.reduced()
        i, k = self.pairs(xr)
        scale = np.ones(i.shape[0])
        if v.vexclude.shape[0]:
            code = i.astype(np.int64) * n + k
            ex = v.vexclude[:, 0].astype(np.int64) * n + v.vexclude[:, 1]
            order = np.argsort(ex)
            pos = np.searchsorted(ex[order], code)
            pos = np.minimum(pos, ex.shape[0] - 1)
            hit = ex[order][pos] == code
            scale[hit] = v.vexclude_scale[order][pos[hit]]
        d = self.image(xr[i] - xr[k])                     # xr = xi - xk, as ehal_cu1
        r = np.sqrt((d * d).sum(1))
        rv = v.radmin[v.jvdw[i], v.jvdw[k]]
        eps = v.epsilon[v.jvdw[i], v.jvdw[k]] * scale
        ghal, dhal = v.ghal, v.dhal
        e, de = self.pair_terms(r, rv, eps)
        ev = float(e.sum())
        nev = int(((scale != 0) & (e != 0)).sum())
        out = dict(ev=ev, nev=nev, npairs=int(i.shape[0]))
        vol = float(abs(np.linalg.det(self.lvec)))
        if v.elrc_vol:
            out["ev"] += v.elrc_vol / vol
        if want_grad:
            f = (de / r)[:, None] * d                      # dE/dx_i on the reduced site
            gred = np.zeros((n, 3))
            np.add.at(gred, i, f)
            np.add.at(gred, k, -f)
            vir = d.T @ f
            # ehalResolveGradient (src/cu/ehal.cu:34-62)
            g = np.zeros((n, 3))
            own = v.ired == np.arange(n)
            g[own] += gred[own]
            h = ~own
            np.add.at(g, np.nonzero(h)[0], gred[h] * v.kred[h][:, None])
            np.add.at(g, v.ired[h], gred[h] * (1.0 - v.kred[h])[:, None])
            if v.vlrc_vol:
                vir = vir + np.eye(3) * (v.vlrc_vol / vol)
            out["grad"] = g
            out["virial"] = vir
        return out
