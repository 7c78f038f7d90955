// Per-interaction math of the AMOEBA valence terms: energy, gradient on every atom of the interaction and the
// internal virial.  Functional forms follow the reference's include/seq/{bond,angle,strbnd,urey,opbend,torsion,
// pitors,tortor}.h (cited per function); the gradients are our own derivations, written around two shared
// building blocks -- d(theta)/d(u,w) of a bond angle and d(phi)/d(a,b,c,d) of a dihedral in Blondel-Karplus form --
// instead of the reference's per-term expansions.
//
// Everything is __host__ __device__ and templated on the arithmetic type so that tests/valmath_host.cpp can run
// the very same code on the CPU (float and double) against the oracle without a GPU.
//
// Conventions: X[k] are the positions of the atoms of one interaction RELATIVE TO ITS FIRST ATOM (all terms are
// translation invariant; the caller subtracts in double).  G[k] receives dE/dX[k].  Angles are in degrees where
// the force-field polynomials want them (radian = 180/pi).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define VM_HD __host__ __device__ __forceinline__
#else
#define VM_HD inline
#endif

namespace vm {
template <class R>
struct V3 {
   R x, y, z;
};
template <class R>
VM_HD V3<R> mk(R x, R y, R z)
{
   V3<R> v;
   v.x = x, v.y = y, v.z = z;
   return v;
}
template <class R>
VM_HD V3<R> operator+(V3<R> a, V3<R> b) { return mk<R>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class R>
VM_HD V3<R> operator-(V3<R> a, V3<R> b) { return mk<R>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class R>
VM_HD V3<R> operator-(V3<R> a) { return mk<R>(-a.x, -a.y, -a.z); }
template <class R>
VM_HD V3<R> operator*(R s, V3<R> a) { return mk<R>(s * a.x, s * a.y, s * a.z); }
template <class R>
VM_HD R dot(V3<R> a, V3<R> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class R>
VM_HD V3<R> cross(V3<R> a, V3<R> b) { return mk<R>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
template <class R>
VM_HD R vmax(R a, R b) { return a > b ? a : b; }
template <class R>
VM_HD R vmin(R a, R b) { return a < b ? a : b; }

template <class R>
struct Consts {     // order of valparams.CONST_NAMES
   R bndunit, cbnd, qbnd, angunit, cang, qang, pang, sang, stbnunit, ureyunit, cury, qury, opbunit, copb, qopb, popb, sopb, torsunit,
      ptorunit, ttorunit;
};

#define VM_RADIAN 57.29577951308232088

// ---- building block 1: theta = angle(u, w) in radians and its derivatives with respect to u and w.
//      d(theta)/du = u x (u x w) / (|u|^2 |u x w|),  d(theta)/dw = -w x (u x w) / (|w|^2 |u x w|)
template <class R>
VM_HD R angle_and_grad(V3<R> u, V3<R> w, V3<R>& dth_du, V3<R>& dth_dw)
{
   const R ru2 = dot(u, u), rw2 = dot(w, w);
   V3<R> p = cross(u, w);
   const R rp = vmax<R>((R)sqrt(dot(p, p)), (R)1e-6);
   R cs = dot(u, w) / (R)sqrt(ru2 * rw2);
   cs = vmin<R>((R)1, vmax<R>((R)-1, cs));
   dth_du = ((R)1 / (ru2 * rp)) * cross(u, p);
   dth_dw = ((R)-1 / (rw2 * rp)) * cross(w, p);
   return (R)acos(cs);
}

// ---- building block 2: dihedral a-b-c-d with Tinker's sign (include/seq/torsion.h:70-95): cos, sin and
//      d(phi)/d(a,b,c,d).  Returns false for a degenerate (collinear) geometry, as the reference skips it.
template <class R>
VM_HD bool dihedral(V3<R> a, V3<R> b, V3<R> c, V3<R> d, R& cs, R& sn, V3<R> dphi[4])
{
   V3<R> ba = b - a, cb = c - b, dc = d - c;
   V3<R> t = cross(ba, cb), u = cross(cb, dc);
   const R rt2 = dot(t, t), ru2 = dot(u, u);
   const R rtru = (R)sqrt(rt2 * ru2);
   if (!(rtru != 0))
      return false;
   const R rcb2 = dot(cb, cb);
   const R rcb = (R)sqrt(rcb2);
   cs = dot(t, u) / rtru;
   sn = dot(cb, cross(t, u)) / (rcb * rtru);
   const R fa = -rcb / rt2, fd = rcb / ru2;
   const R pb = dot(ba, cb) / (rt2 * rcb), pc = dot(dc, cb) / (ru2 * rcb);
   dphi[0] = fa * t;
   dphi[3] = fd * u;
   dphi[1] = (pb - fa) * t + pc * u;
   dphi[2] = (-pb) * t + (-pc - fd) * u;
   return true;
}

// unit * k * dt^2 (1 + c dt + q dt^2 + p dt^3 + s dt^4) and its derivative with respect to dt
template <class R>
VM_HD R poly6(R dt, R k, R unit, R c, R q, R p, R s, R& dedt)
{
   const R dt2 = dt * dt, dt3 = dt2 * dt, dt4 = dt2 * dt2;
   dedt = unit * k * dt * ((R)2 + (R)3 * c * dt + (R)4 * q * dt2 + (R)5 * p * dt3 + (R)6 * s * dt4);
   return unit * k * dt2 * ((R)1 + c * dt + q * dt2 + p * dt3 + s * dt4);
}

// ---- bond stretch, atoms {a,b}: include/seq/bond.h:44-53.  Also Urey-Bradley over {a,c} (include/seq/urey.h:37-44).
template <class R>
VM_HD R stretch(const V3<R>* X, R ideal, R force, R unit, R cub, R qrt, V3<R>* G)
{
   V3<R> ab = X[0] - X[1];
   const R r = (R)sqrt(dot(ab, ab));
   const R dt = r - ideal, dt2 = dt * dt;
   const R e = unit * force * dt2 * ((R)1 + cub * dt + qrt * dt2);
   const R deddt = (R)2 * unit * force * dt * ((R)1 + (R)1.5 * cub * dt + (R)2 * qrt * dt2);
   G[0] = (deddt / r) * ab;
   G[1] = -G[0];
   return e;
}

// ---- angle bend, atoms {a,b,c[,d]}: include/seq/angle.h.  inplane: the centre b is replaced by its projection p
//      onto the plane through a, c, d (lines 156-190); the chain rule through p is carried out explicitly.
template <class R>
VM_HD R angle_bend(const V3<R>* X, bool inplane, R ideal, R force, const Consts<R>& K, V3<R>* G)
{
   const R rad = (R)VM_RADIAN;
   V3<R> dth_du, dth_dw;
   R dedt;
   if (!inplane) {
      V3<R> u = X[0] - X[1], w = X[2] - X[1];
      if (!(dot(u, u) != 0 && dot(w, w) != 0)) {
         G[0] = G[1] = G[2] = mk<R>(0, 0, 0);
         return 0;
      }
      const R th = rad * angle_and_grad(u, w, dth_du, dth_dw);
      const R e = poly6(th - ideal, force, K.angunit, K.cang, K.qang, K.pang, K.sang, dedt);
      dedt *= rad;
      G[0] = dedt * dth_du;
      G[2] = dedt * dth_dw;
      G[1] = -(G[0] + G[2]);
      return e;
   }
   V3<R> a = X[0], b = X[1], c = X[2], d = X[3];
   V3<R> ad = a - d, bd = b - d, cd = c - d;
   V3<R> t = cross(ad, cd);
   const R q = dot(t, t), s = dot(t, bd);
   V3<R> p = b - (s / q) * t;
   V3<R> u = a - p, w = c - p;
   G[0] = G[1] = G[2] = G[3] = mk<R>(0, 0, 0);
   if (!(dot(u, u) != 0 && dot(w, w) != 0))
      return 0;
   const R th = rad * angle_and_grad(u, w, dth_du, dth_dw);
   const R e = poly6(th - ideal, force, K.angunit, K.cang, K.qang, K.pang, K.sang, dedt);
   dedt *= rad;
   V3<R> gu = dedt * dth_du, gw = dedt * dth_dw;
   V3<R> gp = -(gu + gw);
   // p = b - t s/q with s = t.bd, q = t.t, t = ad x cd
   const R gpt = dot(gp, t);
   V3<R> T = (-s / q) * gp + (-gpt / q) * bd + ((R)2 * s * gpt / (q * q)) * t;      // dE/dt
   V3<R> gbd = (-gpt / q) * t;                                                      // dE/d(bd) through s
   V3<R> gad = cross(cd, T), gcd = cross(T, ad);
   G[0] = gu + gad;
   G[2] = gw + gcd;
   G[1] = gp + gbd;
   G[3] = -(gad + gcd + gbd);
   return e;
}

// ---- stretch-bend, atoms {a,b,c}: include/seq/strbnd.h:60-84
template <class R>
VM_HD R stretch_bend(const V3<R>* X, R ideal, R bl1, R bl2, R k1, R k2, R unit, V3<R>* G)
{
   const R rad = (R)VM_RADIAN;
   V3<R> u = X[0] - X[1], w = X[2] - X[1];
   const R ru = (R)sqrt(dot(u, u)), rw = (R)sqrt(dot(w, w));
   G[0] = G[1] = G[2] = mk<R>(0, 0, 0);
   if (!(ru != 0 && rw != 0))
      return 0;
   V3<R> dth_du, dth_dw;
   const R dt = rad * angle_and_grad(u, w, dth_du, dth_dw) - ideal;
   const R dr = k1 * (ru - bl1) + k2 * (rw - bl2);
   G[0] = (unit * k1 * dt / ru) * u + (unit * dr * rad) * dth_du;
   G[2] = (unit * k2 * dt / rw) * w + (unit * dr * rad) * dth_dw;
   G[1] = -(G[0] + G[2]);
   return unit * dr * dt;
}

// ---- out-of-plane bend, atoms {a, b (centre), c, d (out of plane)}: include/seq/opbend.h:84-118
template <class R>
VM_HD R opbend(const V3<R>* X, bool allinger, R force, const Consts<R>& K, V3<R>* G)
{
   const R rad = (R)VM_RADIAN;
   V3<R> ab = X[0] - X[1], cb = X[2] - X[1], db = X[3] - X[1];
   V3<R> m1 = allinger ? X[0] - X[3] : ab;      // the two in-plane vectors whose Gram determinant is cc
   V3<R> m2 = allinger ? X[2] - X[3] : cb;
   const R r1 = dot(m1, m1), r2 = dot(m2, m2), d12 = dot(m1, m2);
   const R cc = r1 * r2 - d12 * d12;
   const R ee = dot(db, cross(ab, cb));
   const R rdb2 = vmax<R>(dot(db, db), (R)1e-4);
   G[0] = G[1] = G[2] = G[3] = mk<R>(0, 0, 0);
   if (!(cc != 0))
      return 0;
   const R den = (R)sqrt(cc * rdb2);
   const R S = ee / den;
   const R sine = vmin<R>((R)1, (R)fabs(S));
   const R ang = rad * (R)asin(sine);
   R dedt;
   const R e = poly6(ang, force, K.opbunit, K.copb, K.qopb, K.popb, K.sopb, dedt);
   const R cosphi = vmax<R>((R)sqrt(vmax<R>((R)1 - S * S, (R)0)), (R)1e-8);
   const R dedS = dedt * rad * (ee < 0 ? (R)-1 : (R)1) / cosphi;
   // dS = d(ee)/den - (S/2) (d(cc)/cc + d(rdb2)/rdb2)
   const R f_ee = dedS / den, f_cc = (R)-0.5 * dedS * S / cc, f_db = (R)-0.5 * dedS * S / rdb2;
   V3<R> g_ab = f_ee * cross(cb, db), g_cb = f_ee * cross(db, ab), g_db = f_ee * cross(ab, cb) + ((R)2 * f_db) * db;
   V3<R> g_m1 = ((R)2 * f_cc) * (r2 * m1 - d12 * m2), g_m2 = ((R)2 * f_cc) * (r1 * m2 - d12 * m1);
   if (allinger) {
      G[0] = g_ab + g_m1;
      G[2] = g_cb + g_m2;
      G[3] = g_db - (g_m1 + g_m2);
      G[1] = -(g_ab + g_cb + g_db);
   } else {
      G[0] = g_ab + g_m1;
      G[2] = g_cb + g_m2;
      G[3] = g_db;
      G[1] = -(G[0] + G[2] + G[3]);
   }
   return e;
}

// ---- torsion, atoms {a,b,c,d}: include/seq/torsion.h:96-130.  prm = 6 x {amplitude, cos(phase), sin(phase)}
template <class R, class P>
VM_HD R torsion(const V3<R>* X, const P* prm, R unit, V3<R>* G)
{
   R cs, sn;
   V3<R> dphi[4];
   G[0] = G[1] = G[2] = G[3] = mk<R>(0, 0, 0);
   if (!dihedral(X[0], X[1], X[2], X[3], cs, sn, dphi))
      return 0;
   R cn = cs, snn = sn, e = 0, dedphi = 0;
   for (int n = 1; n <= 6; ++n) {
      const R v = (R)prm[3 * (n - 1)], c0 = (R)prm[3 * (n - 1) + 1], s0 = (R)prm[3 * (n - 1) + 2];
      e += v * ((R)1 + cn * c0 + snn * s0);
      dedphi += v * (R)n * (cn * s0 - snn * c0);
      const R c1 = cn * cs - snn * sn;
      snn = snn * cs + cn * sn;
      cn = c1;
   }
   dedphi *= unit;
   for (int k = 0; k < 4; ++k)
      G[k] = dedphi * dphi[k];
   return unit * e;
}

// ---- pi-orbital torsion, atoms {a,b,c,d,e,g}: include/seq/pitors.h:62-186.  The dihedral runs over the pseudo-sites
//      p = c + (a-d) x (b-d), c, d, q = d + (e-c) x (g-c).  vir6 receives the reference's own virial expression
//      (lines 178-186), which treats p and q as sites and is NOT sum r (x) g.
template <class R>
VM_HD R pitors(const V3<R>* X, R kpit, R unit, V3<R>* G, R* vir6)
{
   V3<R> a = X[0], b = X[1], c = X[2], d = X[3], e_ = X[4], g_ = X[5];
   V3<R> ad = a - d, bd = b - d, ec = e_ - c, gc = g_ - c;
   V3<R> p = c + cross(ad, bd), q = d + cross(ec, gc);
   R cs, sn;
   V3<R> dphi[4];
   for (int k = 0; k < 6; ++k)
      G[k] = mk<R>(0, 0, 0);
   for (int k = 0; k < 6; ++k)
      vir6[k] = 0;
   if (!dihedral(p, c, d, q, cs, sn, dphi))
      return 0;
   const R cos2 = cs * cs - sn * sn, sin2 = (R)2 * cs * sn;
   const R dedphi = (R)2 * unit * kpit * sin2;
   V3<R> gp = dedphi * dphi[0], gc_ = dedphi * dphi[1], gd_ = dedphi * dphi[2], gq = dedphi * dphi[3];
   V3<R> ga = cross(bd, gp), gb = cross(gp, ad), ge = cross(gc, gq), gg = cross(gq, ec);
   G[0] = ga;
   G[1] = gb;
   G[4] = ge;
   G[5] = gg;
   G[2] = gc_ + gp - (ge + gg);
   G[3] = gd_ + gq - (ga + gb);
   V3<R> dc = d - c, cp = c - p, qd = q - d;
   V3<R> vt = gd_ + gq;
   vir6[0] = dc.x * vt.x + cp.x * gp.x - qd.x * gq.x;
   vir6[1] = dc.y * vt.x + cp.y * gp.x - qd.y * gq.x;
   vir6[2] = dc.z * vt.x + cp.z * gp.x - qd.z * gq.x;
   vir6[3] = dc.y * vt.y + cp.y * gp.y - qd.y * gq.y;
   vir6[4] = dc.z * vt.y + cp.z * gp.y - qd.z * gq.y;
   vir6[5] = dc.z * vt.z + cp.z * gp.z - qd.z * gq.z;
   return unit * kpit * ((R)1 - cos2);
}

// cubic Hermite basis on [0,1] {value at 0, value at 1, slope at 0, slope at 1} and its derivative
template <class R>
VM_HD void hermite(R s, R* h, R* dh)
{
   const R s2 = s * s, s3 = s2 * s;
   h[0] = (R)2 * s3 - (R)3 * s2 + (R)1, h[1] = (R)-2 * s3 + (R)3 * s2, h[2] = s3 - (R)2 * s2 + s, h[3] = s3 - s2;
   dh[0] = (R)6 * s2 - (R)6 * s, dh[1] = (R)-6 * s2 + (R)6 * s, dh[2] = (R)3 * s2 - (R)4 * s + (R)1, dh[3] = (R)3 * s2 - (R)2 * s;
}

struct TorTorGrid {      // one torsion-torsion table inside the flattened arrays (valparams.ValenceTerms)
   int nx, ny, off, xoff, yoff;
};

// ---- torsion-torsion, atoms {a,b,c,d,e} + optional chirality probe: include/seq/tortor.h:150-262.  The bicubic
//      patch is evaluated as a Hermite tensor product of the corner values / slopes / cross slopes (what the 16
//      bcucof coefficients expand to).  chk = position of the probe atom relative to X[0], has_chk = 0 if none.
template <class R, class P>
VM_HD R tortor(const V3<R>* X, bool has_chk, V3<R> chk, TorTorGrid T, const P* ttx, const P* tty, const P* tbf, const P* tbx, const P* tby,
   const P* tbxy, R unit, V3<R>* G)
{
   const R rad = (R)VM_RADIAN;
   R c1, s1, c2, s2;
   V3<R> d1[4], d2[4];
   for (int k = 0; k < 5; ++k)
      G[k] = mk<R>(0, 0, 0);
   if (!dihedral(X[0], X[1], X[2], X[3], c1, s1, d1) || !dihedral(X[1], X[2], X[3], X[4], c2, s2, d2))
      return 0;
   R v1 = rad * (R)atan2(s1, c1), v2 = rad * (R)atan2(s2, c2);
   R sign = 1;
   if (has_chk) {
      V3<R> ac = chk - X[2], bc = X[1] - X[2], dcv = X[3] - X[2];
      // the reference's determinant (lines 219-226): ac . ((c-b) x (d-c)) = ac . (bc x dc) -- same orientation
      const R vol = dot(ac, cross(bc, dcv));
      if (vol < 0)
         sign = -1;
   }
   v1 *= sign, v2 *= sign;
   if (v1 < (R)-180)
      v1 += (R)360;
   if (v1 >= (R)180)
      v1 -= (R)360;
   if (v2 < (R)-180)
      v2 += (R)360;
   if (v2 >= (R)180)
      v2 -= (R)360;
   int xlo = (int)floor((v1 + (R)180) * (R)(T.nx - 1) / (R)360);
   int ylo = (int)floor((v2 + (R)180) * (R)(T.ny - 1) / (R)360);
   xlo = xlo < 0 ? 0 : (xlo > T.nx - 2 ? T.nx - 2 : xlo);
   ylo = ylo < 0 ? 0 : (ylo > T.ny - 2 ? T.ny - 2 : ylo);
   const R x1l = (R)ttx[T.xoff + xlo], x1u = (R)ttx[T.xoff + xlo + 1];
   const R y1l = (R)tty[T.yoff + ylo], y1u = (R)tty[T.yoff + ylo + 1];
   const R dx = x1u - x1l, dy = y1u - y1l;
   const int pos1 = T.off + ylo * T.nx + xlo, pos2 = pos1 + T.nx;
   const int corner[4] = {pos1, pos1 + 1, pos2 + 1, pos2};      // (l,l) (u,l) (u,u) (l,u)
   R ht[4], dht[4], hu[4], dhu[4];
   hermite((v1 - x1l) / dx, ht, dht);
   hermite((v2 - y1l) / dy, hu, dhu);
   const int tv[4] = {0, 1, 1, 0}, ts[4] = {2, 3, 3, 2}, uv[4] = {0, 0, 1, 1}, us[4] = {2, 2, 3, 3};
   R e = 0, de1 = 0, de2 = 0;
   for (int k = 0; k < 4; ++k) {
      const R f = (R)tbf[corner[k]], fx = dx * (R)tbx[corner[k]], fy = dy * (R)tby[corner[k]], fxy = dx * dy * (R)tbxy[corner[k]];
      const R A = f * ht[tv[k]] + fx * ht[ts[k]], Ad = f * dht[tv[k]] + fx * dht[ts[k]];
      const R B = fy * ht[tv[k]] + fxy * ht[ts[k]], Bd = fy * dht[tv[k]] + fxy * dht[ts[k]];
      e += A * hu[uv[k]] + B * hu[us[k]];
      de1 += Ad * hu[uv[k]] + Bd * hu[us[k]];
      de2 += A * dhu[uv[k]] + B * dhu[us[k]];
   }
   const R g1 = sign * unit * rad * de1 / dx, g2 = sign * unit * rad * de2 / dy;      // dE/dphi1, dE/dphi2 (radians)
   G[0] = g1 * d1[0];
   G[1] = g1 * d1[1] + g2 * d2[0];
   G[2] = g1 * d1[2] + g2 * d2[1];
   G[3] = g1 * d1[3] + g2 * d2[2];
   G[4] = g2 * d2[3];
   return unit * e;
}

// internal virial of a translation-invariant interaction from relative positions: {xx, yx, zx, yy, zy, zz}
template <class R>
VM_HD void virial6(const V3<R>* X, const V3<R>* G, int m, R* v)
{
   for (int k = 0; k < 6; ++k)
      v[k] = 0;
   for (int k = 0; k < m; ++k) {
      v[0] += X[k].x * G[k].x, v[1] += X[k].y * G[k].x, v[2] += X[k].z * G[k].x;
      v[3] += X[k].y * G[k].y, v[4] += X[k].z * G[k].y, v[5] += X[k].z * G[k].z;
   }
}
}      // namespace vm
