// Per-pair AMOEBA real-space math, our own formulation:
//
//   U = sum_n G_n(moments, R) * B_{n-1}(r),     R = r_k - r_i,
//
// with the rotational invariants G_1..G_5 of two point multipoles (charge, dipole, traceless
// quadrupole/3 as Tinker stores it) and ONE radial hierarchy B_n, B_{n+1} = -(1/r) dB_n/dr.
// Ewald screening (erfc), Thole damping (lambda_3..lambda_9) and exclusion scaling only change
// the B_n passed in:   B_n = bn_n - (1 - s*lambda_n) * rr_n.
// Gradient and torques follow by differentiating G_n; they were checked term by term against
// the oracle's generic interaction-tensor contraction.  The same physics is spread over
// include/seq/pair_mpole.h:237-442, pair_polar.h:367-742, pair_field.h:124-420 and
// damp.h:8-151 in the reference; nothing below is transcribed from those files.
#pragma once
#include "apx_internal.h"

struct Mpole {
   real c, dx, dy, dz, qxx, qxy, qxz, qyy, qyz, qzz;
};

struct V3 {
   real x, y, z;
};
__device__ __forceinline__ V3 v3(real x, real y, real z)
{
   V3 r = {x, y, z};
   return r;
}
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(real s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ void operator+=(V3& a, V3 b)
{
   a.x += b.x;
   a.y += b.y;
   a.z += b.z;
}
__device__ __forceinline__ void operator-=(V3& a, V3 b)
{
   a.x -= b.x;
   a.y -= b.y;
   a.z -= b.z;
}
__device__ __forceinline__ real dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ V3 dip(const Mpole& m) { return v3(m.dx, m.dy, m.dz); }
// Q v for the symmetric quadrupole
__device__ __forceinline__ V3 qmul(const Mpole& m, V3 v)
{
   return v3(m.qxx * v.x + m.qxy * v.y + m.qxz * v.z, m.qxy * v.x + m.qyy * v.y + m.qyz * v.z, m.qxz * v.x + m.qyz * v.y + m.qzz * v.z);
}

#ifdef APX_DOUBLE
__device__ __forceinline__ real r_exp(real x) { return exp(x); }
__device__ __forceinline__ real r_rsqrt(real x) { return rsqrt(x); }
__device__ __forceinline__ real r_div(real a, real b) { return a / b; }
// erfc(x) given e = exp(-x*x)
__device__ __forceinline__ real r_erfc_e(real x, real) { return erfc(x); }
#else
// Mixed build: hardware exp2 / reciprocal approximations (the reference's CUDA build is compiled
// with --use_fast_math, src/cu/CMakeLists.txt:28-35, which makes the same substitutions).
__device__ __forceinline__ real r_exp(real x) { return __expf(x); }
__device__ __forceinline__ real r_rsqrt(real x) { return rsqrtf(x); }
__device__ __forceinline__ real r_div(real a, real b) { return __fdividef(a, b); }
// erfc(x) = exp(-x^2) * t * P9(t), t = 1/(1 + x/2): our own Chebyshev fit of erfcx(x)/t on
// 0 <= x <= 4.6 (a*cutoff is 3.8 for the 7 A Ewald cutoff); relative error 3e-9 in exact
// arithmetic, 3e-7 in float -- the same as erfcf's documented 4 ulp -- and < 3e-7 up to x = 6.
__device__ __forceinline__ real r_erfc_e(real x, real e)
{
   const real t = __fdividef(1.0f, fmaf(0.5f, x, 1.0f));
   real p = 3.5327550404e-02f;
   p = fmaf(p, t, -2.5197705174e-01f);
   p = fmaf(p, t, 7.4152748764e-01f);
   p = fmaf(p, t, -1.1026462441e+00f);
   p = fmaf(p, t, 7.9686896115e-01f);
   p = fmaf(p, t, -3.1024419607e-01f);
   p = fmaf(p, t, 3.0320297180e-01f);
   p = fmaf(p, t, 2.2082756977e-01f);
   p = fmaf(p, t, 2.8517960792e-01f);
   p = fmaf(p, t, 2.8193334191e-01f);
   return p * t * e;
}
#endif

// minimum image
__device__ __forceinline__ void apx_image(const Box& b, real& dx, real& dy, real& dz)
{
   if (b.orthogonal) {
      dx -= b.lx * rint(dx * b.ilx);
      dy -= b.ly * rint(dy * b.ily);
      dz -= b.lz * rint(dz * b.ilz);
   } else {
      real f1 = dx * b.r[0] + dy * b.r[1] + dz * b.r[2];
      real f2 = dx * b.r[3] + dy * b.r[4] + dz * b.r[5];
      real f3 = dx * b.r[6] + dy * b.r[7] + dz * b.r[8];
      f1 -= rint(f1);
      f2 -= rint(f2);
      f3 -= rint(f3);
      dx = f1 * b.l[0] + f2 * b.l[1] + f3 * b.l[2];
      dy = f1 * b.l[3] + f2 * b.l[4] + f3 * b.l[5];
      dz = f1 * b.l[6] + f2 * b.l[7] + f3 * b.l[8];
   }
}

// ---- radial hierarchies --------------------------------------------------------------------
// rr[n] = (2n-1)!! / r^(2n+1)
template <int N>
__device__ __forceinline__ void radial_coulomb(real rinv, real rr2, real* rr)
{
   rr[0] = rinv;
   #pragma unroll
   for (int j = 1; j < N; ++j)
      rr[j] = (real)(2 * j - 1) * rr[j - 1] * rr2;
}

// Ewald real-space: bn[0] = erfc(a r)/r, upward recursion
template <int N>
__device__ __forceinline__ void radial_ewald(real r, real rinv, real rr2, real aewald, real* bn)
{
   real ra = aewald * r;
   real ex = r_exp(-ra * ra);
   bn[0] = r_erfc_e(ra, ex) * rinv;
   real a2 = 2 * aewald * aewald;
   real pref = (real)0.5641895835477563 / aewald;   // 1/(sqrt(pi) a)
   #pragma unroll
   for (int j = 1; j < N; ++j) {
      pref *= a2;
      bn[j] = ((real)(2 * j - 1) * bn[j - 1] + pref * ex) * rr2;
   }
}

// Thole: returns (1 - lambda_{2n+1}) for n = 1..N-1 in om[1..N-1] (om[0] unused = 0).
// 1 - lambda_3 = e^-x, 1 - lambda_5 = (1+x) e^-x, 1 - lambda_7 = (1 + x + 0.6 x^2) e^-x,
// 1 - lambda_9 = (1 + x + 18/35 x^2 + 9/35 x^3) e^-x,  x = gamma (r / (pd_i pd_k))^3
template <int N>
__device__ __forceinline__ void thole_one_minus_lambda(real r, real pdi, real pdk, real pgamma, real* om)
{
   real dmp = pdi * pdk;
   real ex = 0, x = 0;
   if (dmp != 0) {
      real q = r_div(r, dmp);
      x = pgamma * q * q * q;
      ex = r_exp(-x);
   }
   om[0] = 0;
   if (N > 1) om[1] = ex;
   if (N > 2) om[2] = ex * (1 + x);
   if (N > 3) om[3] = ex * (1 + x + (real)0.6 * x * x);
   if (N > 4) om[4] = ex * (1 + x * (1 + x * ((real)(18.0 / 35.0) + (real)(9.0 / 35.0) * x)));
   if (N > 5) om[5] = 0;
}

// ---- fields ---------------------------------------------------------------------------------
// field of a multipole: E = R*(sgn*c*B1 + d.R B2 + sgn*R.Q.R B3) - B1 d - sgn*2 B2 Q.R
// sgn = -1: source at k, field at i ; sgn = +1: source at i, field at k   (R = r_k - r_i)
__device__ __forceinline__ V3 mpole_field(V3 R, const Mpole& s, real B1, real B2, real B3, real sgn)
{
   V3 d = dip(s);
   V3 q = qmul(s, R);
   real dr = dot3(d, R), qr = dot3(q, R);
   real a = sgn * s.c * B1 + dr * B2 + sgn * qr * B3;
   return a * R - B1 * d - (sgn * 2 * B2) * q;
}

// field of a dipole u at the other site (even in R)
__device__ __forceinline__ V3 dipole_field(V3 R, V3 u, real B1, real B2)
{
   return (B2 * dot3(R, u)) * R - B1 * u;
}

// ---- energy / gradient / torque -------------------------------------------------------------
// full multipole - full multipole.  g = dU/dr_k (= -dU/dr_i), ti/tk = torques.
template <bool DO_G>
__device__ __forceinline__ real pair_mm(V3 R, const Mpole& I, const Mpole& K, const real* B, V3& g, V3& ti, V3& tk)
{
   V3 di = dip(I), dk = dip(K);
   V3 qi = qmul(I, R), qk = qmul(K, R);
   real dir = dot3(di, R), dkr = dot3(dk, R), qir = dot3(qi, R), qkr = dot3(qk, R);
   real dik = dot3(di, dk), qik = dot3(qi, qk), diqk = dot3(di, qk), dkqi = dot3(dk, qi);
   real qiqk = 2 * (I.qxy * K.qxy + I.qxz * K.qxz + I.qyz * K.qyz) + I.qxx * K.qxx + I.qyy * K.qyy + I.qzz * K.qzz;
   real G1 = I.c * K.c;
   real G2 = K.c * dir - I.c * dkr + dik;
   real G3 = I.c * qkr + K.c * qir - dir * dkr + 2 * (dkqi - diqk + qiqk);
   real G4 = dir * qkr - dkr * qir - 4 * qik;
   real G5 = qir * qkr;
   real U = G1 * B[0] + G2 * B[1] + G3 * B[2] + G4 * B[3] + G5 * B[4];
   if (DO_G) {
      V3 Qidk = qmul(I, dk), Qkdi = qmul(K, di), Qiqk = qmul(I, qk), Qkqi = qmul(K, qi);
      real radial = G1 * B[1] + G2 * B[2] + G3 * B[3] + G4 * B[4] + G5 * B[5];
      // sum_n B_{n-1} dG_n/dR, grouped by vector
      real cdi = B[1] * K.c - B[2] * dkr + B[3] * qkr;      // coefficient of d_i
      real cdk = -B[1] * I.c - B[2] * dir - B[3] * qir;     // coefficient of d_k
      real cqi = 2 * (B[2] * K.c - B[3] * dkr + B[4] * qkr); // coefficient of Q_i R
      real cqk = 2 * (B[2] * I.c + B[3] * dir + B[4] * qir); // coefficient of Q_k R
      g = cdi * di + cdk * dk + cqi * qi + cqk * qk + (2 * B[2]) * (Qidk - Qkdi) - (4 * B[3]) * (Qiqk + Qkqi) - radial * R;
      // antisymmetric part of Q_i Q_k (vector dual)
      V3 dqq = v3(I.qxy * K.qxz + I.qyy * K.qyz + I.qyz * K.qzz - I.qxz * K.qxy - I.qyz * K.qyy - I.qzz * K.qyz,
         I.qxz * K.qxx + I.qyz * K.qxy + I.qzz * K.qxz - I.qxx * K.qxz - I.qxy * K.qyz - I.qxz * K.qzz,
         I.qxx * K.qxy + I.qxy * K.qyy + I.qxz * K.qyz - I.qxy * K.qxx - I.qyy * K.qxy - I.qyz * K.qxz);
      // site i
      V3 dUdi = B[1] * (K.c * R + dk) - B[2] * (dkr * R + 2 * qk) + (B[3] * qkr) * R;
      V3 ti_q = (B[2] * K.c - B[3] * dkr + B[4] * qkr) * cross3(qi, R) + B[2] * (cross3(Qidk, R) + cross3(qi, dk))
         - (2 * B[3]) * (cross3(Qiqk, R) + cross3(qi, qk)) + (2 * B[2]) * dqq;
      ti = cross3(dUdi, di) - 2 * ti_q;
      // site k
      V3 dUdk = B[1] * (di - I.c * R) + B[2] * (2 * qi - dir * R) - (B[3] * qir) * R;
      V3 tk_q = (B[2] * I.c + B[3] * dir + B[4] * qir) * cross3(qk, R) - B[2] * (cross3(Qkdi, R) + cross3(qk, di))
         - (2 * B[3]) * (cross3(Qkqi, R) + cross3(qk, qi)) - (2 * B[2]) * dqq;
      tk = cross3(dUdk, dk) - 2 * tk_q;
   }
   return U;
}

// multipole at i  x  bare dipole u at k.  g = dU/dr_k, ti = torque on i.
template <bool DO_G>
__device__ __forceinline__ real pair_mu(V3 R, const Mpole& I, V3 u, const real* B, V3& g, V3& ti)
{
   V3 di = dip(I);
   V3 qi = qmul(I, R);
   real dir = dot3(di, R), qir = dot3(qi, R), ukr = dot3(u, R);
   real G2 = dot3(di, u) - I.c * ukr;
   real G3 = 2 * dot3(u, qi) - dir * ukr;
   real G4 = -ukr * qir;
   real U = G2 * B[1] + G3 * B[2] + G4 * B[3];
   if (DO_G) {
      V3 Qiu = qmul(I, u);
      real radial = G2 * B[2] + G3 * B[3] + G4 * B[4];
      g = (-B[1] * I.c - B[2] * dir - B[3] * qir) * u - (B[2] * ukr) * di + (2 * B[2]) * Qiu - (2 * B[3] * ukr) * qi - radial * R;
      V3 dUdi = B[1] * u - (B[2] * ukr) * R;
      V3 tq = B[2] * (cross3(Qiu, R) + cross3(qi, u)) - (B[3] * ukr) * cross3(qi, R);
      ti = cross3(dUdi, di) - 2 * tq;
   }
   return U;
}

// bare dipole u at i  x  multipole at k.  g = dU/dr_k, tk = torque on k.
template <bool DO_G>
__device__ __forceinline__ real pair_um(V3 R, V3 u, const Mpole& K, const real* B, V3& g, V3& tk)
{
   V3 dk = dip(K);
   V3 qk = qmul(K, R);
   real dkr = dot3(dk, R), qkr = dot3(qk, R), uir = dot3(u, R);
   real G2 = K.c * uir + dot3(u, dk);
   real G3 = -uir * dkr - 2 * dot3(u, qk);
   real G4 = uir * qkr;
   real U = G2 * B[1] + G3 * B[2] + G4 * B[3];
   if (DO_G) {
      V3 Qku = qmul(K, u);
      real radial = G2 * B[2] + G3 * B[3] + G4 * B[4];
      g = (B[1] * K.c - B[2] * dkr + B[3] * qkr) * u - (B[2] * uir) * dk - (2 * B[2]) * Qku + (2 * B[3] * uir) * qk - radial * R;
      V3 dUdk = B[1] * u - (B[2] * uir) * R;
      V3 tq = (B[3] * uir) * cross3(qk, R) - B[2] * (cross3(Qku, R) + cross3(qk, u));
      tk = cross3(dUdk, dk) - 2 * tq;
   }
   return U;
}

// dipole - dipole gradient dU/dr_k
__device__ __forceinline__ V3 pair_uu_grad(V3 R, V3 a, V3 b, const real* B)
{
   real ar = dot3(a, R), br = dot3(b, R);
   real G2 = dot3(a, b), G3 = -ar * br;
   return (-B[2] * br) * a - (B[2] * ar) * b - (G2 * B[2] + G3 * B[3]) * R;
}

// ---- atomics --------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_fixed(fixed_t* p, real v)
{
   atomicAdd(p, (fixed_t)(long long)((double)v * APX_FIXED_SCALE));
}
__device__ __forceinline__ void atomic_real3(real* base, int s, V3 v)
{
   atomicAdd(base + 3 * s, v.x);
   atomicAdd(base + 3 * s + 1, v.y);
   atomicAdd(base + 3 * s + 2, v.z);
}
