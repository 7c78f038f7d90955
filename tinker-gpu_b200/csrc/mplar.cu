// Energy / gradient / torque / virial of the permanent-multipole and polarization terms:
// the fused real-space tile kernel (role of emplar_cu1a/b/c, src/cu/amoeba/emplar.cu:505-544, and
// of empole_cu1 + epolar_cu1 on the ANALYZE path), Ewald self terms (include/seq/emselfamoeba.h),
// reciprocal-space energy/force/torque/virial (src/cu/hippo/empole.cu:202-336,
// src/cu/epolarrecip.cu:9-511), the dot-product polarization energy (src/cu/amoeba/epolar.cu:15-30)
// and the final fixed-point reductions of energy() (src/energy.cpp:319-448).
//
// Differences of organisation from the reference:
//   * one real-space pass computes permanent + polarization terms with all exclusion scales = 1;
//     a second, tiny pass corrects listed pairs with (scale-1) undamped/Thole-only terms;
//   * the permanent-multipole PME round trip is done ONCE per evaluation (inside induce's dfield)
//     and its fmp/fphi are reused by the reciprocal multipole and polarization terms
//     (the reference repeats it in empoleEwaldRecip, hippo/empole.cu:308-317);
//   * energies/virials are reduced warp->atomic into a handful of fixed-point / double scalars
//     instead of 524288-entry buffers that must be re-reduced every call (32 MB virial read).
#include "apx_internal.h"
#include "pairmath.cuh"
#include "rows.cuh"
#include <cmath>
#include <algorithm>

#define FULL 0xffffffffu
#define SHF(v, src) __shfl_sync(FULL, (v), (src))

// dbuf slots (doubles)
enum {
   D_EM_RECIP = 0, D_EP_RECIP = 1, D_EM_SELF = 2, D_EP_DOT = 3, D_EP_SELF = 4,
   D_VIR_TRQ = 8,      // 6
   D_CONV_E = 16,      // 1 + 6 (vir_m)
   D_VIR_MREC = 24,    // 6: atom part of the multipole recip virial
   D_VIR_PREC = 32,    // 6: polarization recip virial (atom part)
   D_VIR_CROSS = 40,   // 6
   D_TOTAL = 48
};
// ebuf slots (fixed point): 0 em_real, 1 ep_real, 2..7 virial of real-space pairs
// cnt: 0 nem, 1 nep

void apx_to_sorted(apx_ctx* c, const double* in_dev, real* out);
void apx_from_sorted(apx_ctx* c, const real* in, double* out_dev);

namespace {
__device__ __forceinline__ int as_int(real w)
{
#ifdef APX_DOUBLE
   return (int)__double_as_longlong(w);
#else
   return __float_as_int(w);
#endif
}
__device__ __forceinline__ Mpole load_mpole(const real4* mp0, const real4* mp1, const real2* mp2, int s)
{
   real4 a = mp0[s], b = mp1[s];
   real2 c = mp2[s];
   Mpole m;
   m.c = a.x, m.dx = a.y, m.dy = a.z, m.dz = a.w;
   m.qxx = b.x, m.qxy = b.y, m.qxz = b.z, m.qyy = b.w, m.qyz = c.x, m.qzz = c.y;
   return m;
}
__device__ __forceinline__ void atomic_fixed3(fixed_t* gx, fixed_t* gy, fixed_t* gz, int s, V3 v)
{
   atomic_fixed(gx + s, v.x);
   atomic_fixed(gy + s, v.y);
   atomic_fixed(gz + s, v.z);
}
__device__ __forceinline__ double warp_sum(double x)
{
   for (int o = 16; o > 0; o >>= 1)
      x += __shfl_xor_sync(FULL, x, o);
   return x;
}
__device__ __forceinline__ void atomic_fixed_d(fixed_t* p, double v)
{
   atomicAdd(p, (fixed_t)(long long)(v * APX_FIXED_SCALE));
}

struct MplarArgs {
   int a0, a1;           // owned sorted range
   Box box;
   real cut2, aewald, f;
   const int* vstart;     // directed neighbor rows (rows.cu)
   const int* rcnt;
   const int* nbr;
   const pos_t* posq;
   const real4* tpj;
   const real* thlval;
   int nj, table;
   const real4* mp0;
   const real4* mp1;
   const real2* mp2;
   const real* ud;
   const real* up;
   int do_m, do_p, mutual, do_e, do_v, do_a, pair_ep, ewald;
   fixed_t* gx;
   fixed_t* gy;
   fixed_t* gz;
   fixed_t* trq;      // [3*npad] fixed point
   fixed_t* ebuf;
   int* cnt;          // [0] unordered pairs counted by the exclusion pass (signed correction)
   int* cnt2;         // [0] DIRECTED pairs visited by the row pass (= 2 x pairs)
};

// One lane group per atom i, one directed pair (i,k) per lane per step.  Each unordered pair is
// visited from both ends: a visit adds half the pair energy / virial and the complete gradient and
// torque of ITS OWN atom i (the k-side torque expressions are dead code here and are dropped by
// the compiler), so nothing is scattered and no atomics are needed for the per-atom sums.
template <bool DO_G, bool EWALD, int G>
__global__ void __launch_bounds__(ROWS_BLOCK) k_mplar_rows(MplarArgs A)
{
   double em = 0, ep = 0, vxx = 0, vxy = 0, vxz = 0, vyy = 0, vyz = 0, vzz = 0;
   int nem = 0;
   ROWS_FOREACH_ATOM(G, A.a0, A.a1, i, l, act)
   {
      const pos_t pi = A.posq[i];
      const real4 qi = A.tpj[i];
      const Mpole mi = load_mpole(A.mp0, A.mp1, A.mp2, i);
      V3 udi = v3(0, 0, 0), upi = v3(0, 0, 0);
      if (A.do_p) {
         udi = v3(A.ud[3 * i], A.ud[3 * i + 1], A.ud[3 * i + 2]);
         upi = v3(A.up[3 * i], A.up[3 * i + 1], A.up[3 * i + 2]);
      }
      const int beg = A.vstart[i];
      const int len = act ? A.rcnt[i] : 0;
      V3 gi = v3(0, 0, 0), ti = v3(0, 0, 0);
      real emr = 0, epr = 0, v0 = 0, v1 = 0, v2 = 0, v3_ = 0, v4 = 0, v5 = 0;
      for (int q = l; q < len; q += G) {
         const int k = A.nbr[beg + q];
         if (k < 0)      // ROW_LISTED_FLAG: the pair belongs to k_mplar_listed
            continue;
         const pos_t pk = A.posq[k];
         const real4 qk = A.tpj[k];
         const Mpole mk = load_mpole(A.mp0, A.mp1, A.mp2, k);
         real dx, dy, dz;
         pair_delta(A.box, pi, pk, dx, dy, dz);
         const real r2 = dx * dx + dy * dy + dz * dz;
         const real rinv = r_rsqrt(r2);
         const real r = r2 * rinv, rr2 = rinv * rinv;
         real rr[6], B[6];
         radial_coulomb<6>(rinv, rr2, rr);
         if (EWALD)
            radial_ewald<6>(r, rinv, rr2, A.aewald, B);
         else {
            #pragma unroll
            for (int j = 0; j < 6; ++j)
               B[j] = rr[j];
         }
         const V3 R = v3(dx, dy, dz);
         V3 g = v3(0, 0, 0), tqi = v3(0, 0, 0);
         ++nem;
         if (A.do_m) {
            V3 g1, t1, t2;
            real U = pair_mm<DO_G>(R, mi, mk, B, g1, t1, t2);
            emr += U;
            if (DO_G) {
               g += g1;
               tqi += t1;
            }
         }
         if (A.do_p) {
            const V3 ukd = v3(A.ud[3 * k], A.ud[3 * k + 1], A.ud[3 * k + 2]);
            const V3 ukp = v3(A.up[3 * k], A.up[3 * k + 1], A.up[3 * k + 2]);
            real om[6];
            const real pg = A.table ? A.thlval[as_int(qi.w) * A.nj + as_int(qk.w)] : min(qi.x, qk.x);
            thole_one_minus_lambda<6>(r, pos_w(pi), pos_w(pk), pg, om);
            #pragma unroll
            for (int j = 1; j < 5; ++j)
               B[j] -= om[j] * rr[j];
            if (A.pair_ep) {
               V3 d0, d1;
               epr += pair_mu<false>(R, mi, ukd, B, d0, d1) + pair_um<false>(R, udi, mk, B, d0, d1);
            }
            if (DO_G) {
               const V3 ubk = (real)0.5 * (ukd + ukp), ubi = (real)0.5 * (udi + upi);
               V3 g1, g2, t1, t2;
               pair_mu<true>(R, mi, ubk, B, g1, t1);
               pair_um<true>(R, ubi, mk, B, g2, t2);
               g += g1 + g2;
               tqi += t1;
               if (A.mutual)
                  g += (real)0.5 * (pair_uu_grad(R, udi, ukp, B) + pair_uu_grad(R, upi, ukd, B));
            }
         }
         if (DO_G) {
            gi -= g;
            ti += tqi;
            if (A.do_v) {
               v0 += R.x * g.x;
               v1 += R.y * g.x + R.x * g.y;
               v2 += R.z * g.x + R.x * g.z;
               v3_ += R.y * g.y;
               v4 += R.z * g.y + R.y * g.z;
               v5 += R.z * g.z;
            }
         }
      }
      if (DO_G) {
         gi = group_sum3<G>(gi);
         ti = group_sum3<G>(ti);
         if (l == 0 && act) {
            atomic_fixed3(A.gx, A.gy, A.gz, i, A.f * gi);
            atomic_fixed3(A.trq, A.trq + 1, A.trq + 2, 3 * i, A.f * ti);
         }
         if (A.do_v) {
            vxx += (double)v0, vxy += (double)v1, vxz += (double)v2;
            vyy += (double)v3_, vyz += (double)v4, vzz += (double)v5;
         }
      }
      em += (double)emr;
      ep += (double)epr;
   }
   // every pair was visited twice
   const double hf = 0.5 * (double)A.f;
   if (A.do_e) {
      em = warp_sum(em);
      ep = warp_sum(ep);
      if ((threadIdx.x & 31) == 0) {
         if (em != 0.0) atomic_fixed_d(&A.ebuf[0], hf * em);
         if (ep != 0.0) atomic_fixed_d(&A.ebuf[1], 0.5 * hf * ep);
      }
   }
   if (A.do_a) {
      for (int o = 16; o > 0; o >>= 1)
         nem += __shfl_xor_sync(FULL, nem, o);
      if ((threadIdx.x & 31) == 0 && nem)
         atomicAdd(&A.cnt2[0], nem);
   }
   if (DO_G && A.do_v) {
      vxx = warp_sum(vxx), vxy = warp_sum(vxy), vxz = warp_sum(vxz);
      vyy = warp_sum(vyy), vyz = warp_sum(vyz), vzz = warp_sum(vzz);
      if ((threadIdx.x & 31) == 0) {
         atomic_fixed_d(&A.ebuf[2], hf * vxx);
         atomic_fixed_d(&A.ebuf[3], 0.5 * hf * vxy);
         atomic_fixed_d(&A.ebuf[4], 0.5 * hf * vxz);
         atomic_fixed_d(&A.ebuf[5], hf * vyy);
         atomic_fixed_d(&A.ebuf[6], 0.5 * hf * vyz);
         atomic_fixed_d(&A.ebuf[7], hf * vzz);
      }
   }
}

// Listed pairs (the exclusion / scaling table mdpuexclude: bonded-range and same-polarization-group pairs), evaluated ONCE
// with their true scale factors and in DOUBLE in both builds:  B_n = bn_n - (1 - s lambda_n) rr_n.
// The row pass skips these pairs (ROW_LISTED_FLAG).  Why not "all scales 1 in the row pass + (s-1) correction" as the field
// kernels do: for a bonded pair at 1 A both parts are ~100 kcal/mol/A and cancel to a few; in float that cancellation alone
// costs 1e-5 kcal/mol/A RMS of force error, the whole north-star tolerance (tests/test_pairmath_host.py measures it on the
// CPU with this very header).  There are ~8 listed pairs per atom against ~140 unlisted ones, so double costs nothing here.
// Separation from the caller-order f64 coordinates through perm[] with a double-precision minimum image.
struct ListedD {
   double l[9], r[9];     // cell vectors / reciprocal vectors (rows)
   const double* xyz;     // [n][3] caller order
   const int* perm;       // sorted slot -> caller index
   const double* sc;      // [nx][4] m, d, p, u of listed pair e (same order as the PairExcl records)
   double cut2, aewald, f;
};

__device__ __forceinline__ pm64::Mpole load_mpole_d(const real4* mp0, const real4* mp1, const real2* mp2, int s)
{
   real4 a = mp0[s], b = mp1[s];
   real2 c = mp2[s];
   pm64::Mpole m;
   m.c = a.x, m.dx = a.y, m.dy = a.z, m.dz = a.w;
   m.qxx = b.x, m.qxy = b.y, m.qxz = b.z, m.qyy = b.w, m.qyz = c.x, m.qzz = c.y;
   return m;
}
__device__ __forceinline__ pm64::V3 load3_d(const real* v, int s) { return pm64::v3((double)v[3 * s], (double)v[3 * s + 1], (double)v[3 * s + 2]); }
__device__ __forceinline__ void atomic_fixed3_d(fixed_t* gx, fixed_t* gy, fixed_t* gz, int s, pm64::V3 v)
{
   atomic_fixed_d(gx + s, v.x);
   atomic_fixed_d(gy + s, v.y);
   atomic_fixed_d(gz + s, v.z);
}

template <bool DO_G>
__global__ void __launch_bounds__(128) k_mplar_listed(int nx, const PairExcl* __restrict__ ex, MplarArgs A, ListedD D)
{
   typedef pm64::V3 W3;
   int e = blockIdx.x * blockDim.x + threadIdx.x;
   double em = 0, ep = 0, v[6] = {0, 0, 0, 0, 0, 0};
   int dn = 0;
   if (e < nx) {
      const PairExcl p = ex[e];
      // several GPUs: a pair is handled by the owner(s) of its atoms, each writing only to its own
      // atom; the owner of p.i also books the pair's energy, virial and count
      const bool own_i = p.i >= A.a0 && p.i < A.a1, own_k = p.k >= A.a0 && p.k < A.a1;
      double Rx = 0, Ry = 0, Rz = 0, r2 = 1e300;
      if (own_i || own_k) {
         const int ci = D.perm[p.i], ck = D.perm[p.k];
         const double d0 = D.xyz[3 * ck] - D.xyz[3 * ci], d1 = D.xyz[3 * ck + 1] - D.xyz[3 * ci + 1], d2 = D.xyz[3 * ck + 2] - D.xyz[3 * ci + 2];
         double f1 = d0 * D.r[0] + d1 * D.r[1] + d2 * D.r[2];
         double f2 = d0 * D.r[3] + d1 * D.r[4] + d2 * D.r[5];
         double f3 = d0 * D.r[6] + d1 * D.r[7] + d2 * D.r[8];
         f1 -= rint(f1), f2 -= rint(f2), f3 -= rint(f3);
         Rx = f1 * D.l[0] + f2 * D.l[1] + f3 * D.l[2];
         Ry = f1 * D.l[3] + f2 * D.l[4] + f3 * D.l[5];
         Rz = f1 * D.l[6] + f2 * D.l[7] + f3 * D.l[8];
         r2 = Rx * Rx + Ry * Ry + Rz * Rz;
      }
      if (r2 <= D.cut2) {
         const double sm = D.sc[4 * e], sd = D.sc[4 * e + 1], sp = D.sc[4 * e + 2], su = D.sc[4 * e + 3];
         const double rinv = pm64::r_rsqrt(r2), r = r2 * rinv, rr2 = rinv * rinv;
         double rr[6], bn[6], B[6];
         pm64::radial_coulomb<6>(rinv, rr2, rr);
         if (A.ewald)
            pm64::radial_ewald<6>(r, rinv, rr2, D.aewald, bn);
         else {
            #pragma unroll
            for (int q = 0; q < 6; ++q)
               bn[q] = rr[q];
         }
         const W3 R = pm64::v3(Rx, Ry, Rz);
         const pm64::Mpole mi = load_mpole_d(A.mp0, A.mp1, A.mp2, p.i), mk = load_mpole_d(A.mp0, A.mp1, A.mp2, p.k);
         W3 g = pm64::v3(0, 0, 0), tqi = pm64::v3(0, 0, 0), tqk = pm64::v3(0, 0, 0);
         if (own_i && (A.do_m ? sm != 0.0 : sp != 0.0))
            dn = 1;
         if (A.do_m) {
            #pragma unroll
            for (int q = 0; q < 6; ++q)
               B[q] = bn[q] - (1.0 - sm) * rr[q];
            W3 g1, t1, t2;
            const double U = pm64::pair_mm<DO_G>(R, mi, mk, B, g1, t1, t2);
            if (own_i)
               em = D.f * U;
            if (DO_G) {
               g += g1;
               tqi += t1;
               tqk += t2;
            }
         }
         if (A.do_p) {
            double om[6], lam[6];
            const real4 qi = A.tpj[p.i], qk = A.tpj[p.k];
            const double pg = A.table ? (double)A.thlval[as_int(qi.w) * A.nj + as_int(qk.w)] : (double)min(qi.x, qk.x);
            pm64::thole_one_minus_lambda<6>(r, (double)pos_w(A.posq[p.i]), (double)pos_w(A.posq[p.k]), pg, om);
            #pragma unroll
            for (int q = 0; q < 6; ++q)
               lam[q] = (1.0 - om[q]) * rr[q];
            const W3 udi = load3_d(A.ud, p.i), upi = load3_d(A.up, p.i), udk = load3_d(A.ud, p.k), upk = load3_d(A.up, p.k);
            // p-scaled: permanent with ud ; d-scaled: permanent with up
            for (int pass = 0; pass < 2; ++pass) {
               const double sc = pass == 0 ? sp : sd;
               #pragma unroll
               for (int q = 0; q < 6; ++q)
                  B[q] = 0.5 * (bn[q] - rr[q] + sc * lam[q]);
               const W3 uk = pass == 0 ? udk : upk, ui = pass == 0 ? udi : upi;
               W3 g1, g2, t1, t2;
               const double U = pm64::pair_mu<DO_G>(R, mi, uk, B, g1, t1) + pm64::pair_um<DO_G>(R, ui, mk, B, g2, t2);
               if (pass == 0 && A.pair_ep && own_i)
                  ep = D.f * U;
               if (DO_G) {
                  g += g1 + g2;
                  tqi += t1;
                  tqk += t2;
               }
            }
            if (DO_G && A.mutual) {
               #pragma unroll
               for (int q = 0; q < 6; ++q)
                  B[q] = 0.5 * (bn[q] - rr[q] + su * lam[q]);
               g += pm64::pair_uu_grad(R, udi, upk, B) + pm64::pair_uu_grad(R, upi, udk, B);
            }
         }
         if (DO_G) {
            g = D.f * g;
            if (own_i) {
               atomic_fixed3_d(A.gx, A.gy, A.gz, p.i, -1.0 * g);
               atomic_fixed3_d(A.trq, A.trq + 1, A.trq + 2, 3 * p.i, D.f * tqi);
            }
            if (own_k) {
               atomic_fixed3_d(A.gx, A.gy, A.gz, p.k, g);
               atomic_fixed3_d(A.trq, A.trq + 1, A.trq + 2, 3 * p.k, D.f * tqk);
            }
            if (A.do_v && own_i) {
               v[0] = R.x * g.x;
               v[1] = 0.5 * (R.y * g.x + R.x * g.y);
               v[2] = 0.5 * (R.z * g.x + R.x * g.z);
               v[3] = R.y * g.y;
               v[4] = 0.5 * (R.z * g.y + R.y * g.z);
               v[5] = R.z * g.z;
            }
         }
      }
   }
   int lane = threadIdx.x & 31;
   if (A.do_e) {
      em = warp_sum(em);
      ep = warp_sum(ep);
      if (lane == 0) {
         if (em != 0.0) atomic_fixed_d(&A.ebuf[0], em);
         if (ep != 0.0) atomic_fixed_d(&A.ebuf[1], ep);
      }
   }
   if (A.do_a) {
      for (int o = 16; o > 0; o >>= 1)
         dn += __shfl_xor_sync(FULL, dn, o);
      if (lane == 0 && dn)
         atomicAdd(&A.cnt[0], dn);
   }
   if (DO_G && A.do_v) {
      #pragma unroll
      for (int q = 0; q < 6; ++q) {
         double x = warp_sum(v[q]);
         if (lane == 0 && x != 0.0)
            atomic_fixed_d(&A.ebuf[2 + q], x);
      }
   }
}

struct RecipX {
   real a[3][3];
   real ftc[6][6];
   real rc[9];        // recip rows
   int nf[3];
};

__device__ __forceinline__ void frac_to_cart10(const RecipX& X, const real* f, real* cphi)
{
   cphi[0] = f[0];
   #pragma unroll
   for (int c = 0; c < 3; ++c)
      cphi[1 + c] = X.a[c][0] * f[1] + X.a[c][1] * f[2] + X.a[c][2] * f[3];
   #pragma unroll
   for (int j = 0; j < 6; ++j) {
      real t = 0;
      #pragma unroll
      for (int k = 0; k < 6; ++k)
         t += X.ftc[k][j] * f[4 + k];
      cphi[4 + j] = t;
   }
}

// torque and virial of a Cartesian multipole (c,d,Q) in a potential with Cartesian derivatives
// cphi = {phi, grad(3), xx,yy,zz,xy,xz,yz}:  tau = -d x grad - 2 dual(Q Phi),
// V = -sym(d (x) grad) - 2 sym(Q Phi)
__device__ __forceinline__ void mpole_in_potential(const Mpole& m, const real* cp, V3& tau, double* v, real scale, bool do_v)
{
   real P[3][3] = {{cp[4], cp[7], cp[8]}, {cp[7], cp[5], cp[9]}, {cp[8], cp[9], cp[6]}};
   real Q[3][3] = {{m.qxx, m.qxy, m.qxz}, {m.qxy, m.qyy, m.qyz}, {m.qxz, m.qyz, m.qzz}};
   real M[3][3];
   #pragma unroll
   for (int a = 0; a < 3; ++a)
      #pragma unroll
      for (int b = 0; b < 3; ++b)
         M[a][b] = Q[a][0] * P[0][b] + Q[a][1] * P[1][b] + Q[a][2] * P[2][b];
   V3 d = v3(m.dx, m.dy, m.dz), g = v3(cp[1], cp[2], cp[3]);
   V3 dxg = cross3(d, g);
   tau = v3(-dxg.x - 2 * (M[1][2] - M[2][1]), -dxg.y - 2 * (M[2][0] - M[0][2]), -dxg.z - 2 * (M[0][1] - M[1][0]));
   if (do_v) {
      v[0] += (double)(scale * (-d.x * g.x - 2 * M[0][0]));
      v[1] += (double)(scale * (-(real)0.5 * (d.x * g.y + d.y * g.x) - (M[0][1] + M[1][0])));
      v[2] += (double)(scale * (-(real)0.5 * (d.x * g.z + d.z * g.x) - (M[0][2] + M[2][0])));
      v[3] += (double)(scale * (-d.y * g.y - 2 * M[1][1]));
      v[4] += (double)(scale * (-(real)0.5 * (d.y * g.z + d.z * g.y) - (M[1][2] + M[2][1])));
      v[5] += (double)(scale * (-d.z * g.z - 2 * M[2][2]));
   }
}

// index tables of the derivative of phi_k w.r.t. fractional coordinate 1,2,3 (tinker deriv1/2/3, 0-based)
__constant__ int c_d1[10] = {1, 4, 7, 8, 10, 15, 17, 13, 14, 19};
__constant__ int c_d2[10] = {2, 7, 5, 9, 13, 11, 18, 15, 19, 16};
__constant__ int c_d3[10] = {3, 8, 9, 6, 14, 16, 12, 19, 17, 18};

__device__ __forceinline__ void block_add(double* vals, int nv, double* out)
{
   // warp reduce each then one atomic per warp
   int lane = threadIdx.x & 31;
   for (int q = 0; q < nv; ++q) {
      double x = warp_sum(vals[q]);
      if (lane == 0 && x != 0.0)
         atomicAdd(&out[q], x);
   }
}

// reciprocal + self multipole terms per atom
template <bool DO_G>
__global__ void k_recip_mpole(int n, RecipX X, real f, real aewald, int do_e, int do_v, const real4* __restrict__ mp0,
   const real4* __restrict__ mp1, const real2* __restrict__ mp2, const real* __restrict__ fmp, const real* __restrict__ fphi,
   fixed_t* gx, fixed_t* gy, fixed_t* gz, fixed_t* trq, double* __restrict__ dbuf)
{
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   double acc[2] = {0, 0};
   double v[6] = {0, 0, 0, 0, 0, 0};
   if (s < n) {
      Mpole m = load_mpole(mp0, mp1, mp2, s);
      const real* fm = fmp + 10 * s;
      const real* fp = fphi + 20 * s;
      if (do_e) {
         real e = 0;
         #pragma unroll
         for (int k = 0; k < 10; ++k)
            e += fm[k] * fp[k];
         acc[0] = (double)((real)0.5 * f * e);
         real a2 = 2 * aewald * aewald;
         real cii = m.c * m.c, dii = m.dx * m.dx + m.dy * m.dy + m.dz * m.dz;
         real qii = 2 * (m.qxy * m.qxy + m.qxz * m.qxz + m.qyz * m.qyz) + m.qxx * m.qxx + m.qyy * m.qyy + m.qzz * m.qzz;
         real fterm = -f * aewald * (real)0.5641895835477563;
         acc[1] = (double)(fterm * (cii + a2 * (dii / 3 + 2 * a2 * qii / 5)));
      }
      if (DO_G) {
         real f1 = 0, f2 = 0, f3 = 0;
         #pragma unroll
         for (int k = 0; k < 10; ++k) {
            f1 += fm[k] * fp[c_d1[k]];
            f2 += fm[k] * fp[c_d2[k]];
            f3 += fm[k] * fp[c_d3[k]];
         }
         f1 *= X.nf[0];
         f2 *= X.nf[1];
         f3 *= X.nf[2];
         V3 h = v3(X.rc[0] * f1 + X.rc[3] * f2 + X.rc[6] * f3, X.rc[1] * f1 + X.rc[4] * f2 + X.rc[7] * f3,
            X.rc[2] * f1 + X.rc[5] * f2 + X.rc[8] * f3);
         atomic_fixed3(gx, gy, gz, s, f * h);
         real cp[10];
         frac_to_cart10(X, fp, cp);
         V3 tau;
         mpole_in_potential(m, cp, tau, v, f, do_v != 0);
         atomic_fixed3(trq, trq + 1, trq + 2, 3 * s, f * tau);
      }
   }
   if (do_e) {
      double x = warp_sum(acc[0]), y = warp_sum(acc[1]);
      if ((threadIdx.x & 31) == 0) {
         atomicAdd(&dbuf[D_EM_RECIP], x);
         atomicAdd(&dbuf[D_EM_SELF], y);
      }
   }
   if (DO_G && do_v)
      block_add(v, 6, dbuf + D_VIR_MREC);
}

// reciprocal + self polarization terms per atom (after apx_pme_uind_fphi)
template <bool DO_G>
__global__ void k_recip_polar(int n, RecipX X, real f, real aewald, int do_e, int do_v, int mutual, const real4* __restrict__ mp0,
   const real4* __restrict__ mp1, const real2* __restrict__ mp2, const real* __restrict__ fmp, const real* __restrict__ fphi,
   const real* __restrict__ ud, const real* __restrict__ up, const real* __restrict__ fphid, const real* __restrict__ fphip,
   const real* __restrict__ fphidp, fixed_t* gx, fixed_t* gy, fixed_t* gz, fixed_t* trq, double* __restrict__ dbuf)
{
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   double acc[2] = {0, 0};
   double v[6] = {0, 0, 0, 0, 0, 0};
   if (s < n) {
      Mpole m = load_mpole(mp0, mp1, mp2, s);
      V3 d = v3(ud[3 * s], ud[3 * s + 1], ud[3 * s + 2]), q = v3(up[3 * s], up[3 * s + 1], up[3 * s + 2]);
      real fd[3], fq[3];
      #pragma unroll
      for (int k = 0; k < 3; ++k) {
         fd[k] = X.a[0][k] * d.x + X.a[1][k] * d.y + X.a[2][k] * d.z;
         fq[k] = X.a[0][k] * q.x + X.a[1][k] * q.y + X.a[2][k] * q.z;
      }
      const real* fp = fphi + 20 * s;
      real a3 = aewald * aewald * aewald * (real)0.5641895835477563;   // a^3/sqrt(pi)
      if (do_e) {
         acc[0] = (double)((real)0.5 * f * (fd[0] * fp[1] + fd[1] * fp[2] + fd[2] * fp[3]));
         acc[1] = (double)(-(real)(2.0 / 3.0) * f * a3 * (m.dx * d.x + m.dy * d.y + m.dz * d.z));
      }
      if (DO_G) {
         const real* fm = fmp + 10 * s;
         const real* pd = fphid + 10 * s;
         const real* pp = fphip + 10 * s;
         const real* ps = fphidp + 20 * s;
         real f1 = 0, f2 = 0, f3 = 0;
         #pragma unroll
         for (int k = 0; k < 3; ++k) {
            int j1 = c_d1[k + 1], j2 = c_d2[k + 1], j3 = c_d3[k + 1];
            real su = fd[k] + fq[k];
            f1 += su * fp[j1];
            f2 += su * fp[j2];
            f3 += su * fp[j3];
            if (mutual) {
               f1 += fd[k] * pp[j1] + fq[k] * pd[j1];
               f2 += fd[k] * pp[j2] + fq[k] * pd[j2];
               f3 += fd[k] * pp[j3] + fq[k] * pd[j3];
            }
         }
         #pragma unroll
         for (int k = 0; k < 10; ++k) {
            f1 += fm[k] * ps[c_d1[k]];
            f2 += fm[k] * ps[c_d2[k]];
            f3 += fm[k] * ps[c_d3[k]];
         }
         f1 *= (real)0.5 * X.nf[0];
         f2 *= (real)0.5 * X.nf[1];
         f3 *= (real)0.5 * X.nf[2];
         V3 h = v3(X.rc[0] * f1 + X.rc[3] * f2 + X.rc[6] * f3, X.rc[1] * f1 + X.rc[4] * f2 + X.rc[7] * f3,
            X.rc[2] * f1 + X.rc[5] * f2 + X.rc[8] * f3);
         atomic_fixed3(gx, gy, gz, s, f * h);
         // permanent multipole in the averaged induced potential (0.5 * (d+p))
         real half[10], cdp[10];
         #pragma unroll
         for (int k = 0; k < 10; ++k)
            half[k] = (real)0.5 * ps[k];
         frac_to_cart10(X, half, cdp);
         V3 tau;
         mpole_in_potential(m, cdp, tau, v, f, do_v != 0);
         V3 ub = (real)0.5 * (d + q);
         V3 dm = v3(m.dx, m.dy, m.dz);
         tau += ((real)(4.0 / 3.0) * a3) * cross3(dm, ub);
         atomic_fixed3(trq, trq + 1, trq + 2, 3 * s, f * tau);
         if (do_v) {
            // induced dipoles in the permanent potential gradient, and the mutual cross terms
            real cp[10];
            frac_to_cart10(X, fp, cp);
            V3 g = v3(cp[1], cp[2], cp[3]);
            V3 su = d + q;
            real w = -(real)0.5 * f;
            v[0] += (double)(w * su.x * g.x);
            v[1] += (double)(w * (real)0.5 * (su.x * g.y + su.y * g.x));
            v[2] += (double)(w * (real)0.5 * (su.x * g.z + su.z * g.x));
            v[3] += (double)(w * su.y * g.y);
            v[4] += (double)(w * (real)0.5 * (su.y * g.z + su.z * g.y));
            v[5] += (double)(w * su.z * g.z);
            if (mutual) {
               V3 gd, gp;
               gd.x = X.a[0][0] * pd[1] + X.a[0][1] * pd[2] + X.a[0][2] * pd[3];
               gd.y = X.a[1][0] * pd[1] + X.a[1][1] * pd[2] + X.a[1][2] * pd[3];
               gd.z = X.a[2][0] * pd[1] + X.a[2][1] * pd[2] + X.a[2][2] * pd[3];
               gp.x = X.a[0][0] * pp[1] + X.a[0][1] * pp[2] + X.a[0][2] * pp[3];
               gp.y = X.a[1][0] * pp[1] + X.a[1][1] * pp[2] + X.a[1][2] * pp[3];
               gp.z = X.a[2][0] * pp[1] + X.a[2][1] * pp[2] + X.a[2][2] * pp[3];
               // M_ab = q_a gd_b + d_a gp_b
               v[0] += (double)(w * (q.x * gd.x + d.x * gp.x));
               v[1] += (double)(w * (real)0.5 * (q.x * gd.y + d.x * gp.y + q.y * gd.x + d.y * gp.x));
               v[2] += (double)(w * (real)0.5 * (q.x * gd.z + d.x * gp.z + q.z * gd.x + d.z * gp.x));
               v[3] += (double)(w * (q.y * gd.y + d.y * gp.y));
               v[4] += (double)(w * (real)0.5 * (q.y * gd.z + d.y * gp.z + q.z * gd.y + d.z * gp.y));
               v[5] += (double)(w * (q.z * gd.z + d.z * gp.z));
            }
         }
      }
   }
   if (do_e) {
      double x = warp_sum(acc[0]), y = warp_sum(acc[1]);
      if ((threadIdx.x & 31) == 0) {
         atomicAdd(&dbuf[D_EP_RECIP], x);
         atomicAdd(&dbuf[D_EP_SELF], y);
      }
   }
   if (DO_G && do_v)
      block_add(v, 6, dbuf + D_VIR_PREC);
}

// E_pol = -1/2 f sum u_d . udir_p / alpha
__global__ void k_ep_dot(int n3, real f, const real4* __restrict__ tpj, const real* __restrict__ ud, const real* __restrict__ udirp,
   double* __restrict__ dbuf)
{
   int q = blockIdx.x * blockDim.x + threadIdx.x;
   double e = 0;
   if (q < n3)
      e = (double)(tpj[q / 3].z * ud[q] * udirp[q]);
   e = warp_sum(e);
   if ((threadIdx.x & 31) == 0 && e != 0.0)
      atomicAdd(&dbuf[D_EP_DOT], -0.5 * (double)f * e);
}

// cmp with the induced dipoles added to the dipole slot (for the cross virial)
__global__ void k_add_dipole(int n, const real4* __restrict__ mp0, const real* __restrict__ u, real4* __restrict__ out)
{
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= n)
      return;
   real4 m = mp0[s];
   m.y += u[3 * s];
   m.z += u[3 * s + 1];
   m.w += u[3 * s + 2];
   out[s] = m;
}

__global__ void k_grad_out(int n, const int* __restrict__ perm, const fixed_t* __restrict__ gx, const fixed_t* __restrict__ gy,
   const fixed_t* __restrict__ gz, double* __restrict__ out)
{
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= n)
      return;
   int i = perm[s];
   const double inv = 1.0 / APX_FIXED_SCALE;
   out[3 * i] = (double)(long long)gx[s] * inv;
   out[3 * i + 1] = (double)(long long)gy[s] * inv;
   out[3 * i + 2] = (double)(long long)gz[s] * inv;
}

__global__ void k_trq_to_real(int n3, const fixed_t* __restrict__ in, real* __restrict__ out)
{
   int q = blockIdx.x * blockDim.x + threadIdx.x;
   if (q < n3)
      out[q] = (real)((double)(long long)in[q] * (1.0 / APX_FIXED_SCALE));
}

RecipX make_recipx(apx_ctx* c)
{
   RecipX X;
   int nf[3] = {c->nfft1, c->nfft2, c->nfft3};
   double a[3][3];
   for (int cc = 0; cc < 3; ++cc)
      for (int f = 0; f < 3; ++f)
         a[cc][f] = nf[f] * (double)c->box.r[3 * f + cc];
   const int qi1[6] = {0, 1, 2, 0, 0, 1}, qi2[6] = {0, 1, 2, 1, 2, 2};
   auto at = [&](int f, int cc) { return a[cc][f]; };
   double ftc[6][6];
   for (int i1 = 0; i1 < 3; ++i1) {
      int k = qi1[i1];
      for (int i2 = 0; i2 < 3; ++i2)
         ftc[i2][i1] = at(qi1[i2], k) * at(qi1[i2], k);
      for (int i2 = 3; i2 < 6; ++i2)
         ftc[i2][i1] = 2 * at(qi1[i2], k) * at(qi2[i2], k);
   }
   for (int i1 = 3; i1 < 6; ++i1) {
      int k = qi1[i1], m = qi2[i1];
      for (int i2 = 0; i2 < 3; ++i2)
         ftc[i2][i1] = at(qi1[i2], k) * at(qi1[i2], m);
      for (int i2 = 3; i2 < 6; ++i2)
         ftc[i2][i1] = at(qi1[i2], k) * at(qi2[i2], m) + at(qi1[i2], m) * at(qi2[i2], k);
   }
   for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
         X.a[i][j] = (real)a[i][j];
   for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j)
         X.ftc[i][j] = (real)ftc[i][j];
   for (int i = 0; i < 9; ++i)
      X.rc[i] = c->box.r[i];
   for (int i = 0; i < 3; ++i)
      X.nf[i] = nf[i];
   return X;
}

} // namespace

#define MP_G 16

void apx_dfield_full(apx_ctx* c, bool want_ev);
void apx_pme_cross_virial(apx_ctx* c, real4* mpa, real4* mpb, double* out6);
void apx_unpack_dp_all(apx_ctx* c, const real4* in, real* d, real* p);

// stage 0: the whole evaluation.  1: enqueue everything up to the copy of the reduced scalars, do not wait (md.cu puts the closing
// kick and the thermostat behind it before it synchronises).  2: the rest of a stage-1 call, after the caller's synchronisation.
static bool energy_once(apx_ctx* c, int vers, bool do_m, bool do_p, apx_energy_result* out, bool do_vdw, bool do_val, bool defer,
   int stage = 0);

// energy(vers) of the electrostatic terms (+ vdW / valence when attached).  The first attempt never waits for the GPU between
// the solver and the energy epilogue (the solver's first batch of iterations is sized from recent solves); should that batch
// not have converged -- the iteration count grew -- the evaluation is simply repeated with a solver that waits for its batches.
// Every accumulator is zeroed at the top of an attempt, so the repetition is exact.
void apx_energy_impl(apx_ctx* c, int vers, bool do_m, bool do_p, apx_energy_result* out, bool do_vdw, bool do_val)
{
   if (!energy_once(c, vers, do_m, do_p, out, do_vdw, do_val, true)) {
      if (!energy_once(c, vers, do_m, do_p, out, do_vdw, do_val, false))
         APX_THROW("energy: the induced-dipole solver did not finish");
   }
}

static bool energy_once(apx_ctx* c, int vers, bool do_m, bool do_p, apx_energy_result* out, bool do_vdw, bool do_val, bool defer,
   int stage)
{
   const int n = c->n;
   const int a0 = c->a0, no = c->a1 - c->a0, n3 = 3 * no;      // per-atom passes run on the owned range
   const bool dist = c->dist.on != 0;
   cudaStream_t st = c->stream;
   const bool do_e = vers & APX_ENERGY, do_g = vers & APX_GRAD, do_v = (vers & APX_VIRIAL) && do_g, do_a = vers & APX_ANALYZ;
   do_m = do_m && c->opt.use_mpole;
   do_p = do_p && c->opt.use_polar;
   const bool ewald = c->opt.use_ewald != 0;
   const bool pair_ep = do_e && do_a;          // ANALYZE: pairwise polarization energy; otherwise dot product
   do_vdw = do_vdw && c->vdw.on;
   do_val = do_val && apx_valence_on(c);
   int iters = 0;
   bool induce_deferred = false;
   if (stage != 2) {
   c->md_forces_valid = 0;                     // the accumulators are rewritten (md.cu sets the flag again after its own calls)
   cudaEventRecord(c->ev2, st);
   if (apx_graph_begin(c, 0x1000)) {
      apx_rotpole(c);      // mpoleInit(vers) runs on every energy() call in the reference (src/amoeba/emplar.cpp:12)
      // ---- zero accumulators
      // gx gy gz trqf ebuf dbuf cnt are contiguous (arena_e, apx_api.cu): one memset
      CUDA_CHECK(cudaMemsetAsync(c->arena_e.p, 0, c->arena_e_bytes, st));
      apx_graph_end(c, 0x1000);
   }
   c->mpole_inited = 1;
   c->mpole_pme_valid = 0;
   // ---- vdW term on its own stream, beside everything below (joins before the reductions)
   // APX_VDW_AT: where the vdW stream forks from the main stream.  0 = here, beside the solver's prologue (permanent field:
   // throughput-bound kernels that the vdW rows slow down); 1 = after the prologue, beside the PCG iterations (latency-bound
   // chains that leave most of every SM idle); 2 = after the solver, beside the energy epilogue
   static const int vdw_at = getenv("APX_VDW_AT") ? atoi(getenv("APX_VDW_AT")) : 0;      // measured on dhfr2 MD: 1.20 / 1.25 / 1.29 ms per step for 0 / 1 / 2
   c->vdw_fork_vers = -1;
   if (do_vdw && (vdw_at == 0 || !do_p || c->dist.on))
      apx_vdw_launch(c, vers);
   else if (do_vdw && vdw_at == 1)
      c->vdw_fork_vers = vers;      // apx_induce_impl forks it between its prologue and its iterations
   // ---- valence terms, likewise (evalence.cu)
   if (do_val)
      apx_valence_launch(c, vers);
   // ---- induced dipoles (also runs the permanent PME round trip -> fmp, fphi, conv E/virial)
   // (with the device-side loop the solve is only enqueued here: nothing waits for the convergence flag, the epilogue below is
   //  enqueued behind it, and the solver's host-side bookkeeping runs after the one synchronisation of this function)
   if (do_p) {
      induce_deferred = apx_induce_impl(c, defer);
      iters = c->stats.pcg_iterations;
   } else if (ewald) {
      apx_pme_mpole(c, true);
   }
   if (do_vdw && do_p && !c->dist.on && vdw_at == 2)
      apx_vdw_launch(c, vers);
   if (c->vdw_fork_vers >= 0) {      // (direct polarization: no iterations to fork beside)
      apx_vdw_launch(c, c->vdw_fork_vers);
      c->vdw_fork_vers = -1;
   }
   if (dist && do_p) {
      // converged dipoles of the neighbours owned by other GPUs
      apx_pack_dp(c, c->uind, c->uinp, c->pk_p);
      apx_dist_halo(c, c->pk_p, st);
      apx_unpack_dp_all(c, c->pk_p, c->uind, c->uinp);
   }
   // ---- real space, reciprocal space, torques: one fixed launch sequence per (vers, terms) -> one CUDA graph
   } else {
      induce_deferred = c->epend_deferred != 0;
      iters = c->stats.pcg_iterations;
   }
   // Behind a deferred solve the region is the body of an IF node keyed on the solver's convergence flag: should the first batch
   // of iterations not have converged, nothing of it runs, the accumulators keep what the vdW / valence terms put there, and
   // the region is simply launched again once the remaining iterations are done (below).
   const bool cond_epilogue = do_p && c->opt.poltyp_mutual && !dist && c->use_graph != 0;
   // md.cu hands in the closing half-kick and the thermostat (c->epi_tail): behind a deferred solve they become part of the SAME
   // IF-node body as the epilogue -- one graph launch less per MD step (the separate tail graph started ~40 us after the
   // epilogue's last kernel, profiles/r02o_trace_md.txt).  They may only run where the epilogue is known to run on converged
   // dipoles exactly once: inside the IF body, or eagerly behind a solve that was waited for; in every other case
   // c->epi_tail_ran stays 0 and the caller launches its own tail.
   const bool with_tail = (bool)c->epi_tail && cond_epilogue;
   const int ekey = 0x4000 | (vers & 0xff) | (do_m ? 0x100 : 0) | (do_p ? 0x200 : 0) | (cond_epilogue ? 0x800 : 0) | (with_tail ? 0x10000 : 0);
   if (stage != 2)
      c->epi_tail_ran = 0;
   auto epilogue = [&]() {
   const bool eager = apx_graph_begin(c, ekey, cond_epilogue ? c->flags.p + 1 : nullptr);
   if (!eager && with_tail && apx_graph_is_conditional(c, ekey))
      c->epi_tail_ran = 1;
   if (eager) {
   MplarArgs A;
   A.a0 = c->a0;
   A.a1 = c->a1;
   A.box = c->box;
   A.cut2 = (real)(c->opt.cutoff * c->opt.cutoff);
   A.aewald = (real)c->opt.aewald;
   A.f = c->f_elec;
   A.vstart = c->rows.vstart;
   A.rcnt = c->rows.cnt;
   A.nbr = c->rows.nbr;
   A.posq = c->posq;
   A.ewald = ewald ? 1 : 0;
   A.tpj = c->tpj;
   A.thlval = c->thlval;
   A.nj = c->opt.njpolar;
   A.table = c->thole_table;
   A.mp0 = c->mp0;
   A.mp1 = c->mp1;
   A.mp2 = c->mp2;
   A.ud = c->uind;
   A.up = c->uinp;
   A.do_m = do_m;
   A.do_p = do_p;
   A.mutual = c->opt.poltyp_mutual;
   A.do_e = do_e;
   A.do_v = do_v;
   A.do_a = do_a;
   A.pair_ep = pair_ep;
   A.gx = c->gx;
   A.gy = c->gy;
   A.gz = c->gz;
   A.trq = c->trqf;
   A.ebuf = c->ebuf;
   A.cnt = c->cnt;
   A.cnt2 = c->cnt.p + 1;
   // The real-space rows (100 us at dhfr2) on the second stream, beside the reciprocal-space part of the epilogue (PME round
   // trip of the converged dipoles + the per-atom reciprocal kernels, 60 us): every accumulator they share is fixed-point or
   // double atomics, so the two branches commute.  They join before the torques are resolved.
   const bool fork_rows = ewald && !dist;
   cudaStream_t rs = fork_rows ? c->stream2 : st;
   if (fork_rows) {
      CUDA_CHECK(cudaEventRecord(c->ev_fork, st));
      CUDA_CHECK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
   }
   if (c->rows.nverlet > 0 && (do_m || do_p) && (do_g || do_e) && !(c->diag_skip & 1)) {
      int grid = rows_grid<MP_G>(c);
      if (do_g && ewald) k_mplar_rows<true, true, MP_G><<<grid, ROWS_BLOCK, 0, rs>>>(A);
      else if (do_g) k_mplar_rows<true, false, MP_G><<<grid, ROWS_BLOCK, 0, rs>>>(A);
      else if (ewald) k_mplar_rows<false, true, MP_G><<<grid, ROWS_BLOCK, 0, rs>>>(A);
      else k_mplar_rows<false, false, MP_G><<<grid, ROWS_BLOCK, 0, rs>>>(A);
      APX_COUNT_LAUNCH(c);
      if (c->nexcl > 0) {
         ListedD D;
         for (int q = 0; q < 9; ++q)
            D.l[q] = c->opt.lvec[q], D.r[q] = c->recip_d[q];
         D.xyz = c->xyz_d, D.perm = c->perm, D.sc = c->excl_sc_d;
         D.cut2 = c->opt.cutoff * c->opt.cutoff, D.aewald = c->opt.aewald, D.f = c->opt.electric / c->opt.dielec;
         int g = (c->nexcl + 127) / 128;
         if (do_g) k_mplar_listed<true><<<g, 128, 0, rs>>>(c->nexcl, c->excl_s, A, D);
         else k_mplar_listed<false><<<g, 128, 0, rs>>>(c->nexcl, c->excl_s, A, D);
         APX_COUNT_LAUNCH(c);
      }
   }
   if (fork_rows)
      CUDA_CHECK(cudaEventRecord(c->ev_join, c->stream2));
   // ---- reciprocal space + self
   if (ewald) {
      RecipX X = make_recipx(c);
      int g = std::max(1, (no + 127) / 128);
      const size_t o = (size_t)a0;
      if (do_m && !(c->diag_skip & 2)) {
         if (do_g)
            k_recip_mpole<true><<<g, 128, 0, st>>>(no, X, c->f_elec, (real)c->opt.aewald, do_e, do_v, c->mp0 + o, c->mp1 + o, c->mp2 + o,
               c->fmp + 10 * o, c->fphi + 20 * o, c->gx + o, c->gy + o, c->gz + o, c->trqf + 3 * o, c->dbuf);
         else
            k_recip_mpole<false><<<g, 128, 0, st>>>(no, X, c->f_elec, (real)c->opt.aewald, do_e, do_v, c->mp0 + o, c->mp1 + o, c->mp2 + o,
               c->fmp + 10 * o, c->fphi + 20 * o, c->gx + o, c->gy + o, c->gz + o, c->trqf + 3 * o, c->dbuf);
         APX_COUNT_LAUNCH(c);
      }
      if (do_p && (do_g || pair_ep) && !(c->diag_skip & 2)) {
         if (do_g)
            apx_pme_uind_fphi(c, c->uind, c->uinp, true);
         if (do_g)
            k_recip_polar<true><<<g, 128, 0, st>>>(no, X, c->f_elec, (real)c->opt.aewald, pair_ep, do_v, c->opt.poltyp_mutual, c->mp0 + o,
               c->mp1 + o, c->mp2 + o, c->fmp + 10 * o, c->fphi + 20 * o, c->uind + 3 * o, c->uinp + 3 * o, c->fphid + 10 * o,
               c->fphip + 10 * o, c->fphidp + 20 * o, c->gx + o, c->gy + o, c->gz + o, c->trqf + 3 * o, c->dbuf);
         else
            k_recip_polar<false><<<g, 128, 0, st>>>(no, X, c->f_elec, (real)c->opt.aewald, pair_ep, do_v, c->opt.poltyp_mutual, c->mp0 + o,
               c->mp1 + o, c->mp2 + o, c->fmp + 10 * o, c->fphi + 20 * o, c->uind + 3 * o, c->uinp + 3 * o, c->fphid + 10 * o,
               c->fphip + 10 * o, c->fphidp + 20 * o, c->gx + o, c->gy + o, c->gz + o, c->trqf + 3 * o, c->dbuf);
         APX_COUNT_LAUNCH(c);
         if (do_v) {
            // (M + up) x (M + ud) structure-factor product
            k_add_dipole<<<g, 128, 0, st>>>(no, c->mp0 + o, c->uinp + 3 * o, c->mpx_a + o);
            k_add_dipole<<<g, 128, 0, st>>>(no, c->mp0 + o, c->uind + 3 * o, c->mpx_b + o);
            c->stats.kernel_launches += 2;
            apx_pme_cross_virial(c, c->mpx_a, c->mpx_b, c->dbuf.p + D_VIR_CROSS);
         }
      }
   }
   if (fork_rows)
      CUDA_CHECK(cudaStreamWaitEvent(st, c->ev_join, 0));
   if (do_p && do_e && !pair_ep) {
      k_ep_dot<<<std::max(1, (n3 + 255) / 256), 256, 0, st>>>(n3, c->f_elec, c->tpj + a0, c->uind + 3 * (size_t)a0, c->udirp + 3 * (size_t)a0,
         c->dbuf);
      APX_COUNT_LAUNCH(c);
   }
   // ---- torques -> forces
   if (do_g) {
      k_trq_to_real<<<std::max(1, (n3 + 255) / 256), 256, 0, st>>>(n3, c->trqf + 3 * (size_t)a0, c->trq + 3 * (size_t)a0);
      APX_COUNT_LAUNCH(c);
      apx_torque(c, do_v);
   }
   if (with_tail) {
      const bool in_if = c->capturing && c->cond_outer != nullptr;
      if (in_if || (!c->capturing && !induce_deferred)) {
         c->epi_tail();
         c->epi_tail_ran = 1;
      }
   }
   apx_graph_end(c, ekey);
   }
   };
   const size_t tail = (size_t)((char*)(c->cnt.p + 4) - (char*)c->ebuf.p);
   if (stage != 2) {
   // (the kick inside the epilogue needs the vdW forces: their stream is joined first; it finished long ago, beside the prologue)
   if (with_tail && do_vdw)
      apx_vdw_join(c, false);
   epilogue();
   if (do_vdw && !with_tail)
      apx_vdw_join(c);
   else if (do_vdw)
      apx_vdw_copy_out(c);
   if (do_val)
      apx_valence_join(c);
   if (dist) {
      // every GPU holds partial sums: forces on frame atoms may belong to a neighbour's slab
      if (do_g)
         apx_dist_allreduce_u64(c, c->gx.p, (size_t)(c->gz.p + c->npad - c->gx.p));
      apx_dist_allreduce_u64(c, c->ebuf.p, 8);
      apx_dist_allreduce_f64(c, c->dbuf.p, D_TOTAL);
      if (do_a)
         apx_dist_allreduce_i32(c, c->cnt.p, 4);
   }
   // ---- reductions to the host (energy.cpp:334-384)
   // ebuf, dbuf and cnt sit back to back at the end of the accumulator arena: ONE copy into pinned memory
   apx_induce_copy_out(c);
   CUDA_CHECK(cudaMemcpyAsync(c->red_h, c->ebuf.p, tail, cudaMemcpyDeviceToHost, st));
   cudaEventRecord(c->ev3, st);
   c->epend_deferred = induce_deferred ? 1 : 0;
   }
   if (stage == 1)
      return true;
   CUDA_CHECK(cudaStreamSynchronize(st));
   if (induce_deferred) {
      if (!apx_induce_finish(c)) {      // iteration count, timings, predictor history, the not-converged error
         // the unawaited first batch of iterations did not converge
         c->stats.energy_retries++;
         if (!apx_graph_is_conditional(c, ekey))
            return false;      // the epilogue ran on unconverged dipoles: the caller repeats the whole evaluation
         apx_induce_resume(c);      // the rest of the iterations, waited for
         epilogue();                // its IF node did not fire the first time: the accumulators are as the other terms left them
         CUDA_CHECK(cudaMemcpyAsync(c->red_h, c->ebuf.p, tail, cudaMemcpyDeviceToHost, st));
         cudaEventRecord(c->ev3, st);
         CUDA_CHECK(cudaStreamSynchronize(st));
      }
      iters = c->stats.pcg_iterations;
   }
   const fixed_t* eb = reinterpret_cast<const fixed_t*>(c->red_h);
   const double* db = reinterpret_cast<const double*>(c->red_h + ((char*)c->dbuf.p - (char*)c->ebuf.p));
   const int* cn = reinterpret_cast<const int*>(c->red_h + ((char*)c->cnt.p - (char*)c->ebuf.p));
   cudaEventElapsedTime(&c->stats.ms_energy, c->ev2, c->ev3);
   auto fx = [](fixed_t v) { return (double)(long long)v / APX_FIXED_SCALE; };
   apx_energy_result r;
   r.em = r.ep = 0;
   r.nem = r.nep = 0;
   if (do_e && do_m)
      r.em = fx(eb[0]) + (ewald ? db[D_EM_RECIP] + db[D_EM_SELF] : 0.0);
   if (do_e && do_p)
      r.ep = pair_ep ? fx(eb[1]) + (ewald ? db[D_EP_RECIP] + db[D_EP_SELF] : 0.0) : db[D_EP_DOT];
   r.esum = r.em + r.ep;
   if (do_a) {
      int npair = cn[0] + cn[1] / 2;
      r.nem = do_m ? npair + (ewald ? n : 0) : 0;
      r.nep = do_p ? npair + (ewald ? n : 0) : 0;
   }
   for (int q = 0; q < 9; ++q)
      r.virial[q] = 0;
   if (do_v) {
      double v6[6];
      for (int q = 0; q < 6; ++q) {
         v6[q] = fx(eb[2 + q]) + db[D_VIR_TRQ + q];
         if (ewald) {
            if (do_m)
               v6[q] += db[D_VIR_MREC + q] + db[D_CONV_E + 1 + q];
            if (do_p)
               v6[q] += db[D_VIR_PREC + q] - db[D_CONV_E + 1 + q] + db[D_VIR_CROSS + q];
         }
      }
      r.virial[0] = v6[0], r.virial[1] = v6[1], r.virial[2] = v6[2];
      r.virial[3] = v6[1], r.virial[4] = v6[3], r.virial[5] = v6[4];
      r.virial[6] = v6[2], r.virial[7] = v6[4], r.virial[8] = v6[5];
   }
   r.pcg_iterations = iters;
   r.pcg_eps = do_p ? c->scal_h[2] : 0.0;
   r.ev = 0;
   r.nev = 0;
   if (do_vdw) {
      apx_vdw_collect(c, vers, &r);
      r.esum += r.ev;
   }
   r.evalence = 0;
   for (int t = 0; t < 8; ++t)
      r.eval_term[t] = 0, r.nval_term[t] = 0;
   if (do_val) {
      apx_valence_result vr;
      apx_valence_collect(c, vers, &vr);
      r.evalence = vr.esum;
      r.esum += vr.esum;
      for (int t = 0; t < 8; ++t)
         r.eval_term[t] = vr.e[t], r.nval_term[t] = vr.count[t];
      if (do_v)
         for (int q = 0; q < 9; ++q)
            r.virial[q] += vr.virial[q];
   }
   apx_valence_set_in_total(c, do_val && do_g ? 1 : 0);
   if (out)
      *out = r;
   return true;
}

void apx_energy_impl_md(apx_ctx* c, int vers, apx_energy_result* out)
{
   apx_energy_impl(c, vers, true, true, out, true, false);
}

// The same evaluation in two halves for the integrator (md.cu): everything enqueued, nothing awaited; then, after the caller has
// put its own work behind it and synchronised, the host-side rest.  The second half returns true when the solver's first batch
// had not converged and the iterations / the epilogue had to be finished there (the caller's conditional work did not run).
void apx_energy_md_enqueue(apx_ctx* c, int vers)
{
   (void)energy_once(c, vers, true, true, nullptr, true, false, true, 1);
}
bool apx_energy_md_collect(apx_ctx* c, int vers, apx_energy_result* out)
{
   const int misses = c->stats.energy_retries;
   if (!energy_once(c, vers, true, true, out, true, false, true, 2)) {
      // (only without conditional graph nodes: the epilogue ran on unconverged dipoles -- evaluate again, waiting for the solver)
      if (!energy_once(c, vers, true, true, out, true, false, false, 0))
         APX_THROW("energy: the induced-dipole solver did not finish");
      return true;
   }
   return c->stats.energy_retries != misses;
}

void apx_grad_to_caller(apx_ctx* c, double* dev_out)
{
   k_grad_out<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->n, c->perm, c->gx, c->gy, c->gz, dev_out);
   APX_COUNT_LAUNCH(c);
   if (apx_valence_in_total(c))      // the last energy() included the valence terms: their gradient is kept in caller order
      apx_valence_grad_out(c, dev_out, true);
}
