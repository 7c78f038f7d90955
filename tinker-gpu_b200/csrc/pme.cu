// PME reciprocal pipeline: order-5 B-spline spreading of multipoles / induced-dipole pairs, 3-D FFT
// (cuFFT C2C, the d-dipoles in the real and the p-dipoles in the imaginary part as the reference
// packs them, src/acc/pme.cpp:147-183), influence-function multiply, and the gather of potential
// derivatives -- gridMpole/gridUind/pmeConv/fphiMpole/fphiUind(2) + the cart<->frac transforms of
// src/pme.cpp:221-351, src/cu/pme.cu:14-1007 and the recip/self field terms of
// src/cu/amoeba/field.cu:13-24,72-104.
//
// Organisation: one WARP per atom.  Lanes 0..2 build the three 1-D spline tables
// (values + 3 derivatives) in shared memory; the 125 stencil points are then covered by the 32
// lanes so that neighbouring lanes touch neighbouring x-addresses.  Cartesian->fractional
// conversion of the multipoles/dipoles is done inside the spread kernel, the fractional->Cartesian
// conversion and the Ewald self term inside the gather kernel, so a ufield round trip is
// memset + spread + FFT + multiply + FFT + gather.  The influence function is tabulated once per
// box (the reference recomputes exp() per grid point on every call, "TODO: store vs recompute",
// src/amoeba/field.cpp:87).
#include "apx_internal.h"
#include "dp.cuh"
#include <cmath>

namespace {
struct Xform {
   real a[3][3];     // a[c][f] = nfft_f * recip_f[c]   (cart -> frac for vectors: v_f = sum_c a[c][f] v_c)
   real ctf[6][6];   // quadrupole cart -> frac
   real ftc[6][6];   // potential second derivatives frac -> cart
};

// --- B-splines ---------------------------------------------------------------------------------
// th[p][l]: l-th derivative of the order-5 spline weight of stencil point p (p = 0..4)
__device__ void bspline5(real w, real th[5][4])
{
   real a2[2] = {1 - w, w};
   real a3[3], a4[4], a5[5];
   // order k from order k-1 :  M_k[j] = ((w + k-1-j) M_{k-1}[j-1] + (j+1-w) M_{k-1}[j]) / (k-1)
   a3[0] = (real)0.5 * (1 - w) * a2[0];
   a3[1] = (real)0.5 * ((w + 1) * a2[0] + (2 - w) * a2[1]);
   a3[2] = (real)0.5 * w * a2[1];
   const real t3 = (real)(1.0 / 3.0);
   a4[0] = t3 * (1 - w) * a3[0];
   a4[1] = t3 * ((w + 2) * a3[0] + (2 - w) * a3[1]);
   a4[2] = t3 * ((w + 1) * a3[1] + (3 - w) * a3[2]);
   a4[3] = t3 * w * a3[2];
   a5[0] = (real)0.25 * (1 - w) * a4[0];
   a5[1] = (real)0.25 * ((w + 3) * a4[0] + (2 - w) * a4[1]);
   a5[2] = (real)0.25 * ((w + 2) * a4[1] + (3 - w) * a4[2]);
   a5[3] = (real)0.25 * ((w + 1) * a4[2] + (4 - w) * a4[3]);
   a5[4] = (real)0.25 * w * a4[3];
   // derivative of an order-k spline = backward difference of the order-(k-1) spline
   real d1[5] = {-a4[0], a4[0] - a4[1], a4[1] - a4[2], a4[2] - a4[3], a4[3]};
   real e3[4] = {-a3[0], a3[0] - a3[1], a3[1] - a3[2], a3[2]};
   real d2[5] = {-e3[0], e3[0] - e3[1], e3[1] - e3[2], e3[2] - e3[3], e3[3]};
   real g2[3] = {-a2[0], a2[0] - a2[1], a2[1]};
   real g3[4] = {-g2[0], g2[0] - g2[1], g2[1] - g2[2], g2[2]};
   real d3[5] = {-g3[0], g3[0] - g3[1], g3[1] - g3[2], g3[2] - g3[3], g3[3]};
   #pragma unroll
   for (int p = 0; p < 5; ++p) {
      th[p][0] = a5[p];
      th[p][1] = d1[p];
      th[p][2] = d2[p];
      th[p][3] = d3[p];
   }
}

struct Stencil {
   int i1, i2, i3;   // first grid index along each axis
};

// Per-atom spline table, filled once per step (the positions do not change inside the CG loop):
// 16 real4 per atom -- [5*d + p] = {theta and its first three derivatives} of stencil point p along
// axis d (15 entries), [15] = the three stencil origins as integer bit patterns (bsplineFill of the
// reference, src/cu/pme.cu, keeps the same information as thetai1..3 + igrid).
__global__ void k_theta_fill(int n, Box b, int n1, int n2, int n3, const pos_t* __restrict__ posq, real4* __restrict__ theta)
{
   int t = blockIdx.x * blockDim.x + threadIdx.x;
   int s = t >> 2, d = t & 3;
   if (s >= n)
      return;
   const pos_t pos = posq[s];
   int nf[3] = {n1, n2, n3};
   int ig[3];
   real ww[3];
#ifdef APX_DOUBLE
   real f[3];
   f[0] = pos.x * b.r[0] + pos.y * b.r[1] + pos.z * b.r[2];
   f[1] = pos.x * b.r[3] + pos.y * b.r[4] + pos.z * b.r[5];
   f[2] = pos.x * b.r[6] + pos.y * b.r[7] + pos.z * b.r[8];
   #pragma unroll
   for (int q = 0; q < 3; ++q) {
      real w = f[q] + (real)0.5;
      w -= floor(w);
      real fr = nf[q] * w;
      int ii = (int)floor(fr);
      if (ii >= nf[q]) ii = nf[q] - 1;      // w == 1-ulp rounding guard
      ww[q] = fr - ii;
      ii -= 4;
      ig[q] = ii < 0 ? ii + nf[q] : ii;
   }
#else
   // 32-bit fractional coordinates: w = f + 1/2 (mod 1) is an integer add that wraps, nfft*w a 32x32 -> 64 bit product whose
   // high word is the grid cell and whose low word is the position inside the cell, exact to 2^-32 of a cell (float positions
   // lose 4e-6 of a cell at nfft = 64, which shows up in the reciprocal-space forces)
   const unsigned qq[3] = {pos.x, pos.y, pos.z};
   #pragma unroll
   for (int q = 0; q < 3; ++q) {
      const unsigned qs = qq[q] + 0x80000000u;
      int ii = (int)__umulhi(qs, (unsigned)nf[q]);
      ww[q] = (real)(qs * (unsigned)nf[q]) * (real)2.3283064365386963e-10;
      ii -= 4;
      ig[q] = ii < 0 ? ii + nf[q] : ii;
   }
#endif
   if (d < 3) {
      real th[5][4];
      bspline5(d == 0 ? ww[0] : (d == 1 ? ww[1] : ww[2]), th);
      #pragma unroll
      for (int p = 0; p < 5; ++p) {
         real4 o;
         o.x = th[p][0], o.y = th[p][1], o.z = th[p][2], o.w = th[p][3];
         theta[16 * (size_t)s + 5 * d + p] = o;
      }
   } else {
      real4 o;
#ifdef APX_DOUBLE
      o.x = __longlong_as_double((long long)ig[0]), o.y = __longlong_as_double((long long)ig[1]);
      o.z = __longlong_as_double((long long)ig[2]);
#else
      o.x = __int_as_float(ig[0]), o.y = __int_as_float(ig[1]), o.z = __int_as_float(ig[2]);
#endif
      o.w = 0;
      theta[16 * (size_t)s + 15] = o;
   }
}

// lanes 0..14 copy the table of atom s into sth[dim][5][4]; every lane returns the stencil origin
__device__ __forceinline__ Stencil load_stencil(const real4* __restrict__ theta, int s, real (*sth)[5][4], int lane)
{
   real4 v = theta[16 * (size_t)s + (lane & 15)];
   if (lane < 15)
      *reinterpret_cast<real4*>(&sth[0][0][0] + 4 * lane) = v;
   Stencil st;
#ifdef APX_DOUBLE
   st.i1 = (int)__double_as_longlong(__shfl_sync(0xffffffffu, v.x, 15));
   st.i2 = (int)__double_as_longlong(__shfl_sync(0xffffffffu, v.y, 15));
   st.i3 = (int)__double_as_longlong(__shfl_sync(0xffffffffu, v.z, 15));
#else
   st.i1 = __float_as_int(__shfl_sync(0xffffffffu, v.x, 15));
   st.i2 = __float_as_int(__shfl_sync(0xffffffffu, v.y, 15));
   st.i3 = __float_as_int(__shfl_sync(0xffffffffu, v.z, 15));
#endif
   __syncwarp();
   return st;
}

__device__ __forceinline__ int wrapi(int i, int n) { return i >= n ? i - n : i; }
// local plane of global plane (i3 + iz): a GPU holds planes zbase .. zbase+nzl-1 (mod n3) of the grid;
// on one GPU zbase = 0 and nzl = n3
__device__ __forceinline__ int zlocal(int i, int n3, int zbase)
{
   int z = wrapi(i, n3) - zbase;
   return z < 0 ? z + n3 : z;
}

// --- spread ------------------------------------------------------------------------------------
// Deterministic spreading (apx_set_pme_fixed_point, APX_PME_FIXED=1): the contributions are rounded to fixed point
// (2^32 per unit in the mixed build, 2^40 in the double build) and summed with 64-bit INTEGER reductions into a shadow grid of
// two int64 per point, which k_fix_to_grid converts into the float grid the FFT reads (and zeroes again).  Integer sums do
// not depend on the order in which the atoms arrive, so the grid -- and with the fixed lane order of the gathers and the row
// kernels everything computed from it -- is bit-reproducible from run to run; the float / vector reductions of the default
// path are not (neither are the reference's, src/cu/pme.cu:14-255).  Cost: four 8-byte reductions where the default path
// issues one 16-byte vector reduction.
#ifdef APX_DOUBLE
#define PME_FIX_SCALE 1099511627776.0      // 2^40
#else
#define PME_FIX_SCALE 4294967296.0         // 2^32
#endif
__device__ __forceinline__ void fix_add(long long* __restrict__ fix, size_t word, real v)
{
   atomicAdd(reinterpret_cast<unsigned long long*>(fix) + word, (unsigned long long)__double2ll_rn((double)v * PME_FIX_SCALE));
}
__global__ void k_fix_to_grid(size_t n, long long* __restrict__ fix, cplx* __restrict__ grid)
{
   size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n)
      return;
   longlong2 v = reinterpret_cast<longlong2*>(fix)[i];
   reinterpret_cast<longlong2*>(fix)[i] = make_longlong2(0, 0);
   cplx g;
   g.x = (real)((double)v.x * (1.0 / PME_FIX_SCALE));
   g.y = (real)((double)v.y * (1.0 / PME_FIX_SCALE));
   grid[i] = g;
}

__global__ void __launch_bounds__(128) k_spread_mpole(int n, Xform X, int n1, int n2, int n3, int zbase, int nzl,
   const real4* __restrict__ theta, const real4* __restrict__ mp0, const real4* __restrict__ mp1, const real2* __restrict__ mp2,
   real* __restrict__ fmp_out, cplx* __restrict__ grid, long long* __restrict__ fix)
{
   __shared__ __align__(16) real sth[4][3][5][4];
   const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
   int s = blockIdx.x * 4 + wib;
   if (s >= n)
      return;
   Stencil st = load_stencil(theta, s, sth[wib], lane);
   real4 m0 = mp0[s], m1 = mp1[s];
   real2 m2 = mp2[s];
   // Cartesian multipole with the off-diagonal quadrupoles doubled (rpoleToCmp), then -> fractional
   real cq[6] = {m1.x, m1.w, m2.y, 2 * m1.y, 2 * m1.z, 2 * m2.x};   // xx yy zz xy xz yz
   real cd[3] = {m0.y, m0.z, m0.w};
   real fm[10];
   fm[0] = m0.x;
   #pragma unroll
   for (int f = 0; f < 3; ++f)
      fm[1 + f] = X.a[0][f] * cd[0] + X.a[1][f] * cd[1] + X.a[2][f] * cd[2];
   #pragma unroll
   for (int j = 0; j < 6; ++j) {
      real t = 0;
      #pragma unroll
      for (int k = 0; k < 6; ++k)
         t += X.ctf[k][j] * cq[k];
      fm[4 + j] = t;
   }
   if (fmp_out && lane < 10) {
      real val = fm[0];
      #pragma unroll
      for (int q = 1; q < 10; ++q)
         if (lane == q)
            val = fm[q];
      fmp_out[10 * s + lane] = val;
   }
   for (int p = lane; p < 125; p += 32) {
      const int iz = p / 25, iy = (p / 5) % 5, ix = p % 5;
      const real* t = sth[wib][0][ix];
      const real* u = sth[wib][1][iy];
      const real* v = sth[wib][2][iz];
      const real t0 = t[0], t1 = t[1], t2 = t[2], u0 = u[0], u1 = u[1], u2 = u[2], v0 = v[0], v1 = v[1], v2 = v[2];
      const real val = fm[0] * t0 * u0 * v0 + fm[1] * t1 * u0 * v0 + fm[2] * t0 * u1 * v0 + fm[3] * t0 * u0 * v1 + fm[4] * t2 * u0 * v0
         + fm[5] * t0 * u2 * v0 + fm[6] * t0 * u0 * v2 + fm[7] * t1 * u1 * v0 + fm[8] * t1 * u0 * v1 + fm[9] * t0 * u1 * v1;
      const int zl = zlocal(st.i3 + iz, n3, zbase);
      if (zl < nzl) {
         const size_t at = ((size_t)zl * n2 + wrapi(st.i2 + iy, n2)) * n1 + wrapi(st.i1 + ix, n1);
         if (fix)
            fix_add(fix, 2 * at, val);
         else
            atomicAdd(&grid[at].x, val);
      }
   }
}

// ---- lane-group kernels for the induced-dipole grids: LG lanes share one atom, 128/LG atoms per
// CTA.  A whole warp per atom (as the multipole kernels above use) spends most of its issue slots on
// index arithmetic and the 5-step shuffle reductions; with 8 lanes per atom each lane keeps 16
// stencil points in flight and a reduction is 3 steps.
#define PME_LG 8
#define PME_APB (128 / PME_LG)      // atoms per CTA

// the LG lanes of a group copy their atom's 15 spline records into sth and return the stencil origin
template <int LG>
__device__ __forceinline__ Stencil load_stencil_lg(const real4* __restrict__ theta, int s, real (*sth)[5][4], int l)
{
   real4 org;
   #pragma unroll
   for (int q = l; q < 16; q += LG) {
      real4 v = theta[16 * (size_t)s + q];
      if (q < 15)
         *reinterpret_cast<real4*>(&sth[0][0][0] + 4 * q) = v;
      else
         org = v;
   }
   // record 15 is read by lane (15 % LG) of the group
   const int src = (threadIdx.x & 31 & ~(LG - 1)) + (15 % LG);
   Stencil st;
#ifdef APX_DOUBLE
   st.i1 = (int)__double_as_longlong(__shfl_sync(0xffffffffu, org.x, src));
   st.i2 = (int)__double_as_longlong(__shfl_sync(0xffffffffu, org.y, src));
   st.i3 = (int)__double_as_longlong(__shfl_sync(0xffffffffu, org.z, src));
#else
   st.i1 = __float_as_int(__shfl_sync(0xffffffffu, org.x, src));
   st.i2 = __float_as_int(__shfl_sync(0xffffffffu, org.y, src));
   st.i3 = __float_as_int(__shfl_sync(0xffffffffu, org.z, src));
#endif
   __syncwarp();
   return st;
}

// spread of a packed (d,p) dipole pair array (dp.cuh): d -> real part, p -> imaginary part
template <int LG>
__global__ void __launch_bounds__(128) k_spread_dp(int n, Xform X, int n1, int n2, int n3, int zbase, int nzl,
   const real4* __restrict__ theta,
   const real4* __restrict__ U, cplx* __restrict__ grid, const int* __restrict__ skip, long long* __restrict__ fix)
{
   if (skip && skip[1])
      return;
   __shared__ __align__(16) real sth[128 / LG][3][5][4];
   const int l = threadIdx.x & (LG - 1), gib = threadIdx.x / LG;
   const int s_ = blockIdx.x * (128 / LG) + gib;
   const bool act = s_ < n;
   const int s = act ? s_ : n - 1;
   Stencil st = load_stencil_lg<LG>(theta, s, sth[gib], l);
   if (!act)
      return;
   V3 d, q;
   load_dp(U, s, d, q);
   real fd[3], fp[3];
   #pragma unroll
   for (int f = 0; f < 3; ++f) {
      fd[f] = X.a[0][f] * d.x + X.a[1][f] * d.y + X.a[2][f] * d.z;
      fp[f] = X.a[0][f] * q.x + X.a[1][f] * q.y + X.a[2][f] * q.z;
   }
   // x fastest across the lanes: neighbouring lanes hit neighbouring addresses
   for (int p = l; p < 125; p += LG) {
      const int iz = p / 25, iy = (p / 5) % 5, ix = p % 5;
      const real* t = sth[gib][0][ix];
      const real* u = sth[gib][1][iy];
      const real* v = sth[gib][2][iz];
      const real w100 = t[1] * u[0] * v[0], w010 = t[0] * u[1] * v[0], w001 = t[0] * u[0] * v[1];
      const real vd = fd[0] * w100 + fd[1] * w010 + fd[2] * w001;
      const real vp = fp[0] * w100 + fp[1] * w010 + fp[2] * w001;
      const int zl = zlocal(st.i3 + iz, n3, zbase);
      if (zl >= nzl)
         continue;
      cplx* g = &grid[((size_t)zl * n2 + wrapi(st.i2 + iy, n2)) * n1 + wrapi(st.i1 + ix, n1)];
      if (fix) {
         fix_add(fix, 2 * (size_t)(g - grid), vd);
         fix_add(fix, 2 * (size_t)(g - grid) + 1, vp);
         continue;
      }
#ifdef APX_DOUBLE
      atomicAdd(&g->x, vd);
      atomicAdd(&g->y, vp);
#else
      atomicAdd(reinterpret_cast<float2*>(g), make_float2(vd, vp));
#endif
   }
}

#ifndef APX_DOUBLE
// ---- second generation of the induced-dipole spread / gather (mixed build, even nfft1) ---------------------------------------
// What ncu said about the kernels above at 1 M atoms (profiles/r01j_water1m_ncu_full_summary.txt): 7-8 % of the HBM roofline with
// 2-3 times the algorithmic DRAM traffic.  Two causes, both removed here:
//   * the per-atom spline table: 256 B per atom and launch, against 16 B of fractional coordinates.  Lanes 0..2 of an atom's
//     group evaluate the order-5 weights and first derivatives of one axis each (40 flops) into shared memory instead;
//   * one 8-byte reduction per grid point and atom.  Grid points that are neighbours in x are neighbours in memory, so the
//     five x-points of a stencil row are covered by THREE 16-byte aligned pairs (one padding point of weight 0): 75 vector
//     reductions (REDG.E.ADD.F32x4) per atom instead of 125, and 75 16-byte loads in the gather.  nfft1 even: a pair never
//     straddles the periodic wrap.
// weights and first derivatives only (dipoles): th[p] = {theta_p, theta'_p}
__device__ __forceinline__ void bspline5_01(real w, real2 th[5])
{
   const real a2_0 = 1 - w, a2_1 = w;
   const real a3_0 = (real)0.5 * (1 - w) * a2_0, a3_1 = (real)0.5 * ((w + 1) * a2_0 + (2 - w) * a2_1), a3_2 = (real)0.5 * w * a2_1;
   const real t3 = (real)(1.0 / 3.0);
   const real a4_0 = t3 * (1 - w) * a3_0, a4_1 = t3 * ((w + 2) * a3_0 + (2 - w) * a3_1), a4_2 = t3 * ((w + 1) * a3_1 + (3 - w) * a3_2),
              a4_3 = t3 * w * a3_2;
   th[0] = make_float2((real)0.25 * (1 - w) * a4_0, -a4_0);
   th[1] = make_float2((real)0.25 * ((w + 3) * a4_0 + (2 - w) * a4_1), a4_0 - a4_1);
   th[2] = make_float2((real)0.25 * ((w + 2) * a4_1 + (3 - w) * a4_2), a4_1 - a4_2);
   th[3] = make_float2((real)0.25 * ((w + 1) * a4_2 + (4 - w) * a4_3), a4_2 - a4_3);
   th[4] = make_float2((real)0.25 * w * a4_3, a4_3);
}

// lanes 0..2 of the group: axis l of atom s -> sth[axis][0..4] (+ two zero pads [5], [6] so that th[ix + 1] is defined for
// ix = -1 .. 5), stencil origin -> sorg[axis]
__device__ __forceinline__ void stencil_fill(const pos_t pos, int l, int n1, int n2, int n3, real2 (*sth)[8], int* sorg)
{
   if (l < 3) {
      const unsigned q = l == 0 ? pos.x : (l == 1 ? pos.y : pos.z);
      const unsigned nf = (unsigned)(l == 0 ? n1 : (l == 1 ? n2 : n3));
      const unsigned qs = q + 0x80000000u;
      int ii = (int)__umulhi(qs, nf);
      const real w = (real)(qs * nf) * (real)2.3283064365386963e-10;
      ii -= 4;
      sorg[l] = ii < 0 ? ii + (int)nf : ii;
      real2 th[5];
      bspline5_01(w, th);
      sth[l][0] = make_float2(0, 0);
      #pragma unroll
      for (int p = 0; p < 5; ++p)
         sth[l][p + 1] = th[p];
      sth[l][6] = make_float2(0, 0);
   }
   __syncwarp();
}

template <int LG>
__global__ void __launch_bounds__(128) k_spread_dp2(int n, Xform X, int n1, int n2, int n3, int zbase, int nzl,
   const pos_t* __restrict__ posq, const real4* __restrict__ U, cplx* __restrict__ grid, const int* __restrict__ skip,
   long long* __restrict__ fix)
{
   if (skip && skip[1])
      return;
   __shared__ real2 sth[128 / LG][3][8];
   __shared__ int sorg[128 / LG][4];
   const int l = threadIdx.x & (LG - 1), gib = threadIdx.x / LG;
   const int s_ = blockIdx.x * (128 / LG) + gib;
   const bool act = s_ < n;
   const int s = act ? s_ : n - 1;
   stencil_fill(posq[s], l, n1, n2, n3, sth[gib], sorg[gib]);
   if (!act)
      return;
   const int i1 = sorg[gib][0], i2 = sorg[gib][1], i3 = sorg[gib][2];
   V3 d, q;
   load_dp(U, s, d, q);
   real fd[3], fp[3];
   #pragma unroll
   for (int f = 0; f < 3; ++f) {
      fd[f] = X.a[0][f] * d.x + X.a[1][f] * d.y + X.a[2][f] * d.z;
      fp[f] = X.a[0][f] * q.x + X.a[1][f] * q.y + X.a[2][f] * q.z;
   }
   const int x0 = i1 & ~1;          // aligned pair holding the first stencil point
   const int off = i1 - x0;         // 0 or 1: stencil index of point x0 + j is j - off
   for (int it = l; it < 75; it += LG) {
      const int row = it / 3, pr = it - 3 * row;
      const int iz = row / 5, iy = row - 5 * iz;
      const int zl = zlocal(i3 + iz, n3, zbase);
      if (zl >= nzl)
         continue;
      const real2 u = sth[gib][1][iy + 1], v = sth[gib][2][iz + 1];
      const int j0 = 2 * pr - off;      // stencil index of the pair's first point: -1 .. 4
      const real2 t0 = sth[gib][0][j0 + 1], t1 = sth[gib][0][j0 + 2];
      const real uv = u.x * v.x, duv = u.y * v.x, udv = u.x * v.y;
      // value of point j:  fd . (t'_j u v, t_j u' v, t_j u v')
      const real a0 = t0.y * uv, b0 = t0.x * duv, c0 = t0.x * udv;
      const real a1 = t1.y * uv, b1 = t1.x * duv, c1 = t1.x * udv;
      float4 val;
      val.x = fd[0] * a0 + fd[1] * b0 + fd[2] * c0;
      val.y = fp[0] * a0 + fp[1] * b0 + fp[2] * c0;
      val.z = fd[0] * a1 + fd[1] * b1 + fd[2] * c1;
      val.w = fp[0] * a1 + fp[1] * b1 + fp[2] * c1;
      int x = x0 + 2 * pr;
      x = x >= n1 ? x - n1 : x;
      cplx* g = &grid[((size_t)zl * n2 + wrapi(i2 + iy, n2)) * n1 + x];
      if (fix) {
         const size_t w = 2 * (size_t)(g - grid);
         fix_add(fix, w, val.x), fix_add(fix, w + 1, val.y), fix_add(fix, w + 2, val.z), fix_add(fix, w + 3, val.w);
         continue;
      }
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(g), "f"(val.x), "f"(val.y), "f"(val.z), "f"(val.w) : "memory");
   }
}

template <int EPI, int LG>
__global__ void __launch_bounds__(128) k_gather_dp2(int n, Xform X, int n1, int n2, int n3, int zbase, int nzl, real selfterm,
   const pos_t* __restrict__ posq, const real4* __restrict__ tpj, const cplx* __restrict__ grid, const real4* __restrict__ U,
   const real4* __restrict__ F, real* __restrict__ out_d, real* __restrict__ out_p, real4* __restrict__ OUT,
   double* __restrict__ slot, const int* __restrict__ skip, const int* __restrict__ itp)
{
   if (skip && skip[1])
      return;
   __shared__ real2 sth[128 / LG][3][8];
   __shared__ int sorg[128 / LG][4];
   const int l = threadIdx.x & (LG - 1), gib = threadIdx.x / LG;
   const int s_ = blockIdx.x * (128 / LG) + gib;
   const bool act = s_ < n;
   const int s = act ? s_ : n - 1;
   double dot_d = 0, dot_p = 0;
   stencil_fill(posq[s], l, n1, n2, n3, sth[gib], sorg[gib]);
   const int i1 = sorg[gib][0], i2 = sorg[gib][1], i3 = sorg[gib][2];
   const int x0 = i1 & ~1, off = i1 - x0;
   real fd[3] = {0, 0, 0}, fp[3] = {0, 0, 0};
   for (int it = l; it < 75; it += LG) {
      const int row = it / 3, pr = it - 3 * row;
      const int iz = row / 5, iy = row - 5 * iz;
      const int zl = min(zlocal(i3 + iz, n3, zbase), nzl - 1);
      int x = x0 + 2 * pr;
      x = x >= n1 ? x - n1 : x;
      const float4 g = __ldg(reinterpret_cast<const float4*>(&grid[((size_t)zl * n2 + wrapi(i2 + iy, n2)) * n1 + x]));
      const real2 u = sth[gib][1][iy + 1], v = sth[gib][2][iz + 1];
      const int j0 = 2 * pr - off;
      const real2 t0 = sth[gib][0][j0 + 1], t1 = sth[gib][0][j0 + 2];
      const real uv = u.x * v.x, duv = u.y * v.x, udv = u.x * v.y;
      const real a = t0.y * g.x + t1.y * g.z, b = t0.x * g.x + t1.x * g.z;      // d grid: sum of t' g and of t g over the pair
      const real ap = t0.y * g.y + t1.y * g.w, bp = t0.x * g.y + t1.x * g.w;    // p grid
      fd[0] += a * uv, fd[1] += b * duv, fd[2] += b * udv;
      fp[0] += ap * uv, fp[1] += bp * duv, fp[2] += bp * udv;
   }
   #pragma unroll
   for (int q = 0; q < 3; ++q) {
      #pragma unroll
      for (int o = LG / 2; o > 0; o >>= 1) {
         fd[q] += __shfl_xor_sync(0xffffffffu, fd[q], o);
         fp[q] += __shfl_xor_sync(0xffffffffu, fp[q], o);
      }
   }
   if (l == 0 && act) {
      V3 ud, up;
      load_dp(U, s, ud, up);
      V3 cd = v3(X.a[0][0] * fd[0] + X.a[0][1] * fd[1] + X.a[0][2] * fd[2], X.a[1][0] * fd[0] + X.a[1][1] * fd[1] + X.a[1][2] * fd[2],
         X.a[2][0] * fd[0] + X.a[2][1] * fd[1] + X.a[2][2] * fd[2]);
      V3 cp = v3(X.a[0][0] * fp[0] + X.a[0][1] * fp[1] + X.a[0][2] * fp[2], X.a[1][0] * fp[0] + X.a[1][1] * fp[1] + X.a[1][2] * fp[2],
         X.a[2][0] * fp[0] + X.a[2][1] * fp[1] + X.a[2][2] * fp[2]);
      V3 ed = selfterm * ud - cd, ep = selfterm * up - cp;
      if (F) {
         V3 a, b;
         load_dp(F, s, a, b);
         ed += a;
         ep += b;
      }
      if (EPI == 0) {
         out_d[3 * s] = ed.x, out_d[3 * s + 1] = ed.y, out_d[3 * s + 2] = ed.z;
         out_p[3 * s] = ep.x, out_p[3 * s + 1] = ep.y, out_p[3 * s + 2] = ep.z;
      } else if (EPI == 1) {
         if (tpj[s].y == 0) {
            ed = v3(0, 0, 0);
            ep = v3(0, 0, 0);
         }
         store_dp(OUT, s, ed, ep);
      } else {
         real pinv = tpj[s].z;
         V3 vd = pinv * ud - ed, vp = pinv * up - ep;
         store_dp(OUT, s, vd, vp);
         dot_d = (double)ud.x * vd.x + (double)ud.y * vd.y + (double)ud.z * vd.z;
         dot_p = (double)up.x * vp.x + (double)up.y * vp.y + (double)up.z * vp.z;
      }
   }
   if (EPI == 2)
      pcg_block_add2(dot_d, dot_p, pcg_slot_of(slot, itp), 2, 3);
}
#endif

// --- influence function ------------------------------------------------------------------------
// (all three kernels below index a slab [k3][k2 in y0..y0+ny)[k1] of the transformed grid: the whole
//  grid on one GPU, the rows this GPU holds after the transpose of the slab FFT otherwise)
__global__ void k_make_qfac(int n1, int n2, int n3, int y0, int ny, Box box, real pterm, real volterm, const real* __restrict__ bs1,
   const real* __restrict__ bs2, const real* __restrict__ bs3, real* __restrict__ qfac)
{
   size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   size_t ntot = (size_t)n1 * ny * n3;
   if (i >= ntot)
      return;
   int k3 = (int)(i / ((size_t)n1 * ny)), j = (int)(i - (size_t)k3 * n1 * ny), k2 = y0 + j / n1, k1 = j % n1;
   int r1 = k1 < (n1 + 1) / 2 ? k1 : k1 - n1;
   int r2 = k2 < (n2 + 1) / 2 ? k2 : k2 - n2;
   int r3 = k3 < (n3 + 1) / 2 ? k3 : k3 - n3;
   double h1 = (double)box.r[0] * r1 + (double)box.r[3] * r2 + (double)box.r[6] * r3;
   double h2 = (double)box.r[1] * r1 + (double)box.r[4] * r2 + (double)box.r[7] * r3;
   double h3 = (double)box.r[2] * r1 + (double)box.r[5] * r2 + (double)box.r[8] * r3;
   double hsq = h1 * h1 + h2 * h2 + h3 * h3;
   double term = -(double)pterm * hsq;
   double e = 0;
   if ((k1 | k2 | k3) != 0 && term > -50.0)
      e = exp(term) / ((double)volterm * hsq * (double)bs1[k1] * (double)bs2[k2] * (double)bs3[k3]);
   qfac[i] = (real)e;
}

__global__ void k_conv(size_t ntot, const real* __restrict__ qfac, cplx* __restrict__ grid)
{
   size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= ntot)
      return;
   real f = qfac[i];
   cplx g = grid[i];
   g.x *= f;
   g.y *= f;
   grid[i] = g;
}

// multiply + reciprocal energy / virial of |Q|^2 (pmeConv<DO_E,DO_V>); out[0]=e, out[1..6]=vxx,vxy,vxz,vyy,vyz,vzz
__global__ void k_conv_ev(int n1, int n2, int n3, int y0, int ny, Box box, real pterm, real felec, const real* __restrict__ qfac,
   cplx* __restrict__ grid, double* __restrict__ out)
{
   size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   size_t ntot = (size_t)n1 * ny * n3;
   double acc[7] = {0, 0, 0, 0, 0, 0, 0};
   if (i < ntot) {
      real f = qfac[i];
      cplx g = grid[i];
      if (f != 0) {
         int k3 = (int)(i / ((size_t)n1 * ny)), j = (int)(i - (size_t)k3 * n1 * ny), k2 = y0 + j / n1, k1 = j % n1;
         int r1 = k1 < (n1 + 1) / 2 ? k1 : k1 - n1;
         int r2 = k2 < (n2 + 1) / 2 ? k2 : k2 - n2;
         int r3 = k3 < (n3 + 1) / 2 ? k3 : k3 - n3;
         double h1 = (double)box.r[0] * r1 + (double)box.r[3] * r2 + (double)box.r[6] * r3;
         double h2 = (double)box.r[1] * r1 + (double)box.r[4] * r2 + (double)box.r[7] * r3;
         double h3 = (double)box.r[2] * r1 + (double)box.r[5] * r2 + (double)box.r[8] * r3;
         double hsq = h1 * h1 + h2 * h2 + h3 * h3;
         double term = -(double)pterm * hsq;
         double struc2 = (double)g.x * g.x + (double)g.y * g.y;
         double eterm = 0.5 * (double)felec * (double)f * struc2;
         double vterm = (2.0 / hsq) * (1.0 - term) * eterm;
         acc[0] = eterm;
         acc[1] = h1 * h1 * vterm - eterm;
         acc[2] = h1 * h2 * vterm;
         acc[3] = h1 * h3 * vterm;
         acc[4] = h2 * h2 * vterm - eterm;
         acc[5] = h2 * h3 * vterm;
         acc[6] = h3 * h3 * vterm - eterm;
      }
      g.x *= f;
      g.y *= f;
      grid[i] = g;
   }
   __shared__ double sh[7][8];
   int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
   #pragma unroll
   for (int q = 0; q < 7; ++q) {
      double x = acc[q];
      for (int o = 16; o > 0; o >>= 1)
         x += __shfl_xor_sync(0xffffffffu, x, o);
      if (lane == 0)
         sh[q][w] = x;
   }
   __syncthreads();
   if (threadIdx.x < 7) {
      double x = 0;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k)
         x += sh[threadIdx.x][k];
      if (x != 0.0)
         atomicAdd(&out[threadIdx.x], x);
   }
}

// structure-factor cross product of two transformed grids (epolarEwaldRecipSelfVirial_cu5)
__global__ void k_cross_virial(int n1, int n2, int n3, int y0, int ny, Box box, real pterm, real felec, const real* __restrict__ qfac,
   const cplx* __restrict__ ga, const cplx* __restrict__ gb, double* __restrict__ out)
{
   size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   size_t ntot = (size_t)n1 * ny * n3;
   double acc[6] = {0, 0, 0, 0, 0, 0};
   if (i < ntot) {
      real f = qfac[i];
      if (f != 0) {
         int k3 = (int)(i / ((size_t)n1 * ny)), j = (int)(i - (size_t)k3 * n1 * ny), k2 = y0 + j / n1, k1 = j % n1;
         int r1 = k1 < (n1 + 1) / 2 ? k1 : k1 - n1;
         int r2 = k2 < (n2 + 1) / 2 ? k2 : k2 - n2;
         int r3 = k3 < (n3 + 1) / 2 ? k3 : k3 - n3;
         double h1 = (double)box.r[0] * r1 + (double)box.r[3] * r2 + (double)box.r[6] * r3;
         double h2 = (double)box.r[1] * r1 + (double)box.r[4] * r2 + (double)box.r[7] * r3;
         double h3 = (double)box.r[2] * r1 + (double)box.r[5] * r2 + (double)box.r[8] * r3;
         double hsq = h1 * h1 + h2 * h2 + h3 * h3;
         double term = -(double)pterm * hsq;
         cplx a = ga[i], b = gb[i];
         double struc2 = (double)a.x * b.x + (double)a.y * b.y;
         double eterm = 0.5 * (double)felec * (double)f * struc2;
         double vterm = (2.0 / hsq) * (1.0 - term) * eterm;
         acc[0] = h1 * h1 * vterm - eterm;
         acc[1] = h1 * h2 * vterm;
         acc[2] = h1 * h3 * vterm;
         acc[3] = h2 * h2 * vterm - eterm;
         acc[4] = h2 * h3 * vterm;
         acc[5] = h3 * h3 * vterm - eterm;
      }
   }
   __shared__ double sh[6][8];
   int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
   #pragma unroll
   for (int q = 0; q < 6; ++q) {
      double x = acc[q];
      for (int o = 16; o > 0; o >>= 1)
         x += __shfl_xor_sync(0xffffffffu, x, o);
      if (lane == 0)
         sh[q][w] = x;
   }
   __syncthreads();
   if (threadIdx.x < 6) {
      double x = 0;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k)
         x += sh[threadIdx.x][k];
      if (x != 0.0)
         atomicAdd(&out[threadIdx.x], x);
   }
}

// --- gather ------------------------------------------------------------------------------------
// 25 lanes own one (iy,iz) row of five x-points each.
// MODE 0: permanent multipoles: fphi[20] stored; field assigned: fd = term*dipole - grad_cart(phi)
// (the ufield gather with its fused epilogues is k_gather_dp below)
// MODE 2: energy step: fphid[10], fphip[10], fphidp[20] stored
template <int MODE>
__global__ void __launch_bounds__(128) k_gather(int n, Xform X, int n1, int n2, int n3, int zbase, int nzl, real selfterm,
   const real4* __restrict__ theta, const cplx* __restrict__ grid, const real4* __restrict__ mp0, const real* __restrict__ ud,
   const real* __restrict__ up, real* __restrict__ out_a, real* __restrict__ out_b, real* __restrict__ out_c,
   const int* __restrict__ skip)
{
   if (skip && skip[1])
      return;
   __shared__ __align__(16) real sth[4][3][5][4];
   const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
   int s = blockIdx.x * 4 + wib;
   if (s >= n)
      return;
   Stencil st = load_stencil(theta, s, sth[wib], lane);
   // row sums over x for derivative orders 0..3, real and imaginary parts
   real tr[4] = {0, 0, 0, 0}, ti[4] = {0, 0, 0, 0};
   real u[4] = {0, 0, 0, 0}, v[4] = {0, 0, 0, 0};
   if (lane < 25) {
      int iy = lane % 5, iz = lane / 5;
      const int zl = min(zlocal(st.i3 + iz, n3, zbase), nzl - 1);
      size_t base = ((size_t)zl * n2 + wrapi(st.i2 + iy, n2)) * n1;
      #pragma unroll
      for (int ix = 0; ix < 5; ++ix) {
         cplx g = grid[base + wrapi(st.i1 + ix, n1)];
         const real* t = sth[wib][0][ix];
         #pragma unroll
         for (int l = 0; l < 4; ++l) {
            tr[l] += g.x * t[l];
            if (MODE != 0)
               ti[l] += g.y * t[l];
         }
      }
      #pragma unroll
      for (int l = 0; l < 4; ++l) {
         u[l] = sth[wib][1][iy][l];
         v[l] = sth[wib][2][iz][l];
      }
   }
#define WSUM(x)                                                                                                          \
   for (int o = 16; o > 0; o >>= 1)                                                                                        \
      x += __shfl_xor_sync(0xffffffffu, x, o)
   if (MODE == 0) {
      // 20 derivative combinations (a,b,c) in the reference's fphi order
      real f[20];
      f[0] = tr[0] * u[0] * v[0];
      f[1] = tr[1] * u[0] * v[0];
      f[2] = tr[0] * u[1] * v[0];
      f[3] = tr[0] * u[0] * v[1];
      f[4] = tr[2] * u[0] * v[0];
      f[5] = tr[0] * u[2] * v[0];
      f[6] = tr[0] * u[0] * v[2];
      f[7] = tr[1] * u[1] * v[0];
      f[8] = tr[1] * u[0] * v[1];
      f[9] = tr[0] * u[1] * v[1];
      f[10] = tr[3] * u[0] * v[0];
      f[11] = tr[0] * u[3] * v[0];
      f[12] = tr[0] * u[0] * v[3];
      f[13] = tr[2] * u[1] * v[0];
      f[14] = tr[2] * u[0] * v[1];
      f[15] = tr[1] * u[2] * v[0];
      f[16] = tr[0] * u[2] * v[1];
      f[17] = tr[1] * u[0] * v[2];
      f[18] = tr[0] * u[1] * v[2];
      f[19] = tr[1] * u[1] * v[1];
      #pragma unroll
      for (int q = 0; q < 20; ++q) {
         WSUM(f[q]);
      }
      if (lane < 20) {
         real val = f[0];
         #pragma unroll
         for (int q = 1; q < 20; ++q)
            if (lane == q)
               val = f[q];
         out_a[20 * s + lane] = val;
      }
      if (lane < 3) {
         // Cartesian gradient of the reciprocal potential, component `lane`
         real cphi = X.a[lane][0] * f[1] + X.a[lane][1] * f[2] + X.a[lane][2] * f[3];
         real4 m0 = mp0[s];
         real d = lane == 0 ? m0.y : (lane == 1 ? m0.z : m0.w);
         out_b[3 * s + lane] = selfterm * d - cphi;
      }
   } else {
      real fd[10], fp[10], fs[20];
      real ts[4];
      #pragma unroll
      for (int l = 0; l < 4; ++l)
         ts[l] = tr[l] + ti[l];
#define COMBO(arr, t)                                                                                                    \
   arr[0] = t[0] * u[0] * v[0];                                                                                            \
   arr[1] = t[1] * u[0] * v[0];                                                                                            \
   arr[2] = t[0] * u[1] * v[0];                                                                                            \
   arr[3] = t[0] * u[0] * v[1];                                                                                            \
   arr[4] = t[2] * u[0] * v[0];                                                                                            \
   arr[5] = t[0] * u[2] * v[0];                                                                                            \
   arr[6] = t[0] * u[0] * v[2];                                                                                            \
   arr[7] = t[1] * u[1] * v[0];                                                                                            \
   arr[8] = t[1] * u[0] * v[1];                                                                                            \
   arr[9] = t[0] * u[1] * v[1];
      COMBO(fd, tr)
      COMBO(fp, ti)
      COMBO(fs, ts)
      fs[10] = ts[3] * u[0] * v[0];
      fs[11] = ts[0] * u[3] * v[0];
      fs[12] = ts[0] * u[0] * v[3];
      fs[13] = ts[2] * u[1] * v[0];
      fs[14] = ts[2] * u[0] * v[1];
      fs[15] = ts[1] * u[2] * v[0];
      fs[16] = ts[0] * u[2] * v[1];
      fs[17] = ts[1] * u[0] * v[2];
      fs[18] = ts[0] * u[1] * v[2];
      fs[19] = ts[1] * u[1] * v[1];
      fd[0] = 0;
      fp[0] = 0;
      #pragma unroll
      for (int q = 0; q < 10; ++q) {
         WSUM(fd[q]);
         WSUM(fp[q]);
      }
      #pragma unroll
      for (int q = 0; q < 20; ++q) {
         WSUM(fs[q]);
      }
      if (lane == 0) {
         #pragma unroll
         for (int q = 0; q < 10; ++q) {
            out_a[10 * s + q] = fd[q];
            out_b[10 * s + q] = fp[q];
         }
         #pragma unroll
         for (int q = 0; q < 20; ++q)
            out_c[20 * s + q] = fs[q];
      }
   }
#undef WSUM
#undef COMBO
}

// Gather of the reciprocal mutual field of a packed dipole pair U, fused with what follows it:
//   field = selfterm*U - grad_cart(phi) + F        (F = real-space field from the row kernel, may be null)
//   EPI 0: plain output  fd, fp [n][3]                               (ufield operator)
//   EPI 1: residual      R = field, zero where alpha == 0            (r0 = -T u0)
//   EPI 2: PCG           V = U/alpha - field ; partial U.V -> slot   (pcgP1 + dots, src/cu/induce.cu)
template <int EPI, int LG>
__global__ void __launch_bounds__(128) k_gather_dp(int n, Xform X, int n1, int n2, int n3, int zbase, int nzl, real selfterm,
   const real4* __restrict__ theta, const real4* __restrict__ tpj, const cplx* __restrict__ grid, const real4* __restrict__ U,
   const real4* __restrict__ F, real* __restrict__ out_d, real* __restrict__ out_p, real4* __restrict__ OUT,
   double* __restrict__ slot, const int* __restrict__ skip, const int* __restrict__ itp)
{
   if (skip && skip[1])
      return;
   __shared__ __align__(16) real sth[128 / LG][3][5][4];
   const int l = threadIdx.x & (LG - 1), gib = threadIdx.x / LG;
   const int s_ = blockIdx.x * (128 / LG) + gib;
   const bool act = s_ < n;
   const int s = act ? s_ : n - 1;
   double dot_d = 0, dot_p = 0;
   Stencil st = load_stencil_lg<LG>(theta, s, sth[gib], l);
   real fd[3] = {0, 0, 0}, fp[3] = {0, 0, 0};
   for (int p = l; p < 125; p += LG) {
      const int iz = p / 25, iy = (p / 5) % 5, ix = p % 5;
      const int zl = min(zlocal(st.i3 + iz, n3, zbase), nzl - 1);
      const cplx g = grid[((size_t)zl * n2 + wrapi(st.i2 + iy, n2)) * n1 + wrapi(st.i1 + ix, n1)];
      const real* t = sth[gib][0][ix];
      const real* u = sth[gib][1][iy];
      const real* v = sth[gib][2][iz];
      const real w100 = t[1] * u[0] * v[0], w010 = t[0] * u[1] * v[0], w001 = t[0] * u[0] * v[1];
      fd[0] += g.x * w100, fd[1] += g.x * w010, fd[2] += g.x * w001;
      fp[0] += g.y * w100, fp[1] += g.y * w010, fp[2] += g.y * w001;
   }
   #pragma unroll
   for (int q = 0; q < 3; ++q) {
      #pragma unroll
      for (int o = LG / 2; o > 0; o >>= 1) {
         fd[q] += __shfl_xor_sync(0xffffffffu, fd[q], o);
         fp[q] += __shfl_xor_sync(0xffffffffu, fp[q], o);
      }
   }
   if (l == 0 && act) {
      V3 ud, up;
      load_dp(U, s, ud, up);
      V3 cd = v3(X.a[0][0] * fd[0] + X.a[0][1] * fd[1] + X.a[0][2] * fd[2], X.a[1][0] * fd[0] + X.a[1][1] * fd[1] + X.a[1][2] * fd[2],
         X.a[2][0] * fd[0] + X.a[2][1] * fd[1] + X.a[2][2] * fd[2]);
      V3 cp = v3(X.a[0][0] * fp[0] + X.a[0][1] * fp[1] + X.a[0][2] * fp[2], X.a[1][0] * fp[0] + X.a[1][1] * fp[1] + X.a[1][2] * fp[2],
         X.a[2][0] * fp[0] + X.a[2][1] * fp[1] + X.a[2][2] * fp[2]);
      V3 ed = selfterm * ud - cd, ep = selfterm * up - cp;
      if (F) {
         V3 a, b;
         load_dp(F, s, a, b);
         ed += a;
         ep += b;
      }
      if (EPI == 0) {
         out_d[3 * s] = ed.x, out_d[3 * s + 1] = ed.y, out_d[3 * s + 2] = ed.z;
         out_p[3 * s] = ep.x, out_p[3 * s + 1] = ep.y, out_p[3 * s + 2] = ep.z;
      } else if (EPI == 1) {
         if (tpj[s].y == 0) {
            ed = v3(0, 0, 0);
            ep = v3(0, 0, 0);
         }
         store_dp(OUT, s, ed, ep);
      } else {
         real pinv = tpj[s].z;
         V3 vd = pinv * ud - ed, vp = pinv * up - ep;
         store_dp(OUT, s, vd, vp);
         dot_d = (double)ud.x * vd.x + (double)ud.y * vd.y + (double)ud.z * vd.z;
         dot_p = (double)up.x * vp.x + (double)up.y * vp.y + (double)up.z * vp.z;
      }
   }
   if (EPI == 2)
      pcg_block_add2(dot_d, dot_p, pcg_slot_of(slot, itp), 2, 3);
}

// --- host side ---------------------------------------------------------------------------------
void bspline_host(double x, int n, double* c)   // tinker/source/pmestuf.f:134-158, 1-based c
{
   c[1] = 1.0 - x;
   c[2] = x;
   for (int k = 3; k <= n; ++k) {
      double denom = 1.0 / (k - 1);
      c[k] = x * c[k - 1] * denom;
      for (int i = 1; i <= k - 2; ++i)
         c[k - i] = ((x + i) * c[k - i - 1] + (k - i - x) * c[k - i]) * denom;
      c[1] = (1.0 - x) * c[1] * denom;
   }
}

std::vector<double> dftmod_host(int nfft, int order)   // pmestuf.f:172-237
{
   std::vector<double> c(order + 2, 0.0), bsarray(nfft, 0.0), mod(nfft, 0.0);
   bspline_host(0.0, order, c.data());
   for (int i = 0; i < order; ++i)
      bsarray[i + 1] = c[i + 1];
   const double pi = M_PI;
   double factor = 2.0 * pi / nfft;
   for (int i = 0; i < nfft; ++i) {
      double s1 = 0, s2 = 0;
      for (int j = 0; j < nfft; ++j) {
         double arg = factor * ((double)i * j);
         s1 += bsarray[j] * cos(arg);
         s2 += bsarray[j] * sin(arg);
      }
      mod[i] = s1 * s1 + s2 * s2;
   }
   const double eps = 1.0e-7;
   if (mod[0] < eps)
      mod[0] = 0.5 * mod[1];
   for (int i = 1; i < nfft - 1; ++i)
      if (mod[i] < eps)
         mod[i] = 0.5 * (mod[i - 1] + mod[i + 1]);
   if (mod[nfft - 1] < eps)
      mod[nfft - 1] = 0.5 * mod[nfft - 2];
   const int jcut = 50;
   for (int i = 1; i <= nfft; ++i) {
      int k = i - 1;
      if (i > nfft / 2)
         k -= nfft;
      double zeta = 1.0;
      if (k != 0) {
         double sum1 = 1.0, sum2 = 1.0;
         double fac = pi * k / nfft;
         for (int j = 1; j <= jcut; ++j) {
            double a1 = fac / (fac + pi * j), a2 = fac / (fac - pi * j);
            sum1 += pow(a1, order) + pow(a2, order);
            sum2 += pow(a1, 2 * order) + pow(a2, 2 * order);
         }
         zeta = sum2 / sum1;
      }
      mod[i - 1] *= zeta * zeta;
   }
   return mod;
}

Xform make_xform(apx_ctx* c)
{
   Xform X;
   double a[3][3];
   int nf[3] = {c->nfft1, c->nfft2, c->nfft3};
   for (int cc = 0; cc < 3; ++cc)
      for (int f = 0; f < 3; ++f)
         a[cc][f] = nf[f] * (double)c->box.r[3 * f + cc];
   const int qi1[6] = {0, 1, 2, 0, 0, 1}, qi2[6] = {0, 1, 2, 1, 2, 2};
   double ctf[6][6], ftc[6][6];
   for (int i1 = 0; i1 < 3; ++i1) {
      int k = qi1[i1];
      for (int i2 = 0; i2 < 6; ++i2)
         ctf[i2][i1] = a[qi1[i2]][k] * a[qi2[i2]][k];
   }
   for (int i1 = 3; i1 < 6; ++i1) {
      int k = qi1[i1], m = qi2[i1];
      for (int i2 = 0; i2 < 6; ++i2)
         ctf[i2][i1] = a[qi1[i2]][k] * a[qi2[i2]][m] + a[qi2[i2]][k] * a[qi1[i2]][m];
   }
   // frac -> cart uses at[f][c] = a[c][f]
   auto at = [&](int f, int cc) { return a[cc][f]; };
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
      for (int j = 0; j < 6; ++j) {
         X.ctf[i][j] = (real)ctf[i][j];
         X.ftc[i][j] = (real)ftc[i][j];
      }
   return X;
}

inline void fft(apx_ctx* c, int dir)
{
#ifdef APX_DOUBLE
   CUFFT_CHECK(cufftExecZ2Z(c->plan, c->qgrid, c->qgrid, dir));
#else
   CUFFT_CHECK(cufftExecC2C(c->plan, c->qgrid, c->qgrid, dir));
#endif
}
inline real selfterm(apx_ctx* c)
{
   double a = c->opt.aewald;
   return (real)(4.0 / 3.0 * a * a * a / sqrt(M_PI));
}
// elements of the grid this GPU holds in real space (all planes, or its slab + halo planes) and of
// the transformed slab it multiplies by the influence function
inline size_t nlocal(apx_ctx* c) { return (size_t)c->nfft1 * c->nfft2 * c->nzl; }
#ifndef APX_DOUBLE
// second-generation dipole spread / gather (pairs of x-points, splines evaluated in the kernel): APX_PME_GEN2=0 for the first
inline bool pme_gen2(const apx_ctx* c)
{
   static const int on = getenv("APX_PME_GEN2") ? atoi(getenv("APX_PME_GEN2")) : 1;
   return on && (c->nfft1 % 2) == 0;
}
#endif
inline size_t nconv(apx_ctx* c) { return (size_t)c->nfft1 * c->qny * c->nfft3; }
inline cplx* conv_grid(apx_ctx* c) { return c->dist.on ? c->dist.tbuf.p : c->qgrid.p; }
} // namespace

void apx_pme_setup(apx_ctx* c)
{
   if (!c->opt.use_ewald)
      return;
   c->nfft1 = c->opt.nfft[0];
   c->nfft2 = c->opt.nfft[1];
   c->nfft3 = c->opt.nfft[2];
   if (c->opt.bsorder != 5)
      APX_THROW("only pme-order 5 is built (the reference hard-codes MAX_BSORDER 5, include/seq/bsplgen.h)");
   if (c->dist.on) {
      apx_dist_pme_setup(c);      // sets zbase, nzl, qy0, qny, the slab-FFT plans and buffers
   } else {
      c->zbase = 0, c->nzl = c->nfft3;
      c->qy0 = 0, c->qny = c->nfft2;
      if (!c->plan_ok) {
#ifdef APX_DOUBLE
         CUFFT_CHECK(cufftPlan3d(&c->plan, c->nfft3, c->nfft2, c->nfft1, CUFFT_Z2Z));
#else
         CUFFT_CHECK(cufftPlan3d(&c->plan, c->nfft3, c->nfft2, c->nfft1, CUFFT_C2C));
#endif
         CUFFT_CHECK(cufftSetStream(c->plan, c->stream));
         c->plan_ok = 1;
      }
      apx_fft64_setup(c);
   }
   c->qgrid.ensure(nlocal(c));
   if (c->pme_fixed)
      apx_pme_fixed_setup(c);
   c->qfac.ensure(nconv(c));
   int nf[3] = {c->nfft1, c->nfft2, c->nfft3};
   DevBuf<real>* bs[3] = {&c->bsmod1, &c->bsmod2, &c->bsmod3};
   for (int d = 0; d < 3; ++d) {
      std::vector<double> m = dftmod_host(nf[d], 5);
      std::vector<real> mr(m.begin(), m.end());
      bs[d]->ensure(nf[d]);
      CUDA_CHECK(cudaMemcpyAsync(bs[d]->p, mr.data(), sizeof(real) * nf[d], cudaMemcpyHostToDevice, c->stream));
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
   }
   double pterm = (M_PI / c->opt.aewald) * (M_PI / c->opt.aewald);
   double volterm = M_PI * (double)c->box.volume;
   size_t K = nconv(c);
   k_make_qfac<<<(unsigned)((K + 255) / 256), 256, 0, c->stream>>>(c->nfft1, c->nfft2, c->nfft3, c->qy0, c->qny, c->box, (real)pterm,
      (real)volterm, c->bsmod1, c->bsmod2, c->bsmod3, c->qfac);
   APX_COUNT_LAUNCH(c);
}

void apx_pme_fixed_setup(apx_ctx* c)
{
   if (!c->opt.use_ewald)
      return;
   const size_t K = nlocal(c);
   if (c->qfix.cap < 2 * K) {
      c->qfix.ensure(2 * K);
      CUDA_CHECK(cudaMemsetAsync(c->qfix.p, 0, c->qfix.cap * sizeof(long long), c->stream));
   }
}

// spline tables of the current positions (owned atoms); called whenever posd changes (nblist.cu)
void apx_pme_fill_theta(apx_ctx* c)
{
   if (!c->opt.use_ewald)
      return;
   const int a0 = c->a0, no = c->a1 - c->a0;
   c->theta.ensure(16 * (size_t)c->npad);
   if (no > 0)
      k_theta_fill<<<(4 * no + 127) / 128, 128, 0, c->stream>>>(no, c->box, c->nfft1, c->nfft2, c->nfft3, c->posq + a0,
         c->theta + 16 * (size_t)a0);
   APX_COUNT_LAUNCH(c);
}

void apx_pme_destroy(apx_ctx* c)
{
   if (c->plan_ok)
      cufftDestroy(c->plan);
   c->plan_ok = 0;
   apx_dist_pme_destroy(c);
}

static void conv(apx_ctx* c, bool want_ev, double* out)
{
   size_t K = nconv(c);
   cplx* g = conv_grid(c);
   if (want_ev) {
      double pterm = (M_PI / c->opt.aewald) * (M_PI / c->opt.aewald);
      k_conv_ev<<<(unsigned)((K + 255) / 256), 256, 0, c->stream>>>(c->nfft1, c->nfft2, c->nfft3, c->qy0, c->qny, c->box, (real)pterm,
         c->f_elec, c->qfac, g, out);
   } else {
      k_conv<<<(unsigned)((K + 255) / 256), 256, 0, c->stream>>>(K, c->qfac, g);
   }
   APX_COUNT_LAUNCH(c);
}

// forward transform of the spread grid / inverse transform back to real space (with the halo planes
// of the slab decomposition when the grid is shared by several GPUs)
static void fft_forward(apx_ctx* c)
{
   if (c->dist.on)
      apx_dist_fft_forward(c, c->dist.tbuf);
   else
      fft(c, CUFFT_FORWARD);
}
static void fft_inverse(apx_ctx* c)
{
   if (c->dist.on)
      apx_dist_fft_inverse(c, c->dist.tbuf);
   else
      fft(c, CUFFT_INVERSE);
}

// fixed-point shadow grid of the deterministic spreading mode (nullptr: float reductions straight into qgrid)
static inline long long* pme_fix(apx_ctx* c) { return c->pme_fixed ? c->qfix.p : nullptr; }
// shadow grid -> qgrid (assigns every point of the local planes, zeroes the shadow grid for the next spread); the solver's
// kernels that return early once converged (skip flag) leave the shadow grid zero, so this stays idempotent
static inline void pme_fix_flush(apx_ctx* c)
{
   if (!c->pme_fixed)
      return;
   const size_t K = nlocal(c);
   k_fix_to_grid<<<(unsigned)((K + 255) / 256), 256, 0, c->stream>>>(K, c->qfix.p, c->qgrid.p);
   APX_COUNT_LAUNCH(c);
}

// permanent multipoles: fills fmp, fphi and ASSIGNS field = recip + self part of dfield.
// dbuf[16] = recip |Q|^2 energy, dbuf[17..22] = its virial (vir_m) when want_ev.
void apx_pme_mpole(apx_ctx* c, bool want_ev)
{
   const int a0 = c->a0, no = c->a1 - c->a0;
   Xform X = make_xform(c);
   CUDA_CHECK(cudaMemsetAsync(c->qgrid.p, 0, nlocal(c) * sizeof(cplx), c->stream));
   if (no > 0)
      k_spread_mpole<<<(no + 3) / 4, 128, 0, c->stream>>>(no, X, c->nfft1, c->nfft2, c->nfft3, c->zbase, c->nzl, c->theta + 16 * (size_t)a0,
         c->mp0 + a0, c->mp1 + a0, c->mp2 + a0, c->fmp + 10 * (size_t)a0, c->qgrid, pme_fix(c));
   APX_COUNT_LAUNCH(c);
   pme_fix_flush(c);
   fft_forward(c);
   if (want_ev)
      CUDA_CHECK(cudaMemsetAsync(c->dbuf.p + 16, 0, 7 * sizeof(double), c->stream));
   conv(c, want_ev, c->dbuf.p + 16);
   fft_inverse(c);
   if (no > 0)
      k_gather<0><<<(no + 3) / 4, 128, 0, c->stream>>>(no, X, c->nfft1, c->nfft2, c->nfft3, c->zbase, c->nzl, selfterm(c),
         c->theta + 16 * (size_t)a0, c->qgrid, c->mp0 + a0, nullptr, nullptr, c->fphi + 20 * (size_t)a0, c->field + 3 * (size_t)a0, nullptr,
         nullptr);
   APX_COUNT_LAUNCH(c);
   c->mpole_pme_valid = 1;
}

// ---- mutual-field operator on packed dipole pairs (dp.cuh) ----
void apx_pme_zero_grid(apx_ctx* c)
{
   CUDA_CHECK(cudaMemsetAsync(c->qgrid.p, 0, nlocal(c) * sizeof(cplx), c->stream));
}

// grid must be zero on entry
void apx_pme_spread_dp(apx_ctx* c, const real4* U)
{
   const int a0 = c->a0, no = c->a1 - c->a0;
   Xform X = make_xform(c);
#ifndef APX_DOUBLE
   if (no > 0 && pme_gen2(c)) {
      k_spread_dp2<PME_LG><<<(no + PME_APB - 1) / PME_APB, 128, 0, c->stream>>>(no, X, c->nfft1, c->nfft2, c->nfft3, c->zbase, c->nzl,
         c->posq + a0, U + 2 * (size_t)a0, c->qgrid, c->skip, pme_fix(c));
      APX_COUNT_LAUNCH(c);
      pme_fix_flush(c);
      return;
   }
#endif
   if (no > 0)
      k_spread_dp<PME_LG><<<(no + PME_APB - 1) / PME_APB, 128, 0, c->stream>>>(no, X, c->nfft1, c->nfft2, c->nfft3, c->zbase, c->nzl,
         c->theta + 16 * (size_t)a0, U + 2 * (size_t)a0, c->qgrid, c->skip, pme_fix(c));
   APX_COUNT_LAUNCH(c);
   pme_fix_flush(c);
}

// forward FFT, influence function, inverse FFT
void apx_pme_convolve(apx_ctx* c)
{
   if (!c->dist.on && apx_fft64_usable(c)) {
      apx_fft64_convolve(c);
      return;
   }
   fft_forward(c);
   conv(c, false, nullptr);
   fft_inverse(c);
}

// epi 0: fd/fp plain out ; 1: OUT = residual ; 2: OUT = Ap with partial dots into slot
void apx_pme_gather_dp(apx_ctx* c, int epi, const real4* U, const real4* F, real* fd, real* fp, real4* OUT, double* slot, const int* itp)
{
   const int a0 = c->a0, no = c->a1 - c->a0;
   Xform X = make_xform(c);
   int g = (no + PME_APB - 1) / PME_APB;
   if (g < 1)
      g = 1;
#define GATHER_DP(E)                                                                                                       \
   k_gather_dp<E, PME_LG><<<g, 128, 0, c->stream>>>(no, X, c->nfft1, c->nfft2, c->nfft3, c->zbase, c->nzl, selfterm(c),           \
      c->theta + 16 * (size_t)a0, c->tpj + a0, c->qgrid, U + 2 * (size_t)a0, F ? F + 2 * (size_t)a0 : nullptr,                 \
      fd ? fd + 3 * (size_t)a0 : nullptr, fp ? fp + 3 * (size_t)a0 : nullptr, OUT ? OUT + 2 * (size_t)a0 : nullptr, slot, c->skip, itp)
#ifndef APX_DOUBLE
#define GATHER_DP2(E)                                                                                                      \
   k_gather_dp2<E, PME_LG><<<g, 128, 0, c->stream>>>(no, X, c->nfft1, c->nfft2, c->nfft3, c->zbase, c->nzl, selfterm(c),          \
      c->posq + a0, c->tpj + a0, c->qgrid, U + 2 * (size_t)a0, F ? F + 2 * (size_t)a0 : nullptr,                               \
      fd ? fd + 3 * (size_t)a0 : nullptr, fp ? fp + 3 * (size_t)a0 : nullptr, OUT ? OUT + 2 * (size_t)a0 : nullptr, slot, c->skip, itp)
   if (no > 0 && pme_gen2(c)) {
      if (epi == 0) GATHER_DP2(0);
      else if (epi == 1) GATHER_DP2(1);
      else GATHER_DP2(2);
      APX_COUNT_LAUNCH(c);
      return;
   }
#undef GATHER_DP2
#endif
   if (no > 0) {
      if (epi == 0) GATHER_DP(0);
      else if (epi == 1) GATHER_DP(1);
      else GATHER_DP(2);
   }
#undef GATHER_DP
   APX_COUNT_LAUNCH(c);
}

// energy step: fphid, fphip (10 each) and fphidp (20) of the converged dipoles
void apx_pme_uind_fphi(apx_ctx* c, const real* ud, const real* up, bool)
{
   const int a0 = c->a0, no = c->a1 - c->a0;
   Xform X = make_xform(c);
   CUDA_CHECK(cudaMemsetAsync(c->qgrid.p, 0, nlocal(c) * sizeof(cplx), c->stream));
   apx_pack_dp(c, ud, up, c->pk_p);
   if (no > 0)
      k_spread_dp<PME_LG><<<(no + PME_APB - 1) / PME_APB, 128, 0, c->stream>>>(no, X, c->nfft1, c->nfft2, c->nfft3, c->zbase, c->nzl,
         c->theta + 16 * (size_t)a0, c->pk_p + 2 * (size_t)a0, c->qgrid, nullptr, pme_fix(c));
   APX_COUNT_LAUNCH(c);
   pme_fix_flush(c);
   apx_pme_convolve(c);
   if (no > 0)
      k_gather<2><<<(no + 3) / 4, 128, 0, c->stream>>>(no, X, c->nfft1, c->nfft2, c->nfft3, c->zbase, c->nzl, selfterm(c),
         c->theta + 16 * (size_t)a0, c->qgrid, nullptr, ud + 3 * (size_t)a0, up + 3 * (size_t)a0, c->fphid + 10 * (size_t)a0,
         c->fphip + 10 * (size_t)a0, c->fphidp + 20 * (size_t)a0, nullptr);
   APX_COUNT_LAUNCH(c);
}

// structure-factor product of the grids of (M + up) and (M + ud): two spreads + two forward FFTs
// (epolarEwaldRecipSelfVirial_cu3..5, src/cu/epolarrecip.cu:476-509)
void apx_pme_cross_virial(apx_ctx* c, real4* mpa, real4* mpb, double* out6)
{
   const int a0 = c->a0, no = c->a1 - c->a0;
   Xform X = make_xform(c);
   size_t K = nconv(c);
   DevBuf<cplx>& second = c->dist.on ? c->dist.tbuf2 : c->qgrid2;
   second.ensure(K);
   CUDA_CHECK(cudaMemsetAsync(c->qgrid.p, 0, nlocal(c) * sizeof(cplx), c->stream));
   if (no > 0)
      k_spread_mpole<<<(no + 3) / 4, 128, 0, c->stream>>>(no, X, c->nfft1, c->nfft2, c->nfft3, c->zbase, c->nzl, c->theta + 16 * (size_t)a0,
         mpa + a0, c->mp1 + a0, c->mp2 + a0, nullptr, c->qgrid, pme_fix(c));
   pme_fix_flush(c);
   fft_forward(c);
   CUDA_CHECK(cudaMemcpyAsync(second.p, conv_grid(c), K * sizeof(cplx), cudaMemcpyDeviceToDevice, c->stream));
   CUDA_CHECK(cudaMemsetAsync(c->qgrid.p, 0, nlocal(c) * sizeof(cplx), c->stream));
   if (no > 0)
      k_spread_mpole<<<(no + 3) / 4, 128, 0, c->stream>>>(no, X, c->nfft1, c->nfft2, c->nfft3, c->zbase, c->nzl, c->theta + 16 * (size_t)a0,
         mpb + a0, c->mp1 + a0, c->mp2 + a0, nullptr, c->qgrid, pme_fix(c));
   pme_fix_flush(c);
   fft_forward(c);
   double pterm = (M_PI / c->opt.aewald) * (M_PI / c->opt.aewald);
   k_cross_virial<<<(unsigned)((K + 255) / 256), 256, 0, c->stream>>>(c->nfft1, c->nfft2, c->nfft3, c->qy0, c->qny, c->box, (real)pterm,
      c->f_elec, c->qfac, conv_grid(c), second, out6);
   c->stats.kernel_launches += 3;
}
