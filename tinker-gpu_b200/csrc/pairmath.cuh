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
//
// The formulas live in pairmath_body.inc and are instantiated twice: in the global namespace for the build's own `real`
// (float in the mixed build), and in namespace pm64 for double -- the mixed build evaluates the few listed (excluded /
// scaled, i.e. bonded-range) pairs in double, where float would lose the north-star force tolerance to cancellation
// (DESIGN.md section 8).  Compiled by g++ as plain host code too (tests/pairmath_host.cpp): the error budget of the mixed
// build is checked on the CPU against the oracle.
#pragma once
#include "apx_internal.h"

#ifdef __CUDACC__
#define PM_FN __device__ __forceinline__
#else
#define PM_FN inline
#endif

#ifdef APX_DOUBLE
#define PM_DOUBLE 1
#else
#define PM_DOUBLE 0
#endif
#include "pairmath_body.inc"
#undef PM_DOUBLE

#ifdef __CUDACC__
// ---- pair separation from the per-step position records (pos_t, apx_internal.h) ------------------------------------------
// mixed build: integer difference of 32-bit fractional coordinates (wraps = minimum image), float only after the subtraction
__device__ __forceinline__ void pair_delta(const Box& b, const pos_t& pi, const pos_t& pk, real& dx, real& dy, real& dz)
{
#ifdef APX_DOUBLE
   dx = pk.x - pi.x, dy = pk.y - pi.y, dz = pk.z - pi.z;
   apx_image(b, dx, dy, dz);
#else
   const real f1 = (real)(int)(pk.x - pi.x), f2 = (real)(int)(pk.y - pi.y), f3 = (real)(int)(pk.z - pi.z);
   // l / 2^32 in two floats: without the low part every separation would carry the same relative rounding error of the
   // cell edge (up to 6e-8) -- a coherent error, the cell 4e-6 A too large, not a random one
   if (b.orthogonal) {
      dx = fmaf(f1, b.q[0], f1 * b.qlo[0]);
      dy = fmaf(f2, b.q[4], f2 * b.qlo[4]);
      dz = fmaf(f3, b.q[8], f3 * b.qlo[8]);
   } else {
      dx = f1 * b.q[0] + f2 * b.q[1] + f3 * b.q[2] + (f1 * b.qlo[0] + f2 * b.qlo[1] + f3 * b.qlo[2]);
      dy = f1 * b.q[3] + f2 * b.q[4] + f3 * b.q[5] + (f1 * b.qlo[3] + f2 * b.qlo[4] + f3 * b.qlo[5]);
      dz = f1 * b.q[6] + f2 * b.q[7] + f3 * b.q[8] + (f1 * b.qlo[6] + f2 * b.qlo[7] + f3 * b.qlo[8]);
   }
#endif
}
// Thole damping length carried in the record
__device__ __forceinline__ real pos_w(const pos_t& p)
{
#ifdef APX_DOUBLE
   return p.w;
#else
   return __uint_as_float(p.w);
#endif
}
// a record stored in / read from a real4 slot (the interleaved neighbour records of field.cu)
__device__ __forceinline__ real4 pos_as_real4(const pos_t& p)
{
#ifdef APX_DOUBLE
   return p;
#else
   return make_float4(__uint_as_float(p.x), __uint_as_float(p.y), __uint_as_float(p.z), __uint_as_float(p.w));
#endif
}
__device__ __forceinline__ pos_t real4_as_pos(const real4& v)
{
#ifdef APX_DOUBLE
   return v;
#else
   return make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
#endif
}
#endif

namespace pm64 {
typedef double real;
#define PM_DOUBLE 1
#include "pairmath_body.inc"
#undef PM_DOUBLE
}
