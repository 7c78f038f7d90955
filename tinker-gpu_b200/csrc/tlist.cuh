// Entry format of the stored-tensor operator (tlist.cu): t = { B1 with the sign of B2 in its lowest mantissa bit, sqrt|B2| R },
// so that  T u = B2 (R.u) R - B1 u = sgn (S.u) S - B1 u.  Shared by the kernels that write entries (tlist.cu, field.cu).
#pragma once
#include "apx_internal.h"

// B1 with the sign of B2 in its lowest mantissa bit
__device__ __forceinline__ real tl_tag(real b1, bool neg)
{
#ifdef APX_DOUBLE
   long long u = __double_as_longlong(b1);
   u = (u & ~1ll) | (neg ? 1ll : 0ll);
   return __longlong_as_double(u);
#else
   unsigned u = __float_as_uint(b1);
   u = (u & ~1u) | (neg ? 1u : 0u);
   return __uint_as_float(u);
#endif
}
// v with its sign flipped when the tag bit of b1 is set
__device__ __forceinline__ real tl_signed(real v, real b1)
{
#ifdef APX_DOUBLE
   return __longlong_as_double(__double_as_longlong(v) ^ (__double_as_longlong(b1) << 63));
#else
   return __uint_as_float(__float_as_uint(v) ^ (__float_as_uint(b1) << 31));
#endif
}
__device__ __forceinline__ real4 tl_pack(real B1, real B2, real dx, real dy, real dz)
{
   const real s = sqrt(fabs(B2));
   real4 t;
   t.x = tl_tag(B1, B2 < 0);
   t.y = s * dx, t.z = s * dy, t.w = s * dz;
   return t;
}


#include "pairmath.cuh"
// load of a tensor entry: read once per application, but read again by the NEXT application -- at dhfr2 size the whole
// tensor list (53 MB) stays in L2 between the 8 applications of an induce(), so no evict-first hint (ld.global.cs)
__device__ __forceinline__ real4 tl_ld(const real4* p)
{
#ifdef APX_DOUBLE
   return *p;
#else
   return __ldg(p);
#endif
}
// ---- apply: F_i = sum_k T_ik (ud_k, up_k) ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tl_apply(const real4 t, const real4 ua, const real4 ub, V3& fd, V3& fp)
{
   const real sd = tl_signed(t.y * ua.x + t.z * ua.y + t.w * ua.z, t.x);
   const real sp = tl_signed(t.y * ua.w + t.z * ub.x + t.w * ub.y, t.x);
   fd.x += sd * t.y - t.x * ua.x;
   fd.y += sd * t.z - t.x * ua.y;
   fd.z += sd * t.w - t.x * ua.z;
   fp.x += sp * t.y - t.x * ua.w;
   fp.y += sp * t.z - t.x * ub.x;
   fp.z += sp * t.w - t.x * ub.y;
}

// the neighbour's packed dipole pair (32 bytes, 32-byte aligned) in ONE 256-bit load (LDG.E.256, new with sm_100): the gathers
// are what bounds this kernel at dhfr2 size -- every lane of a gather touches another cache line and L1 looks up one line per
// cycle, so two 128-bit gathers per entry cost 64 cycles per warp row (22 us per launch at 160 atoms per SM), one costs 32
__device__ __forceinline__ void tl_gather(const real4* __restrict__ U, int k, real4& ua, real4& ub)
{
#ifdef APX_DOUBLE
   ua = U[2 * k];
   ub = U[2 * k + 1];
#else
   asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
       : "=f"(ua.x), "=f"(ua.y), "=f"(ua.z), "=f"(ua.w), "=f"(ub.x), "=f"(ub.y), "=f"(ub.z), "=f"(ub.w)
       : "l"(U + 2 * (size_t)k));
#endif
}

