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

