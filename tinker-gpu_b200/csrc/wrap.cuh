// Wrapping of caller-order f64 coordinates into the cell (shared by nblist.cu and ehal.cu).
#pragma once
#include "apx_internal.h"

// (wx,wy,wz) = position wrapped into the cell, (fx,fy,fz) = its fractional coordinates in [0,1)
__device__ __forceinline__ void wrap_pos(const Box& b, double x, double y, double z, real& wx, real& wy, real& wz, real& fx,
   real& fy, real& fz)
{
   double f1 = x * (double)b.r[0] + y * (double)b.r[1] + z * (double)b.r[2];
   double f2 = x * (double)b.r[3] + y * (double)b.r[4] + z * (double)b.r[5];
   double f3 = x * (double)b.r[6] + y * (double)b.r[7] + z * (double)b.r[8];
   f1 -= floor(f1);
   f2 -= floor(f2);
   f3 -= floor(f3);
   if (f1 >= 1.0) f1 = 0.0;
   if (f2 >= 1.0) f2 = 0.0;
   if (f3 >= 1.0) f3 = 0.0;
   fx = (real)f1;
   fy = (real)f2;
   fz = (real)f3;
   wx = (real)(f1 * (double)b.l[0] + f2 * (double)b.l[1] + f3 * (double)b.l[2]);
   wy = (real)(f1 * (double)b.l[3] + f2 * (double)b.l[4] + f3 * (double)b.l[5]);
   wz = (real)(f1 * (double)b.l[6] + f2 * (double)b.l[7] + f3 * (double)b.l[8]);
}

// the same with the cell in double, + the fractional coordinates as 32-bit integers: q = floor(f 2^32)
__device__ __forceinline__ void wrap_pos_q(const BoxD& b, double x, double y, double z, real& wx, real& wy, real& wz, unsigned& q1,
   unsigned& q2, unsigned& q3)
{
   double f1 = x * b.r[0] + y * b.r[1] + z * b.r[2];
   double f2 = x * b.r[3] + y * b.r[4] + z * b.r[5];
   double f3 = x * b.r[6] + y * b.r[7] + z * b.r[8];
   f1 -= floor(f1);
   f2 -= floor(f2);
   f3 -= floor(f3);
   if (f1 >= 1.0) f1 = 0.0;
   if (f2 >= 1.0) f2 = 0.0;
   if (f3 >= 1.0) f3 = 0.0;
   q1 = (unsigned)(unsigned long long)(f1 * 4294967296.0);
   q2 = (unsigned)(unsigned long long)(f2 * 4294967296.0);
   q3 = (unsigned)(unsigned long long)(f3 * 4294967296.0);
   wx = (real)(f1 * b.l[0] + f2 * b.l[1] + f3 * b.l[2]);
   wy = (real)(f1 * b.l[3] + f2 * b.l[4] + f3 * b.l[5]);
   wz = (real)(f1 * b.l[6] + f2 * b.l[7] + f3 * b.l[8]);
}
