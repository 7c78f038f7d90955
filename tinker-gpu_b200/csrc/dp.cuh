// Packed (d, p) vector pairs of the induced-dipole solver: atom s owns two real4,
//   v[2s] = (d.x, d.y, d.z, p.x),  v[2s+1] = (p.y, p.z, 0, 0),
// so that a neighbour's pair of 3-vectors is two 16-byte loads in the row kernels and the
// vector passes of the solver move whole atoms.  Sub-slotted double accumulators spread the
// grid-wide dot products over PCG_NSUB addresses to keep same-address L2 atomics short.
#pragma once
#include "apx_internal.h"
#include "pairmath.cuh"

#define PCG_NSUB 16
// per-iteration slot: quantity q (0,1 r.z entering ; 2,3 p.Ap ; 4,5 r.r) at [q*PCG_NSUB .. +PCG_NSUB)
#define PCG_NQ 6
#define PCG_SLOT (PCG_NQ * PCG_NSUB)

__device__ __forceinline__ void load_dp(const real4* __restrict__ v, int s, V3& d, V3& p)
{
   real4 a = v[2 * s], b = v[2 * s + 1];
   d = v3(a.x, a.y, a.z);
   p = v3(a.w, b.x, b.y);
}
__device__ __forceinline__ void store_dp(real4* __restrict__ v, int s, V3 d, V3 p)
{
   real4 a, b;
   a.x = d.x, a.y = d.y, a.z = d.z, a.w = p.x;
   b.x = p.y, b.y = p.z, b.z = 0, b.w = 0;
   v[2 * s] = a;
   v[2 * s + 1] = b;
}

// sum of the PCG_NSUB sub-slots of one quantity (every thread reads them; they sit in L2/L1)
__device__ __forceinline__ double pcg_q(const double* __restrict__ slot, int q)
{
   double s = 0;
   #pragma unroll
   for (int k = 0; k < PCG_NSUB; ++k)
      s += slot[q * PCG_NSUB + k];
   return s;
}

// The same sums for a whole CTA: the first warp adds the sub-slots of quantities q0..q0+NQ-1 and
// leaves them in shared memory.  Every thread of the CTA must call this (it synchronises).
template <int NQ>
__device__ __forceinline__ void pcg_q_block(const double* __restrict__ slot, int q0, double* out)
{
   __shared__ double sq_[NQ];
   if (threadIdx.x < 32) {
      #pragma unroll
      for (int j = 0; j < NQ; ++j) {
         double v = (int)threadIdx.x < PCG_NSUB ? slot[(q0 + j) * PCG_NSUB + threadIdx.x] : 0.0;
         #pragma unroll
         for (int o = PCG_NSUB / 2; o > 0; o >>= 1)
            v += __shfl_xor_sync(0xffffffffu, v, o);
         if (threadIdx.x == 0)
            sq_[j] = v;
      }
   }
   __syncthreads();
   #pragma unroll
   for (int j = 0; j < NQ; ++j)
      out[j] = sq_[j];
}

// block-wide sum of two doubles -> one atomic pair per CTA into sub-slot (blockIdx % PCG_NSUB)
__device__ __forceinline__ void pcg_block_add2(double a, double b, double* __restrict__ slot, int qa, int qb)
{
   __shared__ double sh_[2][32];
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
   #pragma unroll
   for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
   }
   if (lane == 0) {
      sh_[0][w] = a;
      sh_[1][w] = b;
   }
   __syncthreads();
   if (threadIdx.x == 0) {
      double x = 0, y = 0;
      const int nw = (blockDim.x + 31) >> 5;
      for (int k = 0; k < nw; ++k) {
         x += sh_[0][k];
         y += sh_[1][k];
      }
      const int sub = blockIdx.x % PCG_NSUB;
      atomicAdd(&slot[qa * PCG_NSUB + sub], x);
      atomicAdd(&slot[qb * PCG_NSUB + sub], y);
   }
}

// scattered correction of a packed pair (exclusion passes only)
__device__ __forceinline__ void atomic_dp(real4* __restrict__ v, int s, V3 d, V3 p)
{
   real* a = reinterpret_cast<real*>(v + 2 * s);
   atomicAdd(a + 0, d.x);
   atomicAdd(a + 1, d.y);
   atomicAdd(a + 2, d.z);
   atomicAdd(a + 3, p.x);
   atomicAdd(a + 4, p.y);
   atomicAdd(a + 5, p.z);
}
