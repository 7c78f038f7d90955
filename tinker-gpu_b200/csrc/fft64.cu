// Fused 64^3 complex FFT -> influence function -> inverse FFT for the PME convolution.
//
// At 64^3 (dhfr2) the grid is 2 MB and lives in L2; cuFFT runs a 3-D C2C transform as three
// launches, so one convolution costs 7 launches of ~3 us each, almost all of it launch and drain
// latency.  Here the round trip is 3 launches:
//
//   k_fft64_xy<-1>   one CTA per z-plane: 64 x-FFTs then 64 y-FFTs through shared memory
//   k_fft64_z_conv   one CTA per y: z-FFT, multiply by the tabulated influence function,
//                    inverse z-FFT -- the grid is read and written once instead of three times
//   k_fft64_xy<+1>   inverse x/y transforms
//
// A 64-point transform is two radix-8 passes (64 = 8 x 8), 8 points per thread in registers,
// one exchange through shared memory between the passes.  Conventions are cuFFT's / the
// reference's (src/cudart/fft.cpp:54-57): forward = exp(-2 pi i jk/N), both directions unnormalised.
// Other grid sizes, and the double-precision build, go through cuFFT (pme.cu).
#include "apx_internal.h"
#include <cmath>

#ifndef APX_DOUBLE
namespace {
__constant__ float2 c_w64[64];      // exp(+2 pi i t / 64)

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// a * exp(S * i * theta) with w = exp(+i theta)
template <int S>
__device__ __forceinline__ float2 cmulw(float2 a, float2 w)
{
   return make_float2(a.x * w.x - (float)S * a.y * w.y, (float)S * a.x * w.y + a.y * w.x);
}
// a * exp(S i pi/2)
template <int S>
__device__ __forceinline__ float2 cmuli(float2 a)
{
   return S > 0 ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

// in-register 8-point DFT, X[m] = sum_k a[k] exp(S 2 pi i k m / 8), natural order in and out
template <int S>
__device__ __forceinline__ void fft8(float2* a)
{
   const float h = 0.70710678118654752f;
   float2 t;
   // span 4
   t = csub(a[0], a[4]); a[0] = cadd(a[0], a[4]); a[4] = t;
   t = csub(a[1], a[5]); a[1] = cadd(a[1], a[5]); a[5] = make_float2(h * (t.x - (float)S * t.y), h * ((float)S * t.x + t.y));
   t = csub(a[2], a[6]); a[2] = cadd(a[2], a[6]); a[6] = cmuli<S>(t);
   t = csub(a[3], a[7]); a[3] = cadd(a[3], a[7]); a[7] = make_float2(h * (-t.x - (float)S * t.y), h * ((float)S * t.x - t.y));
   // span 2
   t = csub(a[0], a[2]); a[0] = cadd(a[0], a[2]); a[2] = t;
   t = csub(a[1], a[3]); a[1] = cadd(a[1], a[3]); a[3] = cmuli<S>(t);
   t = csub(a[4], a[6]); a[4] = cadd(a[4], a[6]); a[6] = t;
   t = csub(a[5], a[7]); a[5] = cadd(a[5], a[7]); a[7] = cmuli<S>(t);
   // span 1
   t = csub(a[0], a[1]); a[0] = cadd(a[0], a[1]); a[1] = t;
   t = csub(a[2], a[3]); a[2] = cadd(a[2], a[3]); a[3] = t;
   t = csub(a[4], a[5]); a[4] = cadd(a[4], a[5]); a[5] = t;
   t = csub(a[6], a[7]); a[6] = cadd(a[6], a[7]); a[7] = t;
   // bit reversal: a1<->a4 , a3<->a6
   t = a[1]; a[1] = a[4]; a[4] = t;
   t = a[3]; a[3] = a[6]; a[6] = t;
}

// First radix-8 pass of a 64-point transform held as a[k] = x[8k + j]: after it a[q] = W64^{jq} Y_j[q].
template <int S>
__device__ __forceinline__ void pass1(float2* a, int j)
{
   fft8<S>(a);
   #pragma unroll
   for (int q = 1; q < 8; ++q)
      a[q] = cmulw<S>(a[q], c_w64[(j * q) & 63]);
}

#define PL 65      // padded row length of the shared plane (float2)

// one CTA (512 threads) per z-plane: x transforms, then y transforms
template <int S>
__global__ void __launch_bounds__(512) k_fft64_xy(float2* __restrict__ grid)
{
   __shared__ float2 pl[64 * PL];
   float2* g = grid + (size_t)blockIdx.x * 4096;
   const int tid = threadIdx.x;
   float2 a[8];
   {
      // rows: 8 consecutive lanes share a row; lane j holds x = 8k + j
      const int row = tid >> 3, j = tid & 7;
      #pragma unroll
      for (int k = 0; k < 8; ++k)
         a[k] = g[row * 64 + 8 * k + j];
      pass1<S>(a, j);
      #pragma unroll
      for (int q = 0; q < 8; ++q)
         pl[row * PL + j * 8 + q] = a[q];
      __syncwarp();
      #pragma unroll
      for (int jj = 0; jj < 8; ++jj)
         a[jj] = pl[row * PL + jj * 8 + j];
      fft8<S>(a);
      __syncwarp();
      #pragma unroll
      for (int p = 0; p < 8; ++p)
         pl[row * PL + j + 8 * p] = a[p];
   }
   __syncthreads();
   {
      // columns: consecutive lanes hold consecutive x; thread (c, j) holds y = 8k + j
      const int c = tid & 63, j = tid >> 6;
      #pragma unroll
      for (int k = 0; k < 8; ++k)
         a[k] = pl[(8 * k + j) * PL + c];
      pass1<S>(a, j);
      __syncthreads();
      #pragma unroll
      for (int q = 0; q < 8; ++q)
         pl[(j * 8 + q) * PL + c] = a[q];
      __syncthreads();
      #pragma unroll
      for (int jj = 0; jj < 8; ++jj)
         a[jj] = pl[(jj * 8 + j) * PL + c];
      fft8<S>(a);
      #pragma unroll
      for (int p = 0; p < 8; ++p)
         g[(j + 8 * p) * 64 + c] = a[p];
   }
}

// one CTA (512 threads) per y: forward z transform, influence function, inverse z transform
__global__ void __launch_bounds__(512) k_fft64_z_conv(float2* __restrict__ grid, const float* __restrict__ qfac)
{
   __shared__ float2 sl[64 * PL];
   const int y = blockIdx.x;
   const int tid = threadIdx.x;
   const int c = tid & 63, j = tid >> 6;
   float2* g = grid + (size_t)y * 64 + c;
   const float* qf = qfac + (size_t)y * 64 + c;
   float2 a[8];
   #pragma unroll
   for (int k = 0; k < 8; ++k)
      a[k] = g[(size_t)(8 * k + j) * 4096];
   pass1<-1>(a, j);
   #pragma unroll
   for (int q = 0; q < 8; ++q)
      sl[(j * 8 + q) * PL + c] = a[q];
   __syncthreads();
   #pragma unroll
   for (int jj = 0; jj < 8; ++jj)
      a[jj] = sl[(jj * 8 + j) * PL + c];
   fft8<-1>(a);
   // a[p] is the coefficient at z = j + 8p: exactly the input layout of the next transform
   #pragma unroll
   for (int p = 0; p < 8; ++p) {
      float f = qf[(size_t)(j + 8 * p) * 4096];
      a[p].x *= f;
      a[p].y *= f;
   }
   pass1<1>(a, j);
   __syncthreads();
   #pragma unroll
   for (int q = 0; q < 8; ++q)
      sl[(j * 8 + q) * PL + c] = a[q];
   __syncthreads();
   #pragma unroll
   for (int jj = 0; jj < 8; ++jj)
      a[jj] = sl[(jj * 8 + j) * PL + c];
   fft8<1>(a);
   #pragma unroll
   for (int p = 0; p < 8; ++p)
      g[(size_t)(j + 8 * p) * 4096] = a[p];
}
} // namespace

bool apx_fft64_usable(const apx_ctx* c)
{
   return c->native_fft && c->nfft1 == 64 && c->nfft2 == 64 && c->nfft3 == 64;
}

void apx_fft64_setup(apx_ctx* c)
{
   float2 w[64];
   for (int t = 0; t < 64; ++t) {
      double th = 2.0 * M_PI * t / 64.0;
      w[t] = make_float2((float)cos(th), (float)sin(th));
   }
   CUDA_CHECK(cudaMemcpyToSymbolAsync(c_w64, w, sizeof(w), 0, cudaMemcpyHostToDevice, c->stream));
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

// grid <- IFFT( qfac * FFT(grid) ), unnormalised
void apx_fft64_convolve(apx_ctx* c)
{
   float2* g = reinterpret_cast<float2*>(c->qgrid.p);
   k_fft64_xy<-1><<<64, 512, 0, c->stream>>>(g);
   k_fft64_z_conv<<<64, 512, 0, c->stream>>>(g, c->qfac);
   k_fft64_xy<1><<<64, 512, 0, c->stream>>>(g);
   c->stats.kernel_launches += 3;
}
#else
bool apx_fft64_usable(const apx_ctx*) { return false; }
void apx_fft64_setup(apx_ctx*) {}
void apx_fft64_convolve(apx_ctx*) {}
#endif
