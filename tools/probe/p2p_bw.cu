// Probe: NVLink bandwidth of SM copy kernels between two B200s, push (remote stores) against pull (remote loads), by grid size;
// both directions at once, like an exchange.  One process, peer access.  nvcc -O3 -arch=sm_100a -o p2p_bw p2p_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int U>
__global__ void __launch_bounds__(256) k_copy(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16)
{
   const size_t per = (n16 + gridDim.x - 1) / gridDim.x;
   const size_t b0 = per * blockIdx.x, b1 = b0 + per < n16 ? b0 + per : n16;
   size_t q = b0 + threadIdx.x;
   const unsigned bd = blockDim.x;
   for (; q + (U - 1) * bd < b1; q += U * bd) {
      uint4 v[U];
      #pragma unroll
      for (int u = 0; u < U; ++u)
         v[u] = __ldcg(src + q + u * bd);
      #pragma unroll
      for (int u = 0; u < U; ++u)
         dst[q + u * bd] = v[u];
   }
   for (; q < b1; q += bd)
      dst[q] = __ldcg(src + q);
}

// TMA bulk copies through shared memory: one thread per CTA moves 16 KB tiles, 4 in flight
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
#define TILE 16384
#define STAGES 4
__global__ void __launch_bounds__(32) k_bulk(const char* __restrict__ src, char* __restrict__ dst, size_t bytes)
{
   extern __shared__ __align__(128) char sm[];
   __shared__ unsigned long long bar[STAGES];
   const size_t ntile = bytes / TILE;
   if (threadIdx.x == 0) {
      for (int s = 0; s < STAGES; ++s)
         asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      unsigned phase[STAGES] = {0, 0, 0, 0};
      size_t t = blockIdx.x;
      // prime
      size_t issued = t;
      int k = 0;
      for (; k < STAGES && issued < ntile; ++k, issued += gridDim.x) {
         asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[k])), "r"((unsigned)TILE) : "memory");
         asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + k * TILE)),
            "l"(src + issued * TILE), "r"((unsigned)TILE), "r"(smem_u32(&bar[k])) : "memory");
      }
      int s = 0;
      for (; t < ntile; t += gridDim.x) {
         unsigned done = 0;
         while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar[s])), "r"(phase[s]) : "memory");
         phase[s] ^= 1;
         asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + t * TILE), "r"(smem_u32(sm + s * TILE)), "r"((unsigned)TILE) : "memory");
         asm volatile("cp.async.bulk.commit_group;" ::: "memory");
         if (issued < ntile) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the store has read the stage
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"((unsigned)TILE) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + s * TILE)),
               "l"(src + issued * TILE), "r"((unsigned)TILE), "r"(smem_u32(&bar[s])) : "memory");
            issued += gridDim.x;
         }
         s = (s + 1) % STAGES;
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
   }
}

int main(int argc, char** argv)
{
   int nd = 0;
   CK(cudaGetDeviceCount(&nd));
   if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
   const size_t bytes = (size_t)(argc > 1 ? atoi(argv[1]) : 18) << 20;
   char* buf[2][2];
   cudaStream_t st[2];
   cudaEvent_t e0[2], e1[2];
   for (int d = 0; d < 2; ++d) {
      CK(cudaSetDevice(d));
      CK(cudaDeviceEnablePeerAccess(1 - d, 0));
      CK(cudaMalloc(&buf[d][0], bytes));
      CK(cudaMalloc(&buf[d][1], bytes));
      CK(cudaMemset(buf[d][0], d + 1, bytes));
      CK(cudaStreamCreate(&st[d]));
      CK(cudaEventCreate(&e0[d]));
      CK(cudaEventCreate(&e1[d]));
      CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * TILE));
   }
   const size_t n16 = bytes / 16;
   auto run = [&](const char* name, int mode, int grid, bool both) {
      // mode 0: push U=4, 1: pull U=4, 2: push U=8, 3: pull U=8, 4: bulk push, 5: bulk pull, 6: memcpy peer
      float best[2] = {1e9f, 1e9f};
      for (int rep = 0; rep < 6; ++rep) {
         for (int d = 0; d < (both ? 2 : 1); ++d) {
            CK(cudaSetDevice(d));
            const bool pull = mode == 1 || mode == 3 || mode == 5;
            const char* src = pull ? buf[1 - d][0] : buf[d][0];
            char* dst = pull ? buf[d][1] : buf[1 - d][1];
            CK(cudaEventRecord(e0[d], st[d]));
            if (mode == 0 || mode == 1)
               k_copy<4><<<grid, 256, 0, st[d]>>>((const uint4*)src, (uint4*)dst, n16);
            else if (mode == 2 || mode == 3)
               k_copy<8><<<grid, 256, 0, st[d]>>>((const uint4*)src, (uint4*)dst, n16);
            else if (mode == 4 || mode == 5)
               k_bulk<<<grid, 32, STAGES * TILE, st[d]>>>(src, dst, bytes);
            else
               CK(cudaMemcpyPeerAsync(buf[1 - d][1], 1 - d, buf[d][0], d, bytes, st[d]));
            CK(cudaEventRecord(e1[d], st[d]));
         }
         for (int d = 0; d < (both ? 2 : 1); ++d) {
            CK(cudaSetDevice(d));
            CK(cudaStreamSynchronize(st[d]));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0[d], e1[d]));
            if (rep > 0 && ms < best[d]) best[d] = ms;
         }
      }
      printf("%-14s grid %4d %s: %7.1f us  %6.1f GB/s", name, grid, both ? "both" : "one ", best[0] * 1e3, bytes / best[0] * 1e-6);
      if (both) printf("   | dev1 %7.1f us %6.1f GB/s", best[1] * 1e3, bytes / best[1] * 1e-6);
      printf("\n");
      fflush(stdout);
   };
   printf("message %zu MB\n", bytes >> 20);
   run("memcpyPeer", 6, 0, false);
   run("memcpyPeer", 6, 0, true);
   const int grids[] = {16, 32, 64, 148, 296, 592};
   const char* names[] = {"push U4", "pull U4", "push U8", "pull U8", "bulk push", "bulk pull"};
   for (int mode = 0; mode < 6; ++mode)
      for (int g : grids) {
         if (mode >= 4 && g > 296) continue;
         run(names[mode], mode, g, false);
         run(names[mode], mode, g, true);
      }
   // verify one bulk result
   CK(cudaSetDevice(0));
   unsigned char h[4];
   CK(cudaMemcpy(h, buf[1][1] + bytes - 4, 4, cudaMemcpyDefault));
   printf("check byte %d (expect 1 or 2)\n", h[0]);
   return 0;
}
