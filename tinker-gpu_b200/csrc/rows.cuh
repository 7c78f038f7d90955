// Lane-group iteration over the directed neighbor rows (rows.cu): G consecutive lanes share one
// atom, 128/G atoms per CTA, CTAs stride over the atoms.
#pragma once
#include "apx_internal.h"
#include "pairmath.cuh"

#define ROWS_BLOCK 128
// Row entries of the electrostatics list carry a flag in the sign bit: the pair is LISTED in the exclusion / scaling table
// (mdpuexclude).  The field kernels mask it off; the fused energy kernel skips flagged entries, because listed pairs are
// evaluated once, with their true scale factors, in double by the exclusion pass (mplar.cu).
#define ROW_INDEX_MASK 0x7fffffff
#define ROW_LISTED_FLAG 0x80000000u

// for (atoms a0 <= i < a1 of this lane group): `i` = atom (clamped to a1-1), `l` = lane in group,
// `act` = i is real.  [a0,a1) is the sorted range this GPU owns (the whole system on one GPU).
// The trip count is uniform across a warp so the body may use full-mask shuffles.
#define ROWS_FOREACH_ATOM(G, a0, a1, i, l, act)                                                                          \
   const int l = threadIdx.x & ((G) - 1);                                                                                \
   for (int i_ = (a0) + blockIdx.x * (ROWS_BLOCK / (G)) + threadIdx.x / (G), w_ = i_ - (threadIdx.x & 31) / (G),           \
            i = min(i_, (a1) - 1), act = i_ < (a1);                                                                      \
        w_ < (a1); i_ += gridDim.x * (ROWS_BLOCK / (G)), w_ += gridDim.x * (ROWS_BLOCK / (G)), i = min(i_, (a1) - 1), act = i_ < (a1))

template <int G>
__device__ __forceinline__ real group_sum(real v)
{
   #pragma unroll
   for (int o = G / 2; o > 0; o >>= 1)
      v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}
template <int G>
__device__ __forceinline__ V3 group_sum3(V3 v)
{
   return v3(group_sum<G>(v.x), group_sum<G>(v.y), group_sum<G>(v.z));
}

template <int G>
inline int rows_grid(const apx_ctx* c, int ctas_per_sm = 32)
{
   int per = ROWS_BLOCK / G;
   int want = (c->a1 - c->a0 + per - 1) / per;
   int cap = c->sm_count * ctas_per_sm;
   return want < 1 ? 1 : (want < cap ? want : cap);
}
