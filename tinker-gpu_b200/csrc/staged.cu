// Staged real-space CG operator: 64-atom groups whose neighbour records are staged in shared memory by bulk copies.
//
// The row kernels of field.cu gather 48 bytes per directed pair through L1: ncu shows them bound by L1 wavefronts at the
// 1 M-atom size (l1tex 86 % of peak, issue slots 54 %; profiles/r01j_water1m_ncu_full_summary.txt), i.e. the FP32 pipe waits
// for the load path.  Here a CTA owns GROUP = 64 consecutive sorted atoms (two Morton blocks, a compact region of space):
//
//   at a list rebuild   k_group_build: the union of the 32-atom blocks its Verlet rows reach ("Verlet j-blocks", ascending), and
//                       the rows re-expressed as 16-bit SLOTS = rank of the block in that list x 32 + lane;
//   every step          k_rows_compact_grp (replaces k_rows_compact): rows cut to the pairs inside the cutoff as before, plus
//                       the ACTIVE j-blocks of the group (those that still hold a partner) and the rows in active-slot form;
//   every operator      k_ufield_staged: ONE elected thread arms an mbarrier with the byte count and issues cp.async.bulk
//                       copies (1-D TMA, one per run of adjacent active blocks, 1.5 KB per block) of the interleaved records
//                       into shared memory; the 256 threads (4 lanes per atom) then walk the rows with 16-bit indices and
//                       three LDS.128 per pair instead of one LDG.32 + three LDG.128 gathers.  Every record is fetched from
//                       L2 once per GROUP (~45 blocks x 1.5 KB for 64 x 280 pairs) instead of once per pair.
//
// Same pair math, same row order and the same 4-lane shuffle reduction per atom as the row kernel: results agree to rounding.
// Used for single-GPU contexts without a per-pair Thole table; everything else keeps the row kernels (field.cu).
#include "apx_internal.h"
#include "pairmath.cuh"
#include "rows.cuh"
#include "dp.cuh"
#include "staged.cuh"
#include <algorithm>

#define FULL 0xffffffffu

namespace {
// ---------------------------------------------------------------------------------------------------------------------------
// list rebuild: Verlet j-blocks of every group + Verlet rows in slot form
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_group_build(int n, int nblk, int nw, const int* __restrict__ vstart, const int* __restrict__ vnbr,
   int* __restrict__ vjb, int* __restrict__ nvjb, unsigned short* __restrict__ vslot, int* __restrict__ oflow)
{
   extern __shared__ unsigned sm_gb[];
   unsigned* bits = sm_gb;            // [nw] one bit per 32-atom block
   unsigned* pre = sm_gb + nw;        // [nw] set bits before this word
   __shared__ int s_part[256];
   const int g = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
   const int i0 = g * SG_GROUP, i1 = min(n, i0 + SG_GROUP);
   for (int q = t; q < nw; q += 256)
      bits[q] = 0;
   __syncthreads();
   for (int i = i0 + w; i < i1; i += 8) {
      const int beg = vstart[i], end = vstart[i + 1];
      for (int q = beg + lane; q < end; q += 32) {
         const int kb = (vnbr[q] & ROW_INDEX_MASK) >> 5;
         atomicOr(&bits[kb >> 5], 1u << (kb & 31));
      }
   }
   __syncthreads();
   // exclusive prefix of the per-word popcounts: 256 chunks
   const int per = (nw + 255) / 256;
   int mine = 0;
   for (int q = t * per; q < min(nw, (t + 1) * per); ++q)
      mine += __popc(bits[q]);
   s_part[t] = mine;
   __syncthreads();
   if (t < 32) {      // one warp scans the 256 partials, 8 per lane
      int v[8], s = 0;
      #pragma unroll
      for (int k = 0; k < 8; ++k) {
         v[k] = s_part[8 * t + k];
         s += v[k];
      }
      int x = s;
      #pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
         int y = __shfl_up_sync(FULL, x, o);
         if (lane >= o)
            x += y;
      }
      int run = x - s;
      #pragma unroll
      for (int k = 0; k < 8; ++k) {
         s_part[8 * t + k] = run;
         run += v[k];
      }
   }
   __syncthreads();
   {
      int run = s_part[t];
      for (int q = t * per; q < min(nw, (t + 1) * per); ++q) {
         pre[q] = run;
         run += __popc(bits[q]);
      }
   }
   __syncthreads();
   const int total = pre[nw - 1] + __popc(bits[nw - 1]);
   if (t == 0) {
      nvjb[g] = min(total, SG_VJB_CAP);
      if (total > SG_VJB_CAP)
         *oflow = 1;
   }
   for (int q = t; q < nw; q += 256) {
      unsigned m = bits[q];
      while (m) {
         const int b = __ffs(m) - 1;
         m &= m - 1;
         const int rank = pre[q] + __popc(bits[q] & ((1u << b) - 1));
         if (rank < SG_VJB_CAP)
            vjb[(size_t)g * SG_VJB_CAP + rank] = q * 32 + b;
      }
   }
   for (int i = i0 + w; i < i1; i += 8) {
      const int beg = vstart[i], end = vstart[i + 1];
      for (int q = beg + lane; q < end; q += 32) {
         const int kraw = vnbr[q];
         const int k = kraw & ROW_INDEX_MASK, kb = k >> 5;
         const int rank = pre[kb >> 5] + __popc(bits[kb >> 5] & ((1u << (kb & 31)) - 1));
         vslot[q] = (unsigned short)(((min(rank, SG_VJB_CAP - 1) << 5) | (k & 31)) | (kraw < 0 ? SG_LISTED : 0));
      }
   }
}

// ---------------------------------------------------------------------------------------------------------------------------
// every step: rows inside the cutoff (preconditioner range first), active j-blocks, rows in active-slot form
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rows_compact_grp(int n, Box b, real cut2, real ucut2, const real4* __restrict__ posd,
   const int* __restrict__ vstart, const unsigned short* __restrict__ vslot, const int* __restrict__ vjb, const int* __restrict__ nvjb,
   int* __restrict__ nbr, unsigned short* __restrict__ nbr16, int* __restrict__ cnt, int* __restrict__ cntu, int* __restrict__ ajb,
   int* __restrict__ najb, unsigned long long* __restrict__ total)
{
   __shared__ int s_vjb[SG_VJB_CAP];
   __shared__ unsigned s_act[SG_VJB_CAP / 32];
   __shared__ unsigned short s_remap[SG_VJB_CAP];
   const int g = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
   const int i0 = g * SG_GROUP, i1 = min(n, i0 + SG_GROUP);
   const int nv = nvjb[g];
   for (int q = t; q < nv; q += 256)
      s_vjb[q] = vjb[(size_t)g * SG_VJB_CAP + q];
   if (t < SG_VJB_CAP / 32)
      s_act[t] = 0;
   __syncthreads();
   const unsigned lt = (1u << lane) - 1;
   for (int i = i0 + w; i < i1; i += 8) {
      const int beg = vstart[i], end = vstart[i + 1];
      const real4 pi = posd[i];
      int out = beg, nu = 0;
      for (int sweep = 0; sweep < 2; ++sweep) {
         if (sweep == 0 && ucut2 <= 0)
            continue;
         for (int q0 = beg; q0 < end; q0 += 32) {
            const int q = q0 + lane;
            const unsigned s = q < end ? vslot[q] : 0;
            const int slot = s & SG_SLOT_MASK;
            const int k = s_vjb[slot >> 5] * 32 + (slot & 31);
            bool ok = false;
            if (q < end) {
               const real4 pk = posd[k];
               real dx = pk.x - pi.x, dy = pk.y - pi.y, dz = pk.z - pi.z;
               apx_image(b, dx, dy, dz);
               const real r2 = dx * dx + dy * dy + dz * dz;
               ok = sweep == 0 ? r2 <= ucut2 : (r2 > ucut2 && r2 <= cut2);
            }
            const unsigned m = __ballot_sync(FULL, ok);
            if (ok) {
               const int o = out + __popc(m & lt);
               nbr[o] = (s & SG_LISTED) ? (int)((unsigned)k | ROW_LISTED_FLAG) : k;
               nbr16[o] = (unsigned short)s;
               atomicOr(&s_act[slot >> 10], 1u << ((slot >> 5) & 31));
            }
            out += __popc(m);
         }
         if (sweep == 0)
            nu = out - beg;
      }
      if (lane == 0) {
         cnt[i] = out - beg;
         cntu[i] = nu;
         if (total) {
            atomicAdd(&total[0], (unsigned long long)(out - beg));
            atomicAdd(&total[1], (unsigned long long)nu);
         }
      }
   }
   __syncthreads();
   if (t < SG_VJB_CAP) {
      int r = 0;
      #pragma unroll
      for (int q = 0; q < SG_VJB_CAP / 32; ++q)
         if (q < (t >> 5))
            r += __popc(s_act[q]);
      r += __popc(s_act[t >> 5] & ((1u << (t & 31)) - 1));
      const bool act = t < nv && ((s_act[t >> 5] >> (t & 31)) & 1u);
      s_remap[t] = (unsigned short)r;
      if (act)
         ajb[(size_t)g * SG_VJB_CAP + r] = s_vjb[t];
      if (t == SG_VJB_CAP - 1)
         najb[g] = r + (act ? 1 : 0);
   }
   __syncthreads();
   for (int i = i0 + w; i < i1; i += 8) {
      const int beg = vstart[i], len = cnt[i];
      for (int q = beg + lane; q < beg + len; q += 32) {
         const unsigned s = nbr16[q];
         const int slot = s & SG_SLOT_MASK;
         nbr16[q] = (unsigned short)((s & SG_LISTED) | (s_remap[slot >> 5] << 5) | (slot & 31));
      }
   }
}

// ---------------------------------------------------------------------------------------------------------------------------
// the operator
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <bool EWALD>
__global__ void __launch_bounds__(SG_THREADS, 2) k_ufield_staged(int n, int cap, Box box, real aewald, const int* __restrict__ vstart,
   const int* __restrict__ cnt, const unsigned short* __restrict__ nbr16, const int* __restrict__ ajb, const int* __restrict__ najb,
   const real4* __restrict__ rec, real4* __restrict__ F, const int* __restrict__ skip)
{
   if (skip && skip[1])
      return;
   extern __shared__ __align__(128) unsigned char sm_uf[];
   real4* srec = reinterpret_cast<real4*>(sm_uf);                                      // [cap * 32][3]
   unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm_uf + (size_t)cap * 32 * 3 * sizeof(real4));
   const int g = blockIdx.x, t = threadIdx.x;
   const int na = najb[g];
   const int nst = min(na, cap);      // blocks beyond the shared-memory capacity are read from global memory (rare)
   const int* jb = ajb + (size_t)g * SG_VJB_CAP;
   constexpr unsigned BLK_BYTES = 32 * 3 * sizeof(real4);
   if (t == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"((unsigned)nst * BLK_BYTES) : "memory");
      // one bulk copy per run of adjacent blocks
      int r0 = 0;
      while (r0 < nst) {
         const int b0 = jb[r0];
         int r1 = r0 + 1;
         while (r1 < nst && jb[r1] == b0 + (r1 - r0))
            ++r1;
         const unsigned bytes = (unsigned)(r1 - r0) * BLK_BYTES;
         asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(srec + (size_t)r0 * 96)),
            "l"(rec + (size_t)b0 * 96), "r"(bytes), "r"(smem_u32(bar))
            : "memory");
         r0 = r1;
      }
   }
   // own atom: SG_LANES lanes per atom
   const int l = t & (SG_LANES - 1);
   const int i_ = g * SG_GROUP + t / SG_LANES;
   const bool act = i_ < n;
   const int i = act ? i_ : n - 1;
   const pos_t pi = real4_as_pos(rec[3 * (size_t)i]);
   const real thi = rec[3 * (size_t)i + 2].z;
   const int beg = vstart[i];
   const int len = act ? cnt[i] : 0;
   __syncthreads();      // the barrier is initialised before anyone waits on it
   {
      unsigned done = 0;
      while (!done)
         asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(bar)) : "memory");
   }
   V3 fdi = v3(0, 0, 0), fpi = v3(0, 0, 0);
   for (int q = l; q < len; q += SG_LANES) {
      const int slot = nbr16[beg + q] & SG_SLOT_MASK;
      real4 pkr, ua, ub;
      if (slot < nst * 32) {
         const real4* r = srec + 3 * slot;
         pkr = r[0], ua = r[1], ub = r[2];
      } else {
         const real4* r = rec + 3 * ((size_t)jb[slot >> 5] * 32 + (slot & 31));
         pkr = r[0], ua = r[1], ub = r[2];
      }
      const pos_t pk = real4_as_pos(pkr);
      real dx, dy, dz;
      pair_delta(box, pi, pk, dx, dy, dz);
      const real r2 = dx * dx + dy * dy + dz * dz;
      const real rinv = r_rsqrt(r2);
      const real r = r2 * rinv, rr2 = rinv * rinv;
      real rr[3], bn[3], om[3];
      radial_coulomb<3>(rinv, rr2, rr);
      if (EWALD)
         radial_ewald<3>(r, rinv, rr2, aewald, bn);
      thole_one_minus_lambda<3>(r, pos_w(pi), pos_w(pk), min(thi, ub.z), om);
      const real B1 = (EWALD ? bn[1] : rr[1]) - om[1] * rr[1];
      const real B2 = (EWALD ? bn[2] : rr[2]) - om[2] * rr[2];
      const V3 R = v3(dx, dy, dz);
      fdi += dipole_field(R, v3(ua.x, ua.y, ua.z), B1, B2);
      fpi += dipole_field(R, v3(ua.w, ub.x, ub.y), B1, B2);
   }
   fdi = group_sum3<SG_LANES>(fdi);
   fpi = group_sum3<SG_LANES>(fpi);
   if (l == 0 && act)
      store_dp(F, i, fdi, fpi);
}
} // namespace

// ---------------------------------------------------------------------------------------------------------------------------
bool apx_staged_usable(const apx_ctx* c)
{
   return c->staged_on && c->grp.ok && !c->dist.on && !c->thole_table && c->use_records;
}

void apx_group_build(apx_ctx* c)
{
   GroupList& G = c->grp;
   G.ok = 0;
   if (!c->staged_on || c->dist.on || c->thole_table)
      return;
   const int n = c->n, nblk = c->nblk, nw = (nblk + 31) / 32;
   const size_t smem = 2 * (size_t)nw * sizeof(unsigned);
   if (smem > 160 * 1024)      // > 20 M atoms: the row kernels stay in charge
      return;
   const int ngrp = (n + SG_GROUP - 1) / SG_GROUP;
   G.ngrp = ngrp;
   G.vjb.ensure((size_t)ngrp * SG_VJB_CAP);
   G.ajb.ensure((size_t)ngrp * SG_VJB_CAP);
   G.nvjb.ensure(ngrp);
   G.najb.ensure(ngrp);
   G.oflow.ensure(1);
   G.vslot.ensure((size_t)c->rows.nverlet + 64);
   G.nbr16.ensure((size_t)c->rows.nverlet + 64);
   CUDA_CHECK(cudaMemsetAsync(G.oflow.p, 0, sizeof(int), c->stream));
   if (smem > 48 * 1024)
      CUDA_CHECK(cudaFuncSetAttribute(k_group_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
   k_group_build<<<ngrp, 256, smem, c->stream>>>(n, nblk, nw, c->rows.vstart, c->rows.vnbr, G.vjb, G.nvjb, G.vslot, G.oflow);
   APX_COUNT_LAUNCH(c);
   int of = 0;
   CUDA_CHECK(cudaMemcpyAsync(&of, G.oflow.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
   CUDA_CHECK(cudaStreamSynchronize(c->stream));      // (the list build synchronises anyway)
   G.ok = of ? 0 : 1;      // a group reaching more than SG_VJB_CAP blocks: the row kernels stay in charge
}

void apx_rows_compact_grouped(apx_ctx* c, bool count)
{
   RowList& L = c->rows;
   GroupList& G = c->grp;
   const real cut = c->list_cutoff;
   const bool sparse = c->opt.use_polar && c->opt.pcgprec && c->opt.usolve_cutoff > 0;
   const real ucut = sparse ? (real)std::min(c->opt.usolve_cutoff, (double)cut) : (real)0;
   if (count)
      CUDA_CHECK(cudaMemsetAsync(L.total.p, 0, 2 * sizeof(unsigned long long), c->stream));
   k_rows_compact_grp<<<G.ngrp, 256, 0, c->stream>>>(c->n, c->box, cut * cut, ucut * ucut, c->posd, L.vstart, G.vslot, G.vjb, G.nvjb, L.nbr,
      G.nbr16, L.cnt, L.cntu, G.ajb, G.najb, count ? L.total.p : nullptr);
   APX_COUNT_LAUNCH(c);
}

void apx_ufield_staged(apx_ctx* c, cudaStream_t st, real4* F)
{
   RowList& L = c->rows;
   GroupList& G = c->grp;
   const int cap = c->staged_cap;
   const size_t smem = (size_t)cap * 32 * 3 * sizeof(real4) + 16;
   static int attr = 0;
   if (attr != cap) {
      CUDA_CHECK(cudaFuncSetAttribute(k_ufield_staged<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CUDA_CHECK(cudaFuncSetAttribute(k_ufield_staged<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = cap;
   }
   if (c->opt.use_ewald)
      k_ufield_staged<true><<<G.ngrp, SG_THREADS, smem, st>>>(c->n, cap, c->box, (real)c->opt.aewald, L.vstart, L.cnt, G.nbr16, G.ajb, G.najb, c->uf_rec, F,
         c->skip);
   else
      k_ufield_staged<false><<<G.ngrp, SG_THREADS, smem, st>>>(c->n, cap, c->box, (real)c->opt.aewald, L.vstart, L.cnt, G.nbr16, G.ajb, G.najb, c->uf_rec,
         F, c->skip);
}
