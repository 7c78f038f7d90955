// Directed per-atom neighbor rows (struct RowList, apx_internal.h).
//
// The reference walks 32x32 tiles of (i-block, k-atom list) pairs and scatters the k-side results
// with atomics (include/ff/spatial.h:15-140, src/cu/amoeba/*_cu1.cc).  At AMOEBA densities only
// ~8 % of the lanes of such a tile hold a pair inside the 7 A cutoff, so the pair kernels here use a
// different structure: every atom owns the full row of its neighbours (both directions of each
// pair are stored), a group of lanes walks one row with every lane on a real pair, the i-side
// sums are reduced with shuffles and written once, and nothing is ever scattered to the k side.
//
//   apx_rows_build    at list rebuild: Verlet rows (cutoff + buffer) from the block bounding boxes
//   apx_rows_compact  every step: rows of the pairs inside the cutoff right now, the ones inside
//                     the preconditioner range (usolve-cutoff) first
#include "apx_internal.h"
#include "pairmath.cuh"
#include "rows.cuh"
#include <cub/cub.cuh>
#include <algorithm>

#define FULL 0xffffffffu

namespace {

// bounding boxes of super-blocks (32 consecutive blocks = 1024 sorted atoms): first level of the
// search, so a warp looks at nblk/32 boxes instead of nblk before it descends
__global__ void k_super_boxes(int nblk, int nsb, const real4* __restrict__ ctr, const real4* __restrict__ ext,
   real4* __restrict__ sctr, real4* __restrict__ sext)
{
   const int sb = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const int lane = threadIdx.x & 31;
   if (sb >= nsb)
      return;
   const int kb = min(sb * 32 + lane, nblk - 1);
   const real4 c = ctr[kb], e = ext[kb];
   real lox = c.x - e.x, hix = c.x + e.x, loy = c.y - e.y, hiy = c.y + e.y, loz = c.z - e.z, hiz = c.z + e.z;
   #pragma unroll
   for (int o = 16; o > 0; o >>= 1) {
      lox = min(lox, __shfl_xor_sync(FULL, lox, o));
      hix = max(hix, __shfl_xor_sync(FULL, hix, o));
      loy = min(loy, __shfl_xor_sync(FULL, loy, o));
      hiy = max(hiy, __shfl_xor_sync(FULL, hiy, o));
      loz = min(loz, __shfl_xor_sync(FULL, loz, o));
      hiz = max(hiz, __shfl_xor_sync(FULL, hiz, o));
   }
   if (lane == 0) {
      real4 oc, oe;
      oc.x = (real)0.5 * (lox + hix), oc.y = (real)0.5 * (loy + hiy), oc.z = (real)0.5 * (loz + hiz), oc.w = 0;
      oe.x = (real)0.5 * (hix - lox), oe.y = (real)0.5 * (hiy - loy), oe.z = (real)0.5 * (hiz - loz), oe.w = 0;
      sctr[sb] = oc;
      sext[sb] = oe;
   }
}

__device__ __forceinline__ bool boxes_within(const Box& b, real4 ci, real4 ei, real4 ck, real4 ek, real range2)
{
   real dx = ck.x - ci.x, dy = ck.y - ci.y, dz = ck.z - ci.z;
   apx_image(b, dx, dy, dz);
   dx = max((real)0, fabs(dx) - ei.x - ek.x);
   dy = max((real)0, fabs(dy) - ei.y - ek.y);
   dz = max((real)0, fabs(dz) - ei.z - ek.z);
   return dx * dx + dy * dy + dz * dz <= range2;
}

// One CTA (4 warps) per i-block.  MODE 0: count only (vcnt).  MODE 1: fill exact rows at vstart.  MODE 2: fill padded slots
// [vstart[s], vstart[s+1]) AND count: a row that outgrows its slot raises *oflow and is not written, the count stays exact.
//
//   candidates   the CTA culls the super-block boxes (128 per round), then the 32 blocks of every surviving super-block (one
//                super-block per warp), and appends the surviving j-blocks IN ASCENDING ORDER to a shared list; every
//                RB_CAP candidates (and at the end) the chunk is processed:
//   pass A       the warps split the CANDIDATES: lane = one j-atom; j-atoms outside the range of the i-block's box and
//                i-atoms outside the range of the j-block's box are dropped before any pair is tested; for the remaining
//                i-atoms (position broadcast from shared memory) one ballot per (candidate, i-atom) gives the 32-bit mask of
//                listed j-atoms, kept in shared memory -- every distance is computed once, by one warp;
//   pass B       the warps split the I-ATOMS: the row length of the chunk is the popcount sum of the atom's masks, and the
//                set bits are written at running offsets (k ascending, as k_rows_flag_listed's binary search expects).
// The previous version gave each of an i-block's 4 warps a quarter of its ATOMS: each warp repeated the whole candidate
// search and walked every candidate once per atom (505 us + 1090 us for the two lists of dhfr2, profiles/r02e_trace_md.txt).
#define RB_THREADS 128
#define RB_CAP 256      // candidate j-blocks per chunk: 32 KB of masks

template <int MODE>
__global__ void __launch_bounds__(RB_THREADS) k_rows_build(int n, int nblk, int nsb, int a0, int a1, Box b, real range,
   const real4* __restrict__ posd, const real4* __restrict__ ctr, const real4* __restrict__ ext, const real4* __restrict__ sctr,
   const real4* __restrict__ sext, int* __restrict__ vcnt, const int* __restrict__ vstart, int* __restrict__ vnbr,
   const int* __restrict__ perm, const int* __restrict__ exoff, const int* __restrict__ exlist, real exr2, int* __restrict__ oflow)
{
   __shared__ real4 s_pi[32];
   __shared__ unsigned s_mask[RB_CAP][32];
   __shared__ int s_cand[RB_CAP];
   __shared__ int s_sb[RB_THREADS];
   __shared__ int s_wcount[4];
   __shared__ int s_row[32];      // running row length of every i-atom
   __shared__ int s_base[32], s_room[32];
   const int ib = a0 / 32 + blockIdx.x;
   if (ib >= nblk || ib * 32 >= a1)
      return;
   const int t = threadIdx.x, lane = t & 31, w = t >> 5;
   const real range2 = range * range;
   const real4 ci = ctr[ib], ei = ext[ib];
   // i-atoms of this block that get a row: inside the system and inside the owned range [a0, a1)
   const int q0 = max(0, a0 - ib * 32), q1 = min(32, min(n, a1) - ib * 32);
   if (t < 32) {
      const int si = ib * 32 + t;
      s_pi[t] = posd[min(si, n - 1)];
      s_row[t] = 0;
      const bool real_i = t >= q0 && t < q1;
      s_base[t] = (MODE != 0 && real_i) ? vstart[si] : 0;
      s_room[t] = (MODE == 2 && real_i) ? vstart[si + 1] - vstart[si] : 0x7fffffff;
   }
   __syncthreads();
   const unsigned lt = (1u << lane) - 1;
   const unsigned qmask = (q1 >= 32 ? 0xffffffffu : ((1u << q1) - 1)) & ~((1u << q0) - 1);
   int ncand = 0;      // uniform across the CTA

   auto process = [&](int nc) {
      // ---- pass A: masks
      for (int c = w; c < nc; c += 4) {
         const int kb = s_cand[c];
         const int s = kb * 32 + lane;
         const real4 pk = posd[min(s, n - 1)];
         bool in = false;
         if (s < n) {
            real dx = pk.x - ci.x, dy = pk.y - ci.y, dz = pk.z - ci.z;
            apx_image(b, dx, dy, dz);
            dx = max((real)0, fabs(dx) - ei.x);
            dy = max((real)0, fabs(dy) - ei.y);
            dz = max((real)0, fabs(dz) - ei.z);
            in = dx * dx + dy * dy + dz * dz <= range2;
         }
         unsigned im = 0;
         if (__ballot_sync(FULL, in)) {
            const real4 ck = ctr[kb], ek = ext[kb], pq = s_pi[lane];
            real dx = pq.x - ck.x, dy = pq.y - ck.y, dz = pq.z - ck.z;
            apx_image(b, dx, dy, dz);
            dx = max((real)0, fabs(dx) - ek.x);
            dy = max((real)0, fabs(dy) - ek.y);
            dz = max((real)0, fabs(dz) - ek.z);
            im = __ballot_sync(FULL, dx * dx + dy * dy + dz * dz <= range2) & qmask;
         }
         unsigned mine = 0;
         while (im) {
            const int q = __ffs(im) - 1;
            im &= im - 1;
            const real4 pq = s_pi[q];
            real dx = pk.x - pq.x, dy = pk.y - pq.y, dz = pk.z - pq.z;
            apx_image(b, dx, dy, dz);
            const real r2q = dx * dx + dy * dy + dz * dz;
            bool ok = in && s != ib * 32 + q && r2q <= range2;
            if (exoff && __any_sync(FULL, ok && r2q <= exr2)) {
               // pairs that never interact (vdW 1-2/1-3 with scale 0) are left out of the rows for good:
               // exclusion is topology, so it is tested at list build, not in the pair kernel
               const int cq = perm[ib * 32 + q];
               const int eb = exoff[cq], ee = exoff[cq + 1];
               if (ok && r2q <= exr2) {
                  const int ck = perm[s];
                  for (int e = eb; e < ee; ++e)
                     if (exlist[e] == ck)
                        ok = false;
               }
            }
            const unsigned m = __ballot_sync(FULL, ok);
            if (lane == q)
               mine = m;
         }
         s_mask[c][lane] = mine;
      }
      __syncthreads();
      // ---- pass B: rows
      for (int q = w * 8; q < w * 8 + 8; ++q) {
         if (!((qmask >> q) & 1u))
            continue;
         int tot = 0;
         for (int c = lane; c < nc; c += 32)
            tot += __popc(s_mask[c][q]);
         #pragma unroll
         for (int o = 16; o > 0; o >>= 1)
            tot += __shfl_xor_sync(FULL, tot, o);
         const int have = s_row[q];
         if (MODE != 0 && tot > 0) {
            if (have + tot <= s_room[q]) {
               int off = s_base[q] + have;
               for (int c = 0; c < nc; ++c) {
                  const unsigned m = s_mask[c][q];
                  if ((m >> lane) & 1u)
                     vnbr[off + __popc(m & lt)] = s_cand[c] * 32 + lane;
                  off += __popc(m);
               }
            } else if (lane == 0) {
               *oflow = 1;
               s_room[q] = -1;      // the row stays unwritten from here on
            }
         }
         __syncwarp();
         if (lane == 0)
            s_row[q] = have + tot;
      }
      __syncthreads();
   };

   for (int sb0 = 0; sb0 < nsb; sb0 += RB_THREADS) {
      // super-blocks within range, ascending
      const int sb = sb0 + t;
      const bool hit = sb < nsb && boxes_within(b, ci, ei, sctr[min(sb, nsb - 1)], sext[min(sb, nsb - 1)], range2);
      const unsigned hm = __ballot_sync(FULL, hit);
      if (lane == 0)
         s_wcount[w] = __popc(hm);
      __syncthreads();
      int before = 0, nsbhit = 0;
      #pragma unroll
      for (int k = 0; k < 4; ++k) {
         before += k < w ? s_wcount[k] : 0;
         nsbhit += s_wcount[k];
      }
      if (hit)
         s_sb[before + __popc(hm & lt)] = sb;
      __syncthreads();
      // their blocks, one super-block per warp and round
      for (int j0 = 0; j0 < nsbhit; j0 += 4) {
         const int j = j0 + w;
         unsigned bm = 0;
         int kb = 0;
         if (j < nsbhit) {
            kb = s_sb[j] * 32 + lane;
            bm = __ballot_sync(FULL, kb < nblk && boxes_within(b, ci, ei, ctr[min(kb, nblk - 1)], ext[min(kb, nblk - 1)], range2));
         }
         if (lane == 0)
            s_wcount[w] = __popc(bm);
         __syncthreads();
         int pre = 0, add = 0;
         #pragma unroll
         for (int k = 0; k < 4; ++k) {
            pre += k < w ? s_wcount[k] : 0;
            add += s_wcount[k];
         }
         if ((bm >> lane) & 1u)
            s_cand[ncand + pre + __popc(bm & lt)] = kb;
         ncand += add;
         __syncthreads();
         if (ncand > RB_CAP - 128) {
            process(ncand);
            ncand = 0;
         }
      }
   }
   if (ncand > 0)
      process(ncand);
   if (MODE != 1 && t < 32 && ((qmask >> t) & 1u))
      vcnt[ib * 32 + t] = s_row[t];
}

// padded slot sizes from the previous build's row lengths (caller order): +12.5 % + 16 entries (slack 0: exactly the old length)
__global__ void k_rows_caps(int n, const int* __restrict__ perm, const int* __restrict__ prev_o, int slack, int* __restrict__ cap)
{
   const int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s > n)
      return;
   int c = 0;
   if (s < n) {
      c = prev_o[perm[s]];
      if (slack)
         c += (c >> 3) + 16;
   }
   cap[s] = c;
}

__global__ void k_rows_save_counts(int n, const int* __restrict__ perm, const int* __restrict__ vcnt, int* __restrict__ prev_o)
{
   const int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s < n)
      prev_o[perm[s]] = vcnt[s];
}

// pack the padded rows: one warp per row, coalesced copy
__global__ void __launch_bounds__(128) k_rows_pack(int n, const int* __restrict__ capstart, const int* __restrict__ vstart, const int* __restrict__ vpad,
   int* __restrict__ vnbr)
{
   const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const int lane = threadIdx.x & 31;
   if (s >= n)
      return;
   const int src = capstart[s], dst = vstart[s], len = vstart[s + 1] - dst;
   for (int q = lane; q < len; q += 32)
      vnbr[dst + q] = vpad[src + q];
}

// one warp per atom.  Rows of up to 32 * RC_R entries (all of them at AMOEBA densities: ~300 entries at 9 A) are read ONCE --
// index, position and class (inside usolve-cutoff / inside the cutoff / outside) of every entry stay in registers, so the
// RC_R dependent index -> position round trips of a lane overlap -- and written in two groups, the preconditioner's pairs
// first.  Longer rows take two sweeps (positions stay in L1 between them).
#define RC_R 16
__global__ void __launch_bounds__(128) k_rows_compact(int a0, int n, Box b, real cut2, real ucut2, const real4* __restrict__ posd,
   const int* __restrict__ vstart, const int* __restrict__ vnbr, int* __restrict__ nbr, int* __restrict__ cnt,
   int* __restrict__ cntu, unsigned long long* __restrict__ total)
{
   const int i = a0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);     // owned atoms a0 <= i < n
   const int lane = threadIdx.x & 31;
   if (i >= n)
      return;
   const int beg = vstart[i], end = vstart[i + 1];
   const real4 pi = posd[i];
   const unsigned lt = (1u << lane) - 1;
   int out = beg;
   int nu = 0;
   if (end - beg <= 32 * RC_R) {
      int kk[RC_R];
      unsigned char cls[RC_R];
      #pragma unroll
      for (int j = 0; j < RC_R; ++j) {
         const int q = beg + 32 * j + lane;
         kk[j] = q < end ? vnbr[q] : 0;      // may carry ROW_LISTED_FLAG: copied as is
      }
      int n1 = 0;
      #pragma unroll
      for (int j = 0; j < RC_R; ++j) {
         const int q = beg + 32 * j + lane;
         cls[j] = 0;
         if (q < end) {
            const real4 pk = posd[kk[j] & ROW_INDEX_MASK];
            real dx = pk.x - pi.x, dy = pk.y - pi.y, dz = pk.z - pi.z;
            apx_image(b, dx, dy, dz);
            const real r2 = dx * dx + dy * dy + dz * dz;
            cls[j] = r2 <= ucut2 ? 1 : (r2 <= cut2 ? 2 : 0);
         }
         n1 += __popc(__ballot_sync(FULL, cls[j] == 1));
      }
      nu = n1;
      int o1 = beg, o2 = beg + n1;
      #pragma unroll
      for (int j = 0; j < RC_R; ++j) {
         if (beg + 32 * j >= end)
            break;
         const unsigned m1 = __ballot_sync(FULL, cls[j] == 1), m2 = __ballot_sync(FULL, cls[j] == 2);
         if (cls[j] == 1)
            nbr[o1 + __popc(m1 & lt)] = kk[j];
         else if (cls[j] == 2)
            nbr[o2 + __popc(m2 & lt)] = kk[j];
         o1 += __popc(m1);
         o2 += __popc(m2);
      }
      out = o2;
   } else {
      for (int sweep = 0; sweep < 2; ++sweep) {
         if (sweep == 0 && ucut2 <= 0)
            continue;
         for (int q0 = beg; q0 < end; q0 += 32) {
            int q = q0 + lane;
            const int k = q < end ? vnbr[q] : 0;
            bool ok = false;
            if (q < end) {
               real4 pk = posd[k & ROW_INDEX_MASK];
               real dx = pk.x - pi.x, dy = pk.y - pi.y, dz = pk.z - pi.z;
               apx_image(b, dx, dy, dz);
               real r2 = dx * dx + dy * dy + dz * dz;
               ok = sweep == 0 ? r2 <= ucut2 : (r2 > ucut2 && r2 <= cut2);
            }
            unsigned m = __ballot_sync(FULL, ok);
            if (ok)
               nbr[out + __popc(m & lt)] = k;
            out += __popc(m);
         }
         if (sweep == 0)
            nu = out - beg;
      }
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
// One thread per listed pair and direction: find k in the (ascending) Verlet row of i and set the flag bit.  Runs once per list
// build, before anything reads the rows; a pair farther apart than the Verlet range is simply not found.
__global__ void k_rows_flag_listed(int nx, const PairExcl* __restrict__ ex, int a0, int a1, const int* __restrict__ vstart,
   int* __restrict__ vnbr)
{
   const int t = blockIdx.x * blockDim.x + threadIdx.x;
   if (t >= 2 * nx)
      return;
   const PairExcl p = ex[t >> 1];
   const int i = (t & 1) ? p.k : p.i, k = (t & 1) ? p.i : p.k;
   if (i < a0 || i >= a1)
      return;
   int lo = vstart[i], hi = vstart[i + 1] - 1;
   while (lo <= hi) {
      const int mid = (lo + hi) >> 1;
      const int v = vnbr[mid] & ROW_INDEX_MASK;      // (two listed pairs never share a slot, so a concurrent flag is harmless)
      if (v == k) {
         vnbr[mid] = (int)((unsigned)k | ROW_LISTED_FLAG);
         return;
      }
      if (v < k)
         lo = mid + 1;
      else
         hi = mid - 1;
   }
}
} // namespace

void apx_rows_build(apx_ctx* c)
{
   apx_rows_build_on(c, c->rows, c->posd, c->blk_ctr, c->blk_ext, c->list_cutoff + c->list_buffer, nullptr, nullptr, 0, true);
   if (c->nexcl > 0) {
      k_rows_flag_listed<<<(2 * c->nexcl + 255) / 256, 256, 0, c->stream>>>(c->nexcl, c->excl_s, c->a0, c->a1, c->rows.vstart, c->rows.vnbr);
      APX_COUNT_LAUNCH(c);
   }
   apx_tlist_reserve(c);    // stored pair tensors of the operator (tlist.cu), same offsets as the rows
   apx_group_build(c);      // 64-atom groups, their j-blocks and slot rows for the staged operator (staged.cu)
}

// Verlet rows of the positions `pos` (sorted order, block boxes ctr/ext) within `range`.  exoff/exlist: optional CSR (caller
// indices) of partners that are never listed; only pairs closer than exrange are tested against it.
void apx_rows_build_on(apx_ctx* c, RowList& L, const real4* pos, const real4* bctr, const real4* bext, real range, const int* exoff,
   const int* exlist, real exrange, bool want_compact)
{
   const int n = c->n, nblk = c->nblk, nsb = (nblk + 31) / 32;
   const int a0 = c->a0, a1 = c->a1;
   const int nib = a1 > a0 ? (a1 + 31) / 32 - a0 / 32 : 0;      // i-blocks that hold owned atoms
   const int grid = std::max(1, nib);      // one CTA per i-block
   const real exr2 = exrange * exrange;
   L.vstart.ensure(n + 1);
   L.vcnt.ensure(n + 1);
   L.cnt.ensure(n);
   L.cntu.ensure(n);
   L.total.ensure(2);
   L.sctr.ensure(nsb);
   L.sext.ensure(nsb);
   k_super_boxes<<<(nsb * 32 + 127) / 128, 128, 0, c->stream>>>(nblk, nsb, bctr, bext, L.sctr, L.sext);
   CUDA_CHECK(cudaMemsetAsync(L.vcnt.p, 0, sizeof(int) * (n + 1), c->stream));
   auto scan = [&](int* in, int* out) {
      size_t need = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, n + 1, c->stream);
      if (need > c->cubtmp.cap)
         c->cubtmp.ensure(need);
      need = c->cubtmp.cap;
      cub::DeviceScan::ExclusiveSum(c->cubtmp.p, need, in, out, n + 1, c->stream);
   };
   // ---- one search pass: padded slots sized from the previous build, fill + count, then a packing copy (APX_ROWS_ONEPASS)
   const bool onepass = c->rows_onepass && L.have_prev && !c->dist.on && a0 == 0 && a1 == n;
   bool filled = false;
   if (onepass) {
      const int slack = c->rows_onepass == 2 ? 0 : 1;
      L.capstart.ensure(n + 1);
      L.oflow.ensure(1);
      L.vpad.ensure((size_t)(slack ? L.prev_total + L.prev_total / 8 + 16ll * n : L.prev_total) + 64);
      k_rows_caps<<<(n + 256) / 256, 256, 0, c->stream>>>(n, c->perm, L.prev_o, slack, L.vcnt);      // vcnt as scratch for the slot sizes
      scan(L.vcnt.p, L.capstart.p);
      CUDA_CHECK(cudaMemsetAsync(L.vcnt.p, 0, sizeof(int) * (n + 1), c->stream));
      CUDA_CHECK(cudaMemsetAsync(L.oflow.p, 0, sizeof(int), c->stream));
      k_rows_build<2><<<grid, RB_THREADS, 0, c->stream>>>(n, nblk, nsb, a0, a1, c->box, range, pos, bctr, bext, L.sctr, L.sext, L.vcnt, L.capstart,
         L.vpad, c->perm, exoff, exlist, exr2, L.oflow);
      c->stats.kernel_launches += 2;
   } else {
      k_rows_build<0><<<grid, RB_THREADS, 0, c->stream>>>(n, nblk, nsb, a0, a1, c->box, range, pos, bctr, bext, L.sctr, L.sext,
         L.vcnt, nullptr, nullptr, c->perm, exoff, exlist, exr2, nullptr);
   }
   scan(L.vcnt.p, L.vstart.p);
   int total = 0, oflow = 0;
   CUDA_CHECK(cudaMemcpyAsync(&total, L.vstart.p + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
   if (onepass)
      CUDA_CHECK(cudaMemcpyAsync(&oflow, L.oflow.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
   if (total < 0)
      APX_THROW("neighbor rows exceed 2^31 entries");
   L.nverlet = total;
   L.vnbr.ensure((size_t)total + 32);
   if (want_compact)
      L.nbr.ensure((size_t)total + 32);
   if (onepass && !oflow) {
      k_rows_pack<<<(n * 32 + 127) / 128, 128, 0, c->stream>>>(n, L.capstart, L.vstart, L.vpad, L.vnbr);
      c->stats.kernel_launches += 1;
      filled = true;
   }
   if (!filled)      // two-pass path, or a row outgrew its slot: the counts are exact either way
      k_rows_build<1><<<grid, RB_THREADS, 0, c->stream>>>(n, nblk, nsb, a0, a1, c->box, range, pos, bctr, bext, L.sctr, L.sext,
         nullptr, L.vstart, L.vnbr, c->perm, exoff, exlist, exr2, nullptr);
   if (c->rows_onepass && !c->dist.on && a0 == 0 && a1 == n) {
      L.prev_o.ensure(n);
      if (!L.have_prev) {
         // what the NEXT (one-pass) build needs is allocated now, with the head-room of its padded slots: a cudaMalloc of
         // tens of MB inside a later MD step costs milliseconds (measured: 10 ms on the first rebuild of a dhfr2 run)
         const int slack = c->rows_onepass == 2 ? 0 : 1;
         L.capstart.ensure(n + 1);
         L.oflow.ensure(1);
         L.vpad.ensure((size_t)(slack ? (long long)total + total / 8 + 16ll * n : (long long)total) + 64);
      }
      k_rows_save_counts<<<(n + 255) / 256, 256, 0, c->stream>>>(n, c->perm, L.vcnt, L.prev_o);
      L.prev_total = total;
      L.have_prev = 1;
      c->stats.kernel_launches += 1;
   }
   c->stats.kernel_launches += 3;
}

void apx_rows_compact(apx_ctx* c, bool count)
{
   if (apx_staged_usable(c)) {
      apx_rows_compact_grouped(c, count);
      return;
   }
   RowList& L = c->rows;
   const int no = c->a1 - c->a0;
   const real cut = c->list_cutoff;
   const bool sparse = c->opt.use_polar && c->opt.pcgprec && c->opt.usolve_cutoff > 0;
   const real ucut = sparse ? (real)std::min(c->opt.usolve_cutoff, (double)cut) : (real)0;
   if (count)
      CUDA_CHECK(cudaMemsetAsync(L.total.p, 0, 2 * sizeof(unsigned long long), c->stream));
   if (no > 0)
      k_rows_compact<<<(no * 32 + 127) / 128, 128, 0, c->stream>>>(c->a0, c->a1, c->box, cut * cut, ucut * ucut, c->posd, L.vstart, L.vnbr,
         L.nbr, L.cnt, L.cntu, count ? L.total.p : nullptr);
   APX_COUNT_LAUNCH(c);
}
