// Spatial sort and neighbor-list refresh (our own structure; plays the role of the reference's
// Spatial / spatialDataInit_cu / spatialCheck_cu, include/ff/spatial.h:15-140, src/cu/spatial.cu:728-960).
//
// * atoms are wrapped into the cell and sorted along a Hilbert curve (30-bit key, cub radix sort);
//   32 consecutive sorted atoms form a block with an axis-aligned bounding box;
// * rows.cu turns the block boxes into per-atom Verlet rows (two passes: count, exclusive scan,
//   fill, so the list has no fixed capacity -- the reference throws when its LSTCAP=48 tiles per
//   block overflow, spatial.cu:708-716) and compacts them to the cutoff every step;
// * the list is rebuilt when any atom moved more than buffer/2 since the last build, the
//   reference's criterion (src/nblist.cpp:521-531).
#include "apx_internal.h"
#include "wrap.cuh"
#include <cub/cub.cuh>

namespace {
__device__ __forceinline__ unsigned spread3(unsigned v)
{
   v &= 0x3ff;
   v = (v | (v << 16)) & 0x030000ff;
   v = (v | (v << 8)) & 0x0300f00f;
   v = (v | (v << 4)) & 0x030c30c3;
   v = (v | (v << 2)) & 0x09249249;
   return v;
}

// Hilbert index of a cell (BITS bits per axis) -- Skilling's transpose algorithm (AIP Conf. Proc. 707, 381 (2004)).  Unlike
// the Morton curve, consecutive cells of a Hilbert curve are always face neighbours, so EVERY run of 32 sorted atoms is
// compact: on dhfr2 the largest block box shrinks from 62 x 62 x 31 A (a Morton run across an octant boundary) to 15 A,
// the mean number of candidate j-blocks per i-block from 168 to 110 at 9 A, the maximum from 430 to 154 -- the list build
// was waiting for those few wide blocks.
template <int BITS>
__device__ __forceinline__ unsigned hilbert3(unsigned x, unsigned y, unsigned z)
{
   unsigned X[3] = {x, y, z};
   const unsigned M = 1u << (BITS - 1);
   #pragma unroll
   for (unsigned Q = M; Q > 1; Q >>= 1) {
      const unsigned P = Q - 1;
      #pragma unroll
      for (int i = 0; i < 3; ++i) {
         if (X[i] & Q)
            X[0] ^= P;
         else {
            const unsigned t = (X[0] ^ X[i]) & P;
            X[0] ^= t;
            X[i] ^= t;
         }
      }
   }
   X[1] ^= X[0];
   X[2] ^= X[1];
   unsigned t = 0;
   #pragma unroll
   for (unsigned Q = M; Q > 1; Q >>= 1)
      if (X[2] & Q)
         t ^= Q - 1;
   X[0] ^= t, X[1] ^= t, X[2] ^= t;
   return (spread3(X[0]) << 2) | (spread3(X[1]) << 1) | spread3(X[2]);
}

// nslab > 1 (several GPUs): the key's top bits are the z-slab of the atom in PME grid coordinates
// (w3 = f3 + 1/2 mod 1, the coordinate k_theta_fill uses), below them a 27-bit Morton code -- every
// GPU's atoms are then one contiguous sorted range and sit on that GPU's planes of the grid
__global__ void k_sortkeys(int n, Box b, int nslab, const double* __restrict__ xyz, unsigned* __restrict__ key, int* __restrict__ val,
   real* __restrict__ w3)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n)
      return;
   real wx, wy, wz, fx, fy, fz;
   wrap_pos(b, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], wx, wy, wz, fx, fy, fz);
   if (nslab <= 1) {
      unsigned qx = min(1023u, (unsigned)(fx * 1024));
      unsigned qy = min(1023u, (unsigned)(fy * 1024));
      unsigned qz = min(1023u, (unsigned)(fz * 1024));
      key[i] = hilbert3<10>(qx, qy, qz);
   } else {
      real w = fz + (real)0.5;
      w -= floor(w);
      if (w >= 1) w = 0;
      unsigned slab = min((unsigned)(nslab - 1), (unsigned)(w * nslab));
      unsigned qx = min(511u, (unsigned)(fx * 512));
      unsigned qy = min(511u, (unsigned)(fy * 512));
      unsigned qz = min(511u, (unsigned)(fz * 512));
      key[i] = (slab << 27) | hilbert3<9>(qx, qy, qz);
      w3[i] = w;
   }
   val[i] = i;
}

// per-step: wrapped positions into sorted slots (also used at rebuild)
__global__ void k_gather_pos(int n, int npad, BoxD b, const double* __restrict__ xyz, const int* __restrict__ perm,
   const real* __restrict__ pdamp, real4* __restrict__ posd, pos_t* __restrict__ posq)
{
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= npad)
      return;
   real4 o;
   unsigned q1 = 0, q2 = 0, q3 = 0;
   if (s < n) {
      int i = perm[s];
      wrap_pos_q(b, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], o.x, o.y, o.z, q1, q2, q3);
      o.w = pdamp[i];
   } else {
      o.x = o.y = o.z = 0;
      o.w = 0;
   }
   posd[s] = o;
#ifndef APX_DOUBLE
   posq[s] = make_uint4(q1, q2, q3, __float_as_uint(o.w));
#endif
}

__global__ void k_gather_static(int n, int npad, const int* __restrict__ perm, int* __restrict__ inv,
   const real* __restrict__ thole, const real* __restrict__ polarity, const int* __restrict__ jpolar, real4* __restrict__ tpj)
{
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= npad)
      return;
   real4 o;
   if (s < n) {
      int i = perm[s];
      inv[i] = s;
      real pol = polarity[i];
      o.x = thole[i];
      o.y = pol;
      o.z = (real)1 / (pol > (real)1e-16 ? pol : (real)1e-16);   // polarity_inv, epolar.cpp:505
#ifdef APX_DOUBLE
      o.w = __longlong_as_double((long long)jpolar[i]);
#else
      o.w = __int_as_float(jpolar[i]);
#endif
   } else {
      o.x = o.y = 0;
      o.z = 1;
      o.w = 0;
   }
   tpj[s] = o;
}

__global__ void k_excl_sorted(int nx, const int* __restrict__ ik, const real* __restrict__ sc, const int* __restrict__ inv,
   PairExcl* __restrict__ out)
{
   int e = blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= nx)
      return;
   PairExcl p;
   p.i = inv[ik[2 * e]];
   p.k = inv[ik[2 * e + 1]];
   p.m = sc[4 * e] - 1;
   p.d = sc[4 * e + 1] - 1;
   p.p = sc[4 * e + 2] - 1;
   p.u = sc[4 * e + 3] - 1;
   out[e] = p;
}

// one warp per block: bounding box centre / half extent
__global__ void k_block_boxes(int n, int nblk, const real4* __restrict__ posd, real4* __restrict__ ctr, real4* __restrict__ ext)
{
   int w = (blockIdx.x * blockDim.x + threadIdx.x) / APX_WARP;
   int lane = threadIdx.x & 31;
   if (w >= nblk)
      return;
   int s = w * 32 + lane;
   bool ok = s < n;
   real4 p = posd[ok ? s : w * 32];      // first atom of the block always exists
   real lox = p.x, hix = p.x, loy = p.y, hiy = p.y, loz = p.z, hiz = p.z;
   #pragma unroll
   for (int o = 16; o > 0; o >>= 1) {
      lox = min(lox, __shfl_xor_sync(0xffffffffu, lox, o));
      hix = max(hix, __shfl_xor_sync(0xffffffffu, hix, o));
      loy = min(loy, __shfl_xor_sync(0xffffffffu, loy, o));
      hiy = max(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
      loz = min(loz, __shfl_xor_sync(0xffffffffu, loz, o));
      hiz = max(hiz, __shfl_xor_sync(0xffffffffu, hiz, o));
   }
   if (lane == 0) {
      real4 c, e;
      c.x = (real)0.5 * (lox + hix);
      c.y = (real)0.5 * (loy + hiy);
      c.z = (real)0.5 * (loz + hiz);
      c.w = 0;
      e.x = (real)0.5 * (hix - lox);
      e.y = (real)0.5 * (hiy - loy);
      e.z = (real)0.5 * (hiz - loz);
      e.w = 0;
      ctr[w] = c;
      ext[w] = e;
   }
}

__global__ void k_check_moved(int n, const double* __restrict__ xyz, const double* __restrict__ ref, double lim2, int* flag)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n)
      return;
   double dx = xyz[3 * i] - ref[3 * i], dy = xyz[3 * i + 1] - ref[3 * i + 1], dz = xyz[3 * i + 2] - ref[3 * i + 2];
   if (dx * dx + dy * dy + dz * dz > lim2)
      *flag = 1;
}

} // namespace

void apx_block_boxes(apx_ctx* c, const real4* pos, real4* ctr, real4* ext)
{
   k_block_boxes<<<(c->nblk * 32 + APX_BLOCK - 1) / APX_BLOCK, APX_BLOCK, 0, c->stream>>>(c->n, c->nblk, pos, ctr, ext);
   APX_COUNT_LAUNCH(c);
}

void apx_update_sorted_positions(apx_ctx* c)
{
   int g = (c->npad + 255) / 256;
   k_gather_pos<<<g, 256, 0, c->stream>>>(c->n, c->npad, apx_box_d(c), c->xyz_d, c->perm, c->pdamp_o, c->posd, c->posq);
   APX_COUNT_LAUNCH(c);
   apx_pme_fill_theta(c);
}

// one thread: the answer of k_check_moved and a sequence number into host-mapped pinned memory (host[6] = moved, host[7] = seq).
// The host of an MD step spins on the sequence number instead of synchronising a stream: the kernel sits in a side branch of
// the step graph, beside the last valence evaluation (md.cu)
__global__ void k_list_publish(const int* __restrict__ flag, double* __restrict__ seq, volatile int* host)
{
   const double s = *seq + 1.0;
   *seq = s;
   host[6] = *flag;
   __threadfence_system();
   host[7] = (int)s;
}

// the moved-more-than-buffer/2 test on stream `st`.  seq == nullptr: its answer is copied to flags_h[0] (valid once `st` has
// got there); otherwise it is published with a sequence number (k_list_publish)
void apx_list_check_enqueue(apx_ctx* c, cudaStream_t st, double* seq)
{
   const int n = c->n;
   const double lim = 0.5 * c->opt.list_buffer;
   CUDA_CHECK(cudaMemsetAsync(c->flags.p, 0, sizeof(int), st));
   k_check_moved<<<(n + 255) / 256, 256, 0, st>>>(n, c->xyz_d, c->xyz_ref, lim * lim, c->flags);
   APX_COUNT_LAUNCH(c);
   if (seq) {
      k_list_publish<<<1, 1, 0, st>>>(c->flags, seq, c->flags_h);
      APX_COUNT_LAUNCH(c);
   } else
      CUDA_CHECK(cudaMemcpyAsync(c->flags_h, c->flags.p, sizeof(int), cudaMemcpyDeviceToHost, st));
}

// known_moved: -1 = run the test here; 0 / 1 = the caller already has its answer (md.cu runs it beside the last valence
// evaluation of the inner RESPA level)
void apx_list_refresh(apx_ctx* c, bool force, int known_moved)
{
   int n = c->n;
   bool rebuild = force || !c->list_valid;
   c->tl_valid = 0, c->tl_p_valid = 0;      // positions changed: the stored pair tensors (tlist.cu) are rebuilt by the next operator application
   if (!rebuild && known_moved >= 0)
      rebuild = known_moved != 0;
   else if (!rebuild) {
      // moved more than buffer/2 since the last build?  (src/nblist.cpp:521-531)
      double lim = 0.5 * c->opt.list_buffer;
      CUDA_CHECK(cudaMemsetAsync(c->flags.p, 0, sizeof(int), c->stream));
      k_check_moved<<<(n + 255) / 256, 256, 0, c->stream>>>(n, c->xyz_d, c->xyz_ref, lim * lim, c->flags);
      APX_COUNT_LAUNCH(c);
      CUDA_CHECK(cudaMemcpyAsync(c->flags_h, c->flags.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
      rebuild = c->flags_h[0] != 0;
   }
   if (!rebuild) {
      // per-step refresh: sorted positions, spline tables, rows cut to the cutoff, vdW sites -- a fixed sequence, one graph
      if (apx_graph_begin(c, 0x5000 | (c->vdw.on ? 1 : 0))) {
         apx_update_sorted_positions(c);
         apx_rows_compact(c, false);
         if (c->vdw.on)
            apx_vdw_refresh(c, false);
         apx_graph_end(c, 0x5000 | (c->vdw.on ? 1 : 0));
      }
      return;
   }
   // the captured graphs hold the row buffers' addresses: they stay valid across a rebuild unless a buffer had to grow
   const void* before[13] = {c->rows.vnbr.p, c->rows.nbr.p, c->rows.vstart.p, c->vdw.rows.vnbr.p, c->vdw.rows.vstart.p, c->cubtmp.p,
      c->grp.vslot.p, c->grp.nbr16.p, c->grp.vjb.p, c->grp.ajb.p, c->grp.ok ? (const void*)c : nullptr, c->tl_T.p, c->tl_P.p};
   cudaEventRecord(c->ev2, c->stream);
   // 1. sort along the Morton curve
   const int nslab = c->dist.on ? c->dist.world : 1;
   if (nslab > 1)
      c->w3.ensure(n);
   k_sortkeys<<<(n + 255) / 256, 256, 0, c->stream>>>(n, c->box, nslab, c->xyz_d, c->sortkey, c->permtmp, c->w3);
   size_t need = 0;
   cub::DeviceRadixSort::SortPairs(nullptr, need, c->sortkey.p, c->sortkey2.p, c->permtmp.p, c->perm.p, n, 0, 30, c->stream);
   if (need > c->cubtmp.cap)
      c->cubtmp.ensure(need);
   need = c->cubtmp.cap;
   cub::DeviceRadixSort::SortPairs(c->cubtmp.p, need, c->sortkey.p, c->sortkey2.p, c->permtmp.p, c->perm.p, n, 0, 30, c->stream);
   // ownership of the sorted order: everything on one GPU, or this GPU's slab + the halo plan
   if (c->dist.on)
      apx_dist_after_sort(c);
   else
      c->a0 = 0, c->a1 = n;
   // 2. sorted copies of per-atom data
   int g = (c->npad + 255) / 256;
   k_gather_pos<<<g, 256, 0, c->stream>>>(n, c->npad, apx_box_d(c), c->xyz_d, c->perm, c->pdamp_o, c->posd, c->posq);
   k_gather_static<<<g, 256, 0, c->stream>>>(n, c->npad, c->perm, c->inv, c->thole_o, c->polarity_o, c->jpolar_o, c->tpj);
   if (c->nexcl)
      k_excl_sorted<<<(c->nexcl + 255) / 256, 256, 0, c->stream>>>(c->nexcl, c->excl_ik, c->excl_sc, c->inv, c->excl_s);
   // 3. block boxes and the Verlet rows
   k_block_boxes<<<(c->nblk * 32 + APX_BLOCK - 1) / APX_BLOCK, APX_BLOCK, 0, c->stream>>>(n, c->nblk, c->posd, c->blk_ctr, c->blk_ext);
   c->stats.kernel_launches += 6;
   apx_rows_build(c);
   apx_pme_fill_theta(c);
   CUDA_CHECK(cudaMemcpyAsync(c->xyz_ref, c->xyz_d, sizeof(double) * 3 * n, cudaMemcpyDeviceToDevice, c->stream));
   // rows of the current step + pair counts inside the cutoffs (roofline accounting only)
   {
      apx_rows_compact(c, true);
      unsigned long long h[2] = {0, 0};
      CUDA_CHECK(cudaMemcpyAsync(h, c->rows.total.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
      cudaEventRecord(c->ev3, c->stream);
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
      c->stats.npairs_m = (long long)(h[0] / 2);
      c->stats.npairs_u = (long long)(h[1] / 2);
   }
   cudaEventElapsedTime(&c->stats.ms_list, c->ev2, c->ev3);
   c->stats.nverlet = c->rows.nverlet;
   c->stats.list_rebuilds++;
   c->list_valid = 1;
   c->mpole_inited = 0;     // sorted multipoles must be regenerated in the new order
   if (c->vdw.on)
      apx_vdw_refresh(c, true);
   const void* after[13] = {c->rows.vnbr.p, c->rows.nbr.p, c->rows.vstart.p, c->vdw.rows.vnbr.p, c->vdw.rows.vstart.p, c->cubtmp.p,
      c->grp.vslot.p, c->grp.nbr16.p, c->grp.vjb.p, c->grp.ajb.p, c->grp.ok ? (const void*)c : nullptr, c->tl_T.p, c->tl_P.p};
   for (int q = 0; q < 13; ++q)
      if (before[q] != after[q]) {
         if (getenv("APX_TRACE_GRAPHS"))
            fprintf(stderr, "[apx] list rebuild moved buffer %d (%p -> %p)\n", q, before[q], after[q]);
         apx_pcg_graphs_invalidate(c);
         break;
      }
}
