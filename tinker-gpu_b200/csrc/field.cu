// Real-space field kernels over the directed neighbor rows (rows.cu): permanent field (dfield),
// mutual field of a dipole pair (ufield, the CG operator) and the short-range preconditioner.
// They stand where dfield_cu1 / ufield_cu1 / sparsePrecond_cu1 stand in the reference
// (src/cu/amoeba/field.cu:37-137, precond.cu:13-43) but are organised differently:
//   * a group of G lanes owns one atom i and walks the row of its neighbours, one real pair per
//     lane per step (the reference's 32x32 tiles keep ~8 % of their lanes busy at this density);
//     the i-side sums are reduced by shuffles and written once -- no atomics, no k-side scatter,
//     and the result does not depend on scheduling;
//   * every pair is evaluated with all exclusion scales = 1 (no per-pair bit masks);
//     the few excluded pairs are corrected afterwards by one thread per listed pair with
//     (scale-1) non-Ewald terms -- B_n = (s-1) lambda_n rr_n in the notation of pairmath.cuh;
//   * because d- and p-scaling only differ on excluded pairs, the row pass accumulates ONE
//     permanent field; the d/p split is made by the exclusion pass.
#include "apx_internal.h"
#include "pairmath.cuh"
#include "rows.cuh"
#include "dp.cuh"
#include "tlist.cuh"

namespace {
__device__ __forceinline__ int as_int(real w)
{
#ifdef APX_DOUBLE
   return (int)__double_as_longlong(w);
#else
   return __float_as_int(w);
#endif
}

// An exclusion pass corrects pairs the row pass visited: membership must be decided exactly as the per-step compaction of
// the rows decides it (rows.cu: k_rows_compact, minimum image of the wrapped float coordinates), not from the more
// accurate separation the math uses -- a pair within rounding of the cutoff must not be corrected without being visited.
__device__ __forceinline__ bool excl_in_rows(const Box& box, const real4* __restrict__ posd, int i, int k, real cut2)
{
   const real4 a = posd[i], b = posd[k];
   real dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
   apx_image(box, dx, dy, dz);
   return dx * dx + dy * dy + dz * dz <= cut2;
}

// -------------------------------------------------------------------------------------------
// ufield: field of (ud, up) at every atom, Ewald real space or plain Thole-damped Coulomb
// -------------------------------------------------------------------------------------------
template <bool EWALD, bool TABLE, int G>
__global__ void __launch_bounds__(ROWS_BLOCK) k_ufield_rows(int a0, int a1, Box box, real aewald, const int* __restrict__ vstart,
   const int* __restrict__ cnt, const int* __restrict__ nbr, const pos_t* __restrict__ posq, const real4* __restrict__ tpj,
   const real* __restrict__ thlval, int nj, const real4* __restrict__ U, real4* __restrict__ F, const int* __restrict__ skip)
{
   if (skip && skip[1])
      return;
   ROWS_FOREACH_ATOM(G, a0, a1, i, l, act)
   {
      const pos_t pi = posq[i];
      const real4 qi = tpj[i];
      const int beg = vstart[i];
      const int len = act ? cnt[i] : 0;
      V3 fdi = v3(0, 0, 0), fpi = v3(0, 0, 0);
      for (int q = l; q < len; q += G) {
         const int k = nbr[beg + q] & ROW_INDEX_MASK;
         const pos_t pk = posq[k];
         const real4 qk = tpj[k];
         V3 a, b;
         load_dp(U, k, a, b);
         real dx, dy, dz;
         pair_delta(box, pi, pk, dx, dy, dz);
         const real r2 = dx * dx + dy * dy + dz * dz;
         const real rinv = r_rsqrt(r2);
         const real r = r2 * rinv, rr2 = rinv * rinv;
         real rr[3], bn[3], om[3];
         radial_coulomb<3>(rinv, rr2, rr);
         if (EWALD)
            radial_ewald<3>(r, rinv, rr2, aewald, bn);
         const real pg = TABLE ? thlval[as_int(qi.w) * nj + as_int(qk.w)] : min(qi.x, qk.x);
         thole_one_minus_lambda<3>(r, pos_w(pi), pos_w(pk), pg, om);
         const real B1 = (EWALD ? bn[1] : rr[1]) - om[1] * rr[1];
         const real B2 = (EWALD ? bn[2] : rr[2]) - om[2] * rr[2];
         const V3 R = v3(dx, dy, dz);
         fdi += dipole_field(R, a, B1, B2);
         fpi += dipole_field(R, b, B1, B2);
      }
      fdi = group_sum3<G>(fdi);
      fpi = group_sum3<G>(fpi);
      if (l == 0 && act)
         store_dp(F, i, fdi, fpi);
   }
}

// ---- record form of the same operator ------------------------------------------------------
// ncu (profiles/r01j_ncu_full_summary.txt, 1 M atoms) shows the kernel above bound by L1 sector traffic
// (l1tex 86 % of peak, FMA pipe 37 %): every neighbour costs four 32-byte sectors -- posd, tpj and the two
// halves of the packed dipoles live in three different arrays.  k_uf_records interleaves what a pair needs
// into ONE 48-byte record per atom, rec[3s] = (x,y,z,pdamp), rec[3s+1] = (d.x,d.y,d.z,p.x),
// rec[3s+2] = (p.y,p.z,thole,polarity): two sectors and three 16-byte loads per neighbour.
__global__ void k_uf_records(int n, const pos_t* __restrict__ posq, const real4* __restrict__ tpj, const real4* __restrict__ U,
   real4* __restrict__ rec, const int* __restrict__ skip)
{
   if (skip && skip[1])
      return;
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= n)
      return;
   const real4 q = tpj[s];
   real4 b = U[2 * s + 1];
   b.z = q.x;
   b.w = q.y;
   rec[3 * s] = pos_as_real4(posq[s]);
   rec[3 * s + 1] = U[2 * s];
   rec[3 * s + 2] = b;
}

template <bool EWALD, int G>
__global__ void __launch_bounds__(ROWS_BLOCK) k_ufield_rows_rec(int a0, int a1, Box box, real aewald, const int* __restrict__ vstart,
   const int* __restrict__ cnt, const int* __restrict__ nbr, const real4* __restrict__ rec, real4* __restrict__ F,
   const int* __restrict__ skip)
{
   if (skip && skip[1])
      return;
   ROWS_FOREACH_ATOM(G, a0, a1, i, l, act)
   {
      const pos_t pi = real4_as_pos(rec[3 * i]);
      const real thi = rec[3 * i + 2].z;
      const int beg = vstart[i];
      const int len = act ? cnt[i] : 0;
      V3 fdi = v3(0, 0, 0), fpi = v3(0, 0, 0);
      for (int q = l; q < len; q += G) {
         const int k = nbr[beg + q] & ROW_INDEX_MASK;
         const pos_t pk = real4_as_pos(rec[3 * k]);
         const real4 ua = rec[3 * k + 1], ub = rec[3 * k + 2];
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
      fdi = group_sum3<G>(fdi);
      fpi = group_sum3<G>(fpi);
      if (l == 0 && act)
         store_dp(F, i, fdi, fpi);
   }
}

// exclusion pass for ufield (only pairs whose u-scale != 1; empty for stock AMOEBA)
template <bool TABLE>
__global__ void k_ufield_excl(int nx, int a0, int a1, Box box, real cut2, const PairExcl* __restrict__ ex, const real4* __restrict__ posd, const pos_t* __restrict__ posq,
   const real4* __restrict__ tpj, const real* __restrict__ thlval, int nj, const real4* __restrict__ U, real4* __restrict__ F)
{
   int e = blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= nx)
      return;
   PairExcl p = ex[e];
   if (p.u == 0)
      return;
   const bool own_i = p.i >= a0 && p.i < a1, own_k = p.k >= a0 && p.k < a1;   // each GPU corrects its own atoms
   if (!own_i && !own_k)
      return;
   if (!excl_in_rows(box, posd, p.i, p.k, cut2))
      return;
   const pos_t pi = posq[p.i], pk = posq[p.k];
   real dx, dy, dz;
   pair_delta(box, pi, pk, dx, dy, dz);
   real r2 = dx * dx + dy * dy + dz * dz;
   real rinv = r_rsqrt(r2), r = r2 * rinv, rr2 = rinv * rinv;
   real rr[3], om[3];
   radial_coulomb<3>(rinv, rr2, rr);
   real4 qi = tpj[p.i], qk = tpj[p.k];
   real pg = TABLE ? thlval[as_int(qi.w) * nj + as_int(qk.w)] : min(qi.x, qk.x);
   thole_one_minus_lambda<3>(r, pos_w(pi), pos_w(pk), pg, om);
   real B1 = p.u * (1 - om[1]) * rr[1], B2 = p.u * (1 - om[2]) * rr[2];
   V3 R = v3(dx, dy, dz);
   V3 udi, upi, udk, upk;
   load_dp(U, p.i, udi, upi);
   load_dp(U, p.k, udk, upk);
   if (own_i)
      atomic_dp(F, p.i, dipole_field(R, udk, B1, B2), dipole_field(R, upk, B1, B2));
   if (own_k)
      atomic_dp(F, p.k, dipole_field(R, udi, B1, B2), dipole_field(R, upi, B1, B2));
}

// -------------------------------------------------------------------------------------------
// dfield: permanent-multipole field; row pass = common part, exclusion pass = d/p split
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ Mpole load_mpole(const real4* mp0, const real4* mp1, const real2* mp2, int s)
{
   real4 a = mp0[s], b = mp1[s];
   real2 c = mp2[s];
   Mpole m;
   m.c = a.x, m.dx = a.y, m.dy = a.z, m.dz = a.w;
   m.qxx = b.x, m.qxy = b.y, m.qxz = b.z, m.qyy = b.w, m.qyz = c.x, m.qzz = c.y;
   return m;
}

template <bool EWALD, bool TABLE, int G>
__global__ void __launch_bounds__(ROWS_BLOCK) k_dfield_rows(int a0, int a1, Box box, real aewald, const int* __restrict__ vstart,
   const int* __restrict__ cnt, const int* __restrict__ nbr, const pos_t* __restrict__ posq, const real4* __restrict__ tpj,
   const real* __restrict__ thlval, int nj, const real4* __restrict__ mp0, const real4* __restrict__ mp1,
   const real2* __restrict__ mp2, real* __restrict__ fd, real* __restrict__ fpd, int assign, real4* __restrict__ T,
   real4* __restrict__ P, const int* __restrict__ cntu)
{
   ROWS_FOREACH_ATOM(G, a0, a1, i, l, act)
   {
      const pos_t pi = posq[i];
      const real4 qi = tpj[i];
      const int beg = vstart[i];
      const int len = act ? cnt[i] : 0;
      const int lenu = (P && act) ? cntu[i] : 0;
      V3 fi = v3(0, 0, 0);
      for (int q = l; q < len; q += G) {
         const int k = nbr[beg + q] & ROW_INDEX_MASK;
         const pos_t pk = posq[k];
         const real4 qk = tpj[k];
         const Mpole mk = load_mpole(mp0, mp1, mp2, k);
         real dx, dy, dz;
         pair_delta(box, pi, pk, dx, dy, dz);
         const real r2 = dx * dx + dy * dy + dz * dz;
         const real rinv = r_rsqrt(r2);
         const real r = r2 * rinv, rr2 = rinv * rinv;
         real rr[4], bn[4], om[4];
         radial_coulomb<4>(rinv, rr2, rr);
         if (EWALD)
            radial_ewald<4>(r, rinv, rr2, aewald, bn);
         const real pg = TABLE ? thlval[as_int(qi.w) * nj + as_int(qk.w)] : min(qi.x, qk.x);
         thole_one_minus_lambda<4>(r, pos_w(pi), pos_w(pk), pg, om);
         const real B1 = (EWALD ? bn[1] : rr[1]) - om[1] * rr[1];
         const real B2 = (EWALD ? bn[2] : rr[2]) - om[2] * rr[2];
         const real B3 = (EWALD ? bn[3] : rr[3]) - om[3] * rr[3];
         fi += mpole_field(v3(dx, dy, dz), mk, B1, B2, B3, (real)-1);
         // the mutual-field operator of the solver that follows needs exactly these B1, B2 and R for this pair, in every one of
         // its applications: store them once (tlist.cu)
         if (T)
            T[beg + q] = tl_pack(B1, B2, dx, dy, dz);
         if (q < lenu) {      // ... and the preconditioner's tensor of the pairs inside usolve-cutoff (the first cntu entries)
            const real pp = qi.y * qk.y;
            P[beg + q] = tl_pack(pp * (1 - om[1]) * rr[1], pp * (1 - om[2]) * rr[2], dx, dy, dz);
         }
      }
      fi = group_sum3<G>(fi);
      if (l == 0 && act) {
         // fd already holds the reciprocal + self field (Ewald) or is assigned here; the (p - d)
         // delta array starts at zero for the exclusion pass that follows
         if (assign)
            fd[3 * i] = fi.x, fd[3 * i + 1] = fi.y, fd[3 * i + 2] = fi.z;
         else
            fd[3 * i] += fi.x, fd[3 * i + 1] += fi.y, fd[3 * i + 2] += fi.z;
         fpd[3 * i] = 0, fpd[3 * i + 1] = 0, fpd[3 * i + 2] = 0;
      }
   }
}

// d-correction goes to fd, (p - d) correction to the delta array fpd
template <bool TABLE>
__global__ void k_dfield_excl(int nx, int a0, int a1, Box box, real cut2, const PairExcl* __restrict__ ex, const real4* __restrict__ posd, const pos_t* __restrict__ posq,
   const real4* __restrict__ tpj, const real* __restrict__ thlval, int nj, const real4* __restrict__ mp0,
   const real4* __restrict__ mp1, const real2* __restrict__ mp2, real* __restrict__ fd, real* __restrict__ fpd)
{
   int e = blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= nx)
      return;
   PairExcl p = ex[e];
   if (p.d == 0 && p.p == 0)
      return;
   const bool own_i = p.i >= a0 && p.i < a1, own_k = p.k >= a0 && p.k < a1;
   if (!own_i && !own_k)
      return;
   if (!excl_in_rows(box, posd, p.i, p.k, cut2))
      return;
   const pos_t pi = posq[p.i], pk = posq[p.k];
   real dx, dy, dz;
   pair_delta(box, pi, pk, dx, dy, dz);
   real r2 = dx * dx + dy * dy + dz * dz;
   real rinv = r_rsqrt(r2), r = r2 * rinv, rr2 = rinv * rinv;
   real rr[4], om[4];
   radial_coulomb<4>(rinv, rr2, rr);
   real4 qi = tpj[p.i], qk = tpj[p.k];
   real pg = TABLE ? thlval[as_int(qi.w) * nj + as_int(qk.w)] : min(qi.x, qk.x);
   thole_one_minus_lambda<4>(r, pos_w(pi), pos_w(pk), pg, om);
   real L1 = (1 - om[1]) * rr[1], L2 = (1 - om[2]) * rr[2], L3 = (1 - om[3]) * rr[3];
   V3 R = v3(dx, dy, dz);
   Mpole mi = load_mpole(mp0, mp1, mp2, p.i), mk = load_mpole(mp0, mp1, mp2, p.k);
   V3 ei = mpole_field(R, mk, L1, L2, L3, (real)-1);   // unit-scale damped Coulomb field at i
   V3 ek = mpole_field(R, mi, L1, L2, L3, (real)1);
   if (p.d != 0) {
      if (own_i) atomic_real3(fd, p.i, p.d * ei);
      if (own_k) atomic_real3(fd, p.k, p.d * ek);
   }
   real dp = p.p - p.d;
   if (dp != 0) {
      if (own_i) atomic_real3(fpd, p.i, dp * ei);
      if (own_k) atomic_real3(fpd, p.k, dp * ek);
   }
}

// -------------------------------------------------------------------------------------------
// sparse preconditioner: z += alpha_i alpha_k T_thole(r) r_k  over the pairs inside usolve-cutoff
// (the first cntu entries of every row)
// -------------------------------------------------------------------------------------------
template <bool TABLE, int G, bool PLIST>
__global__ void __launch_bounds__(ROWS_BLOCK) k_precond_rows(int a0, int a1, int ntot, Box box, real udiag, const int* __restrict__ vstart,
   const int* __restrict__ cntu, const int* __restrict__ nbr, const pos_t* __restrict__ posq, const real4* __restrict__ tpj,
   const real* __restrict__ thlval, int nj, const real4* __restrict__ Rv, real4* __restrict__ Z, double* __restrict__ slot,
   const int* __restrict__ skip, PcgTest T, const real4* __restrict__ PL)
{
   if (skip && skip[1])
      return;
   // Stopping rule of induceMutualPcg1 (src/cu/amoeba/pcg.cu:147-165) on r.r of this iteration: when
   // it is met the kernel applies the peek step u += peek*alpha*r instead of z = M r, and the last
   // CTA raises the stop flag (so no CTA of this grid can see it early).
   bool done = false;
   if (T.itp) {      // device-side loop: this kernel belongs to iteration *itp, its slots follow from it
      T.it = *T.itp;
      T.slot = pcg_slot_of(T.slot, T.itp);
      slot = slot ? pcg_slot_of(slot, T.itp) : nullptr;
   }
   if (T.it > 0) {
      double rr2_[2];
      pcg_q_block<2>(T.slot, 4, rr2_);
      double e = fmax(rr2_[0], rr2_[1]);
      double eps = (double)T.debye * sqrt(e / ntot);
      done = eps < (double)T.poleps;
      if (T.it < T.miniter)
         done = false;
      if (T.it >= T.politer)
         done = true;
      if (blockIdx.x == 0 && threadIdx.x == 0) {
         T.result[0] = eps;
         T.result[1] = (double)T.it;
      }
   }
   if (done) {
      for (int s = a0 + blockIdx.x * blockDim.x + threadIdx.x; s < a1; s += gridDim.x * blockDim.x) {
         V3 rd, rp;
         load_dp(Rv, s, rd, rp);
         real term = T.pcgpeek * tpj[s].y;
         T.ud[3 * s] += term * rd.x, T.ud[3 * s + 1] += term * rd.y, T.ud[3 * s + 2] += term * rd.z;
         T.up[3 * s] += term * rp.x, T.up[3 * s + 1] += term * rp.y, T.up[3 * s + 2] += term * rp.z;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
         __threadfence();
         unsigned t = atomicAdd((unsigned*)&T.flags[3], 1u);
         if (t == gridDim.x - 1) {
            T.flags[2] = T.it;
            T.flags[3] = 0;
            __threadfence();
            T.flags[1] = 1;
            if (T.cond)
               cudaGraphSetConditional((cudaGraphConditionalHandle)T.cond, 0);      // leave the WHILE node (pcg.cu)
         }
      }
      return;
   }
   double dot_d = 0, dot_p = 0;
   ROWS_FOREACH_ATOM(G, a0, a1, i, l, act)
   {
      const pos_t pi = posq[i];
      const real4 qi = tpj[i];
      const int beg = vstart[i];
      const int len = (act && cntu) ? cntu[i] : 0;
      V3 zdi = v3(0, 0, 0), zpi = v3(0, 0, 0);
      if (PLIST) {
         // stored tensors (written by the permanent-field rows of this induce(), tlist.cuh): 16 + 4 bytes streamed and one
         // 32-byte gather per pair instead of recomputing the Thole-damped tensor
         for (int q = l; q < len; q += G) {
            const int k = nbr[beg + q] & ROW_INDEX_MASK;
            const real4 t = tl_ld(PL + beg + q);
            real4 ra, rb;
            tl_gather(Rv, k, ra, rb);
            tl_apply(t, ra, rb, zdi, zpi);
         }
      } else
      for (int q = l; q < len; q += G) {
         const int k = nbr[beg + q] & ROW_INDEX_MASK;
         const pos_t pk = posq[k];
         const real4 qk = tpj[k];
         V3 a, b;
         load_dp(Rv, k, a, b);
         real dx, dy, dz;
         pair_delta(box, pi, pk, dx, dy, dz);
         const real r2 = dx * dx + dy * dy + dz * dz;
         const real rinv = r_rsqrt(r2);
         const real r = r2 * rinv, rr2 = rinv * rinv;
         real rr[3], om[3];
         radial_coulomb<3>(rinv, rr2, rr);
         const real pg = TABLE ? thlval[as_int(qi.w) * nj + as_int(qk.w)] : min(qi.x, qk.x);
         thole_one_minus_lambda<3>(r, pos_w(pi), pos_w(pk), pg, om);
         const real pp = qi.y * qk.y;
         const real B1 = pp * (1 - om[1]) * rr[1], B2 = pp * (1 - om[2]) * rr[2];
         const V3 R = v3(dx, dy, dz);
         zdi += dipole_field(R, a, B1, B2);
         zpi += dipole_field(R, b, B1, B2);
      }
      zdi = group_sum3<G>(zdi);
      zpi = group_sum3<G>(zpi);
      if (l == 0 && act) {
         V3 rd, rp;
         load_dp(Rv, i, rd, rp);
         const real dg = udiag * qi.y;
         zdi += dg * rd;
         zpi += dg * rp;
         store_dp(Z, i, zdi, zpi);
         dot_d += (double)rd.x * zdi.x + (double)rd.y * zdi.y + (double)rd.z * zdi.z;
         dot_p += (double)rp.x * zpi.x + (double)rp.y * zpi.y + (double)rp.z * zpi.z;
      }
   }
   if (slot)
      pcg_block_add2(dot_d, dot_p, slot, 0, 1);
}

template <bool TABLE>
__global__ void k_precond_excl(int nx, int a0, int a1, Box box, real cut2, const PairExcl* __restrict__ ex, const real4* __restrict__ posd, const pos_t* __restrict__ posq,
   const real4* __restrict__ tpj, const real* __restrict__ thlval, int nj, const real4* __restrict__ Rv, real4* __restrict__ Z,
   const int* __restrict__ skip)
{
   if (skip && skip[1])
      return;
   int e = blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= nx)
      return;
   PairExcl p = ex[e];
   if (p.u == 0)
      return;
   const bool own_i = p.i >= a0 && p.i < a1, own_k = p.k >= a0 && p.k < a1;
   if (!own_i && !own_k)
      return;
   if (!excl_in_rows(box, posd, p.i, p.k, cut2))
      return;
   const pos_t pi = posq[p.i], pk = posq[p.k];
   real dx, dy, dz;
   pair_delta(box, pi, pk, dx, dy, dz);
   real r2 = dx * dx + dy * dy + dz * dz;
   real rinv = r_rsqrt(r2), r = r2 * rinv, rr2 = rinv * rinv;
   real rr[3], om[3];
   radial_coulomb<3>(rinv, rr2, rr);
   real4 qi = tpj[p.i], qk = tpj[p.k];
   real pg = TABLE ? thlval[as_int(qi.w) * nj + as_int(qk.w)] : min(qi.x, qk.x);
   thole_one_minus_lambda<3>(r, pos_w(pi), pos_w(pk), pg, om);
   real pp = qi.y * qk.y * p.u;
   real B1 = pp * (1 - om[1]) * rr[1], B2 = pp * (1 - om[2]) * rr[2];
   V3 R = v3(dx, dy, dz);
   V3 rdi, rpi, rdk, rpk;
   load_dp(Rv, p.i, rdi, rpi);
   load_dp(Rv, p.k, rdk, rpk);
   if (own_i)
      atomic_dp(Z, p.i, dipole_field(R, rdk, B1, B2), dipole_field(R, rpk, B1, B2));
   if (own_k)
      atomic_dp(Z, p.k, dipole_field(R, rdi, B1, B2), dipole_field(R, rpi, B1, B2));
}

// partial R.Z after an exclusion pass changed Z (only needed when u-scale exclusions exist)
__global__ void k_dot_dp(int n, const real4* __restrict__ A, const real4* __restrict__ Bv, double* __restrict__ slot,
   const int* __restrict__ skip, const int* __restrict__ itp)
{
   slot = pcg_slot_of(slot, itp);
   if (skip && skip[1])
      return;
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   double x = 0, y = 0;
   if (s < n) {
      V3 ad, ap, bd, bp;
      load_dp(A, s, ad, ap);
      load_dp(Bv, s, bd, bp);
      x = (double)ad.x * bd.x + (double)ad.y * bd.y + (double)ad.z * bd.z;
      y = (double)ap.x * bp.x + (double)ap.y * bp.y + (double)ap.z * bp.z;
   }
   pcg_block_add2(x, y, slot, 0, 1);
}

} // namespace

#define UF_G 16
#define DF_G 16
#define PC_G 8

void apx_ufield_real_dp(apx_ctx* c, cudaStream_t st, const real4* U, real4* F)
{
   RowList& L = c->rows;
   real cut = (real)c->opt.cutoff;
   bool ew = c->opt.use_ewald != 0;
   bool tb = c->thole_table != 0;
   // 8 CTAs per SM: the rows run beside the PME spread/FFT chain of the other stream and must leave
   // room for its CTAs (a grid that fills every SM first delays the spread by the length of a wave)
   int grid = rows_grid<UF_G>(c, c->uf_ctas);
   // unused dynamic shared memory caps how many of these long-lived CTAs one SM can hold, whatever order the
   // hardware fills the SMs in: the FFT CTAs of the other stream (512 threads, 33 KB) must always find room
   const size_t uf_smem = (size_t)c->uf_smem_kb * 1024;
#define LAUNCH_UF(E, T)                                                                                                   \
   k_ufield_rows<E, T, UF_G><<<grid, ROWS_BLOCK, uf_smem, st>>>(c->a0, c->a1, c->box, (real)c->opt.aewald, L.vstart, L.cnt, L.nbr, c->posq, c->tpj,  \
      c->thlval, c->opt.njpolar, U, F, c->skip)
   // stored-tensor operator (tlist.cu): the first application after the positions changed writes the tensors
   const bool tl = apx_tlist_usable(c);
   if (tl && !c->tl_valid)
      apx_tlist_build(c, st);
   // device-time the dominant kernel: one event pair per launch, read back by induce()
   int slot = -1;
   const bool ext_iter = c->capturing && c->uf_ext_iter && c->uf_ev.size() >= 4;      // one iteration of a captured batch (pcg.cu): slots 2,3
   const bool ext = ext_iter || (c->capturing && c->graph_key_open >= 0 && (c->graph_key_open & 0x2000) && c->uf_ev.size() >= 2);
   if (ext) {
      // inside the captured prologue of induce() (slots 0,1) or a captured iteration (slots 2,3): external event nodes, so
      // every replay still times this launch
      slot = ext_iter ? 2 : 0;
      CUDA_CHECK(cudaEventRecordWithFlags(c->uf_ev[slot], st, cudaEventRecordExternal));
   } else if (!c->capturing && c->uf_used + 2 <= (int)c->uf_ev.size()) {
      slot = c->uf_used;
      c->uf_used += 2;
      cudaEventRecord(c->uf_ev[slot], st);
   }
   if (tl) {
      apx_ufield_tlist(c, st, U, F);
   } else if (!tb && c->use_records) {
      // records of every atom a row can reach: the whole system (halo atoms included on several GPUs)
      c->uf_rec.ensure(3 * (size_t)c->npad);
      k_uf_records<<<(c->n + 255) / 256, 256, 0, st>>>(c->n, c->posq, c->tpj, U, c->uf_rec, c->skip);
      APX_COUNT_LAUNCH(c);
      if (apx_staged_usable(c) && c->n >= c->staged_min_atoms)
         apx_ufield_staged(c, st, F);      // records staged in shared memory by bulk copies (staged.cu)
      else if (ew)
         k_ufield_rows_rec<true, UF_G><<<grid, ROWS_BLOCK, uf_smem, st>>>(c->a0, c->a1, c->box, (real)c->opt.aewald, L.vstart, L.cnt, L.nbr, c->uf_rec, F, c->skip);
      else
         k_ufield_rows_rec<false, UF_G><<<grid, ROWS_BLOCK, uf_smem, st>>>(c->a0, c->a1, c->box, (real)c->opt.aewald, L.vstart, L.cnt, L.nbr, c->uf_rec, F, c->skip);
   } else if (ew && tb) LAUNCH_UF(true, true);
   else if (ew) LAUNCH_UF(true, false);
   else if (tb) LAUNCH_UF(false, true);
   else LAUNCH_UF(false, false);
   if (slot >= 0 && ext)
      CUDA_CHECK(cudaEventRecordWithFlags(c->uf_ev[slot + 1], st, cudaEventRecordExternal));
   else if (slot >= 0)
      cudaEventRecord(c->uf_ev[slot + 1], st);
   APX_COUNT_LAUNCH(c);
#undef LAUNCH_UF
   if (c->nexcl_u > 0) {
      int g = (c->nexcl + 127) / 128;
      if (tb)
         k_ufield_excl<true><<<g, 128, 0, st>>>(c->nexcl, c->a0, c->a1, c->box, cut * cut, c->excl_s, c->posd, c->posq, c->tpj, c->thlval, c->opt.njpolar, U, F);
      else
         k_ufield_excl<false><<<g, 128, 0, st>>>(c->nexcl, c->a0, c->a1, c->box, cut * cut, c->excl_s, c->posd, c->posq, c->tpj, c->thlval, c->opt.njpolar, U, F);
      APX_COUNT_LAUNCH(c);
   }
}

// fd accumulates (assign = false) or receives (assign = true) the common field + d corrections; fpd receives the (p - d) delta
// only.  With the stored-tensor operator in use the row pass also writes the pair tensors of the mutual field (tlist.cu).
void apx_dfield_real(apx_ctx* c, cudaStream_t st, real* fd, real* fpd, bool assign)
{
   RowList& L = c->rows;
   real cut = (real)c->opt.cutoff;
   bool ew = c->opt.use_ewald != 0;
   bool tb = c->thole_table != 0;
   int grid = rows_grid<DF_G>(c);
   real4* T = (apx_tlist_usable(c) && c->opt.use_polar && c->opt.poltyp_mutual) ? c->tl_T.p : nullptr;
   real4* P = (T && c->opt.pcgprec && c->opt.usolve_cutoff > 0) ? c->tl_P.p : nullptr;
#define LAUNCH_DF(E, T_)                                                                                                  \
   k_dfield_rows<E, T_, DF_G><<<grid, ROWS_BLOCK, 0, st>>>(c->a0, c->a1, c->box, (real)c->opt.aewald, L.vstart, L.cnt, L.nbr, c->posq,  \
      c->tpj, c->thlval, c->opt.njpolar, c->mp0, c->mp1, c->mp2, fd, fpd, assign ? 1 : 0, T, P, L.cntu)
   // (rows may all be empty for a tiny system: the kernel still initialises fd / fpd)
   if (ew && tb) LAUNCH_DF(true, true);
   else if (ew) LAUNCH_DF(true, false);
   else if (tb) LAUNCH_DF(false, true);
   else LAUNCH_DF(false, false);
   APX_COUNT_LAUNCH(c);
#undef LAUNCH_DF
   if (T)
      c->tl_valid = 1, c->tl_p_valid = P ? 1 : 0;
   if (c->nexcl > 0) {
      int g = (c->nexcl + 127) / 128;
      if (tb)
         k_dfield_excl<true><<<g, 128, 0, st>>>(c->nexcl, c->a0, c->a1, c->box, cut * cut, c->excl_s, c->posd, c->posq, c->tpj, c->thlval,
            c->opt.njpolar, c->mp0, c->mp1, c->mp2, fd, fpd);
      else
         k_dfield_excl<false><<<g, 128, 0, st>>>(c->nexcl, c->a0, c->a1, c->box, cut * cut, c->excl_s, c->posd, c->posq, c->tpj, c->thlval,
            c->opt.njpolar, c->mp0, c->mp1, c->mp2, fd, fpd);
      APX_COUNT_LAUNCH(c);
   }
}

// Z = M R: diagonal (udiag*alpha, or alpha alone without the sparse part) + short-range off-diagonal
// blocks (sparsePrecondApply / diagPrecond, src/amoeba/induce.cpp:12-25); partial R.Z into slot.
void apx_precond_dp(apx_ctx* c, const real4* Rv, real4* Z, double* slot, const PcgTest* test)
{
   PcgTest T;
   if (test)
      T = *test;
   bool sparse = c->opt.pcgprec && c->opt.usolve_cutoff > 0;
   real udiag = sparse ? (real)c->opt.uaccel : (real)1;
   RowList& L = c->rows;
   real cut = (real)c->opt.usolve_cutoff;
   bool tb = c->thole_table != 0;
   bool excl = sparse && c->nexcl_u > 0;
   int grid = rows_grid<PC_G>(c);
   const int* cu = sparse ? L.cntu.p : nullptr;
   double* s1 = excl ? nullptr : slot;
   const bool pl = sparse && apx_tlist_usable(c) && c->tl_p_valid && c->tl_P.p;
   if (pl)
      k_precond_rows<false, PC_G, true><<<grid, ROWS_BLOCK, 0, c->stream>>>(c->a0, c->a1, c->n, c->box, udiag, L.vstart, cu, L.nbr, c->posq, c->tpj, c->thlval,
         c->opt.njpolar, Rv, Z, s1, c->skip, T, c->tl_P);
   else if (tb)
      k_precond_rows<true, PC_G, false><<<grid, ROWS_BLOCK, 0, c->stream>>>(c->a0, c->a1, c->n, c->box, udiag, L.vstart, cu, L.nbr, c->posq, c->tpj, c->thlval,
         c->opt.njpolar, Rv, Z, s1, c->skip, T, nullptr);
   else
      k_precond_rows<false, PC_G, false><<<grid, ROWS_BLOCK, 0, c->stream>>>(c->a0, c->a1, c->n, c->box, udiag, L.vstart, cu, L.nbr, c->posq, c->tpj, c->thlval,
         c->opt.njpolar, Rv, Z, s1, c->skip, T, nullptr);
   APX_COUNT_LAUNCH(c);
   if (excl) {
      int g = (c->nexcl + 127) / 128;
      if (tb)
         k_precond_excl<true><<<g, 128, 0, c->stream>>>(c->nexcl, c->a0, c->a1, c->box, cut * cut, c->excl_s, c->posd, c->posq, c->tpj, c->thlval, c->opt.njpolar,
            Rv, Z, c->skip);
      else
         k_precond_excl<false><<<g, 128, 0, c->stream>>>(c->nexcl, c->a0, c->a1, c->box, cut * cut, c->excl_s, c->posd, c->posq, c->tpj, c->thlval, c->opt.njpolar,
            Rv, Z, c->skip);
      APX_COUNT_LAUNCH(c);
      if (slot) {
         k_dot_dp<<<(c->a1 - c->a0 + 127) / 128, 128, 0, c->stream>>>(c->a1 - c->a0, Rv + 2 * c->a0, Z + 2 * c->a0, slot, c->skip, T.itp);
         APX_COUNT_LAUNCH(c);
      }
   }
}

// plain-array front end (apx_precond of the C ABI)
void apx_precond_apply(apx_ctx* c, const real* rd, const real* rp, real* zd, real* zp)
{
   apx_pack_dp(c, rd, rp, c->pk_r);
   if (c->dist.on)
      apx_dist_halo(c, c->pk_r, c->stream);
   apx_precond_dp(c, c->pk_r, c->pk_z, nullptr);
   apx_unpack_dp(c, c->pk_z, zd, zp);
}
