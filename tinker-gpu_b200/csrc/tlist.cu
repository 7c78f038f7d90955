// Stored-tensor form of the real-space CG operator (ufieldEwaldReal_cu / ufieldNonEwald_cu of the reference,
// src/cu/amoeba/field.cu:122-137; the preconditioner of src/cu/amoeba/precond.cu:28-43).
//
// Inside one induce() the positions do not move, so the dipole-dipole tensor of a pair,
//      T_ik u = B2 (R.u) R - B1 u,     B1 = bn1 - (1-lambda3) rr1,  B2 = bn2 - (1-lambda5) rr2,
// is the same in every application of the operator: 7-11 per induce().  The row kernels of field.cu recompute it each time
// (erfc, two exponentials, rsqrt: ~130 warp instructions per directed pair; ncu: issue 54 %, L1 86 %, FMA 37 % of peak -- the
// operator was the dominant kernel at 6-13 % of the FP32 roofline).  Here the FIRST application of an induce() writes one real4
// per directed pair, in the layout of the compacted rows (rows.nbr):
//      t = { B1 , sqrt|B2| R }       with the sign of B2 in the mantissa LSB of B1 (B2 < 0 only for heavily damped close pairs)
// so that  T u = sgn (S.u) S - B1 u  costs 18 FMA, and every later application STREAMS 16 + 4 bytes per directed pair
// (coalesced: a lane group reads consecutive entries of its atom's row) plus a 32-byte gather of the neighbour's packed
// dipole pair.  The operator turns from an instruction-issue-bound kernel into a bandwidth-bound sparse matrix-vector
// product: L2-resident at dhfr2 (3.3 M directed pairs = 53 MB), HBM-bound at 1 M atoms (146 M directed pairs = 2.3 GB).
// The preconditioner gets the same treatment for the pairs inside usolve-cutoff (the first cntu entries of every row).
//
// Results agree with the row kernels to rounding (same B1/B2 arithmetic, one extra square root and product per component);
// i-side sums are reduced by shuffles in a fixed order, no atomics.
#include "apx_internal.h"
#include "pairmath.cuh"
#include "rows.cuh"
#include "dp.cuh"
#include "tlist.cuh"
#include <algorithm>

namespace {
__device__ __forceinline__ int tl_as_int(real w)
{
#ifdef APX_DOUBLE
   return (int)__double_as_longlong(w);
#else
   return __float_as_int(w);
#endif
}
// ---- build: one pass over the compacted rows, same lane groups as the operator -----------------------------------------------
template <bool EWALD, bool TABLE, bool PRECOND, int G>
__global__ void __launch_bounds__(ROWS_BLOCK) k_tlist_build(int a0, int a1, Box box, real aewald, const int* __restrict__ vstart,
   const int* __restrict__ cnt, const int* __restrict__ cntu, const int* __restrict__ nbr, const pos_t* __restrict__ posq,
   const real4* __restrict__ tpj, const real* __restrict__ thlval, int nj, real4* __restrict__ T, real4* __restrict__ P)
{
   ROWS_FOREACH_ATOM(G, a0, a1, i, l, act)
   {
      const pos_t pi = posq[i];
      const real4 qi = tpj[i];
      const int beg = vstart[i];
      const int len = act ? cnt[i] : 0;
      const int lenu = (PRECOND && act) ? cntu[i] : 0;
      for (int q = l; q < len; q += G) {
         const int k = nbr[beg + q] & ROW_INDEX_MASK;
         const pos_t pk = posq[k];
         const real4 qk = tpj[k];
         real dx, dy, dz;
         pair_delta(box, pi, pk, dx, dy, dz);
         const real r2 = dx * dx + dy * dy + dz * dz;
         const real rinv = r_rsqrt(r2);
         const real r = r2 * rinv, rr2 = rinv * rinv;
         real rr[3], bn[3], om[3];
         radial_coulomb<3>(rinv, rr2, rr);
         if (EWALD)
            radial_ewald<3>(r, rinv, rr2, aewald, bn);
         const real pg = TABLE ? thlval[tl_as_int(qi.w) * nj + tl_as_int(qk.w)] : min(qi.x, qk.x);
         thole_one_minus_lambda<3>(r, pos_w(pi), pos_w(pk), pg, om);
         const real B1 = (EWALD ? bn[1] : rr[1]) - om[1] * rr[1];
         const real B2 = (EWALD ? bn[2] : rr[2]) - om[2] * rr[2];
         T[beg + q] = tl_pack(B1, B2, dx, dy, dz);
         if (PRECOND && q < lenu) {
            const real pp = qi.y * qk.y;
            P[beg + q] = tl_pack(pp * (1 - om[1]) * rr[1], pp * (1 - om[2]) * rr[2], dx, dy, dz);
         }
      }
   }
}

// Memory-level parallelism.  ptxas sinks every load of an unrolled body next to its first use (40 registers, one entry in
// flight per lane; volatile asm loads and compiler barriers do not stop it), which serialises 2 x UNROLL memory round trips per
// lane.  MLP = true makes every FMA of the body depend on ALL its loads through real data flow: the loaded words are OR-ed
// together, AND-ed with a kernel argument that is zero at run time, and the (zero) result is XOR-ed into the first factor of
// each entry -- 9 logic instructions per 4 entries, and all 16 loads are in flight before the first FMA can issue.
template <int G, int UNROLL, bool MLP>
__device__ __forceinline__ void tl_row(const int* __restrict__ nbr, const real4* __restrict__ T, const real4* __restrict__ U, int beg,
   int len, int l, unsigned zero, V3& fd, V3& fp)
{
   int q = l;
   for (; q + (UNROLL - 1) * G < len; q += UNROLL * G) {
      int k[UNROLL];
      real4 t[UNROLL], ua[UNROLL], ub[UNROLL];
      #pragma unroll
      for (int j = 0; j < UNROLL; ++j)
         k[j] = nbr[beg + q + j * G] & ROW_INDEX_MASK;
      #pragma unroll
      for (int j = 0; j < UNROLL; ++j)
         t[j] = tl_ld(T + beg + q + j * G);
      #pragma unroll
      for (int j = 0; j < UNROLL; ++j)
         tl_gather(U, k[j], ua[j], ub[j]);
#ifndef APX_DOUBLE
      if (MLP) {
         unsigned x = 0;
         #pragma unroll
         for (int j = 0; j < UNROLL; ++j)
            x |= __float_as_uint(t[j].y) | __float_as_uint(ua[j].x);
         x &= zero;
         #pragma unroll
         for (int j = 0; j < UNROLL; ++j)
            t[j].y = __uint_as_float(__float_as_uint(t[j].y) ^ x);
      }
#endif
      #pragma unroll
      for (int j = 0; j < UNROLL; ++j)
         tl_apply(t[j], ua[j], ub[j], fd, fp);
   }
   for (; q < len; q += G) {
      const int k = nbr[beg + q] & ROW_INDEX_MASK;
      const real4 t = tl_ld(T + beg + q);
      real4 ua, ub;
      tl_gather(U, k, ua, ub);
      tl_apply(t, ua, ub, fd, fp);
   }
}

template <int G, int UNROLL, bool MLP>
__global__ void __launch_bounds__(ROWS_BLOCK) k_ufield_tl(int a0, int a1, const int* __restrict__ vstart, const int* __restrict__ cnt,
   const int* __restrict__ nbr, const real4* __restrict__ T, const real4* __restrict__ U, real4* __restrict__ F,
   const int* __restrict__ skip, unsigned zero)
{
   if (skip && skip[1])
      return;
   ROWS_FOREACH_ATOM(G, a0, a1, i, l, act)
   {
      const int beg = vstart[i];
      const int len = act ? cnt[i] : 0;
      V3 fd = v3(0, 0, 0), fp = v3(0, 0, 0);
      tl_row<G, UNROLL, MLP>(nbr, T, U, beg, len, l, zero, fd, fp);
      fd = group_sum3<G>(fd);
      fp = group_sum3<G>(fp);
      if (l == 0 && act)
         store_dp(F, i, fd, fp);
   }
}
} // namespace

bool apx_tlist_usable(const apx_ctx* c)
{
   return c->tlist_on && c->tl_T.p != nullptr;
}

void apx_tlist_reserve(apx_ctx* c)
{
   if (!c->tlist_on)
      return;
   c->tl_T.ensure((size_t)c->rows.nverlet + 32);
   if (c->opt.use_polar && c->opt.pcgprec && c->opt.usolve_cutoff > 0)
      c->tl_P.ensure((size_t)c->rows.nverlet + 32);
   c->tl_valid = 0, c->tl_p_valid = 0;
}

#define TL_G 8
void apx_tlist_build(apx_ctx* c, cudaStream_t st)
{
   RowList& L = c->rows;
   const bool ew = c->opt.use_ewald != 0, tb = c->thole_table != 0;
   const int grid = rows_grid<TL_G>(c, 16);
#define LAUNCH_TB(E, T_)                                                                                                  \
   k_tlist_build<E, T_, false, TL_G><<<grid, ROWS_BLOCK, 0, st>>>(c->a0, c->a1, c->box, (real)c->opt.aewald, L.vstart, L.cnt, L.cntu, L.nbr, c->posq, \
      c->tpj, c->thlval, c->opt.njpolar, c->tl_T, c->tl_P)
   if (ew && tb) LAUNCH_TB(true, true);
   else if (ew) LAUNCH_TB(true, false);
   else if (tb) LAUNCH_TB(false, true);
   else LAUNCH_TB(false, false);
#undef LAUNCH_TB
   APX_COUNT_LAUNCH(c);
   c->tl_valid = 1;
}

void apx_ufield_tlist(apx_ctx* c, cudaStream_t st, const real4* U, real4* F)
{
   RowList& L = c->rows;
   // APX_TL_MODE (A/B on hardware): bit 0 = forced memory-level parallelism, bit 1 = 16 lanes per atom instead of 8
   static const int mode = getenv("APX_TL_MODE") ? atoi(getenv("APX_TL_MODE")) : 1;
   // CTAs per SM (APX_TL_CTAS): the operator runs beside the spread -> FFT chain of the main stream, which is the critical
   // path and whose 512-thread FFT CTAs need a quarter of an SM's registers at once; a grid that fills every SM makes them
   // wait for it to drain (profiles/r02g_trace_md.txt: the forward FFT started 8 us late and ran at a third of its speed)
   // (large systems are bandwidth bound in every kernel: there the operator wants all the bytes in flight it can get)
   static const int ctas_env = getenv("APX_TL_CTAS") ? std::max(1, atoi(getenv("APX_TL_CTAS"))) : 0;
   // Decomposed runs: NO cap -- a capped grid is a persistent one (every CTA strides over atoms until the kernel ends), and
   // the exchange kernels and FFTs of the main stream then wait for the whole operator although their stream has priority
   // (profiles/r02m_trace_water1m_n2.txt: 326 us of operator in front of a 600 us PME chain that is mostly NVLink traffic).
   // Short-lived CTAs hand their SM slots to the higher-priority work as they retire.
   const int ctas = ctas_env ? ctas_env : (c->dist.on ? (1 << 20) : (c->n >= 200000 ? 16 : 6));
#define LAUNCH_TL(G_, M_)                                                                                                 \
   k_ufield_tl<G_, 4, M_><<<rows_grid<G_>(c, ctas), ROWS_BLOCK, 0, st>>>(c->a0, c->a1, L.vstart, L.cnt, L.nbr, c->tl_T, U, F, c->skip, 0u)
   switch (mode & 3) {
   case 0: LAUNCH_TL(8, false); break;
   case 1: LAUNCH_TL(8, true); break;
   case 2: LAUNCH_TL(16, false); break;
   default: LAUNCH_TL(16, true); break;
   }
#undef LAUNCH_TL
}
