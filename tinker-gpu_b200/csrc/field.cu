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

namespace {
__device__ __forceinline__ int as_int(real w)
{
#ifdef APX_DOUBLE
   return (int)__double_as_longlong(w);
#else
   return __float_as_int(w);
#endif
}

// -------------------------------------------------------------------------------------------
// ufield: field of (ud, up) at every atom, Ewald real space or plain Thole-damped Coulomb
// -------------------------------------------------------------------------------------------
template <bool EWALD, bool TABLE, int G>
__global__ void __launch_bounds__(ROWS_BLOCK) k_ufield_rows(int n, Box box, real aewald, const int* __restrict__ vstart,
   const int* __restrict__ cnt, const int* __restrict__ nbr, const real4* __restrict__ posd, const real4* __restrict__ tpj,
   const real* __restrict__ thlval, int nj, const real* __restrict__ ud, const real* __restrict__ up, real* __restrict__ fd,
   real* __restrict__ fp, const int* __restrict__ skip)
{
   if (skip && skip[1])
      return;
   ROWS_FOREACH_ATOM(G, n, i, l, act)
   {
      const real4 pi = posd[i];
      const real4 qi = tpj[i];
      const int beg = vstart[i];
      const int len = act ? cnt[i] : 0;
      V3 fdi = v3(0, 0, 0), fpi = v3(0, 0, 0);
      for (int q = l; q < len; q += G) {
         const int k = nbr[beg + q];
         const real4 pk = posd[k];
         const real4 qk = tpj[k];
         const V3 a = v3(ud[3 * k], ud[3 * k + 1], ud[3 * k + 2]);
         const V3 b = v3(up[3 * k], up[3 * k + 1], up[3 * k + 2]);
         real dx = pk.x - pi.x, dy = pk.y - pi.y, dz = pk.z - pi.z;
         apx_image(box, dx, dy, dz);
         const real r2 = dx * dx + dy * dy + dz * dz;
         const real rinv = r_rsqrt(r2);
         const real r = r2 * rinv, rr2 = rinv * rinv;
         real rr[3], bn[3], om[3];
         radial_coulomb<3>(rinv, rr2, rr);
         if (EWALD)
            radial_ewald<3>(r, rinv, rr2, aewald, bn);
         const real pg = TABLE ? thlval[as_int(qi.w) * nj + as_int(qk.w)] : min(qi.x, qk.x);
         thole_one_minus_lambda<3>(r, pi.w, pk.w, pg, om);
         const real B1 = (EWALD ? bn[1] : rr[1]) - om[1] * rr[1];
         const real B2 = (EWALD ? bn[2] : rr[2]) - om[2] * rr[2];
         const V3 R = v3(dx, dy, dz);
         fdi += dipole_field(R, a, B1, B2);
         fpi += dipole_field(R, b, B1, B2);
      }
      fdi = group_sum3<G>(fdi);
      fpi = group_sum3<G>(fpi);
      if (l == 0 && act) {
         fd[3 * i] += fdi.x, fd[3 * i + 1] += fdi.y, fd[3 * i + 2] += fdi.z;
         fp[3 * i] += fpi.x, fp[3 * i + 1] += fpi.y, fp[3 * i + 2] += fpi.z;
      }
   }
}

// exclusion pass for ufield (only pairs whose u-scale != 1; empty for stock AMOEBA)
template <bool TABLE>
__global__ void k_ufield_excl(int nx, Box box, real cut2, const PairExcl* __restrict__ ex, const real4* __restrict__ posd,
   const real4* __restrict__ tpj, const real* __restrict__ thlval, int nj, const real* __restrict__ ud, const real* __restrict__ up,
   real* __restrict__ fd, real* __restrict__ fp)
{
   int e = blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= nx)
      return;
   PairExcl p = ex[e];
   if (p.u == 0)
      return;
   real4 pi = posd[p.i], pk = posd[p.k];
   real dx = pk.x - pi.x, dy = pk.y - pi.y, dz = pk.z - pi.z;
   apx_image(box, dx, dy, dz);
   real r2 = dx * dx + dy * dy + dz * dz;
   if (r2 > cut2)
      return;
   real rinv = r_rsqrt(r2), r = r2 * rinv, rr2 = rinv * rinv;
   real rr[3], om[3];
   radial_coulomb<3>(rinv, rr2, rr);
   real4 qi = tpj[p.i], qk = tpj[p.k];
   real pg = TABLE ? thlval[as_int(qi.w) * nj + as_int(qk.w)] : min(qi.x, qk.x);
   thole_one_minus_lambda<3>(r, pi.w, pk.w, pg, om);
   real B1 = p.u * (1 - om[1]) * rr[1], B2 = p.u * (1 - om[2]) * rr[2];
   V3 R = v3(dx, dy, dz);
   V3 udi = v3(ud[3 * p.i], ud[3 * p.i + 1], ud[3 * p.i + 2]), upi = v3(up[3 * p.i], up[3 * p.i + 1], up[3 * p.i + 2]);
   V3 udk = v3(ud[3 * p.k], ud[3 * p.k + 1], ud[3 * p.k + 2]), upk = v3(up[3 * p.k], up[3 * p.k + 1], up[3 * p.k + 2]);
   atomic_real3(fd, p.i, dipole_field(R, udk, B1, B2));
   atomic_real3(fp, p.i, dipole_field(R, upk, B1, B2));
   atomic_real3(fd, p.k, dipole_field(R, udi, B1, B2));
   atomic_real3(fp, p.k, dipole_field(R, upi, B1, B2));
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
__global__ void __launch_bounds__(ROWS_BLOCK) k_dfield_rows(int n, Box box, real aewald, const int* __restrict__ vstart,
   const int* __restrict__ cnt, const int* __restrict__ nbr, const real4* __restrict__ posd, const real4* __restrict__ tpj,
   const real* __restrict__ thlval, int nj, const real4* __restrict__ mp0, const real4* __restrict__ mp1,
   const real2* __restrict__ mp2, real* __restrict__ fd)
{
   ROWS_FOREACH_ATOM(G, n, i, l, act)
   {
      const real4 pi = posd[i];
      const real4 qi = tpj[i];
      const int beg = vstart[i];
      const int len = act ? cnt[i] : 0;
      V3 fi = v3(0, 0, 0);
      for (int q = l; q < len; q += G) {
         const int k = nbr[beg + q];
         const real4 pk = posd[k];
         const real4 qk = tpj[k];
         const Mpole mk = load_mpole(mp0, mp1, mp2, k);
         real dx = pk.x - pi.x, dy = pk.y - pi.y, dz = pk.z - pi.z;
         apx_image(box, dx, dy, dz);
         const real r2 = dx * dx + dy * dy + dz * dz;
         const real rinv = r_rsqrt(r2);
         const real r = r2 * rinv, rr2 = rinv * rinv;
         real rr[4], bn[4], om[4];
         radial_coulomb<4>(rinv, rr2, rr);
         if (EWALD)
            radial_ewald<4>(r, rinv, rr2, aewald, bn);
         const real pg = TABLE ? thlval[as_int(qi.w) * nj + as_int(qk.w)] : min(qi.x, qk.x);
         thole_one_minus_lambda<4>(r, pi.w, pk.w, pg, om);
         const real B1 = (EWALD ? bn[1] : rr[1]) - om[1] * rr[1];
         const real B2 = (EWALD ? bn[2] : rr[2]) - om[2] * rr[2];
         const real B3 = (EWALD ? bn[3] : rr[3]) - om[3] * rr[3];
         fi += mpole_field(v3(dx, dy, dz), mk, B1, B2, B3, (real)-1);
      }
      fi = group_sum3<G>(fi);
      if (l == 0 && act)
         fd[3 * i] += fi.x, fd[3 * i + 1] += fi.y, fd[3 * i + 2] += fi.z;
   }
}

// d-correction goes to fd, (p - d) correction to the delta array fpd
template <bool TABLE>
__global__ void k_dfield_excl(int nx, Box box, real cut2, const PairExcl* __restrict__ ex, const real4* __restrict__ posd,
   const real4* __restrict__ tpj, const real* __restrict__ thlval, int nj, const real4* __restrict__ mp0,
   const real4* __restrict__ mp1, const real2* __restrict__ mp2, real* __restrict__ fd, real* __restrict__ fpd)
{
   int e = blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= nx)
      return;
   PairExcl p = ex[e];
   if (p.d == 0 && p.p == 0)
      return;
   real4 pi = posd[p.i], pk = posd[p.k];
   real dx = pk.x - pi.x, dy = pk.y - pi.y, dz = pk.z - pi.z;
   apx_image(box, dx, dy, dz);
   real r2 = dx * dx + dy * dy + dz * dz;
   if (r2 > cut2)
      return;
   real rinv = r_rsqrt(r2), r = r2 * rinv, rr2 = rinv * rinv;
   real rr[4], om[4];
   radial_coulomb<4>(rinv, rr2, rr);
   real4 qi = tpj[p.i], qk = tpj[p.k];
   real pg = TABLE ? thlval[as_int(qi.w) * nj + as_int(qk.w)] : min(qi.x, qk.x);
   thole_one_minus_lambda<4>(r, pi.w, pk.w, pg, om);
   real L1 = (1 - om[1]) * rr[1], L2 = (1 - om[2]) * rr[2], L3 = (1 - om[3]) * rr[3];
   V3 R = v3(dx, dy, dz);
   Mpole mi = load_mpole(mp0, mp1, mp2, p.i), mk = load_mpole(mp0, mp1, mp2, p.k);
   V3 ei = mpole_field(R, mk, L1, L2, L3, (real)-1);   // unit-scale damped Coulomb field at i
   V3 ek = mpole_field(R, mi, L1, L2, L3, (real)1);
   if (p.d != 0) {
      atomic_real3(fd, p.i, p.d * ei);
      atomic_real3(fd, p.k, p.d * ek);
   }
   real dp = p.p - p.d;
   if (dp != 0) {
      atomic_real3(fpd, p.i, dp * ei);
      atomic_real3(fpd, p.k, dp * ek);
   }
}

// -------------------------------------------------------------------------------------------
// sparse preconditioner: z += alpha_i alpha_k T_thole(r) r_k  over the pairs inside usolve-cutoff
// (the first cntu entries of every row)
// -------------------------------------------------------------------------------------------
template <bool TABLE, int G>
__global__ void __launch_bounds__(ROWS_BLOCK) k_precond_rows(int n, Box box, const int* __restrict__ vstart,
   const int* __restrict__ cntu, const int* __restrict__ nbr, const real4* __restrict__ posd, const real4* __restrict__ tpj,
   const real* __restrict__ thlval, int nj, const real* __restrict__ rd, const real* __restrict__ rp, real* __restrict__ zd,
   real* __restrict__ zp, const int* __restrict__ skip)
{
   if (skip && skip[1])
      return;
   ROWS_FOREACH_ATOM(G, n, i, l, act)
   {
      const real4 pi = posd[i];
      const real4 qi = tpj[i];
      const int beg = vstart[i];
      const int len = act ? cntu[i] : 0;
      V3 zdi = v3(0, 0, 0), zpi = v3(0, 0, 0);
      for (int q = l; q < len; q += G) {
         const int k = nbr[beg + q];
         const real4 pk = posd[k];
         const real4 qk = tpj[k];
         const V3 a = v3(rd[3 * k], rd[3 * k + 1], rd[3 * k + 2]);
         const V3 b = v3(rp[3 * k], rp[3 * k + 1], rp[3 * k + 2]);
         real dx = pk.x - pi.x, dy = pk.y - pi.y, dz = pk.z - pi.z;
         apx_image(box, dx, dy, dz);
         const real r2 = dx * dx + dy * dy + dz * dz;
         const real rinv = r_rsqrt(r2);
         const real r = r2 * rinv, rr2 = rinv * rinv;
         real rr[3], om[3];
         radial_coulomb<3>(rinv, rr2, rr);
         const real pg = TABLE ? thlval[as_int(qi.w) * nj + as_int(qk.w)] : min(qi.x, qk.x);
         thole_one_minus_lambda<3>(r, pi.w, pk.w, pg, om);
         const real pp = qi.y * qk.y;
         const real B1 = pp * (1 - om[1]) * rr[1], B2 = pp * (1 - om[2]) * rr[2];
         const V3 R = v3(dx, dy, dz);
         zdi += dipole_field(R, a, B1, B2);
         zpi += dipole_field(R, b, B1, B2);
      }
      zdi = group_sum3<G>(zdi);
      zpi = group_sum3<G>(zpi);
      if (l == 0 && act) {
         zd[3 * i] += zdi.x, zd[3 * i + 1] += zdi.y, zd[3 * i + 2] += zdi.z;
         zp[3 * i] += zpi.x, zp[3 * i + 1] += zpi.y, zp[3 * i + 2] += zpi.z;
      }
   }
}

template <bool TABLE>
__global__ void k_precond_excl(int nx, Box box, real cut2, const PairExcl* __restrict__ ex, const real4* __restrict__ posd,
   const real4* __restrict__ tpj, const real* __restrict__ thlval, int nj, const real* __restrict__ rd, const real* __restrict__ rp,
   real* __restrict__ zd, real* __restrict__ zp)
{
   int e = blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= nx)
      return;
   PairExcl p = ex[e];
   if (p.u == 0)
      return;
   real4 pi = posd[p.i], pk = posd[p.k];
   real dx = pk.x - pi.x, dy = pk.y - pi.y, dz = pk.z - pi.z;
   apx_image(box, dx, dy, dz);
   real r2 = dx * dx + dy * dy + dz * dz;
   if (r2 > cut2)
      return;
   real rinv = r_rsqrt(r2), r = r2 * rinv, rr2 = rinv * rinv;
   real rr[3], om[3];
   radial_coulomb<3>(rinv, rr2, rr);
   real4 qi = tpj[p.i], qk = tpj[p.k];
   real pg = TABLE ? thlval[as_int(qi.w) * nj + as_int(qk.w)] : min(qi.x, qk.x);
   thole_one_minus_lambda<3>(r, pi.w, pk.w, pg, om);
   real pp = qi.y * qk.y * p.u;
   real B1 = pp * (1 - om[1]) * rr[1], B2 = pp * (1 - om[2]) * rr[2];
   V3 R = v3(dx, dy, dz);
   V3 rdi = v3(rd[3 * p.i], rd[3 * p.i + 1], rd[3 * p.i + 2]), rpi = v3(rp[3 * p.i], rp[3 * p.i + 1], rp[3 * p.i + 2]);
   V3 rdk = v3(rd[3 * p.k], rd[3 * p.k + 1], rd[3 * p.k + 2]), rpk = v3(rp[3 * p.k], rp[3 * p.k + 1], rp[3 * p.k + 2]);
   atomic_real3(zd, p.i, dipole_field(R, rdk, B1, B2));
   atomic_real3(zp, p.i, dipole_field(R, rpk, B1, B2));
   atomic_real3(zd, p.k, dipole_field(R, rdi, B1, B2));
   atomic_real3(zp, p.k, dipole_field(R, rpi, B1, B2));
}

__global__ void k_diag_precond(int n3, real udiag, const real4* __restrict__ tpj, const real* __restrict__ rd,
   const real* __restrict__ rp, real* __restrict__ zd, real* __restrict__ zp)
{
   int q = blockIdx.x * blockDim.x + threadIdx.x;
   if (q >= n3)
      return;
   real pol = udiag * tpj[q / 3].y;
   zd[q] = pol * rd[q];
   zp[q] = pol * rp[q];
}

} // namespace

#define UF_G 16
#define DF_G 16
#define PC_G 8

void apx_ufield_real(apx_ctx* c, const real* ud, const real* up, real* fd, real* fp)
{
   RowList& L = c->rows;
   real cut = (real)c->opt.cutoff;
   bool ew = c->opt.use_ewald != 0;
   bool tb = c->thole_table != 0;
   int grid = rows_grid<UF_G>(c);
#define LAUNCH_UF(E, T)                                                                                                   \
   k_ufield_rows<E, T, UF_G><<<grid, ROWS_BLOCK, 0, c->stream>>>(c->n, c->box, (real)c->opt.aewald, L.vstart, L.cnt, L.nbr, c->posd,  \
      c->tpj, c->thlval, c->opt.njpolar, ud, up, fd, fp, c->skip)
   if (L.nverlet > 0) {
      // device-time the dominant kernel: one event pair per launch, read back by induce()
      int slot = -1;
      if (c->uf_used + 2 <= (int)c->uf_ev.size()) {
         slot = c->uf_used;
         c->uf_used += 2;
         cudaEventRecord(c->uf_ev[slot], c->stream);
      }
      if (ew && tb) LAUNCH_UF(true, true);
      else if (ew) LAUNCH_UF(true, false);
      else if (tb) LAUNCH_UF(false, true);
      else LAUNCH_UF(false, false);
      if (slot >= 0)
         cudaEventRecord(c->uf_ev[slot + 1], c->stream);
      APX_COUNT_LAUNCH(c);
   }
#undef LAUNCH_UF
   if (c->nexcl_u > 0) {
      int g = (c->nexcl + 127) / 128;
      if (tb)
         k_ufield_excl<true><<<g, 128, 0, c->stream>>>(c->nexcl, c->box, cut * cut, c->excl_s, c->posd, c->tpj, c->thlval,
            c->opt.njpolar, ud, up, fd, fp);
      else
         k_ufield_excl<false><<<g, 128, 0, c->stream>>>(c->nexcl, c->box, cut * cut, c->excl_s, c->posd, c->tpj, c->thlval,
            c->opt.njpolar, ud, up, fd, fp);
      APX_COUNT_LAUNCH(c);
   }
}

// fd accumulates the common field + d corrections; fpd receives the (p - d) delta only
void apx_dfield_real(apx_ctx* c, real* fd, real* fpd)
{
   RowList& L = c->rows;
   real cut = (real)c->opt.cutoff;
   bool ew = c->opt.use_ewald != 0;
   bool tb = c->thole_table != 0;
   int grid = rows_grid<DF_G>(c);
#define LAUNCH_DF(E, T)                                                                                                   \
   k_dfield_rows<E, T, DF_G><<<grid, ROWS_BLOCK, 0, c->stream>>>(c->n, c->box, (real)c->opt.aewald, L.vstart, L.cnt, L.nbr, c->posd,  \
      c->tpj, c->thlval, c->opt.njpolar, c->mp0, c->mp1, c->mp2, fd)
   if (L.nverlet > 0) {
      if (ew && tb) LAUNCH_DF(true, true);
      else if (ew) LAUNCH_DF(true, false);
      else if (tb) LAUNCH_DF(false, true);
      else LAUNCH_DF(false, false);
      APX_COUNT_LAUNCH(c);
   }
#undef LAUNCH_DF
   if (c->nexcl > 0) {
      int g = (c->nexcl + 127) / 128;
      if (tb)
         k_dfield_excl<true><<<g, 128, 0, c->stream>>>(c->nexcl, c->box, cut * cut, c->excl_s, c->posd, c->tpj, c->thlval,
            c->opt.njpolar, c->mp0, c->mp1, c->mp2, fd, fpd);
      else
         k_dfield_excl<false><<<g, 128, 0, c->stream>>>(c->nexcl, c->box, cut * cut, c->excl_s, c->posd, c->tpj, c->thlval,
            c->opt.njpolar, c->mp0, c->mp1, c->mp2, fd, fpd);
      APX_COUNT_LAUNCH(c);
   }
}

// z = M r.  If diag_done the caller already wrote the diagonal part (fused PCG update kernel).
void apx_precond_apply(apx_ctx* c, const real* rd, const real* rp, real* zd, real* zp, bool diag_done)
{
   bool sparse = c->opt.pcgprec && c->opt.usolve_cutoff > 0;
   if (!diag_done) {
      int n3 = 3 * c->n;
      real udiag = sparse ? (real)c->opt.uaccel : (real)1;
      k_diag_precond<<<(n3 + 255) / 256, 256, 0, c->stream>>>(n3, udiag, c->tpj, rd, rp, zd, zp);
      APX_COUNT_LAUNCH(c);
   }
   if (!sparse)
      return;
   RowList& L = c->rows;
   real cut = (real)c->opt.usolve_cutoff;
   bool tb = c->thole_table != 0;
   if (L.nverlet > 0) {
      int grid = rows_grid<PC_G>(c);
      if (tb)
         k_precond_rows<true, PC_G><<<grid, ROWS_BLOCK, 0, c->stream>>>(c->n, c->box, L.vstart, L.cntu, L.nbr, c->posd, c->tpj, c->thlval,
            c->opt.njpolar, rd, rp, zd, zp, c->skip);
      else
         k_precond_rows<false, PC_G><<<grid, ROWS_BLOCK, 0, c->stream>>>(c->n, c->box, L.vstart, L.cntu, L.nbr, c->posd, c->tpj, c->thlval,
            c->opt.njpolar, rd, rp, zd, zp, c->skip);
      APX_COUNT_LAUNCH(c);
   }
   if (c->nexcl_u > 0) {
      int g = (c->nexcl + 127) / 128;
      if (tb)
         k_precond_excl<true><<<g, 128, 0, c->stream>>>(c->nexcl, c->box, cut * cut, c->excl_s, c->posd, c->tpj, c->thlval,
            c->opt.njpolar, rd, rp, zd, zp);
      else
         k_precond_excl<false><<<g, 128, 0, c->stream>>>(c->nexcl, c->box, cut * cut, c->excl_s, c->posd, c->tpj, c->thlval,
            c->opt.njpolar, rd, rp, zd, zp);
      APX_COUNT_LAUNCH(c);
   }
}
