// Preconditioned conjugate-gradient solver for the mutual induced dipoles, d and p right-hand
// sides in lock step: same recurrences, guess, peek step and stopping rule as
// induceMutualPcg1_cu (src/cu/amoeba/pcg.cu:14-185) and the pcg* vector kernels
// (src/cu/induce.cu:17-232), restructured for launch/HBM economy.
//
// One iteration is a chain of 4 kernels of ours around the FFT instead of the reference's
// ~14 kernels + 6 cuBLAS dots + a blocking host read:
//
//   K1  k_pcg_dir                     p = z + b p
//       k_spread_dp (pme.cu), FFT, influence function, inverse FFT (fft64.cu / cuFFT)
//   K2  k_ufield_rows      (field.cu) real-space field of p -- on a second, lower-priority stream,
//                                     beside the spread and the FFTs
//   K3  k_gather_dp<2>     (pme.cu)   recip+self field, + real-space field, Ap = p/alpha - field, p.Ap
//   K4  k_pcg_update                  u += a p ; r -= a Ap ; r.r ; PME grid zeroed for the next spread
//   K5  k_precond_rows     (field.cu) eps test on r.r: peek + stop flag, or z = M r (diagonal + sparse rows), r.z
//
//   * vectors are packed (d,p) pairs, 32 B per atom (dp.cuh);
//   * all scalars stay on the device in per-iteration slots of 16 sub-slots each (no zeroing
//     races, no cuBLAS, short same-address atomic chains);
//   * NO per-iteration host synchronisation: the host enqueues a batch of iterations sized from
//     the previous solve; K5 raises a device flag when eps < poleps (and iter >= miniter) and
//     applies the peek step; every later kernel of the batch returns immediately when the flag
//     is up.  The host reads the flag once per batch from pinned memory.
#include "apx_internal.h"
#include "dp.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace {
// udir = alpha E_d ; udirp = alpha (E_d + delta_p) ; fieldp = E_d + delta ; initial guess u = udir.
// P receives the packed guess (or the packed field when there is no guess: r0 = E).
__global__ void k_udir(int n, const real4* __restrict__ tpj, real* __restrict__ fd, const real* __restrict__ fr, real* __restrict__ fpd,
   real* __restrict__ udir, real* __restrict__ udirp, real* __restrict__ uind, real* __restrict__ uinp, real4* __restrict__ P,
   int guess)
{
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= n)
      return;
   real pol = tpj[s].y;
   V3 ed = v3(fd[3 * s], fd[3 * s + 1], fd[3 * s + 2]);
   if (fr) {
      // the real-space part was computed beside the reciprocal part (apx_dfield_full): the sum is the field
      ed += v3(fr[3 * s], fr[3 * s + 1], fr[3 * s + 2]);
      fd[3 * s] = ed.x, fd[3 * s + 1] = ed.y, fd[3 * s + 2] = ed.z;
   }
   V3 ep = ed + v3(fpd[3 * s], fpd[3 * s + 1], fpd[3 * s + 2]);
   fpd[3 * s] = ep.x, fpd[3 * s + 1] = ep.y, fpd[3 * s + 2] = ep.z;
   V3 a = pol * ed, b = pol * ep;
   udir[3 * s] = a.x, udir[3 * s + 1] = a.y, udir[3 * s + 2] = a.z;
   udirp[3 * s] = b.x, udirp[3 * s + 1] = b.y, udirp[3 * s + 2] = b.z;
   V3 z = v3(0, 0, 0);
   V3 u0 = guess ? a : z, u1 = guess ? b : z;
   uind[3 * s] = u0.x, uind[3 * s + 1] = u0.y, uind[3 * s + 2] = u0.z;
   uinp[3 * s] = u1.x, uinp[3 * s + 1] = u1.y, uinp[3 * s + 2] = u1.z;
   if (guess)
      store_dp(P, s, a, b);
   else {
      if (pol == 0) {
         ed = z;
         ep = z;
      }
      store_dp(P, s, ed, ep);
   }
}

// K1: direction update p = z + b p (pcgP3, src/cu/induce.cu:172-190); b = r.z(new) / r.z(old) from the
// sub-slotted sums, b = 0 with p = 0 on the first iteration
__global__ void __launch_bounds__(256) k_pcg_dir(int n, int* __restrict__ flags, real4* __restrict__ P,
   const real4* __restrict__ Z, const double* __restrict__ slot_prev, const double* __restrict__ slot_cur, int device_loop)
{
   if (flags[1])
      return;
   int it = 0;
   if (device_loop) {
      // device-side loop: the iteration number lives in flags[4]; this kernel opens iteration flags[4] + 1 (slot_cur = slot of
      // iteration 1) and its last CTA publishes the new number for the other kernels of the iteration -- every CTA has read
      // the old one before it takes its ticket
      it = flags[4] + 1;
      slot_cur += (size_t)PCG_SLOT * (it - 1);
      slot_prev = it >= 2 ? slot_cur - PCG_SLOT : nullptr;
   }
   real b = 0, bp = 0;
   if (slot_prev) {
      double o[2], c2[2];
      pcg_q_block<2>(slot_prev, 0, o);
      __syncthreads();
      pcg_q_block<2>(slot_cur, 0, c2);
      b = o[0] != 0.0 ? (real)(c2[0] / o[0]) : (real)0;
      bp = o[1] != 0.0 ? (real)(c2[1] / o[1]) : (real)0;
   }
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s < n) {
      V3 zd, zp, pd, pp;
      load_dp(Z, s, zd, zp);
      if (slot_prev) {
         load_dp(P, s, pd, pp);
         zd += b * pd;
         zp += bp * pp;
      }
      store_dp(P, s, zd, zp);
   }
   if (device_loop) {
      __syncthreads();
      if (threadIdx.x == 0) {
         const unsigned t = atomicAdd((unsigned*)&flags[5], 1u);
         if (t == gridDim.x - 1) {
            flags[5] = 0;
            flags[4] = it;
         }
      }
   }
}

// K4: u += a p ; r -= a Ap (zero where alpha == 0) ; partial r.r ; zero the PME grid
__global__ void __launch_bounds__(256) k_pcg_update(int n, const int* __restrict__ flags, const real4* __restrict__ tpj,
   const real4* __restrict__ P, const real4* __restrict__ V, real4* __restrict__ R, real* __restrict__ ud, real* __restrict__ up,
   double* __restrict__ slot, real4* __restrict__ grid4, size_t ngrid4, const int* __restrict__ itp)
{
   if (flags[1])
      return;
   slot = pcg_slot_of(slot, itp);
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   double e1 = 0, e2 = 0;
   double q4[4];
   pcg_q_block<4>(slot, 0, q4);
   if (s < n) {
      real a = q4[2] != 0.0 ? (real)(q4[0] / q4[2]) : (real)0;
      real ap = q4[3] != 0.0 ? (real)(q4[1] / q4[3]) : (real)0;
      V3 pd, pp, vd, vp, rd, rp;
      load_dp(P, s, pd, pp);
      load_dp(V, s, vd, vp);
      load_dp(R, s, rd, rp);
      ud[3 * s] += a * pd.x, ud[3 * s + 1] += a * pd.y, ud[3 * s + 2] += a * pd.z;
      up[3 * s] += ap * pp.x, up[3 * s + 1] += ap * pp.y, up[3 * s + 2] += ap * pp.z;
      rd = rd - a * vd;
      rp = rp - ap * vp;
      if (tpj[s].y == 0) {
         rd = v3(0, 0, 0);
         rp = v3(0, 0, 0);
      }
      store_dp(R, s, rd, rp);
      e1 = (double)rd.x * rd.x + (double)rd.y * rd.y + (double)rd.z * rd.z;
      e2 = (double)rp.x * rp.x + (double)rp.y * rp.y + (double)rp.z * rp.z;
   }
   real4 zero4;
   zero4.x = zero4.y = zero4.z = zero4.w = 0;
   for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < ngrid4; q += (size_t)gridDim.x * blockDim.x)
      grid4[q] = zero4;
   pcg_block_add2(e1, e2, slot, 4, 5);
}

__global__ void k_pack_dp(int n, const real* __restrict__ d, const real* __restrict__ p, real4* __restrict__ out)
{
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= n)
      return;
   store_dp(out, s, v3(d[3 * s], d[3 * s + 1], d[3 * s + 2]), v3(p[3 * s], p[3 * s + 1], p[3 * s + 2]));
}
__global__ void k_unpack_dp(int n, const real4* __restrict__ in, real* __restrict__ d, real* __restrict__ p)
{
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= n)
      return;
   V3 a, b;
   load_dp(in, s, a, b);
   d[3 * s] = a.x, d[3 * s + 1] = a.y, d[3 * s + 2] = a.z;
   p[3 * s] = b.x, p[3 * s + 1] = b.y, p[3 * s + 2] = b.z;
}

// ---- induced-dipole predictors (ulspredSave / ulspredSum, src/amoeba/induce.cpp:27-69, src/cu/upredict.cu).
// The reference keeps 16 (ASPC) or 6 (GEAR) separate [n][3] arrays per dipole set and passes all of them
// to one kernel; here the ring is one buffer of packed (d,p) pairs in caller order, the coefficient of
// every ring slot arrives by value, and the extrapolated guess is written in the three layouts the
// solver wants (uind, uinp, packed direction buffer) in the same pass.
struct UpredCoef {
   real c[16];
   int m;
};
__global__ void k_upred_save(int n, const int* __restrict__ perm, const real* __restrict__ ud, const real* __restrict__ up,
   real4* __restrict__ slot)
{
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= n)
      return;
   store_dp(slot, perm[s], v3(ud[3 * s], ud[3 * s + 1], ud[3 * s + 2]), v3(up[3 * s], up[3 * s + 1], up[3 * s + 2]));
}
__global__ void k_upred_sum(int n, int ntot, const int* __restrict__ perm, const real4* __restrict__ hist, UpredCoef K,
   real* __restrict__ ud, real* __restrict__ up, real4* __restrict__ P)
{
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= n)
      return;
   const int i = perm[s];
   V3 a = v3(0, 0, 0), b = v3(0, 0, 0);
   for (int k = 0; k < K.m; ++k) {
      V3 hd, hp;
      load_dp(hist + (size_t)2 * ntot * k, i, hd, hp);
      a += K.c[k] * hd;
      b += K.c[k] * hp;
   }
   ud[3 * s] = a.x, ud[3 * s + 1] = a.y, ud[3 * s + 2] = a.z;
   up[3 * s] = b.x, up[3 * s + 1] = b.y, up[3 * s + 2] = b.z;
   store_dp(P, s, a, b);
}
// r0 = (udir - u0)/alpha + field(u0)   (pcgRsd0V2 + pcgRsd0, src/cu/induce.cu:46-58,17-33); R holds field(u0)
__global__ void k_rsd0_pred(int n, const real4* __restrict__ tpj, const real* __restrict__ udir, const real* __restrict__ udirp,
   const real* __restrict__ ud, const real* __restrict__ up, real4* __restrict__ R)
{
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= n)
      return;
   V3 rd, rp;
   load_dp(R, s, rd, rp);
   const real pinv = tpj[s].z;
   rd += pinv * v3(udir[3 * s] - ud[3 * s], udir[3 * s + 1] - ud[3 * s + 1], udir[3 * s + 2] - ud[3 * s + 2]);
   rp += pinv * v3(udirp[3 * s] - up[3 * s], udirp[3 * s + 1] - up[3 * s + 1], udirp[3 * s + 2] - up[3 * s + 2]);
   if (tpj[s].y == 0) {
      rd = v3(0, 0, 0);
      rp = v3(0, 0, 0);
   }
   store_dp(R, s, rd, rp);
}

// caller order (f64) <-> sorted order (real)
__global__ void k_to_sorted(int n, const int* __restrict__ perm, const double* __restrict__ in, real* __restrict__ out)
{
   int q = blockIdx.x * blockDim.x + threadIdx.x;
   if (q >= 3 * n)
      return;
   int s = q / 3, c = q - 3 * s;
   out[q] = (real)in[3 * perm[s] + c];
}
__global__ void k_from_sorted(int n, const int* __restrict__ perm, const real* __restrict__ in, double* __restrict__ out)
{
   int q = blockIdx.x * blockDim.x + threadIdx.x;
   if (q >= 3 * n)
      return;
   int s = q / 3, c = q - 3 * s;
   out[3 * perm[s] + c] = (double)in[q];
}
} // namespace

void apx_to_sorted(apx_ctx* c, const double* in_dev, real* out)
{
   k_to_sorted<<<(3 * c->n + 255) / 256, 256, 0, c->stream>>>(c->n, c->perm, in_dev, out);
   APX_COUNT_LAUNCH(c);
}
void apx_from_sorted(apx_ctx* c, const real* in, double* out_dev)
{
   k_from_sorted<<<(3 * c->n + 255) / 256, 256, 0, c->stream>>>(c->n, c->perm, in, out_dev);
   APX_COUNT_LAUNCH(c);
}

// (owned range only; apx_unpack_dp_all also writes the halo atoms)
void apx_pack_dp(apx_ctx* c, const real* d, const real* p, real4* out)
{
   const int a0 = c->a0, no = c->a1 - c->a0;
   k_pack_dp<<<(no + 255) / 256, 256, 0, c->stream>>>(no, d + 3 * a0, p + 3 * a0, out + 2 * a0);
   APX_COUNT_LAUNCH(c);
}
void apx_unpack_dp(apx_ctx* c, const real4* in, real* d, real* p)
{
   const int a0 = c->a0, no = c->a1 - c->a0;
   k_unpack_dp<<<(no + 255) / 256, 256, 0, c->stream>>>(no, in + 2 * a0, d + 3 * a0, p + 3 * a0);
   APX_COUNT_LAUNCH(c);
}
void apx_unpack_dp_all(apx_ctx* c, const real4* in, real* d, real* p)
{
   k_unpack_dp<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->n, in, d, p);
   APX_COUNT_LAUNCH(c);
}

// full dfield into c->field (d) and c->fieldp (p), then udir/udirp and the initial guess
void apx_dfield_full(apx_ctx* c, bool want_ev)
{
   const int a0 = c->a0, no = c->a1 - c->a0;
   const real* fr = nullptr;
   if (c->opt.use_ewald && !c->dist.on) {
      // the real-space rows (second stream) beside the PME round trip of the multipoles (main stream): 60 us and 130 us at
      // dhfr2 that used to run one after the other; the rows ASSIGN their own buffer, k_udir adds the two parts
      c->field_rs.ensure(3 * (size_t)c->npad);
      fr = c->field_rs.p;
      CUDA_CHECK(cudaEventRecord(c->ev_fork, c->stream));
      CUDA_CHECK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
      apx_dfield_real(c, c->stream2, c->field_rs, c->fieldp, true);
      CUDA_CHECK(cudaEventRecord(c->ev_join, c->stream2));
      apx_pme_mpole(c, want_ev);                 // ASSIGNS c->field = recip + self
      CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
   } else {
      if (c->opt.use_ewald)
         apx_pme_mpole(c, want_ev);              // ASSIGNS c->field = recip + self
      apx_dfield_real(c, c->stream, c->field, c->fieldp, !c->opt.use_ewald);      // adds to (Ewald) or assigns (no Ewald) field
   }
   k_udir<<<(no + 255) / 256, 256, 0, c->stream>>>(no, c->tpj + a0, c->field + 3 * a0, fr ? fr + 3 * a0 : nullptr, c->fieldp + 3 * a0,
      c->udir + 3 * a0, c->udirp + 3 * a0, c->uind + 3 * a0, c->uinp + 3 * a0, c->pk_p + 2 * a0, c->opt.pcgguess ? 1 : 0);
   APX_COUNT_LAUNCH(c);
}

// Real-space rows of U on the second stream while the main stream runs spread (optional) + FFTs.
// On return the main stream has waited for the rows; c->pk_f holds the real-space field.
static void field_of_dp(apx_ctx* c, const real4* U, bool spread)
{
   cudaStream_t st = c->stream;
   if (c->dist.on)
      apx_dist_halo(c, const_cast<real4*>(U), st);      // neighbours' dipoles from the GPUs that own them
   // Decomposed runs, APX_DIST_FORK_LATE=1: the operator starts only when the spread has FINISHED.  Measured at 8 GPUs
   // (profiles/r02y_*, r02z_*): started together, the exchange after the spread sees its last CTA pass the peers' arrival flags
   // 27-120 us after kernel entry; forked late that drops to 10 us -- and the same ~100 us reappear at the forward transpose,
   // behind the 2-D FFTs, whose wide CTAs do not fit into the slots the operator's narrow CTAs free one at a time (stream
   // priority picks among CTAs that FIT).  6.65 ms per step either way, so the default stays the earlier start.
   static const int fork_late = getenv("APX_DIST_FORK_LATE") ? atoi(getenv("APX_DIST_FORK_LATE")) : 0;
   const bool late = c->dist.on && fork_late && c->opt.use_ewald && spread;
   if (!late)
      CUDA_CHECK(cudaEventRecord(c->ev_fork, st));
   // the spread heads the critical path (spread -> FFTs -> gather) and goes in first: launched after the rows, its CTAs waited
   // 13 us for the operator's grid to drain (profiles/r02f_trace_md.txt); the operator is only needed by the gather
   if (c->opt.use_ewald && spread)
      apx_pme_spread_dp(c, U);
   if (late)
      CUDA_CHECK(cudaEventRecord(c->ev_fork, st));
   CUDA_CHECK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
   apx_ufield_real_dp(c, c->stream2, U, c->pk_f);
   CUDA_CHECK(cudaEventRecord(c->ev_join, c->stream2));
   if (c->opt.use_ewald)
      apx_pme_convolve(c);
   CUDA_CHECK(cudaStreamWaitEvent(st, c->ev_join, 0));
}

// field of the dipole pair (ud, up): ufield() of the reference (src/amoeba/field.cpp:111-117)
void apx_ufield_full(apx_ctx* c, const real* ud, const real* up, real* fd, real* fp)
{
   apx_pack_dp(c, ud, up, c->pk_p);
   if (c->opt.use_ewald)
      apx_pme_zero_grid(c);
   field_of_dp(c, c->pk_p, true);
   if (c->opt.use_ewald)
      apx_pme_gather_dp(c, 0, c->pk_p, c->pk_f, fd, fp, nullptr, nullptr);
   else
      apx_unpack_dp(c, c->pk_f, fd, fp);
}

__global__ void k_mask_dp(int n, const real4* __restrict__ tpj, real4* __restrict__ V)
{
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s < n && tpj[s].y == 0)
      store_dp(V, s, v3(0, 0, 0), v3(0, 0, 0));
}
__global__ void k_nonewald_ap(int n, const int* __restrict__ flags, const real4* __restrict__ tpj, const real4* __restrict__ P,
   const real4* __restrict__ F, real4* __restrict__ V, double* __restrict__ slot, const int* __restrict__ itp)
{
   if (flags[1])
      return;
   slot = pcg_slot_of(slot, itp);
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   double x = 0, y = 0;
   if (s < n) {
      V3 pd, pp, fd, fp;
      load_dp(P, s, pd, pp);
      load_dp(F, s, fd, fp);
      real pinv = tpj[s].z;
      V3 vd = pinv * pd - fd, vp = pinv * pp - fp;
      store_dp(V, s, vd, vp);
      x = (double)pd.x * vd.x + (double)pd.y * vd.y + (double)pd.z * vd.z;
      y = (double)pp.x * vp.x + (double)pp.y * vp.y + (double)pp.z * vp.z;
   }
   pcg_block_add2(x, y, slot, 2, 3);
}

void apx_upred_configure(apx_ctx* c, int polpred)
{
   if (polpred == 3)       // LSQR: the reference's CUDA back end has no such kernel either (upredict.cu:204-206)
      APX_THROW("polar-predict LSQR is not available on the GPU (ulspredSumLSQR_cu missing in the reference as well)");
   if (polpred < 0 || polpred > 2)
      APX_THROW("unknown polar-predict kind");
   c->opt.polpred = polpred;
   c->maxualt = polpred == 1 ? 16 : (polpred == 2 ? 6 : 0);      // src/amoeba/epolar.cpp:449-468
   c->nualt = 0;
   if (c->maxualt)
      c->upred_hist.ensure((size_t)2 * c->n * c->maxualt);
}

// ulspredSave: the converged dipoles go into ring slot nualt % maxualt
static void upred_save(apx_ctx* c)
{
   const int m = c->maxualt;
   if (!m)
      return;
   int a0 = c->a0, no = c->a1 - c->a0;
   if (c->dist.on) {
      // atoms change slabs at list rebuilds: every rank keeps the history of ALL atoms
      apx_dist_share_owned(c, c->uind.p, 3 * sizeof(real));
      apx_dist_share_owned(c, c->uinp.p, 3 * sizeof(real));
      a0 = 0, no = c->n;
   }
   const int pos = c->nualt % m;
   k_upred_save<<<(no + 255) / 256, 256, 0, c->stream>>>(no, c->perm + a0, c->uind + 3 * a0, c->uinp + 3 * a0,
      c->upred_hist + (size_t)2 * c->n * pos);
   APX_COUNT_LAUNCH(c);
   c->nualt++;
   if (c->nualt > 2 * m)
      c->nualt -= m;
}

static bool trace_graphs()
{
   static const int on = getenv("APX_TRACE_GRAPHS") ? atoi(getenv("APX_TRACE_GRAPHS")) : 0;
   return on != 0;
}

void apx_pcg_graphs_invalidate(apx_ctx* c)
{
   if (trace_graphs())
      fprintf(stderr, "[apx] graphs invalidated (%zu pcg, %zu step)\n", c->graphs.size(), c->step_graphs.size());
   for (auto& g : c->graphs)
      cudaGraphExecDestroy(g.exec);
   c->graphs.clear();
   if (c->loop_exec)
      cudaGraphExecDestroy(c->loop_exec);
   c->loop_exec = nullptr;
   for (auto& kv : c->step_graphs)
      if (kv.second.exec)
         cudaGraphExecDestroy(kv.second.exec);
   c->step_graphs.clear();
}

__global__ void k_set_cond(const int* __restrict__ flag, cudaGraphConditionalHandle h)
{
   cudaGraphSetConditional(h, *flag != 0 ? 1u : 0u);
}

bool apx_graph_is_conditional(apx_ctx* c, int key)
{
   auto it = c->step_graphs.find(key);
   return it != c->step_graphs.end() && it->second.exec && it->second.conditional;
}

bool apx_graph_begin(apx_ctx* c, int key, const int* cond_flag)
{
   c->graph_key_open = -1;
   if (!c->use_graph || c->dist.on || c->capturing)
      return true;
   apx_ctx::StepGraph& G = c->step_graphs[key];
   if (G.exec) {
      CUDA_CHECK(cudaGraphLaunch(G.exec, c->stream));
      c->stats.kernel_launches += G.launches;
      return false;
   }
   if (!G.warm) {
      G.warm = 1;      // eager once: plans, workspaces and scratch buffers get allocated outside any capture
      return true;
   }
   if (trace_graphs())
      fprintf(stderr, "[apx] capturing step graph 0x%x%s\n", key, cond_flag ? " (conditional)" : "");
   c->graph_key_open = key;
   c->graph_launches_before = c->stats.kernel_launches;
   c->cond_outer = nullptr;
   G.conditional = 0;
   if (cond_flag) {
      // outer graph: [k_set_cond] -> [IF node]; the region is captured into the IF node's body
      cudaGraph_t g = nullptr;
      cudaGraphConditionalHandle h = 0;
      bool ok = cudaGraphCreate(&g, 0) == cudaSuccess && cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault) == cudaSuccess;
      cudaGraphNode_t first = nullptr, cnode = nullptr;
      if (ok) {
         ok = cudaStreamBeginCaptureToGraph(c->stream, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
         if (ok) {
            k_set_cond<<<1, 1, 0, c->stream>>>(cond_flag, h);
            cudaGraph_t out = nullptr;
            ok = cudaStreamEndCapture(c->stream, &out) == cudaSuccess;
         }
      }
      if (ok) {
         size_t nn = 1;
         ok = cudaGraphGetNodes(g, &first, &nn) == cudaSuccess && nn == 1;
      }
      cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
      if (ok) {
         np.conditional.handle = h;
         np.conditional.type = cudaGraphCondTypeIf;
         np.conditional.size = 1;
         ok = cudaGraphAddNode(&cnode, g, &first, 1, &np) == cudaSuccess;
      }
      if (ok)
         ok = cudaStreamBeginCaptureToGraph(c->stream, np.conditional.phGraph_out[0], nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
      if (ok) {
         c->cond_outer = g;
         c->capturing = 1;
         G.conditional = 1;
         c->cond_nodes_ok = 1;
         return true;
      }
      c->cond_nodes_ok = 0;
      (void)cudaGetLastError();
      if (g)
         cudaGraphDestroy(g);
      if (trace_graphs())
         fprintf(stderr, "[apx] conditional graph nodes unavailable: plain graph for 0x%x\n", key);
   }
   c->capturing = 1;
   CUDA_CHECK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
   return true;
}

void apx_graph_end(apx_ctx* c, int key)
{
   if (c->graph_key_open != key)
      return;
   c->graph_key_open = -1;
   cudaGraph_t graph = nullptr;
   cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
   c->capturing = 0;
   if (e != cudaSuccess)
      APX_THROW(std::string("step graph capture failed: ") + cudaGetErrorString(e));
   apx_ctx::StepGraph& G = c->step_graphs[key];
   G.launches = c->stats.kernel_launches - c->graph_launches_before;
   if (c->cond_outer) {      // `graph` is the body of the IF node: the outer graph is what gets instantiated
      graph = c->cond_outer;
      c->cond_outer = nullptr;
   }
   CUDA_CHECK(cudaGraphInstantiate(&G.exec, graph, 0));
   cudaGraphDestroy(graph);
   CUDA_CHECK(cudaGraphLaunch(G.exec, c->stream));      // the captured work has not run yet
}

static void upred_save(apx_ctx* c);
static void induce_epilogue(apx_ctx* c, int used, bool predict)
{
   if (c->uf_iter_timed && c->uf_ev.size() >= 4 && used >= 2) {
      // graph path: the operator launch of the second iteration (external event nodes of the batch graph) -- the launch
      // of the prologue runs beside the vdW rows and is not what the other 7 look like
      float ms = 0;
      cudaEventElapsedTime(&ms, c->uf_ev[2], c->uf_ev[3]);
      c->stats.ms_ufield_real = ms;
   } else {
      // mean device time of the real-space ufield launches that did work (speculative launches after
      // convergence return immediately and are excluded): 1 for r0 + one per iteration
      int used_pairs = std::min(c->uf_used / 2, used + ((c->opt.pcgguess || predict) ? 1 : 0));
      float tot = 0;
      for (int k = 0; k < used_pairs; ++k) {
         float ms = 0;
         cudaEventElapsedTime(&ms, c->uf_ev[2 * k], c->uf_ev[2 * k + 1]);
         tot += ms;
      }
      c->stats.ms_ufield_real = used_pairs ? tot / used_pairs : 0;
   }
   c->stats.pcg_iterations = used;
   c->last_iters = used;
   // length of the first batch of the next solves: the largest count seen, lowered by one after 50 solves that needed less
   if (used >= c->pcg_n) {
      c->pcg_n = used;
      c->pcg_n_slack = 0;
   } else if (++c->pcg_n_slack >= 50) {
      c->pcg_n -= 1;
      c->pcg_n_slack = 0;
   }
   c->stats.ms_induce = 0;
   cudaEventElapsedTime(&c->stats.ms_induce, c->ev0, c->ev1);
   c->scal_h[2] = c->scal_h[0];
   upred_save(c);
   if (used >= c->opt.politer && !(c->scal_h[0] < c->opt.poleps))
      APX_THROW("INDUCE  --  Warning, Induced Dipoles are not Converged");   // pcg.cu:180-184
}

// defer = true (energy(), mplar.cu): when the iterations run as the device-side loop nothing here waits for the GPU -- the caller
// enqueues what follows the solve, synchronises once, and then calls apx_induce_finish.  Returns true when a finish is pending.
static bool induce_core(apx_ctx* c, int mode);
bool apx_induce_impl(apx_ctx* c, bool defer) { return induce_core(c, defer ? 1 : 0); }
// the deferred first batch did not converge (apx_induce_finish returned false): more iterations, in batches the host waits for
void apx_induce_resume(apx_ctx* c) { (void)induce_core(c, 2); }

// mode 0: solve and wait.  1: deferred (see above).  2: resume after an unconverged deferred batch -- no prologue
static bool induce_core(apx_ctx* c, int mode)
{
   const bool defer = mode == 1, resume = mode == 2;
   const int n = c->n, n3 = 3 * n;
   const int a0 = c->a0, no = c->a1 - c->a0;      // owned range: every per-atom pass below runs on it
   const int g1 = (no + 255) / 256;
   const bool dist = c->dist.on != 0;
   cudaStream_t st = c->stream;
   const bool ewald = c->opt.use_ewald != 0;
   if (!resume) {
      if (!c->mpole_inited)
         apx_rotpole(c);
      if (c->uf_ev.empty()) {
         c->uf_ev.resize(64);
         for (auto& e : c->uf_ev)
            CUDA_CHECK(cudaEventCreate(&e));
      }
      c->uf_used = 0;
      c->uf_iter_timed = 0;
      c->tl_valid = 0, c->tl_p_valid = 0;      // one tensor build per induce(), by its first operator application: the launch sequence
                                               // (and with it the captured graphs) does not depend on what ran before
      cudaEventRecord(c->ev0, st);
      c->stats.pcg_iterations = 0;
      c->induced_valid = 1;
   }
   // the predictor replaces the direct guess once its ring is full
   const bool predict = resume ? c->induce_pending_predict != 0 : (c->maxualt > 0 && c->nualt >= c->maxualt);
   if (!resume && !c->opt.poltyp_mutual) {
      // DIRECT polarization: u = alpha E
      (void)n3;
      apx_dfield_full(c, true);
      CUDA_CHECK(cudaMemcpyAsync(c->uind.p + 3 * a0, c->udir.p + 3 * a0, sizeof(real) * 3 * no, cudaMemcpyDeviceToDevice, st));
      CUDA_CHECK(cudaMemcpyAsync(c->uinp.p + 3 * a0, c->udirp.p + 3 * a0, sizeof(real) * 3 * no, cudaMemcpyDeviceToDevice, st));
      cudaEventRecord(c->ev1, st);
      CUDA_CHECK(cudaStreamSynchronize(st));
      cudaEventElapsedTime(&c->stats.ms_induce, c->ev0, c->ev1);
      return false;
   }
   const int politer = c->opt.politer;
   const int miniter = std::min(3, n);
   static_assert(PCG_SLOT == 96, "arena_p in apx_api.cu is sized for 96 doubles per iteration");
   double* result = c->scal.p + (size_t)PCG_SLOT * (politer + 3);
   const size_t ngrid4 = ewald ? (size_t)c->nfft1 * c->nfft2 * c->nzl * sizeof(cplx) / sizeof(real4) : 0;
   // graph region: permanent field, direct dipoles, first residual, first preconditioner application -- a fixed launch
   // sequence unless the predictor supplies this step's coefficients by value
   const bool graphable = !predict;
   const int gkey = 0x2000 | (c->opt.pcgguess ? 1 : 0);
   if (!resume && (!graphable || apx_graph_begin(c, gkey))) {
      apx_dfield_full(c, true);
      CUDA_CHECK(cudaMemsetAsync(c->arena_p.p, 0, c->arena_p_bytes, st));      // scal + flags
      // the predictor replaces the direct guess once its ring is full (pcg.cu:26-31)
      if (predict) {
         static const double aspc[16] = {62. / 17., -310. / 51., 2170. / 323., -2329. / 400., 1701. / 409., -806. / 323., 1024. / 809.,
            -479. / 883., 257. / 1316., -434. / 7429., 191. / 13375., -62. / 22287., 3. / 7217., -3. / 67015., 2. / 646323.,
            -1. / 9694845.};
         static const double gear[6] = {6., -15., 20., -15., 6., -1.};
         UpredCoef K;
         K.m = c->maxualt;
         for (int k = 0; k < K.m; ++k) {
            int age = ((c->nualt - 1 - k) % K.m + K.m) % K.m;      // ring slot k holds the solution of this age
            K.c[k] = (real)(c->opt.polpred == 1 ? aspc[age] : gear[age]);
         }
         k_upred_sum<<<g1, 256, 0, st>>>(no, n, c->perm + a0, c->upred_hist, K, c->uind + 3 * a0, c->uinp + 3 * a0, c->pk_p + 2 * a0);
         APX_COUNT_LAUNCH(c);
      }
      // r0 = -T u0  (pcgguess; k_udir left u0 packed in pk_p) or E (no guess; k_udir left E packed in pk_p)
      if (c->opt.pcgguess || predict) {
         if (ewald)
            apx_pme_zero_grid(c);
         field_of_dp(c, c->pk_p, true);
         if (ewald)
            apx_pme_gather_dp(c, 1, c->pk_p, c->pk_f, nullptr, nullptr, c->pk_r, nullptr);
         else {
            CUDA_CHECK(cudaMemcpyAsync(c->pk_r.p + 2 * a0, c->pk_f.p + 2 * a0, sizeof(real4) * 2 * no, cudaMemcpyDeviceToDevice, st));
            k_mask_dp<<<g1, 256, 0, st>>>(no, c->tpj + a0, c->pk_r + 2 * a0);
            APX_COUNT_LAUNCH(c);
         }
         if (predict) {
            k_rsd0_pred<<<g1, 256, 0, st>>>(no, c->tpj + a0, c->udir + 3 * a0, c->udirp + 3 * a0, c->uind + 3 * a0, c->uinp + 3 * a0,
               c->pk_r + 2 * a0);
            APX_COUNT_LAUNCH(c);
         }
      } else {
         CUDA_CHECK(cudaMemcpyAsync(c->pk_r.p + 2 * a0, c->pk_p.p + 2 * a0, sizeof(real4) * 2 * no, cudaMemcpyDeviceToDevice, st));
      }
      // z0 = M r0, r0.z0 -> slot of iteration 1 (whose K1 sets p = z0)
      if (dist)
         apx_dist_halo(c, c->pk_r, st);
      apx_precond_dp(c, c->pk_r, c->pk_z, c->scal.p);
      if (dist)
         apx_dist_allreduce_f64(c, c->scal.p, 2 * PCG_NSUB);
      if (ewald)
         apx_pme_zero_grid(c);
      if (graphable)
         apx_graph_end(c, gkey);
   }
   if (ewald && !resume)
      c->mpole_pme_valid = 1;
   // the prologue applied the operator (and with it built the pair tensors) whether it ran eagerly or as a replayed graph:
   // the host flag must say so, or the capture of the first iteration would bake a second build into its graph
   if (apx_tlist_usable(c))
      c->tl_valid = 1, c->tl_p_valid = c->tl_P.p ? 1 : 0;      // (written by the permanent-field rows of the prologue, field.cu)
   if (c->vdw_fork_vers >= 0 && !resume) {
      // energy() asked for the vdW term to start here: beside the iterations, whose short dependent kernels leave the SMs idle
      apx_vdw_launch(c, c->vdw_fork_vers);
      c->vdw_fork_vers = -1;
   }
   if (!resume && graphable && c->use_graph && !dist && (c->opt.pcgguess) && c->uf_used == 0)
      c->uf_used = 2;      // the r0 operator launch inside the graph was timed through external event nodes
   int iter = resume ? c->induce_iter_launched : 0;
   bool done = false;
   c->skip = c->flags.p;
   int batch = std::max(1, std::min(c->last_iters, politer));
   PcgTest T;
   T.miniter = miniter, T.politer = politer;
   T.poleps = (real)c->opt.poleps, T.debye = (real)4.803206802, T.pcgpeek = (real)c->opt.pcgpeek;
   T.result = result, T.flags = c->flags, T.ud = c->uind, T.up = c->uinp;
   // one iteration.  loop = false: iteration `it`, slots addressed by the host.  loop = true: the body of the device-side loop --
   // the iteration number lives in flags[4] (k_pcg_dir advances it), every kernel finds its slots from it
   auto enqueue_iteration = [&](int it, bool loop, unsigned long long cond) {
      const int* itp = loop ? c->flags.p + 4 : nullptr;
      double* slot = loop ? c->scal.p : c->scal.p + (size_t)PCG_SLOT * (it - 1);
      k_pcg_dir<<<g1, 256, 0, st>>>(no, c->flags, c->pk_p + 2 * a0, c->pk_z + 2 * a0, (!loop && it >= 2) ? slot - PCG_SLOT : nullptr, slot,
         loop ? 1 : 0);
      APX_COUNT_LAUNCH(c);
      field_of_dp(c, c->pk_p, true);
      if (ewald)
         apx_pme_gather_dp(c, 2, c->pk_p, c->pk_f, nullptr, nullptr, c->pk_v, slot, itp);
      else {
         k_nonewald_ap<<<g1, 256, 0, st>>>(no, c->flags, c->tpj + a0, c->pk_p + 2 * a0, c->pk_f + 2 * a0, c->pk_v + 2 * a0, slot, itp);
         APX_COUNT_LAUNCH(c);
      }
      if (dist)
         apx_dist_allreduce_f64(c, slot + 2 * PCG_NSUB, 2 * PCG_NSUB);        // p.Ap over all GPUs
      k_pcg_update<<<g1, 256, 0, st>>>(no, c->flags, c->tpj + a0, c->pk_p + 2 * a0, c->pk_v + 2 * a0, c->pk_r + 2 * a0,
         c->uind + 3 * a0, c->uinp + 3 * a0, slot, (real4*)c->qgrid.p, ngrid4, itp);
      APX_COUNT_LAUNCH(c);
      if (dist) {
         apx_dist_allreduce_f64(c, slot + 4 * PCG_NSUB, 2 * PCG_NSUB);        // r.r
         apx_dist_halo(c, c->pk_r, st);                                       // residuals of the preconditioner's neighbours
      }
      T.it = loop ? 0 : it;
      T.slot = slot;
      T.itp = itp;
      T.cond = cond;
      apx_precond_dp(c, c->pk_r, c->pk_z, slot + PCG_SLOT, &T);
      if (dist)
         apx_dist_allreduce_f64(c, slot + PCG_SLOT, 2 * PCG_NSUB);            // r.z entering the next iteration
   };
   // ---- the loop on the device: ONE graph whose only node is a conditional WHILE node with one iteration as its body.  The
   // preconditioner kernel of the iteration that meets the stopping rule sets the condition to 0 (field.cu: k_precond_rows);
   // no batch sizes to guess, no launches after convergence, no graph per (first iteration, batch length), and the host does
   // not have to look at the convergence flag before it enqueues what follows.  First solve of a context: host-driven and
   // eager (plans, scratch buffers), like every other graph region.
   bool use_loop = c->use_graph && c->use_loop && !dist && !resume;
   if (use_loop && !c->loop_exec && c->loop_warm) {
      if (trace_graphs())
         fprintf(stderr, "[apx] capturing the pcg loop body\n");
      // (a driver without conditional nodes: the first failing call switches this context to the generic batches below)
      cudaGraph_t g = nullptr;
      cudaGraphConditionalHandle h = 0;
      cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
      cudaGraphNode_t node;
      bool ok = cudaGraphCreate(&g, 0) == cudaSuccess && cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault) == cudaSuccess;
      if (ok) {
         np.conditional.handle = h;
         np.conditional.type = cudaGraphCondTypeWhile;
         np.conditional.size = 1;
         ok = cudaGraphAddNode(&node, g, nullptr, 0, &np) == cudaSuccess;
      }
      cudaGraph_t body = ok ? np.conditional.phGraph_out[0] : nullptr;
      ok = ok && cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
      if (ok) {
         const int before = c->stats.kernel_launches;
         c->capturing = 1;
         enqueue_iteration(0, true, (unsigned long long)h);
         cudaGraph_t out = nullptr;
         ok = cudaStreamEndCapture(st, &out) == cudaSuccess;
         c->capturing = 0;
         c->loop_launches = c->stats.kernel_launches - before;
         c->stats.kernel_launches = before;
         ok = ok && cudaGraphInstantiate(&c->loop_exec, g, 0) == cudaSuccess;
      }
      if (g)
         cudaGraphDestroy(g);
      if (!ok) {
         (void)cudaGetLastError();
         c->loop_exec = nullptr;
         c->use_loop = 0;
         if (trace_graphs())
            fprintf(stderr, "[apx] conditional graph nodes unavailable: generic iteration batches instead\n");
      }
   }
   use_loop = use_loop && c->use_loop;
   if (use_loop && c->loop_exec) {
      CUDA_CHECK(cudaGraphLaunch(c->loop_exec, st));
      CUDA_CHECK(cudaMemcpyAsync(c->flags_h, c->flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaMemcpyAsync(c->scal_h, result, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
      cudaEventRecord(c->ev1, st);
      c->skip = nullptr;
      c->induce_pending = 1;
      c->induce_pending_predict = predict ? 1 : 0;
      if (defer)
         return true;
      CUDA_CHECK(cudaStreamSynchronize(st));
      (void)apx_induce_finish(c);
      return false;
   }
   // ---- default: iterations in batches, every batch ONE graph.  After the first (eager) solve of a context the iterations are
   // "generic" -- the iteration number lives on the device, so the graph of an N-iteration batch serves any first iteration:
   // one capture per batch length instead of one per (first iteration, length), which used to put a 1-8 ms capture into MD
   // steps whenever the iteration count moved (profiles/r02g: steps 5-6 of the timed window).  The first batch has the
   // largest iteration count seen recently (c->pcg_n); kernels of iterations past convergence return at once.  With defer
   // the host does not wait for that batch: the caller enqueues the energy epilogue behind it and synchronises once.
   const bool generic = resume || (c->use_graph && !dist && !use_loop && c->loop_warm);
   c->loop_warm = 1;
   if (resume)
      batch = 2;
   else if (generic) {
      batch = std::max(1, std::min(c->pcg_n > 0 ? c->pcg_n : c->last_iters, politer));
      static const int forced = getenv("APX_PCG_FIRST_BATCH") ? atoi(getenv("APX_PCG_FIRST_BATCH")) : 0;      // tests: force the retry path
      if (forced > 0)
         batch = std::min(forced, politer);
   }
   bool first_batch = !resume;
   while (!done) {
      int nit = std::min(batch, politer - iter);
      if (generic) {
         auto find_or_capture = [&](int len) -> apx_ctx::PcgGraph* {
            for (auto& g : c->graphs)
               if (g.it0 == 0 && g.nit == len)
                  return &g;
            if (trace_graphs())
               fprintf(stderr, "[apx] capturing generic pcg graph (%d iterations)\n", len);
            int before = c->stats.kernel_launches;
            c->capturing = 1;
            CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            for (int b = 0; b < len; ++b) {
               c->uf_ext_iter = b == std::min(1, len - 1) ? 1 : 0;      // the second iteration's operator launch is timed (steady state)
               enqueue_iteration(0, true, 0);
            }
            c->uf_ext_iter = 0;
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamEndCapture(st, &graph);
            c->capturing = 0;
            if (e != cudaSuccess)
               APX_THROW(std::string("PCG graph capture failed: ") + cudaGetErrorString(e));
            apx_ctx::PcgGraph ng;
            ng.it0 = 0, ng.nit = len;
            ng.launches = c->stats.kernel_launches - before;
            c->stats.kernel_launches = before;
            CUDA_CHECK(cudaGraphInstantiate(&ng.exec, graph, 0));
            cudaGraphDestroy(graph);
            c->graphs.push_back(ng);
            return &c->graphs.back();
         };
         if (first_batch && c->graphs.empty()) {
            // the iteration count of an MD run wanders by one: the neighbouring batch lengths are captured now, with the first,
            // instead of 2 ms into some later step
            c->graphs.reserve(16);
            if (nit + 1 <= politer)
               find_or_capture(nit + 1);
            if (nit > 1)
               find_or_capture(nit - 1);
            find_or_capture(2);      // (the continuation batches of a solve that outgrows its first batch)
         }
         apx_ctx::PcgGraph* G = find_or_capture(nit);
         if (first_batch)
            c->uf_iter_timed = 1;
         CUDA_CHECK(cudaGraphLaunch(G->exec, st));
         c->stats.kernel_launches += G->launches;
         iter += nit;
      } else {
         for (int b = 0; b < nit; ++b)
            enqueue_iteration(++iter, false, 0);
      }
      const bool leave = generic && defer && first_batch && iter < politer;
      // (an unawaited batch: the caller copies the flags out AFTER what it enqueues behind the solve -- apx_induce_copy_out --
      // so that its graph follows the last iteration kernel directly instead of two 3 us copies and their gaps)
      if (!leave) {
         CUDA_CHECK(cudaMemcpyAsync(c->flags_h, c->flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
         CUDA_CHECK(cudaMemcpyAsync(c->scal_h, result, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
      }
      cudaEventRecord(c->ev1, st);
      if (leave) {
         c->induce_copy_pending = 1;
         c->skip = nullptr;
         c->induce_pending = 2;
         c->induce_pending_predict = predict ? 1 : 0;
         c->induce_iter_launched = iter;
         return true;
      }
      first_batch = false;
      CUDA_CHECK(cudaStreamSynchronize(st));
      done = c->flags_h[1] != 0 || iter >= politer;
      batch = 2;
   }
   c->skip = nullptr;
   induce_epilogue(c, c->flags_h[2] > 0 ? c->flags_h[2] : iter, predict);
   return false;
}

// convergence flag, iteration count and final residual of an unawaited batch -> pinned memory (enqueued by the caller behind its
// own work; a no-op when the solver has already copied them)
void apx_induce_copy_out(apx_ctx* c)
{
   if (!c->induce_copy_pending)
      return;
   c->induce_copy_pending = 0;
   double* result = c->scal.p + (size_t)PCG_SLOT * (c->opt.politer + 3);
   CUDA_CHECK(cudaMemcpyAsync(c->flags_h, c->flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
   CUDA_CHECK(cudaMemcpyAsync(c->scal_h, result, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
}

// what follows the solve on the host once its results are visible (after a stream synchronisation).  false: the deferred first
// batch did not converge -- the caller redoes the evaluation with a solve that waits for its batches
bool apx_induce_finish(apx_ctx* c)
{
   if (!c->induce_pending)
      return true;
   const int kind = c->induce_pending;
   c->induce_pending = 0;
   if (kind == 2 && c->flags_h[1] == 0)
      return false;      // (apx_induce_resume carries on; its bookkeeping raises pcg_n)
   const int used = c->flags_h[2] > 0 ? c->flags_h[2] : c->opt.politer;
   if (kind == 1)
      c->stats.kernel_launches += c->loop_launches * used;
   induce_epilogue(c, used, c->induce_pending_predict != 0);
   return true;
}
