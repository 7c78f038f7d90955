// Preconditioned conjugate-gradient solver for the mutual induced dipoles, d and p right-hand
// sides in lock step: same recurrences, guess, peek step and stopping rule as
// induceMutualPcg1_cu (src/cu/amoeba/pcg.cu:14-185) and the pcg* vector kernels
// (src/cu/induce.cu:17-232), restructured for launch/HBM economy:
//
//   * 3 fused vector passes per iteration instead of 6 kernels + 6 cuBLAS dots + 2 copies:
//       A  vec = p/alpha - field ; partial p.vec
//       B  u += a p ; r -= a vec ; z = udiag*alpha*r (diagonal of the preconditioner) ; partial r.r
//       C  partial r.z   and   D  p = z + b p  /  convergence test / peek
//     (C and D need a grid-wide reduction between them, hence two launches);
//   * all scalars stay on the device in per-iteration slots (no zeroing races, no cuBLAS);
//   * NO per-iteration host synchronisation: the host enqueues a batch of iterations sized from
//     the previous solve; kernel D raises a device flag when  eps < poleps (and iter >= miniter)
//     and applies the peek step; every later kernel of the batch returns immediately when the
//     flag is up.  The host reads the flag once per batch from pinned memory.
#include "apx_internal.h"
#include <algorithm>
#include <cmath>

namespace {
constexpr int SLOT = 8;     // doubles per iteration: [0,1] r.z(prev) [2,3] p.Ap [6,7] r.r ; r.z(new) -> next slot [0,1]

__device__ __forceinline__ void block_sum2(double a, double b, double* out_a, double* out_b)
{
   __shared__ double sh[2][8];
   int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
   for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
   }
   if (lane == 0) {
      sh[0][w] = a;
      sh[1][w] = b;
   }
   __syncthreads();
   if (threadIdx.x == 0) {
      double x = 0, y = 0;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
         x += sh[0][k];
         y += sh[1][k];
      }
      atomicAdd(out_a, x);
      atomicAdd(out_b, y);
   }
}

// udir = alpha E_d ; udirp = alpha (E_d + delta_p) ; fieldp = E_d + delta ; initial guess u = udir
__global__ void k_udir(int n3, const real4* __restrict__ tpj, const real* __restrict__ fd, real* __restrict__ fpd,
   real* __restrict__ udir, real* __restrict__ udirp, real* __restrict__ uind, real* __restrict__ uinp, int guess)
{
   int q = blockIdx.x * blockDim.x + threadIdx.x;
   if (q >= n3)
      return;
   real pol = tpj[q / 3].y;
   real ed = fd[q], ep = ed + fpd[q];
   fpd[q] = ep;
   real a = pol * ed, b = pol * ep;
   udir[q] = a;
   udirp[q] = b;
   uind[q] = guess ? a : 0;
   uinp[q] = guess ? b : 0;
}

// r0: zero where alpha == 0 ; z = udiag alpha r (diagonal part)
__global__ void k_rsd0(int n3, real udiag, const real4* __restrict__ tpj, real* __restrict__ rd, real* __restrict__ rp,
   real* __restrict__ zd, real* __restrict__ zp)
{
   int q = blockIdx.x * blockDim.x + threadIdx.x;
   if (q >= n3)
      return;
   real pol = tpj[q / 3].y;
   real a = rd[q], b = rp[q];
   if (pol == 0) {
      a = 0;
      b = 0;
      rd[q] = 0;
      rp[q] = 0;
   }
   zd[q] = udiag * pol * a;
   zp[q] = udiag * pol * b;
}

// p = z ; r.z -> slot0[0,1]
__global__ void k_init_conj(int n3, const real* __restrict__ rd, const real* __restrict__ rp, const real* __restrict__ zd,
   const real* __restrict__ zp, real* __restrict__ cd, real* __restrict__ cp, double* __restrict__ slot)
{
   int q = blockIdx.x * blockDim.x + threadIdx.x;
   double a = 0, b = 0;
   if (q < n3) {
      real z1 = zd[q], z2 = zp[q];
      cd[q] = z1;
      cp[q] = z2;
      a = (double)rd[q] * z1;
      b = (double)rp[q] * z2;
   }
   block_sum2(a, b, &slot[0], &slot[1]);
}

// pass A
__global__ void k_pass_a(int n3, const int* __restrict__ flags, const real4* __restrict__ tpj, const real* __restrict__ cd,
   const real* __restrict__ cp, const real* __restrict__ fd, const real* __restrict__ fp, real* __restrict__ vd,
   real* __restrict__ vp, double* __restrict__ slot)
{
   if (flags[1])
      return;
   int q = blockIdx.x * blockDim.x + threadIdx.x;
   double a = 0, b = 0;
   if (q < n3) {
      real pinv = tpj[q / 3].z;
      real c1 = cd[q], c2 = cp[q];
      real v1 = pinv * c1 - fd[q], v2 = pinv * c2 - fp[q];
      vd[q] = v1;
      vp[q] = v2;
      a = (double)c1 * v1;
      b = (double)c2 * v2;
   }
   block_sum2(a, b, &slot[2], &slot[3]);
}

// pass B
__global__ void k_pass_b(int n3, const int* __restrict__ flags, real udiag, const real4* __restrict__ tpj,
   const real* __restrict__ cd, const real* __restrict__ cp, const real* __restrict__ vd, const real* __restrict__ vp,
   real* __restrict__ ud, real* __restrict__ up, real* __restrict__ rd, real* __restrict__ rp, real* __restrict__ zd,
   real* __restrict__ zp, double* __restrict__ slot)
{
   if (flags[1])
      return;
   int q = blockIdx.x * blockDim.x + threadIdx.x;
   double e1 = 0, e2 = 0;
   if (q < n3) {
      double pa = slot[2], pb = slot[3];
      real a = pa != 0.0 ? (real)(slot[0] / pa) : (real)0;
      real ap = pb != 0.0 ? (real)(slot[1] / pb) : (real)0;
      real pol = tpj[q / 3].y;
      ud[q] += a * cd[q];
      up[q] += ap * cp[q];
      real r1 = rd[q] - a * vd[q], r2 = rp[q] - ap * vp[q];
      if (pol == 0) {
         r1 = 0;
         r2 = 0;
      }
      rd[q] = r1;
      rp[q] = r2;
      zd[q] = udiag * pol * r1;
      zp[q] = udiag * pol * r2;
      e1 = (double)r1 * r1;
      e2 = (double)r2 * r2;
   }
   block_sum2(e1, e2, &slot[6], &slot[7]);
}

// pass C: r.z into the NEXT slot
__global__ void k_pass_c(int n3, const int* __restrict__ flags, const real* __restrict__ rd, const real* __restrict__ rp,
   const real* __restrict__ zd, const real* __restrict__ zp, double* __restrict__ slot)
{
   if (flags[1])
      return;
   int q = blockIdx.x * blockDim.x + threadIdx.x;
   double a = 0, b = 0;
   if (q < n3) {
      a = (double)rd[q] * zd[q];
      b = (double)rp[q] * zp[q];
   }
   block_sum2(a, b, &slot[SLOT + 0], &slot[SLOT + 1]);
}

// pass D: convergence test; either p = z + b p, or the peek step + raise the flag
__global__ void k_pass_d(int n3, int n, int iter, int miniter, int politer, real poleps, real debye, real pcgpeek, int* flags,
   double* __restrict__ result, const real4* __restrict__ tpj, real* __restrict__ cd, real* __restrict__ cp,
   const real* __restrict__ zd, const real* __restrict__ zp, real* __restrict__ ud, real* __restrict__ up,
   const real* __restrict__ rd, const real* __restrict__ rp, const double* __restrict__ slot)
{
   if (flags[1])
      return;
   int q = blockIdx.x * blockDim.x + threadIdx.x;
   double e = fmax(slot[6], slot[7]);
   double eps = (double)debye * sqrt(e / n);
   bool done = eps < (double)poleps;
   if (iter < miniter)
      done = false;
   if (iter >= politer)
      done = true;
   if (q < n3) {
      if (done) {
         real term = pcgpeek * tpj[q / 3].y;
         ud[q] += term * rd[q];
         up[q] += term * rp[q];
      } else {
         double s0 = slot[0], s1 = slot[1];
         real b = s0 != 0.0 ? (real)(slot[SLOT] / s0) : (real)0;
         real bp = s1 != 0.0 ? (real)(slot[SLOT + 1] / s1) : (real)0;
         cd[q] = zd[q] + b * cd[q];
         cp[q] = zp[q] + bp * cp[q];
      }
   }
   __syncthreads();
   if (q == 0) {
      result[0] = eps;
      result[1] = (double)iter;
      if (done) {
         __threadfence();
         flags[2] = iter;
      }
   }
   // the flag itself is raised by a 1-thread tail kernel so no thread of THIS grid can see it early
}

__global__ void k_raise_flag(int* flags)
{
   if (flags[2] > 0)
      flags[1] = 1;
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

// full dfield into c->field (d) and c->fieldp (p), then udir/udirp and the initial guess
void apx_dfield_full(apx_ctx* c, bool want_ev)
{
   int n3 = 3 * c->n;
   if (c->opt.use_ewald) {
      apx_pme_mpole(c, want_ev);                 // ASSIGNS c->field = recip + self
   } else {
      CUDA_CHECK(cudaMemsetAsync(c->field.p, 0, sizeof(real) * n3, c->stream));
   }
   CUDA_CHECK(cudaMemsetAsync(c->fieldp.p, 0, sizeof(real) * n3, c->stream));
   apx_dfield_real(c, c->field, c->fieldp);
   k_udir<<<(n3 + 255) / 256, 256, 0, c->stream>>>(n3, c->tpj, c->field, c->fieldp, c->udir, c->udirp, c->uind, c->uinp,
      c->opt.pcgguess ? 1 : 0);
   APX_COUNT_LAUNCH(c);
}

void apx_ufield_full(apx_ctx* c, const real* ud, const real* up, real* fd, real* fp)
{
   int n3 = 3 * c->n;
   if (c->opt.use_ewald) {
      apx_pme_ufield(c, ud, up, fd, fp, nullptr, nullptr, nullptr);
   } else {
      CUDA_CHECK(cudaMemsetAsync(fd, 0, sizeof(real) * n3, c->stream));
      CUDA_CHECK(cudaMemsetAsync(fp, 0, sizeof(real) * n3, c->stream));
   }
   apx_ufield_real(c, ud, up, fd, fp);
}

void apx_induce_impl(apx_ctx* c)
{
   const int n = c->n, n3 = 3 * n;
   const int g3 = (n3 + 255) / 256;
   cudaStream_t st = c->stream;
   if (!c->mpole_inited)
      apx_rotpole(c);
   if (c->uf_ev.empty()) {
      c->uf_ev.resize(64);
      for (auto& e : c->uf_ev)
         CUDA_CHECK(cudaEventCreate(&e));
   }
   c->uf_used = 0;
   cudaEventRecord(c->ev0, st);
   apx_dfield_full(c, true);
   c->stats.pcg_iterations = 0;
   c->induced_valid = 1;
   if (!c->opt.poltyp_mutual) {
      // DIRECT polarization: u = alpha E
      CUDA_CHECK(cudaMemcpyAsync(c->uind.p, c->udir.p, sizeof(real) * n3, cudaMemcpyDeviceToDevice, st));
      CUDA_CHECK(cudaMemcpyAsync(c->uinp.p, c->udirp.p, sizeof(real) * n3, cudaMemcpyDeviceToDevice, st));
      cudaEventRecord(c->ev1, st);
      CUDA_CHECK(cudaStreamSynchronize(st));
      cudaEventElapsedTime(&c->stats.ms_induce, c->ev0, c->ev1);
      return;
   }
   const bool sparse = c->opt.pcgprec && c->opt.usolve_cutoff > 0;
   const real udiag = sparse ? (real)c->opt.uaccel : (real)1;
   const int politer = c->opt.politer;
   const int miniter = std::min(3, n);
   size_t nscal = (size_t)SLOT * (politer + 3) + 8;
   c->scal.ensure(nscal);
   CUDA_CHECK(cudaMemsetAsync(c->scal.p, 0, nscal * sizeof(double), st));
   CUDA_CHECK(cudaMemsetAsync(c->flags.p, 0, 4 * sizeof(int), st));
   double* result = c->scal.p + (size_t)SLOT * (politer + 3);

   // r0 = -T u0  (pcgguess) or E (no guess)
   if (c->opt.pcgguess) {
      apx_ufield_full(c, c->uind, c->uinp, c->rsd, c->rsdp);
   } else {
      CUDA_CHECK(cudaMemcpyAsync(c->rsd.p, c->field.p, sizeof(real) * n3, cudaMemcpyDeviceToDevice, st));
      CUDA_CHECK(cudaMemcpyAsync(c->rsdp.p, c->fieldp.p, sizeof(real) * n3, cudaMemcpyDeviceToDevice, st));
   }
   k_rsd0<<<g3, 256, 0, st>>>(n3, udiag, c->tpj, c->rsd, c->rsdp, c->zrsd, c->zrsdp);
   APX_COUNT_LAUNCH(c);
   apx_precond_apply(c, c->rsd, c->rsdp, c->zrsd, c->zrsdp, true);
   k_init_conj<<<g3, 256, 0, st>>>(n3, c->rsd, c->rsdp, c->zrsd, c->zrsdp, c->conj, c->conjp, c->scal.p);
   APX_COUNT_LAUNCH(c);

   int iter = 0;
   bool done = false;
   c->skip = c->flags.p;
   int batch = std::max(1, std::min(c->last_iters, politer));
   while (!done) {
      for (int b = 0; b < batch && iter < politer; ++b) {
         ++iter;
         double* slot = c->scal.p + (size_t)SLOT * (iter - 1);
         apx_ufield_full(c, c->conj, c->conjp, c->field, c->fieldp);
         k_pass_a<<<g3, 256, 0, st>>>(n3, c->flags, c->tpj, c->conj, c->conjp, c->field, c->fieldp, c->vec, c->vecp, slot);
         k_pass_b<<<g3, 256, 0, st>>>(n3, c->flags, udiag, c->tpj, c->conj, c->conjp, c->vec, c->vecp, c->uind, c->uinp, c->rsd, c->rsdp,
            c->zrsd, c->zrsdp, slot);
         apx_precond_apply(c, c->rsd, c->rsdp, c->zrsd, c->zrsdp, true);
         k_pass_c<<<g3, 256, 0, st>>>(n3, c->flags, c->rsd, c->rsdp, c->zrsd, c->zrsdp, slot);
         k_pass_d<<<g3, 256, 0, st>>>(n3, n, iter, miniter, politer, (real)c->opt.poleps, (real)4.803206802, (real)c->opt.pcgpeek,
            c->flags, result, c->tpj, c->conj, c->conjp, c->zrsd, c->zrsdp, c->uind, c->uinp, c->rsd, c->rsdp, slot);
         k_raise_flag<<<1, 1, 0, st>>>(c->flags);
         c->stats.kernel_launches += 5;
      }
      CUDA_CHECK(cudaMemcpyAsync(c->flags_h, c->flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaMemcpyAsync(c->scal_h, result, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
      cudaEventRecord(c->ev1, st);
      CUDA_CHECK(cudaStreamSynchronize(st));
      done = c->flags_h[1] != 0 || iter >= politer;
      batch = 2;
   }
   c->skip = nullptr;
   {
      // mean device time of the real-space ufield launches that did work (speculative launches after
      // convergence return immediately and are excluded: only the first `used+1` pairs count)
      int used_pairs = std::min(c->uf_used / 2, (c->flags_h[2] > 0 ? c->flags_h[2] : iter) + 1);
      float tot = 0;
      for (int k = 0; k < used_pairs; ++k) {
         float ms = 0;
         cudaEventElapsedTime(&ms, c->uf_ev[2 * k], c->uf_ev[2 * k + 1]);
         tot += ms;
      }
      c->stats.ms_ufield_real = used_pairs ? tot / used_pairs : 0;
   }
   int used = c->flags_h[2] > 0 ? c->flags_h[2] : iter;
   c->stats.pcg_iterations = used;
   c->last_iters = used;
   c->stats.ms_induce = 0;
   cudaEventElapsedTime(&c->stats.ms_induce, c->ev0, c->ev1);
   c->scal_h[2] = c->scal_h[0];
   if (used >= politer && !(c->scal_h[0] < c->opt.poleps))
      APX_THROW("INDUCE  --  Warning, Induced Dipoles are not Converged");   // pcg.cu:180-184
}
