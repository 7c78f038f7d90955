// Valence (bonded) terms of the AMOEBA force fields on the GPU: bond, angle (harmonic and in-plane), stretch-bend,
// Urey-Bradley, out-of-plane bend, torsion, pi-orbital torsion, torsion-torsion -- SURVEY.md section 8f rank 3.
// Replaces evalence_cu1 (src/cu/evalence.cu:17-330) and the e*Data uploads of src/bonded/*.cpp.
//
// One fused launch walks all interactions (one thread each, the eight term lists back to back).  The interaction
// math lives in valmath.cuh / valterms.cuh and is shared with the CPU harness of the tests; it runs in double in
// both builds (tests/test_valence_math.py: float misses the 1e-5 kcal/mol/A force tolerance on stiff bonds, and 48 k
// interactions cost nothing).  Forces go to a fixed-point accumulator of their own, in the CALLER's atom order: it is
// the "fast" gradient of the RESPA integrator (md.cu) and is added to the electrostatics / vdW gradient on read-out.
// HBM traffic per launch: 24 B per referenced atom position (L2-resident), <= 76 B parameters per interaction,
// 24 B atomics per atom touched -- ~10 MB at dhfr2, a latency-bound launch (roofline: HBM).
#include "apx_internal.h"
#include "valpack.h"
#include <cstring>

typedef double vreal;

struct ValState {
   int on = 0, n = 0, total = 0;
   int count[vm::T_COUNT] = {0};       // interactions per term as attached (before the use[] switches)
   int use[vm::T_COUNT] = {0};
   vm::ValDev<vreal> D;                // device pointers
   DevBuf<int> i_bnd, i_ang, i_angtyp, i_sb, i_ury, i_opb, i_tors, i_pit, i_tt, i_ttchk, i_ttgrid;
   DevBuf<vreal> p_b, p_a, p_s, p_u, p_o, p_t, p_p, ttx, tty, tbf, tbx, tby, tbxy;
   DevBuf<vm::TorTorGrid> grids;
   DevBuf<fixed_t> vg;                 // [3][n] fixed-point gradient, caller order
   DevBuf<fixed_t> vbuf;               // [0..7] term energies, [8..13] virial xx yx zx yy zy zz
   cudaStream_t stream = nullptr;      // beside induce() when part of apx_energy()
   cudaEvent_t ev_go = nullptr, ev_done = nullptr;
   int in_total = 0;                   // the last apx_energy() included these terms: apx_get_gradient adds vg
};

namespace {
constexpr int VAL_RED_OFF = 2560;      // landing zone inside apx_ctx::red_h (4096 B pinned)

struct DevAcc {
   unsigned long long* sh_e;           // [8] shared fixed-point term energies of this block
   fixed_t* vg;
   int n;
   double v6[6];
   __device__ __forceinline__ void energy(int term, double e) { atomicAdd(&sh_e[term], (unsigned long long)(long long)(e * APX_FIXED_SCALE)); }
   __device__ __forceinline__ void grad(int atom, double x, double y, double z)
   {
      atomicAdd(&vg[atom], (fixed_t)(long long)(x * APX_FIXED_SCALE));
      atomicAdd(&vg[n + atom], (fixed_t)(long long)(y * APX_FIXED_SCALE));
      atomicAdd(&vg[2 * n + atom], (fixed_t)(long long)(z * APX_FIXED_SCALE));
   }
   __device__ __forceinline__ void virial(const double* v)
   {
#pragma unroll
      for (int k = 0; k < 6; ++k)
         v6[k] += v[k];
   }
};

template <bool DO_E, bool DO_G, bool DO_V>
__global__ void __launch_bounds__(128) k_valence(vm::ValDev<vreal> D, int total, int n, const double* __restrict__ xyz, fixed_t* __restrict__ vg,
   fixed_t* __restrict__ vbuf)
{
   __shared__ unsigned long long sh[14];
   if (threadIdx.x < 14)
      sh[threadIdx.x] = 0;
   __syncthreads();
   DevAcc acc;
   acc.sh_e = sh, acc.vg = vg, acc.n = n;
#pragma unroll
   for (int k = 0; k < 6; ++k)
      acc.v6[k] = 0;
   for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x)
      vm::eval_interaction<vreal>(D, idx, xyz, DO_G, DO_V, acc);
   if (DO_V) {
#pragma unroll
      for (int k = 0; k < 6; ++k) {
         double v = acc.v6[k];
         for (int o = 16; o > 0; o >>= 1)
            v += __shfl_down_sync(0xffffffffu, v, o);
         if ((threadIdx.x & 31) == 0)
            atomicAdd(&sh[8 + k], (unsigned long long)(long long)(v * APX_FIXED_SCALE));
      }
   }
   __syncthreads();
   const int lo = DO_E ? 0 : 8, hi = DO_V ? 14 : 8;
   if ((int)threadIdx.x >= lo && (int)threadIdx.x < hi && sh[threadIdx.x] != 0)
      atomicAdd(&vbuf[threadIdx.x], sh[threadIdx.x]);
}

// out[i][xyz] (+)= vg in double, caller order
__global__ void k_valence_grad_out(int n, const fixed_t* __restrict__ vg, double* __restrict__ out, int accumulate)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n)
      return;
   const double inv = 1.0 / APX_FIXED_SCALE;
   for (int k = 0; k < 3; ++k) {
      double g = (double)(long long)vg[(size_t)k * n + i] * inv;
      out[3 * i + k] = accumulate ? out[3 * i + k] + g : g;
   }
}

template <class T>
const T* upload(apx_ctx* c, DevBuf<T>& b, const std::vector<T>& h)
{
   b.ensure(h.size() + 1);
   if (!h.empty())
      CUDA_CHECK(cudaMemcpyAsync(b.p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, c->stream));
   return b.p;
}
}      // namespace

void apx_valence_attach_impl(apx_ctx* c, const apx_valence* v)
{
   if (v->n != c->n)
      APX_THROW("apx_valence_attach: atom count differs from the system's");
   if (c->dist.on)
      APX_THROW("apx_valence_attach: the valence terms are built for single-GPU contexts (replicas for small systems)");
   const int cnt[vm::T_COUNT] = {v->nbond, v->nangle, v->nstrbnd, v->nurey, v->nopbend, v->ntors, v->npitors, v->ntortor};
   const int* lists[vm::T_COUNT] = {v->ibnd, v->iang, v->isb, v->iury, v->iopb, v->itors, v->ipit, v->itt};
   const int width[vm::T_COUNT] = {2, 4, 3, 3, 4, 4, 6, 5};
   for (int t = 0; t < vm::T_COUNT; ++t)
      for (long long q = 0; q < (long long)cnt[t] * width[t]; ++q) {
         const int a = lists[t][q];
         const bool optional = (t == vm::T_ANGLE && q % 4 == 3);      // out-of-plane atom: -1 when unused
         if (a >= v->n || (a < 0 && !optional))
            APX_THROW("apx_valence_attach: atom index out of range in term list " + std::to_string(t));
      }
   for (int i = 0; i < v->nangle; ++i)
      if (v->angtyp[i] == 1 && (v->iang[4 * i + 3] < 0 || v->iang[4 * i + 3] == v->iang[4 * i + 1]))
         APX_THROW("apx_valence_attach: in-plane angle without an out-of-plane atom");
      else if (v->angtyp[i] != 0 && v->angtyp[i] != 1)
         APX_THROW("apx_valence_attach: only HARMONIC and IN-PLANE angles are built");
   for (int i = 0; i < v->ntortor; ++i)
      if (v->tt_grid[i] < 0 || v->tt_grid[i] >= v->ngrid || v->tt_chk[i] >= v->n)
         APX_THROW("apx_valence_attach: torsion-torsion grid / probe index out of range");
   if (!c->val)
      c->val = new ValState;
   ValState& S = *c->val;
   ValPacked<vreal> P(*v);
   S.n = v->n;
   for (int t = 0; t < vm::T_COUNT; ++t)
      S.count[t] = cnt[t], S.use[t] = v->use[t];
   vm::ValDev<vreal>& D = S.D;
   for (int t = 0; t <= vm::T_COUNT; ++t)
      D.off[t] = P.off[t];
   S.total = P.off[vm::T_COUNT];
   D.ibnd = upload(c, S.i_bnd, P.ibnd), D.bprm = upload(c, S.p_b, P.bprm);
   D.iang = upload(c, S.i_ang, P.iang), D.aprm = upload(c, S.p_a, P.aprm), D.angtyp = upload(c, S.i_angtyp, P.angtyp);
   D.isb = upload(c, S.i_sb, P.isb), D.sprm = upload(c, S.p_s, P.sprm);
   D.iury = upload(c, S.i_ury, P.iury), D.uprm = upload(c, S.p_u, P.uprm);
   D.iopb = upload(c, S.i_opb, P.iopb), D.oprm = upload(c, S.p_o, P.oprm);
   D.itors = upload(c, S.i_tors, P.itors), D.tprm = upload(c, S.p_t, P.tprm);
   D.ipit = upload(c, S.i_pit, P.ipit), D.pprm = upload(c, S.p_p, P.pprm);
   D.itt = upload(c, S.i_tt, P.itt), D.ttchk = upload(c, S.i_ttchk, P.ttchk), D.ttgrid = upload(c, S.i_ttgrid, P.ttgrid);
   D.grids = upload(c, S.grids, P.grids);
   D.ttx = upload(c, S.ttx, P.ttx), D.tty = upload(c, S.tty, P.tty), D.tbf = upload(c, S.tbf, P.tbf);
   D.tbx = upload(c, S.tbx, P.tbx), D.tby = upload(c, S.tby, P.tby), D.tbxy = upload(c, S.tbxy, P.tbxy);
   D.K = P.K, D.opbtyp = P.opbtyp;
   S.vg.ensure(3 * (size_t)S.n);
   S.vbuf.ensure(16);
   CUDA_CHECK(cudaMemsetAsync(S.vg.p, 0, sizeof(fixed_t) * 3 * (size_t)S.n, c->stream));
   CUDA_CHECK(cudaMemsetAsync(S.vbuf.p, 0, sizeof(fixed_t) * 16, c->stream));
   CUDA_CHECK(cudaStreamSynchronize(c->stream));      // P's host vectors go out of scope
   if (!S.stream) {
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);
      CUDA_CHECK(cudaStreamCreateWithPriority(&S.stream, cudaStreamNonBlocking, lo));
      CUDA_CHECK(cudaEventCreateWithFlags(&S.ev_go, cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreateWithFlags(&S.ev_done, cudaEventDisableTiming));
   }
   S.on = 1;
   S.in_total = 0;
   // the captured step graphs know nothing of this term: rebuild them
   apx_pcg_graphs_invalidate(c);
}

bool apx_valence_on(const apx_ctx* c) { return c->val && c->val->on; }

// enqueue on `st`: zero (optionally) and evaluate.  zero_grad = false when the caller's kernel has already cleared vg.
void apx_valence_enqueue(apx_ctx* c, int vers, cudaStream_t st, bool zero_grad)
{
   ValState& S = *c->val;
   const bool do_e = vers & APX_ENERGY, do_g = vers & APX_GRAD, do_v = (vers & APX_VIRIAL) && do_g;
   if (do_g && zero_grad)
      CUDA_CHECK(cudaMemsetAsync(S.vg.p, 0, sizeof(fixed_t) * 3 * (size_t)S.n, st));
   if (do_e || do_v)
      CUDA_CHECK(cudaMemsetAsync(S.vbuf.p, 0, sizeof(fixed_t) * 16, st));
   if (S.total == 0 || !(do_e || do_g))
      return;
   const int block = 128;
   const int grid = (S.total + block - 1) / block;
#define LAUNCH_VAL(E_, G_, V_) k_valence<E_, G_, V_><<<grid, block, 0, st>>>(S.D, S.total, S.n, c->xyz_d, S.vg, S.vbuf)
   if (do_e && do_g && do_v) LAUNCH_VAL(true, true, true);
   else if (do_e && do_g) LAUNCH_VAL(true, true, false);
   else if (do_g && do_v) LAUNCH_VAL(false, true, true);
   else if (do_g) LAUNCH_VAL(false, true, false);
   else LAUNCH_VAL(true, false, false);
#undef LAUNCH_VAL
   APX_COUNT_LAUNCH(c);
}

// as part of apx_energy(): fork from the main stream, run beside the solver
void apx_valence_launch(apx_ctx* c, int vers)
{
   ValState& S = *c->val;
   CUDA_CHECK(cudaEventRecord(S.ev_go, c->stream));
   CUDA_CHECK(cudaStreamWaitEvent(S.stream, S.ev_go, 0));
   apx_valence_enqueue(c, vers, S.stream, true);
   CUDA_CHECK(cudaEventRecord(S.ev_done, S.stream));
}

void apx_valence_join(apx_ctx* c)
{
   ValState& S = *c->val;
   CUDA_CHECK(cudaStreamWaitEvent(c->stream, S.ev_done, 0));
   apx_valence_fetch(c, c->stream);
}

// copy the reduced scalars to the pinned landing zone (valid once `st` is synchronised)
void apx_valence_fetch(apx_ctx* c, cudaStream_t st)
{
   CUDA_CHECK(cudaMemcpyAsync(c->red_h + VAL_RED_OFF, c->val->vbuf.p, sizeof(fixed_t) * 16, cudaMemcpyDeviceToHost, st));
}

void apx_valence_collect(apx_ctx* c, int vers, apx_valence_result* r)
{
   ValState& S = *c->val;
   const bool do_e = vers & APX_ENERGY, do_g = vers & APX_GRAD, do_v = (vers & APX_VIRIAL) && do_g;
   const fixed_t* hb = reinterpret_cast<const fixed_t*>(c->red_h + VAL_RED_OFF);
   auto fx = [](fixed_t v) { return (double)(long long)v / APX_FIXED_SCALE; };
   memset(r, 0, sizeof(*r));
   for (int t = 0; t < vm::T_COUNT; ++t) {
      r->count[t] = S.use[t] ? S.count[t] : 0;
      if (do_e) {
         r->e[t] = fx(hb[t]);
         r->esum += r->e[t];
      }
   }
   if (do_v) {
      const double xx = fx(hb[8]), yx = fx(hb[9]), zx = fx(hb[10]), yy = fx(hb[11]), zy = fx(hb[12]), zz = fx(hb[13]);
      const double m[9] = {xx, yx, zx, yx, yy, zy, zx, zy, zz};
      for (int q = 0; q < 9; ++q)
         r->virial[q] = m[q];
   }
}

void apx_valence_set_in_total(apx_ctx* c, int on)
{
   if (c->val)
      c->val->in_total = on;
}

bool apx_valence_in_total(const apx_ctx* c) { return c->val && c->val->on && c->val->in_total; }

fixed_t* apx_valence_grad_buffer(apx_ctx* c) { return c->val->vg.p; }

void apx_valence_grad_out(apx_ctx* c, double* dev_out, bool accumulate)
{
   ValState& S = *c->val;
   k_valence_grad_out<<<(S.n + 255) / 256, 256, 0, c->stream>>>(S.n, S.vg, dev_out, accumulate ? 1 : 0);
   APX_COUNT_LAUNCH(c);
}

void apx_valence_destroy(apx_ctx* c)
{
   if (!c->val)
      return;
   ValState& S = *c->val;
   if (S.stream) {
      cudaStreamSynchronize(S.stream);
      cudaStreamDestroy(S.stream);
      cudaEventDestroy(S.ev_go);
      cudaEventDestroy(S.ev_done);
   }
   DevBuf<int>* ib[] = {&S.i_bnd, &S.i_ang, &S.i_angtyp, &S.i_sb, &S.i_ury, &S.i_opb, &S.i_tors, &S.i_pit, &S.i_tt, &S.i_ttchk, &S.i_ttgrid};
   for (auto* b : ib)
      b->release();
   DevBuf<vreal>* rb[] = {&S.p_b, &S.p_a, &S.p_s, &S.p_u, &S.p_o, &S.p_t, &S.p_p, &S.ttx, &S.tty, &S.tbf, &S.tbx, &S.tby, &S.tbxy};
   for (auto* b : rb)
      b->release();
   S.grids.release();
   S.vg.release();
   S.vbuf.release();
   delete c->val;
   c->val = nullptr;
}
