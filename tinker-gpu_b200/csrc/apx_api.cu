// extern "C" entry points of libapx (include/apx.h).  Every call converts C++ exceptions into an
// error code + apx_last_error(), the C image of the reference's TINKER_THROW (include/tool/error.h).
#include "apx_internal.h"
#include <cmath>
#include <cstring>
#include <cstdlib>

void apx_to_sorted(apx_ctx* c, const double* in_dev, real* out);
void apx_from_sorted(apx_ctx* c, const real* in, double* out_dev);
void apx_dfield_full(apx_ctx* c, bool want_ev);
void apx_grad_to_caller(apx_ctx* c, double* dev_out);
struct ApxComm;
ApxComm* apx_make_nccl_comm(int rank, int world, const char* lib, const void* unique_id);
ApxComm* apx_make_local_comm(int rank, int world, void* hub);
ApxComm* apx_make_direct_comm(int rank, int world, const void* job_id);

static thread_local std::string g_err;

void apx_set_last_error(const std::string& msg) { g_err = msg; }

void apx_throw(const char* file, int line, const std::string& msg)
{
   const char* base = strrchr(file, '/');
   throw ApxError(std::string(base ? base + 1 : file) + ":" + std::to_string(line) + ": " + msg);
}

#define API_BEGIN try {
#define API_END                                                                                                          \
   }                                                                                                                       \
   catch (const std::exception& e)                                                                                         \
   {                                                                                                                       \
      g_err = e.what();                                                                                                    \
      return 1;                                                                                                            \
   }                                                                                                                       \
   return 0;

namespace {
void invert3(const double* m, double* inv, double& det)
{
   det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
   double id = 1.0 / det;
   inv[0] = (m[4] * m[8] - m[5] * m[7]) * id;
   inv[1] = (m[2] * m[7] - m[1] * m[8]) * id;
   inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
   inv[3] = (m[5] * m[6] - m[3] * m[8]) * id;
   inv[4] = (m[0] * m[8] - m[2] * m[6]) * id;
   inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
   inv[6] = (m[3] * m[7] - m[4] * m[6]) * id;
   inv[7] = (m[1] * m[6] - m[0] * m[7]) * id;
   inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

void set_box(apx_ctx* c, const double* lvec)
{
   double inv[9], det;
   invert3(lvec, inv, det);
   Box& b = c->box;
   for (int i = 0; i < 9; ++i) {
      b.l[i] = (real)lvec[i];
      b.r[i] = (real)inv[i];
      b.q[i] = (real)(lvec[i] / 4294967296.0);
      b.qlo[i] = (real)(lvec[i] / 4294967296.0 - (double)b.q[i]);
   }
   b.lx = (real)lvec[0];
   b.ly = (real)lvec[4];
   b.lz = (real)lvec[8];
   b.ilx = (real)(1.0 / lvec[0]);
   b.ily = (real)(1.0 / lvec[4]);
   b.ilz = (real)(1.0 / lvec[8]);
   double off = fabs(lvec[1]) + fabs(lvec[2]) + fabs(lvec[3]) + fabs(lvec[5]) + fabs(lvec[6]) + fabs(lvec[7]);
   b.orthogonal = off < 1e-12 ? 1 : 0;
   b.volume = (real)fabs(det);
   memcpy(c->opt.lvec, lvec, sizeof(double) * 9);
   memcpy(c->recip_d, inv, sizeof(double) * 9);
}

template <class T, class S>
void upload(apx_ctx* c, DevBuf<T>& dst, const S* src, size_t count)
{
   std::vector<T> tmp(count);
   for (size_t i = 0; i < count; ++i)
      tmp[i] = (T)src[i];
   dst.ensure(count);
   CUDA_CHECK(cudaMemcpyAsync(dst.p, tmp.data(), sizeof(T) * count, cudaMemcpyHostToDevice, c->stream));
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

void require_gpu(int device)
{
   int count = 0;
   cudaError_t e = cudaGetDeviceCount(&count);
   if (e != cudaSuccess || count <= 0)
      APX_THROW("no CUDA device available: libapx has no CPU fallback");
   if (device < 0 || device >= count)
      APX_THROW("CUDA device index out of range");
}

void h2d(apx_ctx* c, DevBuf<double>& dst, const double* src, size_t count)
{
   dst.ensure(count);
   // stage through pinned memory so the copy is a true async DMA inside timed regions
   size_t bytes = count * sizeof(double);
   if (bytes > c->pin_bytes) {
      if (c->pin_a)
         cudaFreeHost(c->pin_a);
      CUDA_CHECK(cudaMallocHost(&c->pin_a, bytes));
      c->pin_bytes = bytes;
   }
   memcpy(c->pin_a, src, bytes);
   CUDA_CHECK(cudaMemcpyAsync(dst.p, c->pin_a, bytes, cudaMemcpyHostToDevice, c->stream));
}

void d2h(apx_ctx* c, double* dst, const double* src_dev, size_t count)
{
   size_t bytes = count * sizeof(double);
   if (bytes > c->pin_bytes) {
      if (c->pin_a)
         cudaFreeHost(c->pin_a);
      CUDA_CHECK(cudaMallocHost(&c->pin_a, bytes));
      c->pin_bytes = bytes;
   }
   CUDA_CHECK(cudaMemcpyAsync(c->pin_a, src_dev, bytes, cudaMemcpyDeviceToHost, c->stream));
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
   memcpy(dst, c->pin_a, bytes);
}

void out_sorted3(apx_ctx* c, const real* src_sorted, double* host_out)
{
   if (c->dist.on)      // collective: every rank returns the whole array
      apx_dist_share_owned(c, const_cast<real*>(src_sorted), 3 * sizeof(real));
   c->io_a.ensure(3 * (size_t)c->n);
   apx_from_sorted(c, src_sorted, c->io_a);
   d2h(c, host_out, c->io_a, 3 * (size_t)c->n);
}

void ensure_ready(apx_ctx* c)
{
   if (!c->list_valid)
      apx_list_refresh(c, true);
   if (!c->mpole_inited)
      apx_rotpole(c);
}
} // namespace

extern "C" {
#pragma GCC visibility push(default)

const char* apx_last_error(void) { return g_err.c_str(); }
const char* apx_version(void) { return "apx 0.1 (" APX_PREC_NAME ")"; }
int apx_precision_bytes(void) { return (int)sizeof(real); }

static void create_impl(const apx_system* sys, int device, int rank, int world, const char* transport, const void* handle,
   const char* nccl_lib, apx_ctx** out)
{
   if (!sys || !out)
      APX_THROW("null argument");
   require_gpu(device);
   CUDA_CHECK(cudaSetDevice(device));
   apx_ctx* c = new apx_ctx();
   *out = c;
   c->device = device;
   if (world > 1) {
      if (rank < 0 || rank >= world || world > 8)
         APX_THROW("rank/world out of range (1..8 GPUs)");
      if (!sys->use_ewald)
         APX_THROW("the multi-GPU decomposition is built for PME systems");
      if (fabs(sys->lvec[2]) + fabs(sys->lvec[5]) + fabs(sys->lvec[6]) + fabs(sys->lvec[7]) > 1e-12)
         APX_THROW("z-slab decomposition needs the third cell vector along z");
      c->dist.on = 1;
      c->dist.rank = rank;
      c->dist.world = world;
      std::string t = transport ? transport : "nccl";
      if (t == "nccl")
         c->dist.comm = apx_make_nccl_comm(rank, world, nccl_lib, handle);
      else if (t == "direct")
         c->dist.comm = apx_make_direct_comm(rank, world, handle);
      else if (t == "local")
         c->dist.comm = apx_make_local_comm(rank, world, const_cast<void*>(handle));
      else
         APX_THROW("unknown transport " + t);
   }
   c->opt = *sys;
   c->n = sys->n;
   if (c->n <= 0)
      APX_THROW("system has no atoms");
   c->nblk = (c->n + 31) / 32;
   c->npad = c->nblk * 32;
   cudaDeviceProp prop;
   CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
   c->sm_count = prop.multiProcessorCount;
   // the latency-bound PME chain outranks the throughput-bound real-space rows it overlaps with
   int prio_lo = 0, prio_hi = 0;
   CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
   CUDA_CHECK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi));
   CUDA_CHECK(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prio_lo));
   CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
   CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
   CUDA_CHECK(cudaEventCreate(&c->ev0));
   CUDA_CHECK(cudaEventCreate(&c->ev1));
   CUDA_CHECK(cudaEventCreate(&c->ev2));
   CUDA_CHECK(cudaEventCreate(&c->ev3));
   CUDA_CHECK(cudaMallocHost(&c->flags_h, 8 * sizeof(int)));
   CUDA_CHECK(cudaMallocHost(&c->scal_h, 8 * sizeof(double)));
   CUDA_CHECK(cudaMallocHost(&c->red_h, 4096));
   memset(&c->stats, 0, sizeof(c->stats));
   c->stats.npairs_m = -1;
   c->f_elec = (real)(sys->electric / sys->dielec);
   if (const char* e = getenv("APX_PME_FIXED"))
      c->pme_fixed = atoi(e) ? 1 : 0;
   if (const char* e = getenv("APX_NO_NATIVE_FFT"))
      c->native_fft = atoi(e) ? 0 : 1;
   if (const char* e = getenv("APX_ROWS_ONEPASS"))
      c->rows_onepass = atoi(e);
   if (const char* e = getenv("APX_NO_RECORDS"))
      c->use_records = atoi(e) ? 0 : 1;
   if (const char* e = getenv("APX_UF_CTAS"))
      c->uf_ctas = std::max(1, atoi(e));
   if (const char* e = getenv("APX_UF_SMEM"))
      c->uf_smem_kb = std::max(0, std::min(40, atoi(e)));
   if (const char* e = getenv("APX_LOOP"))
      c->use_loop = atoi(e) ? 1 : 0;
   if (const char* e = getenv("APX_TLIST"))
      c->tlist_on = atoi(e) ? 1 : 0;
   if (const char* e = getenv("APX_STAGED"))
      c->staged_on = atoi(e) ? 1 : 0;
   if (const char* e = getenv("APX_STAGED_CAP"))
      c->staged_cap = std::max(1, std::min(144, atoi(e)));
   if (const char* e = getenv("APX_STAGED_MIN"))
      c->staged_min_atoms = atoi(e);
   if (const char* e = getenv("APX_DIAG_SKIP"))
      c->diag_skip = atoi(e);
   if (const char* e = getenv("APX_NO_GRAPH"))
      c->use_graph = atoi(e) ? 0 : 1;
   set_box(c, sys->lvec);
   if (sys->cutoff > 0.5 * std::min(std::min(sys->lvec[0], sys->lvec[4]), sys->lvec[8]) + 1e-9 && sys->cutoff < 1e6)
      APX_THROW("real-space cutoff exceeds half the box edge (minimum image would fail)");

   const int n = c->n;
   upload(c, c->xyz_d, sys->xyz, 3 * (size_t)n);
   c->xyz_ref.ensure(3 * (size_t)n);
   upload(c, c->zaxis, sys->zaxis, 4 * (size_t)n);
   upload(c, c->pole, sys->pole, 10 * (size_t)n);
   upload(c, c->polarity_o, sys->polarity, n);
   upload(c, c->thole_o, sys->thole, n);
   upload(c, c->pdamp_o, sys->pdamp, n);
   upload(c, c->jpolar_o, sys->jpolar, n);
   upload(c, c->thlval, sys->thlval, (size_t)sys->njpolar * sys->njpolar);
   // per-pair Thole lookup only needed if the table is not min(thole_i, thole_k) (polpair records)
   c->thole_table = 0;
   {
      std::vector<double> tj(sys->njpolar, -1.0);
      for (int i = 0; i < n; ++i)
         tj[sys->jpolar[i]] = sys->thole[i];
      for (int a = 0; a < sys->njpolar && !c->thole_table; ++a)
         for (int b = 0; b < sys->njpolar; ++b) {
            if (tj[a] < 0 || tj[b] < 0)
               continue;
            double want = std::min(tj[a], tj[b]);
            if (fabs(sys->thlval[a * sys->njpolar + b] - want) > 1e-12) {
               c->thole_table = 1;
               break;
            }
         }
   }
   c->nexcl = sys->nmdpu;
   c->nexcl_u = 0;
   if (c->nexcl > 0) {
      upload(c, c->excl_ik, sys->mdpu_ik, 2 * (size_t)c->nexcl);
      upload(c, c->excl_sc, sys->mdpu_scale, 4 * (size_t)c->nexcl);
      upload(c, c->excl_sc_d, sys->mdpu_scale, 4 * (size_t)c->nexcl);
      c->excl_s.ensure(c->nexcl);
      for (int e = 0; e < c->nexcl; ++e)
         if (sys->mdpu_scale[4 * e + 3] != 1.0)
            c->nexcl_u++;
   }
   const size_t np = c->npad;
   c->perm.ensure(np);
   c->inv.ensure(np);
   c->sortkey.ensure(np);
   c->sortkey2.ensure(np);
   c->permtmp.ensure(np);
   c->cubtmp.ensure(1 << 20);
   c->posd.ensure(np);
#ifdef APX_DOUBLE
   c->posq = c->posd.p;
#else
   c->posq_buf.ensure(np);
   c->posq = c->posq_buf.p;
#endif
   c->tpj.ensure(np);
   c->mp0.ensure(np);
   c->mp1.ensure(np);
   c->mp2.ensure(np);
   c->mpx_a.ensure(np);
   c->mpx_b.ensure(np);
   c->blk_ctr.ensure(c->nblk);
   c->blk_ext.ensure(c->nblk);
   DevBuf<real>* vecs[] = {&c->field, &c->fieldp, &c->udir, &c->udirp, &c->uind, &c->uinp, &c->rsd, &c->rsdp, &c->zrsd, &c->zrsdp,
      &c->conj, &c->conjp, &c->vec, &c->vecp, &c->trq};
   for (auto* v : vecs) {
      v->ensure(3 * np);
      CUDA_CHECK(cudaMemset(v->p, 0, sizeof(real) * 3 * np));
   }
   DevBuf<real4>* pks[] = {&c->pk_p, &c->pk_r, &c->pk_z, &c->pk_v, &c->pk_f};
   for (auto* v : pks) {
      v->ensure(2 * np);
      CUDA_CHECK(cudaMemset(v->p, 0, sizeof(real4) * 2 * np));
   }
   c->fphi.ensure(20 * np);
   c->fmp.ensure(10 * np);
   c->fphid.ensure(10 * np);
   c->fphip.ensure(10 * np);
   c->fphidp.ensure(20 * np);
   {
      // arenas (apx_internal.h): typed views into two allocations; cap = 0 marks "not owned"
      auto carve = [](char*& cur, size_t bytes) {
         char* r = cur;
         cur += (bytes + 255) / 256 * 256;
         return r;
      };
      size_t pad = 256 * 8;
      c->arena_e_bytes = sizeof(fixed_t) * (6 * np + 8) + sizeof(double) * 64 + sizeof(int) * 4 + pad;
      c->arena_e.ensure(c->arena_e_bytes);
      char* cur = c->arena_e.p;
      c->gx.p = (fixed_t*)carve(cur, sizeof(fixed_t) * np);
      c->gy.p = (fixed_t*)carve(cur, sizeof(fixed_t) * np);
      c->gz.p = (fixed_t*)carve(cur, sizeof(fixed_t) * np);
      c->trqf.p = (fixed_t*)carve(cur, sizeof(fixed_t) * 3 * np);
      c->ebuf.p = (fixed_t*)carve(cur, sizeof(fixed_t) * 8);
      c->dbuf.p = (double*)carve(cur, sizeof(double) * 64);
      c->cnt.p = (int*)carve(cur, sizeof(int) * 4);
      c->arena_e_bytes = (size_t)(cur - c->arena_e.p);
      size_t nscal = (size_t)96 * (sys->politer + 3) + 8;     // PCG_SLOT doubles per iteration (dp.cuh)
      c->arena_p_bytes = sizeof(double) * nscal + sizeof(int) * 8 + pad;
      c->arena_p.ensure(c->arena_p_bytes);
      cur = c->arena_p.p;
      c->scal.p = (double*)carve(cur, sizeof(double) * nscal);
      c->flags.p = (int*)carve(cur, sizeof(int) * 8);
      c->arena_p_bytes = (size_t)(cur - c->arena_p.p);
      CUDA_CHECK(cudaMemset(c->arena_e.p, 0, c->arena_e_bytes));
      CUDA_CHECK(cudaMemset(c->arena_p.p, 0, c->arena_p_bytes));
   }
   c->io_a.ensure(3 * np);
   c->io_b.ensure(3 * np);
   c->uf_rec.ensure(3 * np);
   c->list_cutoff = (real)std::min(sys->cutoff, 1.0e6);
   c->list_buffer = (real)sys->list_buffer;
   c->a0 = 0, c->a1 = c->n;
   apx_upred_configure(c, sys->polpred);
   apx_pme_setup(c);
   apx_list_refresh(c, true);
}

int apx_create(const apx_system* sys, int device, apx_ctx** out)
{
   API_BEGIN
   create_impl(sys, device, 0, 1, nullptr, nullptr, nullptr, out);
   API_END
}

int apx_create_dist(const apx_system* sys, int device, int rank, int world, const char* transport, const void* handle,
   const char* nccl_lib, apx_ctx** out)
{
   API_BEGIN
   create_impl(sys, device, rank, world, transport, handle, nccl_lib, out);
   API_END
}

int apx_get_dist_info(apx_ctx* c, int* info /* [8]: rank, world, a0, a1, halo atoms, pz, hl, hu */)
{
   API_BEGIN
   info[0] = c->dist.rank, info[1] = c->dist.world, info[2] = c->a0, info[3] = c->a1;
   info[4] = (int)c->dist.halo_atoms, info[5] = c->dist.pz, info[6] = c->dist.hl, info[7] = c->dist.hu;
   API_END
}

void apx_destroy(apx_ctx* c)
{
   if (!c)
      return;
   cudaSetDevice(c->device);
   cudaStreamSynchronize(c->stream);
   cudaStreamSynchronize(c->stream2);
   apx_pcg_graphs_invalidate(c);
   apx_pme_destroy(c);
   apx_vdw_destroy(c);
   apx_valence_destroy(c);
   apx_md_destroy(c);
   apx_dist_destroy(c);
   // views into the arenas are not owned
   c->gx.p = c->gy.p = c->gz.p = c->trqf.p = c->ebuf.p = nullptr;
   c->dbuf.p = nullptr, c->cnt.p = nullptr, c->scal.p = nullptr, c->flags.p = nullptr;
   c->arena_e.release(), c->arena_p.release();
   if (c->flags_h) cudaFreeHost(c->flags_h);
   if (c->scal_h) cudaFreeHost(c->scal_h);
   if (c->red_h) cudaFreeHost(c->red_h);
   if (c->pin_a) cudaFreeHost(c->pin_a);
   cudaEventDestroy(c->ev0);
   cudaEventDestroy(c->ev1);
   cudaEventDestroy(c->ev2);
   cudaEventDestroy(c->ev3);
   // device buffers are released with the context (process-lifetime objects in practice)
   DevBuf<real>* vecs[] = {&c->field, &c->fieldp, &c->udir, &c->udirp, &c->uind, &c->uinp, &c->rsd, &c->rsdp, &c->zrsd, &c->zrsdp,
      &c->conj, &c->conjp, &c->vec, &c->vecp, &c->trq, &c->fphi, &c->fmp, &c->fphid, &c->fphip, &c->fphidp, &c->pole,
      &c->polarity_o, &c->thole_o, &c->pdamp_o, &c->thlval, &c->excl_sc, &c->qfac, &c->bsmod1, &c->bsmod2, &c->bsmod3};
   for (auto* v : vecs)
      v->release();
   c->excl_sc_d.release();
   c->xyz_d.release(), c->xyz_ref.release(), c->zaxis.release(), c->jpolar_o.release(), c->excl_ik.release();
   c->perm.release(), c->inv.release(), c->sortkey.release(), c->sortkey2.release(), c->permtmp.release(), c->cubtmp.release();
   c->posq_buf.release();
   c->posd.release(), c->tpj.release(), c->mp0.release(), c->mp1.release(), c->mp2.release(), c->mpx_a.release(), c->mpx_b.release();
   c->blk_ctr.release(), c->blk_ext.release(), c->excl_s.release(), c->flags.release(), c->scal.release();
   c->rows.vstart.release(), c->rows.vcnt.release(), c->rows.vnbr.release(), c->rows.nbr.release();
   c->rows.prev_o.release(), c->rows.capstart.release(), c->rows.vpad.release(), c->rows.oflow.release();
   c->rows.cnt.release(), c->rows.cntu.release(), c->rows.total.release();
   c->grp.vjb.release(), c->grp.nvjb.release(), c->grp.ajb.release(), c->grp.najb.release(), c->grp.vslot.release(), c->grp.nbr16.release(),
      c->grp.oflow.release();
   c->qfix.release();
   c->qgrid.release(), c->qgrid2.release(), c->gx.release(), c->gy.release(), c->gz.release(), c->trqf.release();
   c->ebuf.release(), c->dbuf.release(), c->cnt.release(), c->io_a.release(), c->io_b.release(), c->io_c.release(), c->io_d.release();
   c->theta.release(), c->pk_p.release(), c->pk_r.release(), c->pk_z.release(), c->pk_v.release(), c->pk_f.release();
   if (c->ev_dev_in) {
      cudaEventDestroy(c->ev_dev_in);
      cudaEventDestroy(c->ev_dev_out);
   }
   cudaEventDestroy(c->ev_fork);
   cudaEventDestroy(c->ev_join);
   cudaStreamDestroy(c->stream2);
   cudaStreamDestroy(c->stream);
   delete c;
}

int apx_set_positions(apx_ctx* c, const double* xyz)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   h2d(c, c->xyz_d, xyz, 3 * (size_t)c->n);
   c->mpole_inited = 0;
   c->mpole_pme_valid = 0;
   c->induced_valid = 0;
   c->md_forces_valid = 0;
   apx_list_refresh(c, false);
   API_END
}

int apx_set_box(apx_ctx* c, const double lvec[9])
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   set_box(c, lvec);
   apx_pcg_graphs_invalidate(c);
   apx_pme_setup(c);
   c->mpole_inited = 0;
   c->md_forces_valid = 0;
   apx_list_refresh(c, true);
   API_END
}

int apx_mpole_init(apx_ctx* c)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   if (!c->list_valid)
      apx_list_refresh(c, true);
   apx_rotpole(c);
   API_END
}

int apx_get_rpole(apx_ctx* c, double* rpole)
{
   API_BEGIN
   ensure_ready(c);
   const int n = c->n;
   std::vector<real4> a(n), b(n);
   std::vector<real2> d(n);
   std::vector<int> perm(n);
   CUDA_CHECK(cudaMemcpyAsync(a.data(), c->mp0.p, sizeof(real4) * n, cudaMemcpyDeviceToHost, c->stream));
   CUDA_CHECK(cudaMemcpyAsync(b.data(), c->mp1.p, sizeof(real4) * n, cudaMemcpyDeviceToHost, c->stream));
   CUDA_CHECK(cudaMemcpyAsync(d.data(), c->mp2.p, sizeof(real2) * n, cudaMemcpyDeviceToHost, c->stream));
   CUDA_CHECK(cudaMemcpyAsync(perm.data(), c->perm.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
   for (int s = 0; s < n; ++s) {
      double* o = rpole + 10 * (size_t)perm[s];
      o[0] = a[s].x, o[1] = a[s].y, o[2] = a[s].z, o[3] = a[s].w;
      o[4] = b[s].x, o[5] = b[s].w, o[6] = d[s].y;    // xx yy zz
      o[7] = b[s].y, o[8] = b[s].z, o[9] = d[s].x;    // xy xz yz
   }
   API_END
}

int apx_dfield(apx_ctx* c, double* field, double* fieldp)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   ensure_ready(c);
   apx_dfield_full(c, false);
   out_sorted3(c, c->field, field);
   out_sorted3(c, c->fieldp, fieldp);
   API_END
}

int apx_ufield(apx_ctx* c, const double* uind, const double* uinp, double* field, double* fieldp)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   ensure_ready(c);
   h2d(c, c->io_a, uind, 3 * (size_t)c->n);
   apx_to_sorted(c, c->io_a, c->conj);
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
   h2d(c, c->io_a, uinp, 3 * (size_t)c->n);
   apx_to_sorted(c, c->io_a, c->conjp);
   apx_ufield_full(c, c->conj, c->conjp, c->vec, c->vecp);
   out_sorted3(c, c->vec, field);
   out_sorted3(c, c->vecp, fieldp);
   API_END
}

int apx_precond(apx_ctx* c, const double* rsd, const double* rsdp, double* zrsd, double* zrsdp)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   ensure_ready(c);
   h2d(c, c->io_a, rsd, 3 * (size_t)c->n);
   apx_to_sorted(c, c->io_a, c->rsd);
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
   h2d(c, c->io_a, rsdp, 3 * (size_t)c->n);
   apx_to_sorted(c, c->io_a, c->rsdp);
   apx_precond_apply(c, c->rsd, c->rsdp, c->zrsd, c->zrsdp);
   out_sorted3(c, c->zrsd, zrsd);
   out_sorted3(c, c->zrsdp, zrsdp);
   API_END
}

int apx_induce(apx_ctx* c)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   ensure_ready(c);
   apx_induce_impl(c);
   API_END
}

int apx_get_uind(apx_ctx* c, double* uind, double* uinp)
{
   API_BEGIN
   if (!c->induced_valid)
      APX_THROW("apx_get_uind before apx_induce / apx_energy");
   out_sorted3(c, c->uind, uind);
   out_sorted3(c, c->uinp, uinp);
   API_END
}

int apx_get_udir(apx_ctx* c, double* udir, double* udirp)
{
   API_BEGIN
   if (!c->induced_valid)
      APX_THROW("apx_get_udir before apx_induce / apx_energy");
   out_sorted3(c, c->udir, udir);
   out_sorted3(c, c->udirp, udirp);
   API_END
}

int apx_energy(apx_ctx* c, int vers, apx_energy_result* out)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   ensure_ready(c);
   apx_energy_impl(c, vers, true, true, out, true, true);
   API_END
}

int apx_valence_attach(apx_ctx* c, const apx_valence* v)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
   if (!v)
      APX_THROW("apx_valence_attach: null description");
   apx_valence_attach_impl(c, v);
   API_END
}

// the bonded terms alone (energy(vers) with only valence potentials switched on)
int apx_evalence(apx_ctx* c, int vers, apx_valence_result* out)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   if (!apx_valence_on(c))
      APX_THROW("apx_evalence: no valence terms attached (apx_valence_attach)");
   c->md_forces_valid = 0;
   apx_valence_enqueue(c, vers, c->stream, true);
   apx_valence_fetch(c, c->stream);
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
   apx_valence_result r;
   apx_valence_collect(c, vers, &r);
   if (out)
      *out = r;
   API_END
}

int apx_get_valence_gradient(apx_ctx* c, double* grad)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   if (!apx_valence_on(c))
      APX_THROW("apx_get_valence_gradient: no valence terms attached");
   c->io_a.ensure(3 * (size_t)c->n);
   apx_valence_grad_out(c, c->io_a, false);
   d2h(c, grad, c->io_a, 3 * (size_t)c->n);
   API_END
}

int apx_md_init(apx_ctx* c, const double* mass, const double* vel, const apx_md_config* cfg)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   if (!mass || !cfg)
      APX_THROW("apx_md_init: masses and configuration are required");
   ensure_ready(c);
   apx_md_init_impl(c, mass, vel, cfg);
   API_END
}

int apx_md_steps(apx_ctx* c, int nsteps, apx_md_report* out)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   if (nsteps < 0)
      APX_THROW("apx_md_steps: negative step count");
   apx_md_steps_impl(c, nsteps, out);
   API_END
}

int apx_md_get_state(apx_ctx* c, double* xyz, double* vel)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   apx_md_get_state_impl(c, xyz, vel);
   API_END
}

int apx_md_set_state(apx_ctx* c, const double* xyz, const double* vel, int forces_valid)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   apx_md_set_state_impl(c, xyz, vel, forces_valid);
   API_END
}

int apx_vdw_attach(apx_ctx* c, const apx_vdw* v)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
   apx_vdw_attach_impl(c, v);
   API_END
}

// evdw(vers) alone: zero the accumulators, ehal, reductions (src/evdw.cpp:472-530)
int apx_evdw(apx_ctx* c, int vers, apx_energy_result* out)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   if (!c->vdw.on)
      APX_THROW("apx_evdw: no vdW term attached (apx_vdw_attach)");
   ensure_ready(c);
   c->md_forces_valid = 0;
   CUDA_CHECK(cudaMemsetAsync(c->arena_e.p, 0, c->arena_e_bytes, c->stream));
   apx_vdw_launch(c, vers);
   apx_vdw_join(c);
   if (c->dist.on && (vers & APX_GRAD))
      apx_dist_allreduce_u64(c, c->gx.p, (size_t)(c->gz.p + c->npad - c->gx.p));
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
   apx_energy_result r;
   memset(&r, 0, sizeof(r));
   apx_vdw_collect(c, vers, &r);
   r.esum = r.ev;
   if (out)
      *out = r;
   API_END
}

int apx_empole(apx_ctx* c, int vers, apx_energy_result* out)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   ensure_ready(c);
   apx_energy_impl(c, vers, true, false, out);
   API_END
}

int apx_epolar(apx_ctx* c, int vers, apx_energy_result* out)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   ensure_ready(c);
   apx_energy_impl(c, vers, false, true, out);
   API_END
}

int apx_get_gradient(apx_ctx* c, double* grad)
{
   API_BEGIN
   c->io_a.ensure(3 * (size_t)c->n);
   apx_grad_to_caller(c, c->io_a);
   d2h(c, grad, c->io_a, 3 * (size_t)c->n);
   API_END
}

int apx_pme_mpole_fphi(apx_ctx* c, double* fphi)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   ensure_ready(c);
   if (!c->opt.use_ewald)
      APX_THROW("PME operator called on a non-Ewald system");
   apx_pme_mpole(c, false);
   if (c->dist.on)
      apx_dist_share_owned(c, c->fphi.p, 20 * sizeof(real));
   const int n = c->n;
   std::vector<real> h(20 * (size_t)n);
   std::vector<int> perm(n);
   CUDA_CHECK(cudaMemcpyAsync(h.data(), c->fphi.p, sizeof(real) * 20 * n, cudaMemcpyDeviceToHost, c->stream));
   CUDA_CHECK(cudaMemcpyAsync(perm.data(), c->perm.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
   for (int s = 0; s < n; ++s)
      for (int q = 0; q < 20; ++q)
         fphi[20 * (size_t)perm[s] + q] = h[20 * (size_t)s + q];
   API_END
}

int apx_pme_uind_fphi(apx_ctx* c, const double* uind, const double* uinp, double* f1, double* f2)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   ensure_ready(c);
   if (!c->opt.use_ewald)
      APX_THROW("PME operator called on a non-Ewald system");
   h2d(c, c->io_a, uind, 3 * (size_t)c->n);
   apx_to_sorted(c, c->io_a, c->conj);
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
   h2d(c, c->io_a, uinp, 3 * (size_t)c->n);
   apx_to_sorted(c, c->io_a, c->conjp);
   apx_pme_uind_fphi(c, c->conj, c->conjp, true);
   if (c->dist.on) {
      apx_dist_share_owned(c, c->fphid.p, 10 * sizeof(real));
      apx_dist_share_owned(c, c->fphip.p, 10 * sizeof(real));
   }
   const int n = c->n;
   std::vector<real> a(10 * (size_t)n), b(10 * (size_t)n);
   std::vector<int> perm(n);
   CUDA_CHECK(cudaMemcpyAsync(a.data(), c->fphid.p, sizeof(real) * 10 * n, cudaMemcpyDeviceToHost, c->stream));
   CUDA_CHECK(cudaMemcpyAsync(b.data(), c->fphip.p, sizeof(real) * 10 * n, cudaMemcpyDeviceToHost, c->stream));
   CUDA_CHECK(cudaMemcpyAsync(perm.data(), c->perm.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
   for (int s = 0; s < n; ++s)
      for (int q = 0; q < 10; ++q) {
         f1[10 * (size_t)perm[s] + q] = a[10 * (size_t)s + q];
         f2[10 * (size_t)perm[s] + q] = b[10 * (size_t)s + q];
      }
   API_END
}

int apx_get_stats(apx_ctx* c, apx_stats* out)
{
   API_BEGIN
   *out = c->stats;
   API_END
}

int apx_stats_reset(apx_ctx* c)
{
   API_BEGIN
   c->stats.kernel_launches = 0;
   c->stats.list_rebuilds = 0;
   c->stats.energy_retries = 0;
   API_END
}

int apx_upred_set(apx_ctx* c, int polpred)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
   apx_upred_configure(c, polpred);
   API_END
}

int apx_upred_count(apx_ctx* c, int* nualt, int* maxualt)
{
   API_BEGIN
   if (nualt)
      *nualt = c->nualt;
   if (maxualt)
      *maxualt = c->maxualt;
   API_END
}

int apx_set_native_fft(apx_ctx* c, int on)
{
   API_BEGIN
   c->native_fft = on ? 1 : 0;
   apx_pcg_graphs_invalidate(c);
   API_END
}

int apx_set_pme_fixed_point(apx_ctx* c, int on)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
   c->pme_fixed = on ? 1 : 0;
   if (c->pme_fixed)
      apx_pme_fixed_setup(c);
   apx_pcg_graphs_invalidate(c);      // the captured launch sequences contain (or lack) the conversion kernel
   API_END
}

int apx_pme_convolve_grid(apx_ctx* c, const double* in, double* out)
{
   API_BEGIN
   CUDA_CHECK(cudaSetDevice(c->device));
   if (!c->opt.use_ewald)
      APX_THROW("PME operator called on a non-Ewald system");
   if (c->dist.on)
      APX_THROW("apx_pme_convolve_grid takes a whole grid: single-GPU contexts only");
   size_t K = (size_t)c->nfft1 * c->nfft2 * c->nfft3;
   std::vector<cplx> h(K);
   for (size_t i = 0; i < K; ++i) {
      h[i].x = (real)in[2 * i];
      h[i].y = (real)in[2 * i + 1];
   }
   CUDA_CHECK(cudaMemcpyAsync(c->qgrid.p, h.data(), K * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
   apx_pme_convolve(c);
   CUDA_CHECK(cudaMemcpyAsync(h.data(), c->qgrid.p, K * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
   for (size_t i = 0; i < K; ++i) {
      out[2 * i] = (double)h[i].x;
      out[2 * i + 1] = (double)h[i].y;
   }
   API_END
}

void* apx_stream(apx_ctx* c) { return (void*)c->stream; }

int apx_synchronize(apx_ctx* c)
{
   API_BEGIN
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
   API_END
}
#pragma GCC visibility pop
} // extern "C"
