// Device-pointer entry points of the C ABI (include/apx.h, names ending in _dev): what the reference's *_cu operators receive
// are DEVICE arrays in the caller's atom order -- real (*)[3] vectors (src/amoeba/field.cpp:8-117, induce.cpp:12-73), separate
// x / y / z coordinate arrays (include/ff/atom.h:39-45), fixed-point gradient and energy / virial buffers that are ACCUMULATED
// into (include/ff/energybuffer.h, src/energy.cpp:333-446) -- so a drop-in adapter must neither stage through host memory nor
// overwrite.  Each call is ordered after the work already enqueued on the caller's stream and visible to what the caller
// enqueues next (event fork / join with the library stream); nothing here blocks the host except where the operator itself
// returns host scalars (apx_induce's convergence read).
#include "apx_internal.h"
#include <cstring>

void apx_dfield_full(apx_ctx* c, bool want_ev);
void apx_grad_to_caller(apx_ctx* c, double* dev_out);

namespace {
template <class T>
__global__ void k_in3(int n, const int* __restrict__ perm, const T* __restrict__ in, real* __restrict__ out)
{
   const int q = blockIdx.x * blockDim.x + threadIdx.x;
   if (q >= 3 * n)
      return;
   const int s = q / 3, c = q - 3 * s;
   out[q] = (real)in[3 * (size_t)perm[s] + c];
}
template <class T>
__global__ void k_out3(int n, const int* __restrict__ perm, const real* __restrict__ in, T* __restrict__ out)
{
   const int q = blockIdx.x * blockDim.x + threadIdx.x;
   if (q >= 3 * n)
      return;
   const int s = q / 3, c = q - 3 * s;
   out[3 * (size_t)perm[s] + c] = (T)in[q];
}
template <class T>
__global__ void k_xyz_in(int n, const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ z, double* __restrict__ xyz)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n)
      return;
   xyz[3 * (size_t)i] = (double)x[i], xyz[3 * (size_t)i + 1] = (double)y[i], xyz[3 * (size_t)i + 2] = (double)z[i];
}
// g[i] += dE/dx_i from the sorted fixed-point accumulators (+ the valence gradient kept in caller order)
template <class T>
__device__ __forceinline__ void add_one(T* p, fixed_t v)
{
   *p += (T)((double)(long long)v * (1.0 / APX_FIXED_SCALE));
}
template <>
__device__ __forceinline__ void add_one<fixed_t>(fixed_t* p, fixed_t v)
{
   *p += v;      // two's-complement fixed point: the integer sum IS the sum
}
template <class T>
__global__ void k_add_grad(int n, const int* __restrict__ inv, const fixed_t* __restrict__ gx, const fixed_t* __restrict__ gy,
   const fixed_t* __restrict__ gz, const fixed_t* __restrict__ vg, T* __restrict__ ox, T* __restrict__ oy, T* __restrict__ oz)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n)
      return;
   const int s = inv[i];
   fixed_t a = gx[s], b = gy[s], c = gz[s];
   if (vg)
      a += vg[i], b += vg[(size_t)n + i], c += vg[2 * (size_t)n + i];
   add_one<T>(ox + i, a);
   add_one<T>(oy + i, b);
   add_one<T>(oz + i, c);
}
struct Scalars {
   double v[16];
   int n;
};
template <class T>
__global__ void k_add_scalars(T* dst, Scalars S)
{
   const int q = threadIdx.x;
   if (q < S.n) {
      if (sizeof(T) == sizeof(fixed_t) && T(-1) > T(0))
         dst[q] += (T)(fixed_t)(long long)(S.v[q] * APX_FIXED_SCALE);
      else
         dst[q] += (T)S.v[q];
   }
}

void check_elem(int elem_bytes)
{
   if (elem_bytes != 4 && elem_bytes != 8)
      APX_THROW("_dev entry point: elem_bytes must be 4 (float) or 8 (double)");
}
void single_gpu(apx_ctx* c)
{
   if (c->dist.on)
      APX_THROW("_dev entry points take whole arrays: single-GPU contexts only");
}

// the library stream waits for the caller's stream ... and the caller's stream for the library's
struct Ordered {
   apx_ctx* c;
   cudaStream_t caller;
   Ordered(apx_ctx* c_, void* stream)
      : c(c_)
      , caller((cudaStream_t)stream)
   {
      CUDA_CHECK(cudaSetDevice(c->device));
      if (!c->ev_dev_in) {
         CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_dev_in, cudaEventDisableTiming));
         CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_dev_out, cudaEventDisableTiming));
      }
      CUDA_CHECK(cudaEventRecord(c->ev_dev_in, caller));
      CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_dev_in, 0));
   }
   void done()
   {
      CUDA_CHECK(cudaEventRecord(c->ev_dev_out, c->stream));
      CUDA_CHECK(cudaStreamWaitEvent(caller, c->ev_dev_out, 0));
   }
};

void in3(apx_ctx* c, const void* src, int eb, real* dst)
{
   const int g = (3 * c->n + 255) / 256;
   if (eb == 4)
      k_in3<float><<<g, 256, 0, c->stream>>>(c->n, c->perm, (const float*)src, dst);
   else
      k_in3<double><<<g, 256, 0, c->stream>>>(c->n, c->perm, (const double*)src, dst);
   APX_COUNT_LAUNCH(c);
}
void out3(apx_ctx* c, const real* src, int eb, void* dst)
{
   if (!dst)
      return;
   const int g = (3 * c->n + 255) / 256;
   if (eb == 4)
      k_out3<float><<<g, 256, 0, c->stream>>>(c->n, c->perm, src, (float*)dst);
   else
      k_out3<double><<<g, 256, 0, c->stream>>>(c->n, c->perm, src, (double*)dst);
   APX_COUNT_LAUNCH(c);
}
void ready(apx_ctx* c)
{
   if (!c->list_valid)
      apx_list_refresh(c, true);
   if (!c->mpole_inited)
      apx_rotpole(c);
}
} // namespace

void apx_set_last_error(const std::string& msg);      // apx_api.cu

#define DEV_BEGIN try {
#define DEV_END                                                                                                          \
   }                                                                                                                       \
   catch (const std::exception& e)                                                                                         \
   {                                                                                                                       \
      apx_set_last_error(e.what());                                                                                        \
      return 1;                                                                                                            \
   }                                                                                                                       \
   return 0;

extern "C" {
#pragma GCC visibility push(default)

int apx_set_positions_dev(apx_ctx* c, const void* x, const void* y, const void* z, int elem_bytes, void* stream)
{
   DEV_BEGIN
   check_elem(elem_bytes);
   single_gpu(c);
   Ordered o(c, stream);
   const int g = (c->n + 255) / 256;
   if (elem_bytes == 4)
      k_xyz_in<float><<<g, 256, 0, c->stream>>>(c->n, (const float*)x, (const float*)y, (const float*)z, c->xyz_d);
   else
      k_xyz_in<double><<<g, 256, 0, c->stream>>>(c->n, (const double*)x, (const double*)y, (const double*)z, c->xyz_d);
   APX_COUNT_LAUNCH(c);
   c->mpole_inited = 0;
   c->mpole_pme_valid = 0;
   c->induced_valid = 0;
   c->md_forces_valid = 0;
   apx_list_refresh(c, false);
   o.done();
   DEV_END
}

int apx_dfield_dev(apx_ctx* c, void* field, void* fieldp, int elem_bytes, void* stream)
{
   DEV_BEGIN
   check_elem(elem_bytes);
   single_gpu(c);
   Ordered o(c, stream);
   ready(c);
   apx_dfield_full(c, false);
   out3(c, c->field, elem_bytes, field);
   out3(c, c->fieldp, elem_bytes, fieldp);
   o.done();
   DEV_END
}

int apx_ufield_dev(apx_ctx* c, const void* uind, const void* uinp, void* field, void* fieldp, int elem_bytes, void* stream)
{
   DEV_BEGIN
   check_elem(elem_bytes);
   single_gpu(c);
   Ordered o(c, stream);
   ready(c);
   in3(c, uind, elem_bytes, c->conj);
   in3(c, uinp, elem_bytes, c->conjp);
   apx_ufield_full(c, c->conj, c->conjp, c->vec, c->vecp);
   out3(c, c->vec, elem_bytes, field);
   out3(c, c->vecp, elem_bytes, fieldp);
   o.done();
   DEV_END
}

int apx_precond_dev(apx_ctx* c, const void* rsd, const void* rsdp, void* zrsd, void* zrsdp, int elem_bytes, void* stream)
{
   DEV_BEGIN
   check_elem(elem_bytes);
   single_gpu(c);
   Ordered o(c, stream);
   ready(c);
   in3(c, rsd, elem_bytes, c->rsd);
   in3(c, rsdp, elem_bytes, c->rsdp);
   apx_precond_apply(c, c->rsd, c->rsdp, c->zrsd, c->zrsdp);
   out3(c, c->zrsd, elem_bytes, zrsd);
   out3(c, c->zrsdp, elem_bytes, zrsdp);
   o.done();
   DEV_END
}

int apx_induce_dev(apx_ctx* c, void* uind, void* uinp, void* udir, void* udirp, int elem_bytes, void* stream)
{
   DEV_BEGIN
   check_elem(elem_bytes);
   single_gpu(c);
   Ordered o(c, stream);
   ready(c);
   apx_induce_impl(c);
   out3(c, c->uind, elem_bytes, uind);
   out3(c, c->uinp, elem_bytes, uinp);
   out3(c, c->udir, elem_bytes, udir);
   out3(c, c->udirp, elem_bytes, udirp);
   o.done();
   DEV_END
}

// the dipoles of the last induce() / energy() again (no solve)
int apx_get_uind_dev(apx_ctx* c, void* uind, void* uinp, void* udir, void* udirp, int elem_bytes, void* stream)
{
   DEV_BEGIN
   check_elem(elem_bytes);
   single_gpu(c);
   if (!c->induced_valid)
      APX_THROW("apx_get_uind_dev before apx_induce / apx_energy");
   Ordered o(c, stream);
   out3(c, c->uind, elem_bytes, uind);
   out3(c, c->uinp, elem_bytes, uinp);
   out3(c, c->udir, elem_bytes, udir);
   out3(c, c->udirp, elem_bytes, udirp);
   o.done();
   DEV_END
}

int apx_add_gradient_dev(apx_ctx* c, void* gx, void* gy, void* gz, int kind, void* stream)
{
   DEV_BEGIN
   single_gpu(c);
   if (kind != APX_DEV_FIXED && kind != APX_DEV_F32 && kind != APX_DEV_F64)
      APX_THROW("apx_add_gradient_dev: kind must be APX_DEV_FIXED, APX_DEV_F32 or APX_DEV_F64");
   Ordered o(c, stream);
   const int g = (c->n + 255) / 256;
   const fixed_t* vg = apx_valence_in_total(c) ? apx_valence_grad_buffer(c) : nullptr;
   if (kind == APX_DEV_FIXED)
      k_add_grad<fixed_t><<<g, 256, 0, c->stream>>>(c->n, c->inv, c->gx, c->gy, c->gz, vg, (fixed_t*)gx, (fixed_t*)gy, (fixed_t*)gz);
   else if (kind == APX_DEV_F32)
      k_add_grad<float><<<g, 256, 0, c->stream>>>(c->n, c->inv, c->gx, c->gy, c->gz, vg, (float*)gx, (float*)gy, (float*)gz);
   else
      k_add_grad<double><<<g, 256, 0, c->stream>>>(c->n, c->inv, c->gx, c->gy, c->gz, vg, (double*)gx, (double*)gy, (double*)gz);
   APX_COUNT_LAUNCH(c);
   o.done();
   DEV_END
}

int apx_add_scalars_dev(apx_ctx* c, void* dst, const double* vals, int count, int kind, void* stream)
{
   DEV_BEGIN
   if (count < 0 || count > 16)
      APX_THROW("apx_add_scalars_dev: at most 16 values per call");
   if (kind != APX_DEV_FIXED && kind != APX_DEV_I32 && kind != APX_DEV_F32 && kind != APX_DEV_F64)
      APX_THROW("apx_add_scalars_dev: kind must be APX_DEV_FIXED, APX_DEV_I32, APX_DEV_F32 or APX_DEV_F64");
   Ordered o(c, stream);
   Scalars S;
   S.n = count;
   memcpy(S.v, vals, sizeof(double) * count);
   if (kind == APX_DEV_FIXED)
      k_add_scalars<fixed_t><<<1, 32, 0, c->stream>>>((fixed_t*)dst, S);
   else if (kind == APX_DEV_I32)
      k_add_scalars<int><<<1, 32, 0, c->stream>>>((int*)dst, S);
   else if (kind == APX_DEV_F32)
      k_add_scalars<float><<<1, 32, 0, c->stream>>>((float*)dst, S);
   else
      k_add_scalars<double><<<1, 32, 0, c->stream>>>((double*)dst, S);
   APX_COUNT_LAUNCH(c);
   o.done();
   DEV_END
}
#pragma GCC visibility pop
}
