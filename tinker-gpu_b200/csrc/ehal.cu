// Buffered 14-7 van der Waals term (Halgren) -- SURVEY.md §8f rank 1, the first row after the
// electrostatics path.  Same physics as ehal_cu (src/cu/ehal.cu:125-160, src/cu/ehal_cu1.cc,
// include/seq/pair_hal.h:52-92): hydrogen sites reduced along their bond, class-pair radmin/epsilon
// tables, quintic switch between taper and cutoff, forces handed back from the reduced sites to
// the atoms (ehalResolveGradient, ehal.cu:34-62), virial on the reduced coordinates.
//
// Structure (not the reference's 32x32 tiles + exclusion bitmasks + k-side atomics):
//   * the reduced sites are sorted exactly like the atoms (same permutation as the electrostatics
//     list, so one sort serves both) and get their own block boxes and directed Verlet rows
//     (rows.cu, range cutoff + buffer).  Pairs with scale 0 (1-2, 1-3) are dropped while the rows
//     are BUILT -- exclusion is topology -- so the pair kernel has no exclusion logic at all;
//     pairs with any other scale get a (scale-1) correction pass (none in AMOEBA);
//   * one lane group walks one site's row with every lane on a listed pair, the class-pair table
//     sits in shared memory, energies/virial are reduced per CTA into 2^32 fixed point, forces are
//     reduced with shuffles and scattered once per site (own atom + parent atom);
//   * the whole term runs on its own low-priority stream BESIDE the induced-dipole solver, which
//     is latency bound and leaves most of the SMs idle.
#include "apx_internal.h"
#include "rows.cuh"
#include "wrap.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

#define FULL 0xffffffffu
#define EH_G 16

namespace {
__device__ __forceinline__ int bits_int(real v)
{
#ifdef APX_DOUBLE
   return (int)__double_as_longlong(v);
#else
   return __float_as_int(v);
#endif
}
__device__ __forceinline__ real int_bits(int v)
{
#ifdef APX_DOUBLE
   return __longlong_as_double((long long)v);
#else
   return __int_as_float(v);
#endif
}

// reduced sites in sorted order, every step: xred = kred (x_i - x_iv) + x_iv  (ehalReduceXyz_cu1, src/cu/ehal.cu:13-27)
__global__ void k_vdw_sites(int n, int npad, Box b, const double* __restrict__ xyz, const int* __restrict__ perm,
   const int* __restrict__ ired, const real* __restrict__ kred, const int* __restrict__ jvdw, real4* __restrict__ pred)
{
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= npad)
      return;
   real4 o;
   if (s < n) {
      const int i = perm[s], iv = ired[i];
      const double rdn = (double)kred[i];
      const double x = rdn * (xyz[3 * i] - xyz[3 * iv]) + xyz[3 * iv];
      const double y = rdn * (xyz[3 * i + 1] - xyz[3 * iv + 1]) + xyz[3 * iv + 1];
      const double z = rdn * (xyz[3 * i + 2] - xyz[3 * iv + 2]) + xyz[3 * iv + 2];
      real fx, fy, fz;
      wrap_pos(b, x, y, z, o.x, o.y, o.z, fx, fy, fz);
      o.w = int_bits(jvdw[i]);
   } else {
      o.x = o.y = o.z = 0;
      o.w = int_bits(0);
   }
   pred[s] = o;
}

__global__ void k_vdw_static(int n, const int* __restrict__ perm, const int* __restrict__ inv, const int* __restrict__ ired,
   const real* __restrict__ kred, int* __restrict__ ired_s, real* __restrict__ kred_s)
{
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= n)
      return;
   const int i = perm[s];
   ired_s[s] = inv[ired[i]];
   kred_s[s] = kred[i];
}

__global__ void k_vdw_excl_sorted(int nx, const int* __restrict__ ik, const real* __restrict__ sc, const int* __restrict__ inv,
   VdwExcl* __restrict__ out)
{
   int e = blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= nx)
      return;
   VdwExcl p;
   p.i = inv[ik[2 * e]];
   p.k = inv[ik[2 * e + 1]];
   p.s = sc[e] - 1;
   out[e] = p;
}

struct HalPrm {
   real cut, off, off2, ghal, dhal, c1d, c1g, rswinv;      // c1d = (1+dhal)^7, c1g = 1+ghal, rswinv = 1/(cut-off)
};

#ifdef APX_DOUBLE
__device__ __forceinline__ real r_rcp(real x) { return 1.0 / x; }
__device__ __forceinline__ void r_sqrt_pair(real r2, real& r, real& rinv)
{
   r = sqrt(r2);
   rinv = 1.0 / r;
}
#else
// 1/x for the two denominators of the 14-7 function (both > 0.0049): hardware reciprocal + one Newton step, without the special-case
// paths of the correctly rounded __frcp_rn (8 instructions and a branch each, twice per listed pair)
__device__ __forceinline__ real r_rcp(real x)
{
   real y;
   asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
   return fmaf(y, fmaf(-x, y, 1.0f), y);
}
// r enters the energy in the 7th power: the 2-ulp hardware rsqrt gets one Newton step (two FMAs)
__device__ __forceinline__ void r_sqrt_pair(real r2, real& r, real& rinv)
{
   const real y = rsqrtf(r2);
   rinv = y * fmaf(-0.5f * r2 * y, y, 1.5f);
   r = r2 * rinv;
}
#endif

// pair_hal_v2 (include/seq/pair_hal.h:52-92) with vlambda = 1; rvinv = 1/radmin rounded once on the host
// (1 for classes without vdW, whose eps is 0)
template <bool DO_G>
__device__ __forceinline__ void pair_hal(const HalPrm& P, real r, real rvinv, real eps, real& e, real& de)
{
   const real rho = r * rvinv;
   const real rho2 = rho * rho, rho6 = rho2 * rho2 * rho2, rho7 = rho6 * rho;
   const real a = rho + P.dhal, a2 = a * a, a6 = a2 * a2 * a2, a7 = a6 * a;
   const real s1 = r_rcp(a7);
   const real s2 = r_rcp(rho7 + P.ghal);
   const real t1 = P.c1d * s1, t2 = P.c1g * s2;
   e = eps * t1 * (t2 - 2);
   if (DO_G) {
      const real dt1 = -7 * a6 * t1 * s1;
      const real dt2 = -7 * rho6 * t2 * s2;
      de = eps * (dt1 * (t2 - 2) + t1 * dt2) * rvinv;
   }
   if (r > P.cut) {
      // switchTaper5 (include/math/switch.h:23-32)
      const real x = (r - P.off) * P.rswinv;
      const real x2 = x * x;
      const real taper = x2 * x * (6 * x2 - 15 * x + 10);
      if (DO_G) {
         const real w = x * (1 - x);
         const real dtaper = 30 * w * w * P.rswinv;
         de = e * dtaper + de * taper;
      }
      e *= taper;
   }
}

__device__ __forceinline__ long long block_sum_ll(long long x, long long* sm)
{
   for (int o = 16; o > 0; o >>= 1)
      x += __shfl_xor_sync(FULL, x, o);
   const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
   __syncthreads();
   if (l == 0)
      sm[w] = x;
   __syncthreads();
   long long t = 0;
   if (threadIdx.x == 0)
      for (int q = 0; q < (int)(blockDim.x >> 5); ++q)
         t += sm[q];
   return t;      // valid in thread 0
}

__device__ __forceinline__ void add_fixed_d(fixed_t* p, double v)
{
   atomicAdd(p, (fixed_t)(long long)(v * APX_FIXED_SCALE));
}

template <int G, bool DO_G, bool DO_V, bool DO_A>
__global__ void __launch_bounds__(ROWS_BLOCK) k_ehal_rows(int a0, int a1, Box box, HalPrm P, int nj, const real2* __restrict__ tab,
   const int* __restrict__ vstart, const int* __restrict__ vnbr, const real4* __restrict__ pred, const int* __restrict__ ired_s,
   const real* __restrict__ kred_s, fixed_t* gx, fixed_t* gy, fixed_t* gz, fixed_t* vbuf, int* vcnt, int do_e)
{
   extern __shared__ __align__(16) unsigned char smraw[];
   real2* stab = reinterpret_cast<real2*>(smraw);
   __shared__ long long red[ROWS_BLOCK / 32];
   for (int q = threadIdx.x; q < nj * nj; q += blockDim.x)
      stab[q] = tab[q];
   __syncthreads();
   // per-site sums are converted to 2^32 fixed point before they meet the sums of other sites: integer addition is
   // associative, so energy and virial do not depend on the grid size or on which CTA handled a site (they still
   // depend on the order of a site's row, i.e. on the sort -- which differs between one GPU and several)
   long long esum = 0;
   long long vs[6] = {0, 0, 0, 0, 0, 0};
   int cnt = 0;
   ROWS_FOREACH_ATOM(G, a0, a1, i, l, act)
   {
      const real4 pi = pred[i];
      const int ji = bits_int(pi.w) * nj;
      const int beg = vstart[i];
      const int end = act ? vstart[i + 1] : beg;
      V3 f = v3(0, 0, 0);
      real ei = 0;
      real vxx = 0, vyx = 0, vzx = 0, vyy = 0, vzy = 0, vzz = 0;
      for (int q = beg + l; q < end; q += G) {
         const int k = vnbr[q];
         const real4 pk = pred[k];
         real dx = pi.x - pk.x, dy = pi.y - pk.y, dz = pi.z - pk.z;
         apx_image(box, dx, dy, dz);
         const real r2 = dx * dx + dy * dy + dz * dz;
         if (r2 <= P.off2) {
            const real2 t = stab[ji + bits_int(pk.w)];
            real r, rinv;
            r_sqrt_pair(r2, r, rinv);
            real e, de = 0;
            pair_hal<DO_G>(P, r, t.x, t.y, e, de);
            ei += e;
            if (DO_A)
               cnt += e != 0 ? 1 : 0;
            if (DO_G) {
               de *= rinv;
               const real fx = de * dx, fy = de * dy, fz = de * dz;
               f += v3(fx, fy, fz);
               if (DO_V) {
                  vxx += dx * fx, vyx += dy * fx, vzx += dz * fx;
                  vyy += dy * fy, vzy += dz * fy, vzz += dz * fz;
               }
            }
         }
      }
      if (do_e) {
         ei = group_sum<G>(ei);
         if (l == 0 && act)
            esum += (long long)((double)ei * (0.5 * APX_FIXED_SCALE));      // every pair is seen from both of its sites
      }
      if (DO_V) {
         real v6[6] = {vxx, vyx, vzx, vyy, vzy, vzz};
         #pragma unroll
         for (int q = 0; q < 6; ++q) {
            const real s = group_sum<G>(v6[q]);
            if (l == 0 && act)
               vs[q] += (long long)((double)s * (0.5 * APX_FIXED_SCALE));
         }
      }
      if (DO_G) {
         f = group_sum3<G>(f);
         if (l == 0 && act) {
            // ehalResolveGradient (src/cu/ehal.cu:34-62): a reduced site shares its force with the parent atom
            const int iv = ired_s[i];
            if (iv == i) {
               atomic_fixed(gx + i, f.x), atomic_fixed(gy + i, f.y), atomic_fixed(gz + i, f.z);
            } else {
               const real kr = kred_s[i], kv = 1 - kr;
               atomic_fixed(gx + i, kr * f.x), atomic_fixed(gy + i, kr * f.y), atomic_fixed(gz + i, kr * f.z);
               atomic_fixed(gx + iv, kv * f.x), atomic_fixed(gy + iv, kv * f.y), atomic_fixed(gz + iv, kv * f.z);
            }
         }
      }
   }
   if (do_e) {
      long long t = block_sum_ll(esum, red);
      if (threadIdx.x == 0 && t != 0)
         atomicAdd(vbuf, (fixed_t)t);
   }
   if (DO_V) {
      #pragma unroll
      for (int q = 0; q < 6; ++q) {
         long long t = block_sum_ll(vs[q], red);
         if (threadIdx.x == 0 && t != 0)
            atomicAdd(vbuf + 1 + q, (fixed_t)t);
      }
   }
   if (DO_A) {
      for (int o = 16; o > 0; o >>= 1)
         cnt += __shfl_xor_sync(FULL, cnt, o);
      if ((threadIdx.x & 31) == 0 && cnt)
         atomicAdd(vcnt, cnt);
   }
}

// pairs whose scale is neither 0 nor 1: add (scale-1) x the pair, each rank for the sites it owns
template <bool DO_G>
__global__ void k_ehal_excl(int nx, const VdwExcl* __restrict__ ex, int a0, int a1, Box box, HalPrm P, int nj,
   const real2* __restrict__ tab, const real4* __restrict__ pred, const int* __restrict__ ired_s, const real* __restrict__ kred_s,
   fixed_t* gx, fixed_t* gy, fixed_t* gz, fixed_t* vbuf, int do_e, int do_v)
{
   int q = blockIdx.x * blockDim.x + threadIdx.x;
   if (q >= nx)
      return;
   const VdwExcl p = ex[q];
   const bool own_i = p.i >= a0 && p.i < a1, own_k = p.k >= a0 && p.k < a1;
   if (!own_i && !own_k)
      return;
   const real4 pi = pred[p.i], pk = pred[p.k];
   real dx = pi.x - pk.x, dy = pi.y - pk.y, dz = pi.z - pk.z;
   apx_image(box, dx, dy, dz);
   const real r2 = dx * dx + dy * dy + dz * dz;
   if (r2 > P.off2)
      return;
   const real2 t = tab[bits_int(pi.w) * nj + bits_int(pk.w)];
   real r, rinv;
   r_sqrt_pair(r2, r, rinv);
   real e, de = 0;
   pair_hal<DO_G>(P, r, t.x, p.s * t.y, e, de);
   const double share = (own_i ? 0.5 : 0.0) + (own_k ? 0.5 : 0.0);
   if (do_e)
      add_fixed_d(vbuf, share * (double)e);
   if (DO_G) {
      de *= rinv;
      const real f[3] = {de * dx, de * dy, de * dz};
      fixed_t* g[3] = {gx, gy, gz};
      for (int side = 0; side < 2; ++side) {
         const int s = side ? p.k : p.i;
         if (!(side ? own_k : own_i))
            continue;
         const real sg = side ? (real)-1 : (real)1;
         const int iv = ired_s[s];
         const real kr = iv == s ? (real)1 : kred_s[s];
         for (int cdim = 0; cdim < 3; ++cdim) {
            atomic_fixed(g[cdim] + s, sg * kr * f[cdim]);
            if (iv != s)
               atomic_fixed(g[cdim] + iv, sg * (1 - kr) * f[cdim]);
         }
      }
      if (do_v) {
         const real d[3] = {dx, dy, dz};
         add_fixed_d(vbuf + 1, share * (double)(d[0] * f[0]));
         add_fixed_d(vbuf + 2, share * (double)(d[1] * f[0]));
         add_fixed_d(vbuf + 3, share * (double)(d[2] * f[0]));
         add_fixed_d(vbuf + 4, share * (double)(d[1] * f[1]));
         add_fixed_d(vbuf + 5, share * (double)(d[2] * f[1]));
         add_fixed_d(vbuf + 6, share * (double)(d[2] * f[2]));
      }
   }
}

template <class T, class S>
void upload_as(DevBuf<T>& dst, const S* src, size_t count, cudaStream_t st)
{
   std::vector<T> h(count);
   for (size_t q = 0; q < count; ++q)
      h[q] = (T)src[q];
   dst.ensure(count + 1);
   CUDA_CHECK(cudaMemcpyAsync(dst.p, h.data(), sizeof(T) * count, cudaMemcpyHostToDevice, st));
   CUDA_CHECK(cudaStreamSynchronize(st));
}

HalPrm make_prm(const VdwState& V)
{
   HalPrm P;
   P.cut = V.cut, P.off = V.off, P.off2 = V.off * V.off;
   P.ghal = V.ghal, P.dhal = V.dhal;
   P.c1d = (real)pow(1.0 + (double)V.dhal, 7.0);
   P.c1g = (real)(1.0 + (double)V.ghal);
   P.rswinv = V.cut < V.off ? (real)(1.0 / ((double)V.cut - (double)V.off)) : (real)0;
   return P;
}
} // namespace

void apx_vdw_attach_impl(apx_ctx* c, const apx_vdw* v)
{
   if (!v || v->n != c->n)
      APX_THROW("apx_vdw_attach: the vdW description must cover the context's atoms");
   if (v->njvdw <= 0 || !v->ired || !v->kred || !v->jvdw || !v->radmin || !v->epsilon)
      APX_THROW("apx_vdw_attach: null argument");
   const double minedge = std::min(std::min(c->opt.lvec[0], c->opt.lvec[4]), c->opt.lvec[8]);
   if (v->cutoff < 1e6 && v->cutoff > 0.5 * minedge + 1e-9)
      APX_THROW("vdw-cutoff exceeds half the box edge (minimum image would fail)");
   VdwState& V = c->vdw;
   const int n = c->n;
   cudaStream_t st = c->stream;
   V.nj = v->njvdw;
   upload_as(V.ired_o, v->ired, n, st);
   upload_as(V.jvdw_o, v->jvdw, n, st);
   upload_as(V.kred_o, v->kred, n, st);
   for (int i = 0; i < n; ++i)
      if (v->ired[i] < 0 || v->ired[i] >= n || v->jvdw[i] < 0 || v->jvdw[i] >= V.nj)
         APX_THROW("apx_vdw_attach: index out of range");
   {
      std::vector<real2> t((size_t)V.nj * V.nj);
      for (size_t q = 0; q < t.size(); ++q) {
         t[q].x = v->radmin[q] > 0 ? (real)(1.0 / v->radmin[q]) : (real)1;
         t[q].y = (real)v->epsilon[q];
      }
      V.tab.ensure(t.size() + 1);
      CUDA_CHECK(cudaMemcpyAsync(V.tab.p, t.data(), sizeof(real2) * t.size(), cudaMemcpyHostToDevice, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
   }
   // exclusions: scale 0 -> CSR of never-listed partners; scale 1 -> nothing; otherwise correction pairs
   std::vector<int> off(n + 1, 0), lst, xik;
   std::vector<real> xsc;
   for (int e = 0; e < v->nvexclude; ++e) {
      const int i = v->vexclude[2 * e], k = v->vexclude[2 * e + 1];
      if (i < 0 || k < 0 || i >= n || k >= n || i == k)
         APX_THROW("apx_vdw_attach: bad exclusion pair");
      const double s = v->vexclude_scale[e];
      if (s == 0.0) {
         off[i + 1]++;
         off[k + 1]++;
      } else if (s != 1.0) {
         xik.push_back(i);
         xik.push_back(k);
         xsc.push_back((real)s);
      }
   }
   for (int i = 0; i < n; ++i)
      off[i + 1] += off[i];
   lst.resize(off[n] + 1);
   {
      std::vector<int> fill(off.begin(), off.end() - 1);
      for (int e = 0; e < v->nvexclude; ++e)
         if (v->vexclude_scale[e] == 0.0) {
            const int i = v->vexclude[2 * e], k = v->vexclude[2 * e + 1];
            lst[fill[i]++] = k;
            lst[fill[k]++] = i;
         }
   }
   upload_as(V.exoff, off.data(), off.size(), st);
   upload_as(V.exlist, lst.data(), lst.size(), st);
   V.nxs = (int)xsc.size();
   if (V.nxs) {
      upload_as(V.xs_ik, xik.data(), xik.size(), st);
      upload_as(V.xs_sc, xsc.data(), xsc.size(), st);
      V.xs_s.ensure(V.nxs);
   }
   // 1-2..1-5 partners sit at most four bonds apart; reduced sites only move inwards
   V.exrange = (real)8.0;
   V.cut = (real)std::min(v->taper, 1.0e6), V.off = (real)std::min(v->cutoff, 1.0e6);
   V.ghal = (real)v->ghal, V.dhal = (real)v->dhal;
   V.elrc_vol = v->elrc_vol, V.vlrc_vol = v->vlrc_vol;
   const size_t np = c->npad;
   V.pred.ensure(np);
   V.ired_s.ensure(np);
   V.kred_s.ensure(np);
   V.ctr.ensure(c->nblk);
   V.ext.ensure(c->nblk);
   V.vbuf.ensure(8);
   V.vcnt.ensure(2);
   if (!V.stream) {
      int lo = 0, hi = 0;
      CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CUDA_CHECK(cudaStreamCreateWithPriority(&V.stream, cudaStreamNonBlocking, lo));
      CUDA_CHECK(cudaEventCreateWithFlags(&V.ev_go, cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreateWithFlags(&V.ev_done, cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreate(&V.t0));
      CUDA_CHECK(cudaEventCreate(&V.t1));
   }
   const size_t smem = sizeof(real2) * (size_t)V.nj * V.nj;
   if (smem > 200 * 1024)
      APX_THROW("too many vdW classes for the shared-memory pair table");
   if (smem > 40 * 1024) {
#define SET_SMEM(G_, A_, B_, C_)                                                                                          \
   CUDA_CHECK(cudaFuncSetAttribute(k_ehal_rows<EH_G, G_, A_, B_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
      SET_SMEM(false, false, false, 0);
      SET_SMEM(false, false, true, 0);
      SET_SMEM(true, false, false, 0);
      SET_SMEM(true, true, false, 0);
      SET_SMEM(true, false, true, 0);
      SET_SMEM(true, true, true, 0);
#undef SET_SMEM
   }
   V.on = 1;
   c->list_valid = 0;       // the next refresh builds the vdW rows together with the electrostatics rows
}

void apx_vdw_refresh(apx_ctx* c, bool rebuilt)
{
   VdwState& V = c->vdw;
   const int n = c->n;
   cudaStream_t st = c->stream;
   k_vdw_sites<<<(c->npad + 255) / 256, 256, 0, st>>>(n, c->npad, c->box, c->xyz_d, c->perm, V.ired_o, V.kred_o, V.jvdw_o, V.pred);
   APX_COUNT_LAUNCH(c);
   if (!rebuilt)
      return;
   k_vdw_static<<<(n + 255) / 256, 256, 0, st>>>(n, c->perm, c->inv, V.ired_o, V.kred_o, V.ired_s, V.kred_s);
   APX_COUNT_LAUNCH(c);
   if (V.nxs) {
      k_vdw_excl_sorted<<<(V.nxs + 255) / 256, 256, 0, st>>>(V.nxs, V.xs_ik, V.xs_sc, c->inv, V.xs_s);
      APX_COUNT_LAUNCH(c);
   }
   apx_block_boxes(c, V.pred, V.ctr, V.ext);
   // The Verlet range is NOT clipped to half the cell: a row lists partner k once, whichever image is nearest, and the pair
   // kernel takes the minimum image of the current positions, so off <= L/2 < off + buffer loses no pair -- whereas a clipped
   // range with the unclipped buffer/2 rebuild criterion could (a pair just outside the clipped range may enter the cutoff
   // between rebuilds).  The block-box test is a lower bound on the minimum-image distance for any range.
   const real range = V.off + (real)c->opt.list_buffer;
   apx_rows_build_on(c, V.rows, V.pred, V.ctr, V.ext, range, V.exoff, V.exlist, V.exrange, false);
   c->stats.nverlet_vdw = V.rows.nverlet;
}

void apx_vdw_launch(apx_ctx* c, int vers)
{
   VdwState& V = c->vdw;
   const bool do_e = vers & APX_ENERGY, do_g = vers & APX_GRAD, do_v = (vers & APX_VIRIAL) && do_g, do_a = vers & APX_ANALYZ;
   cudaStream_t vs = V.stream;
   CUDA_CHECK(cudaEventRecord(V.ev_go, c->stream));
   CUDA_CHECK(cudaStreamWaitEvent(vs, V.ev_go, 0));
   CUDA_CHECK(cudaMemsetAsync(V.vbuf.p, 0, sizeof(fixed_t) * 8, vs));
   CUDA_CHECK(cudaMemsetAsync(V.vcnt.p, 0, sizeof(int) * 2, vs));
   const HalPrm P = make_prm(V);
   const size_t smem = sizeof(real2) * (size_t)V.nj * V.nj;
   const int grid = rows_grid<EH_G>(c, 8);      // leaves room for the solver's CTAs on every SM
   cudaEventRecord(V.t0, vs);
   if (V.rows.nverlet > 0 && (do_e || do_g)) {
#define LAUNCH_EH(G_, V_, A_)                                                                                             \
   k_ehal_rows<EH_G, G_, V_, A_><<<grid, ROWS_BLOCK, smem, vs>>>(c->a0, c->a1, c->box, P, V.nj, V.tab, V.rows.vstart, V.rows.vnbr, \
      V.pred, V.ired_s, V.kred_s, c->gx, c->gy, c->gz, V.vbuf, V.vcnt, do_e ? 1 : 0)
      if (do_g && do_v && do_a) LAUNCH_EH(true, true, true);
      else if (do_g && do_v) LAUNCH_EH(true, true, false);
      else if (do_g && do_a) LAUNCH_EH(true, false, true);
      else if (do_g) LAUNCH_EH(true, false, false);
      else if (do_a) LAUNCH_EH(false, false, true);
      else LAUNCH_EH(false, false, false);
#undef LAUNCH_EH
      APX_COUNT_LAUNCH(c);
   }
   cudaEventRecord(V.t1, vs);
   if (V.nxs > 0) {
      const int g = (V.nxs + 127) / 128;
      if (do_g)
         k_ehal_excl<true><<<g, 128, 0, vs>>>(V.nxs, V.xs_s, c->a0, c->a1, c->box, P, V.nj, V.tab, V.pred, V.ired_s, V.kred_s, c->gx, c->gy,
            c->gz, V.vbuf, do_e, do_v);
      else
         k_ehal_excl<false><<<g, 128, 0, vs>>>(V.nxs, V.xs_s, c->a0, c->a1, c->box, P, V.nj, V.tab, V.pred, V.ired_s, V.kred_s, c->gx, c->gy,
            c->gz, V.vbuf, do_e, do_v);
      APX_COUNT_LAUNCH(c);
   }
   CUDA_CHECK(cudaEventRecord(V.ev_done, vs));
}

// copies = false: only the stream dependency (the caller puts more work in front of the copies and calls apx_vdw_copy_out itself)
void apx_vdw_join(apx_ctx* c, bool copies)
{
   VdwState& V = c->vdw;
   CUDA_CHECK(cudaStreamWaitEvent(c->stream, V.ev_done, 0));
   if (c->dist.on) {
      apx_dist_allreduce_u64(c, V.vbuf.p, 8);
      apx_dist_allreduce_i32(c, V.vcnt.p, 2);
   }
   if (copies)
      apx_vdw_copy_out(c);
}

void apx_vdw_copy_out(apx_ctx* c)
{
   VdwState& V = c->vdw;
   // pinned landing zone behind the electrostatics scalars (apx_ctx::red_h)
   CUDA_CHECK(cudaMemcpyAsync(c->red_h + 2048, V.vbuf.p, sizeof(fixed_t) * 8, cudaMemcpyDeviceToHost, c->stream));
   CUDA_CHECK(cudaMemcpyAsync(c->red_h + 2048 + 64, V.vcnt.p, sizeof(int) * 2, cudaMemcpyDeviceToHost, c->stream));
}

// call with the main stream synchronised after apx_vdw_join
void apx_vdw_collect(apx_ctx* c, int vers, apx_energy_result* r)
{
   VdwState& V = c->vdw;
   const bool do_e = vers & APX_ENERGY, do_g = vers & APX_GRAD, do_v = (vers & APX_VIRIAL) && do_g, do_a = vers & APX_ANALYZ;
   const fixed_t* hb = reinterpret_cast<const fixed_t*>(c->red_h + 2048);
   const int* hc = reinterpret_cast<const int*>(c->red_h + 2048 + 64);
   cudaEventElapsedTime(&c->stats.ms_ehal, V.t0, V.t1);
   auto fx = [](fixed_t v) { return (double)(long long)v / APX_FIXED_SCALE; };
   const double vol = fabs((double)c->box.volume);
   r->ev = 0;
   r->nev = 0;
   if (do_e)
      r->ev = fx(hb[0]) + (V.elrc_vol != 0 ? V.elrc_vol / vol : 0.0);      // src/evdw.cpp:493-499
   if (do_a)
      r->nev = hc[0] / 2;
   if (do_v) {
      const double xx = fx(hb[1]), yx = fx(hb[2]), zx = fx(hb[3]), yy = fx(hb[4]), zy = fx(hb[5]), zz = fx(hb[6]);
      const double corr = V.vlrc_vol != 0 ? V.vlrc_vol / vol : 0.0;      // src/evdw.cpp:500-511
      r->virial[0] += xx + corr, r->virial[1] += yx, r->virial[2] += zx;
      r->virial[3] += yx, r->virial[4] += yy + corr, r->virial[5] += zy;
      r->virial[6] += zx, r->virial[7] += zy, r->virial[8] += zz + corr;
   }
}

void apx_vdw_destroy(apx_ctx* c)
{
   VdwState& V = c->vdw;
   if (V.stream) {
      cudaStreamSynchronize(V.stream);
      cudaStreamDestroy(V.stream);
      cudaEventDestroy(V.ev_go), cudaEventDestroy(V.ev_done), cudaEventDestroy(V.t0), cudaEventDestroy(V.t1);
      V.stream = nullptr;
   }
   V.ired_o.release(), V.jvdw_o.release(), V.kred_o.release(), V.tab.release(), V.exoff.release(), V.exlist.release();
   V.xs_ik.release(), V.xs_sc.release(), V.xs_s.release(), V.pred.release(), V.ired_s.release(), V.kred_s.release();
   V.ctr.release(), V.ext.release(), V.vbuf.release(), V.vcnt.release();
   V.rows.vstart.release(), V.rows.vcnt.release(), V.rows.vnbr.release(), V.rows.nbr.release(), V.rows.cnt.release();
   V.rows.prev_o.release(), V.rows.capstart.release(), V.rows.vpad.release(), V.rows.oflow.release();
   V.rows.cntu.release(), V.rows.sctr.release(), V.rows.sext.release(), V.rows.total.release();
   V.on = 0;
}
