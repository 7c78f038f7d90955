// Local-frame handling: chirality check + rotation of the local-frame multipoles into the lab
// frame (chkpole_cu / rotpole_cu, src/cu/amoeba/rotpole.cu:7-37, math include/seq/rotpole.h:9-223)
// and the conversion of per-site torques into forces on the frame-defining atoms
// (torque_cu, src/cu/amoeba/torque.cu:8-383; Tinker torque.f).
//
// Frame geometry is evaluated in double from the caller-order f64 coordinates (frames never use
// periodic images: molecules are whole), results land directly in the SORTED SoA multipole arrays.
#include "apx_internal.h"
#include <algorithm>

namespace {
struct v3 {
   double x, y, z;
};
__device__ __forceinline__ v3 mk(double x, double y, double z)
{
   v3 r = {x, y, z};
   return r;
}
__device__ __forceinline__ v3 operator+(v3 a, v3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ v3 operator-(v3 a, v3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ v3 operator*(double s, v3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ double dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ v3 cross(v3 a, v3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ v3 unit(v3 a) { return rsqrt(dot(a, a)) * a; }
__device__ __forceinline__ v3 ld(const double* p, int i) { return mk(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }

enum { F_NONE = 0, F_ZONLY = 1, F_ZTHENX = 2, F_BISECTOR = 3, F_ZBISECT = 4, F_3FOLD = 5 };

// chirality: flip the y-dependent components when the signed volume disagrees with the sign
// stored in yaxis (a Z-then-X frame with a third, chirality-defining atom)
__global__ void k_chkpole(int n, const double* __restrict__ xyz, int* __restrict__ zaxis, real* __restrict__ pole)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n)
      return;
   int k = zaxis[4 * i + 2];
   if (zaxis[4 * i + 3] != F_ZTHENX || k == 0)
      return;
   int id = abs(k) - 1;
   v3 pd = ld(xyz, id);
   v3 a = ld(xyz, i) - pd, b = ld(xyz, zaxis[4 * i]) - pd, c = ld(xyz, zaxis[4 * i + 1]) - pd;
   double vol = dot(a, cross(b, c));
   if ((k < 0 && vol > 0) || (k > 0 && vol < 0)) {
      zaxis[4 * i + 2] = -k;
      pole[10 * i + 2] = -pole[10 * i + 2];   // dy
      pole[10 * i + 7] = -pole[10 * i + 7];   // qxy
      pole[10 * i + 9] = -pole[10 * i + 9];   // qyz
   }
}

__device__ __forceinline__ v3 default_x(v3 z)
{
   // Z-Only frames: any direction not parallel to z (Tinker rotmat picks by |z.x| > 0.866)
   return fabs(z.x) > 0.866 ? mk(0, 1, 0) : mk(1, 0, 0);
}

__global__ void k_rotpole(int n, const double* __restrict__ xyz, const int* __restrict__ perm, const int* __restrict__ zaxis,
   const real* __restrict__ pole, real4* __restrict__ mp0, real4* __restrict__ mp1, real2* __restrict__ mp2)
{
   int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= n)
      return;
   int i = perm[s];
   int iz = zaxis[4 * i], ix = zaxis[4 * i + 1], iy = abs(zaxis[4 * i + 2]) - 1, axe = zaxis[4 * i + 3];
   v3 ex = mk(1, 0, 0), ey = mk(0, 1, 0), ez = mk(0, 0, 1);
   if (axe != F_NONE) {
      v3 p = ld(xyz, i);
      ez = unit(ld(xyz, iz) - p);
      ex = (axe == F_ZONLY) ? default_x(ez) : unit(ld(xyz, ix) - p);
      if (axe == F_BISECTOR) {
         ez = unit(ez + ex);
      } else if (axe == F_ZBISECT) {
         v3 t = unit(ld(xyz, iy) - p);
         ex = unit(ex + t);
      } else if (axe == F_3FOLD) {
         v3 t = unit(ld(xyz, iy) - p);
         ez = unit(ez + ex + t);
      }
      ex = unit(ex - dot(ex, ez) * ez);
      ey = cross(ez, ex);
   }
   const real* pl = pole + 10 * i;
   double dl[3] = {pl[1], pl[2], pl[3]};
   double ql[3][3] = {{pl[4], pl[7], pl[8]}, {pl[7], pl[5], pl[9]}, {pl[8], pl[9], pl[6]}};
   double A[3][3] = {{ex.x, ex.y, ex.z}, {ey.x, ey.y, ey.z}, {ez.x, ez.y, ez.z}};   // rows = local axes in lab frame
   double dg[3], qg[3][3];
   for (int a = 0; a < 3; ++a) {
      dg[a] = dl[0] * A[0][a] + dl[1] * A[1][a] + dl[2] * A[2][a];
      for (int b = 0; b < 3; ++b) {
         double t = 0;
         for (int k = 0; k < 3; ++k)
            for (int m = 0; m < 3; ++m)
               t += A[k][a] * A[m][b] * ql[k][m];
         qg[a][b] = t;
      }
   }
   real4 o0, o1;
   real2 o2;
   o0.x = pl[0];
   o0.y = (real)dg[0];
   o0.z = (real)dg[1];
   o0.w = (real)dg[2];
   o1.x = (real)qg[0][0];
   o1.y = (real)qg[0][1];
   o1.z = (real)qg[0][2];
   o1.w = (real)qg[1][1];
   o2.x = (real)qg[1][2];
   o2.y = (real)qg[2][2];
   mp0[s] = o0;
   mp1[s] = o1;
   mp2[s] = o2;
}

__global__ void k_zero_pad(int n, int npad, real4* __restrict__ mp0, real4* __restrict__ mp1, real2* __restrict__ mp2)
{
   int s = n + blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= npad)
      return;
   real4 z4 = {0, 0, 0, 0};
   real2 z2 = {0, 0};
   mp0[s] = z4;
   mp1[s] = z4;
   mp2[s] = z2;
}

__device__ __forceinline__ void add_fixed(fixed_t* g, int s, double v)
{
   atomicAdd(&g[s], (fixed_t)(long long)(v * APX_FIXED_SCALE));
}

template <bool DO_V>
__global__ void k_torque(int a0, int n, const double* __restrict__ xyz, const int* __restrict__ perm, const int* __restrict__ inv,
   const int* __restrict__ zaxis, const real* __restrict__ trq, fixed_t* __restrict__ gx, fixed_t* __restrict__ gy,
   fixed_t* __restrict__ gz, double* __restrict__ vir)
{
   int s = a0 + blockIdx.x * blockDim.x + threadIdx.x;      // owned atoms a0 <= s < n
   double v[6] = {0, 0, 0, 0, 0, 0};
   if (s < n) {
      int i = perm[s];
      int axe = zaxis[4 * i + 3];
      if (axe != F_NONE) {
         int ia = zaxis[4 * i], ic = zaxis[4 * i + 1], id = abs(zaxis[4 * i + 2]) - 1;
         v3 p = ld(xyz, i);
         v3 t = mk(trq[3 * s], trq[3 * s + 1], trq[3 * s + 2]);
         v3 du_ = ld(xyz, ia) - p;
         double lu = sqrt(dot(du_, du_));
         v3 u = (1.0 / lu) * du_;
         v3 vv;
         double lv = 1.0;
         if (axe != F_ZONLY) {
            v3 d = ld(xyz, ic) - p;
            lv = sqrt(dot(d, d));
            vv = (1.0 / lv) * d;
         } else {
            vv = default_x(u);
         }
         v3 w;
         double lw = 1.0;
         if (axe == F_ZBISECT || axe == F_3FOLD) {
            v3 d = ld(xyz, id) - p;
            lw = sqrt(dot(d, d));
            w = (1.0 / lw) * d;
         } else {
            w = unit(cross(u, vv));
         }
         // work done by an infinitesimal rotation about each axis
         double pu = -dot(t, u), pv = -dot(t, vv), pw = -dot(t, w);
         v3 fz = mk(0, 0, 0), fx = mk(0, 0, 0), fy = mk(0, 0, 0);
         if (axe == F_ZONLY || axe == F_ZTHENX || axe == F_BISECTOR) {
            v3 nuv = unit(cross(vv, u)), nuw = unit(cross(w, u));
            double c = dot(u, vv);
            double sn = sqrt(1.0 - c * c);
            if (axe == F_ZONLY) {
               fz = (pv / (lu * sn)) * nuv + (pw / lu) * nuw;
            } else if (axe == F_ZTHENX) {
               fz = (pv / (lu * sn)) * nuv + (pw / lu) * nuw;
               fx = (-pu / (lv * sn)) * nuv;
            } else {
               v3 nvw = unit(cross(w, vv));
               fz = (pv / (lu * sn)) * nuv + (0.5 * pw / lu) * nuw;
               fx = (-pu / (lv * sn)) * nuv + (0.5 * pw / lv) * nvw;
            }
         } else if (axe == F_ZBISECT) {
            v3 r = unit(vv + w);
            v3 sx = unit(cross(u, r));
            v3 nur = unit(cross(r, u)), nus = unit(cross(sx, u));
            double cur = dot(u, r);
            double sur = sqrt(1.0 - cur * cur);
            double cvs = dot(vv, sx), cws = dot(w, sx);
            double svs = sqrt(1.0 - cvs * cvs), sws = sqrt(1.0 - cws * cws);
            v3 t1 = unit(vv - cvs * sx), t2 = unit(w - cws * sx);
            double c1 = dot(u, t1), c2 = dot(u, t2);
            double denom = sqrt(1.0 - c1 * c1) + sqrt(1.0 - c2 * c2);
            double pr = -dot(t, r), ps = -dot(t, sx);
            fz = (pr / (lu * sur)) * nur + (ps / lu) * nus;
            fx = (pu / (lv * denom)) * (svs * sx - cvs * t1);
            fy = (pu / (lw * denom)) * (sws * sx - cws * t2);
         } else {   // 3-Fold: each arm gets the torque about the bisector of the other two
            v3 pp = u + vv + w;
            double lp = sqrt(dot(pp, pp));
            pp = (1.0 / lp) * pp;
            v3 arms[3] = {u, vv, w};
            double lens[3] = {lu, lv, lw};
            v3 out[3];
            for (int a = 0; a < 3; ++a) {
               v3 c = arms[a];
               v3 r = unit(arms[(a + 1) % 3] + arms[(a + 2) % 3]);
               double crc = dot(r, c);
               double src = sqrt(1.0 - crc * crc);
               v3 dl = unit(cross(r, c));
               v3 ep = cross(dl, c);
               double pr = -dot(t, r), pd = -dot(t, dl);
               out[a] = (pr / (lens[a] * src)) * dl + (pd * dot(c, pp) / (lens[a] * lp)) * ep;
            }
            fz = out[0];
            fx = out[1];
            fy = out[2];
         }
         int sa = inv[ia];
         add_fixed(gx, sa, fz.x);
         add_fixed(gy, sa, fz.y);
         add_fixed(gz, sa, fz.z);
         v3 fb = fz + fx + fy;
         add_fixed(gx, s, -fb.x);
         add_fixed(gy, s, -fb.y);
         add_fixed(gz, s, -fb.z);
         if (axe != F_ZONLY) {
            int sc = inv[ic];
            add_fixed(gx, sc, fx.x);
            add_fixed(gy, sc, fx.y);
            add_fixed(gz, sc, fx.z);
         }
         if (axe == F_ZBISECT || axe == F_3FOLD) {
            int sd = inv[id];
            add_fixed(gx, sd, fy.x);
            add_fixed(gy, sd, fy.y);
            add_fixed(gz, sd, fy.z);
         }
         if (DO_V) {
            v3 rz = ld(xyz, ia) - p;
            v3 rx = (ic >= 0) ? ld(xyz, ic) - p : mk(0, 0, 0);
            v3 ry = (id >= 0) ? ld(xyz, id) - p : mk(0, 0, 0);
            v[0] = rx.x * fx.x + ry.x * fy.x + rz.x * fz.x;
            v[1] = 0.5 * (rx.y * fx.x + ry.y * fy.x + rz.y * fz.x + rx.x * fx.y + ry.x * fy.y + rz.x * fz.y);
            v[2] = 0.5 * (rx.z * fx.x + ry.z * fy.x + rz.z * fz.x + rx.x * fx.z + ry.x * fy.z + rz.x * fz.z);
            v[3] = rx.y * fx.y + ry.y * fy.y + rz.y * fz.y;
            v[4] = 0.5 * (rx.z * fx.y + ry.z * fy.y + rz.z * fz.y + rx.y * fx.z + ry.y * fy.z + rz.y * fz.z);
            v[5] = rx.z * fx.z + ry.z * fy.z + rz.z * fz.z;
         }
      }
   }
   if (DO_V) {
      // warp reduce then one atomic per warp
      #pragma unroll
      for (int q = 0; q < 6; ++q) {
         double x = v[q];
         for (int o = 16; o > 0; o >>= 1)
            x += __shfl_xor_sync(0xffffffffu, x, o);
         if ((threadIdx.x & 31) == 0 && x != 0.0)
            atomicAdd(&vir[q], x);
      }
   }
}
} // namespace

void apx_rotpole(apx_ctx* c)
{
   int n = c->n;
   k_chkpole<<<(n + 255) / 256, 256, 0, c->stream>>>(n, c->xyz_d, c->zaxis, c->pole);
   k_rotpole<<<(n + 127) / 128, 128, 0, c->stream>>>(n, c->xyz_d, c->perm, c->zaxis, c->pole, c->mp0, c->mp1, c->mp2);
   if (c->npad > n)
      k_zero_pad<<<1, 32, 0, c->stream>>>(n, c->npad, c->mp0, c->mp1, c->mp2);
   c->stats.kernel_launches += 3;
   c->mpole_inited = 1;
   c->mpole_pme_valid = 0;
}

// dbuf layout: see mplar.cu (vir_trq accumulates into dbuf[8..13])
void apx_torque(apx_ctx* c, bool do_v)
{
   int no = std::max(1, c->a1 - c->a0);
   double* vir = c->dbuf.p + 8;
   if (do_v)
      k_torque<true><<<(no + 127) / 128, 128, 0, c->stream>>>(c->a0, c->a1, c->xyz_d, c->perm, c->inv, c->zaxis, c->trq, c->gx, c->gy, c->gz, vir);
   else
      k_torque<false><<<(no + 127) / 128, 128, 0, c->stream>>>(c->a0, c->a1, c->xyz_d, c->perm, c->inv, c->zaxis, c->trq, c->gx, c->gy, c->gz, vir);
   APX_COUNT_LAUNCH(c);
}
