// oracle/_ref/libref_realspace.so -- the REFERENCE'S OWN real-space pair arithmetic of the AMOEBA electrostatics path,
// executed on the CPU.  TEST INFRASTRUCTURE ONLY.
//
// This driver (our code) includes include/seq/pair_mpole.h, pair_polar.h and pair_field.h where they lie under
// /root/reference (-DTINKER_DOUBLE_PRECISION, oracle/Makefile; nothing is copied into this repository) and calls
// pair_mpole / pair_polar / pair_dfield / pair_ufield over a pair list handed in by the test, with the reference's own
// two-step treatment of scaled pairs: every pair inside the cutoff with the EWALD form and scale 1
// (src/acc/amoeba/empoleewald.cpp:60-76, epolarewald.cpp:112-125, fieldewald.cpp:75-106, 296-322), then the pairs whose
// scale differs from 1 once more with the NON_EWALD form and (scale - 1) (empoleewald.cpp:140-176, epolarewald.cpp:
// 197-290, fieldewald.cpp:135-200, 350-410).  The torque of the polarization term is assembled from the ufld / dufld
// sums as epolarewald.cpp:330-354 does.  What it pins: the oracle's real-space energies, forces, torques and fields at
// dhfr2 size (1.6 M pairs) against the reference's arithmetic.
#include "ff/amoeba/mpole.h"
#include "seq/bsplgen.h"
#include "seq/pair_field.h"
#include "seq/pair_hal.h"
#include "seq/pair_mpole.h"
#include "seq/pair_polar.h"
#include <cstring>
#include <vector>

using namespace tinker;

namespace {
struct Atom {
   real c, dx, dy, dz, qxx, qxy, qxz, qyy, qyz, qzz;
};
inline Atom load(const double* rp, int i)
{
   const double* p = rp + 10 * (size_t)i;      // MPL_PME order: c, x, y, z, xx, yy, zz, xy, xz, yz (include/ff/amoeba/mpole.h:6-15)
   return Atom{(real)p[MPL_PME_0], (real)p[MPL_PME_X], (real)p[MPL_PME_Y], (real)p[MPL_PME_Z], (real)p[MPL_PME_XX], (real)p[MPL_PME_XY],
      (real)p[MPL_PME_XZ], (real)p[MPL_PME_YY], (real)p[MPL_PME_YZ], (real)p[MPL_PME_ZZ]};
}
#define MP(a) a.c, a.dx, a.dy, a.dz, a.qxx, a.qxy, a.qxz, a.qyy, a.qyz, a.qzz

// pairwise virial sums {xx, xy, xz, yy, yz, zz} of the multipole and polarization forces, formed as empoleewald.cpp:100-107 and
// epolarewald.cpp:163-170 do; null = not wanted (set by ref_realspace_virial around a ref_realspace_eval call)
double *s_vm = nullptr, *s_vp = nullptr;
inline void add_virial(double* v, real xr, real yr, real zr, real fx, real fy, real fz)
{
   v[0] += -xr * fx, v[1] += -0.5f * (yr * fx + xr * fy), v[2] += -0.5f * (zr * fx + xr * fz);
   v[3] += -yr * fy, v[4] += -0.5f * (zr * fy + yr * fz), v[5] += -zr * fz;
}

template <class ETYP>
void one_pair(int i, int k, real r2, real xr, real yr, real zr, real ms, real ds, real ps, real us, const Atom& A, const Atom& B,
   const double* pdamp, real pga, const double* ud, const double* up, real f, real aewald, double* em, double* ep, double* gm, double* tm,
   double* gp, double* ufld, double* dufld, double* fd, double* fp, double* ufd, double* ufp)
{
   const real pdi = pdamp[i], pdk = pdamp[k];
   real e;
   // ---- multipole energy / force / torque
   PairMPoleGrad pg;
   pair_mpole<true, true, ETYP>(r2, xr, yr, zr, ms, MP(A), MP(B), f, aewald, e, pg);
   *em += e;
   gm[3 * i] += pg.frcx, gm[3 * i + 1] += pg.frcy, gm[3 * i + 2] += pg.frcz;
   gm[3 * k] -= pg.frcx, gm[3 * k + 1] -= pg.frcy, gm[3 * k + 2] -= pg.frcz;
   if (s_vm)
      add_virial(s_vm, xr, yr, zr, pg.frcx, pg.frcy, pg.frcz);
   for (int q = 0; q < 3; ++q)
      tm[3 * i + q] += pg.ttmi[q], tm[3 * k + q] += pg.ttmk[q];
   // ---- permanent field, d and p scalings
   real3 fid = make_real3(0, 0, 0), fip = make_real3(0, 0, 0), fkd = make_real3(0, 0, 0), fkp = make_real3(0, 0, 0);
   pair_dfield<ETYP>(r2, xr, yr, zr, ds, ps, MP(A), pdi, pga, MP(B), pdk, pga, aewald, fid, fip, fkd, fkp);
   fd[3 * i] += fid.x, fd[3 * i + 1] += fid.y, fd[3 * i + 2] += fid.z;
   fp[3 * i] += fip.x, fp[3 * i + 1] += fip.y, fp[3 * i + 2] += fip.z;
   fd[3 * k] += fkd.x, fd[3 * k + 1] += fkd.y, fd[3 * k + 2] += fkd.z;
   fp[3 * k] += fkp.x, fp[3 * k + 1] += fkp.y, fp[3 * k + 2] += fkp.z;
   if (!ud)
      return;
   // ---- mutual field of the dipoles handed in
   fid = fip = fkd = fkp = make_real3(0, 0, 0);
   pair_ufield<ETYP>(r2, xr, yr, zr, us, (real)ud[3 * i], (real)ud[3 * i + 1], (real)ud[3 * i + 2], (real)up[3 * i], (real)up[3 * i + 1],
      (real)up[3 * i + 2], pdi, pga, (real)ud[3 * k], (real)ud[3 * k + 1], (real)ud[3 * k + 2], (real)up[3 * k], (real)up[3 * k + 1], (real)up[3 * k + 2],
      pdk, pga, aewald, fid, fip, fkd, fkp);
   ufd[3 * i] += fid.x, ufd[3 * i + 1] += fid.y, ufd[3 * i + 2] += fid.z;
   ufp[3 * i] += fip.x, ufp[3 * i + 1] += fip.y, ufp[3 * i + 2] += fip.z;
   ufd[3 * k] += fkd.x, ufd[3 * k + 1] += fkd.y, ufd[3 * k + 2] += fkd.z;
   ufp[3 * k] += fkp.x, ufp[3 * k + 1] += fkp.y, ufp[3 * k + 2] += fkp.z;
   // ---- polarization energy / force / torque pieces (f carries the factor one half, epolarewald.cpp:34)
   PairPolarGrad pp;
   pair_polar<true, true, ETYP>(r2, xr, yr, zr, ds, ps, us, MP(A), (real)ud[3 * i], (real)ud[3 * i + 1], (real)ud[3 * i + 2], (real)up[3 * i],
      (real)up[3 * i + 1], (real)up[3 * i + 2], pdi, pga, MP(B), (real)ud[3 * k], (real)ud[3 * k + 1], (real)ud[3 * k + 2], (real)up[3 * k], (real)up[3 * k + 1],
      (real)up[3 * k + 2], pdk, pga, (real)0.5 * f, aewald, e, pp);
   *ep += e;
   gp[3 * i] += pp.frcx, gp[3 * i + 1] += pp.frcy, gp[3 * i + 2] += pp.frcz;
   gp[3 * k] -= pp.frcx, gp[3 * k + 1] -= pp.frcy, gp[3 * k + 2] -= pp.frcz;
   if (s_vp)
      add_virial(s_vp, xr, yr, zr, pp.frcx, pp.frcy, pp.frcz);
   for (int q = 0; q < 3; ++q)
      ufld[3 * i + q] += pp.ufldi[q], ufld[3 * k + q] += pp.ufldk[q];
   for (int q = 0; q < 6; ++q)
      dufld[6 * i + q] += pp.dufldi[q], dufld[6 * k + q] += pp.dufldk[q];
}
}

// pairs: pi[p] < pk[p] or any order; R[p] = x_k - x_i after the minimum-image shift; scale[p] = {m, d, p, u}; pga[p] = the pair's
// Thole width (thlval[jpolar_i][jpolar_k]); ud / up may be null (multipole + permanent field only).  All outputs are zeroed here.
extern "C" int ref_realspace_eval(int n, long long npair, const int* pi, const int* pk, const double* R, const double* scale, const double* rpole,
   const double* pdamp, const double* pga, const double* ud, const double* up, double f, double aewald, int ewald, double* em, double* ep, double* gm,
   double* tm, double* gp, double* tp, double* fd, double* fp, double* ufd, double* ufp)
{
   *em = *ep = 0;
   for (double* a : {gm, tm, gp, tp, fd, fp, ufd, ufp})
      std::memset(a, 0, sizeof(double) * 3 * (size_t)n);
   std::vector<double> ufld(3 * (size_t)n, 0.0), dufld(6 * (size_t)n, 0.0);
   for (long long p = 0; p < npair; ++p) {
      const int i = pi[p], k = pk[p];
      const real xr = R[3 * p], yr = R[3 * p + 1], zr = R[3 * p + 2];
      const real r2 = xr * xr + yr * yr + zr * zr;
      const Atom A = load(rpole, i), B = load(rpole, k);
      const double* s = scale + 4 * p;
      if (ewald) {
         one_pair<EWALD>(i, k, r2, xr, yr, zr, 1, 1, 1, 1, A, B, pdamp, (real)pga[p], ud, up, (real)f, (real)aewald, em, ep, gm, tm, gp, ufld.data(),
            dufld.data(), fd, fp, ufd, ufp);
         if (s[0] != 1 || s[1] != 1 || s[2] != 1 || s[3] != 1)
            one_pair<NON_EWALD>(i, k, r2, xr, yr, zr, (real)(s[0] - 1), (real)(s[1] - 1), (real)(s[2] - 1), (real)(s[3] - 1), A, B, pdamp, (real)pga[p], ud,
               up, (real)f, 0, em, ep, gm, tm, gp, ufld.data(), dufld.data(), fd, fp, ufd, ufp);
      } else {
         one_pair<NON_EWALD>(i, k, r2, xr, yr, zr, (real)s[0], (real)s[1], (real)s[2], (real)s[3], A, B, pdamp, (real)pga[p], ud, up, (real)f, 0, em, ep, gm,
            tm, gp, ufld.data(), dufld.data(), fd, fp, ufd, ufp);
      }
   }
   if (ud)
      for (int i = 0; i < n; ++i) {      // src/acc/amoeba/epolarewald.cpp:330-354
         const Atom a = load(rpole, i);
         const double* u = &ufld[3 * (size_t)i];
         const double* d = &dufld[6 * (size_t)i];
         tp[3 * i] = a.dz * u[1] - a.dy * u[2] + a.qxz * d[1] - a.qxy * d[3] + 2 * a.qyz * (d[2] - d[5]) + (a.qzz - a.qyy) * d[4];
         tp[3 * i + 1] = a.dx * u[2] - a.dz * u[0] - a.qyz * d[1] + a.qxy * d[4] + 2 * a.qxz * (d[5] - d[0]) + (a.qxx - a.qzz) * d[3];
         tp[3 * i + 2] = a.dy * u[0] - a.dx * u[1] + a.qyz * d[3] - a.qxz * d[4] + 2 * a.qxy * (d[0] - d[2]) + (a.qyy - a.qxx) * d[1];
      }
   return 0;
}

// Ask the next ref_realspace_eval calls to add the pairwise virial sums into vm6 / vp6 (zeroed here; null switches it off again).
extern "C" void ref_realspace_virial(double* vm6, double* vp6)
{
   s_vm = vm6, s_vp = vp6;
   if (vm6)
      std::memset(vm6, 0, 6 * sizeof(double));
   if (vp6)
      std::memset(vp6, 0, 6 * sizeof(double));
}

// B-spline weights and their first three derivatives of order-5 PME at fractional offsets w[0..m): bsplgen<4> of
// include/seq/bsplgen.h, out[m][5][4] (what bsplineFill / the spread and gather loops of src/acc/pme.cpp:60-110 evaluate).
extern "C" int ref_bspline5(int m, const double* w, double* out)
{
   for (int q = 0; q < m; ++q) {
      real th[5 * 4];
      bsplgen<4>((real)w[q], th, 5);
      for (int j = 0; j < 20; ++j)
         out[20 * (size_t)q + j] = th[j];
   }
   return 0;
}

// Buffered 14-7 energy and dE/dr of m pairs at distances r with pair parameters rv (radmin) / eps (already scaled):
// pair_hal_v2<true, 1> of include/seq/pair_hal.h (lambda = 1: no soft core), tapered between evcut and evoff.
extern "C" int ref_hal(long long m, const double* r, const double* rv, const double* eps, double evcut, double evoff, double ghal, double dhal,
   double* e_out, double* de_out)
{
   for (long long q = 0; q < m; ++q) {
      real e, de;
      pair_hal_v2<true, 1>((real)r[q], 1, (real)rv[q], (real)eps[q], (real)evcut, (real)evoff, 1, (real)ghal, (real)dhal, 5, (real)0.7, e, de);
      e_out[q] = e, de_out[q] = de;
   }
   return 0;
}

// Real-space fields alone (the two pair sweeps of the induced-dipole solver): mode 0 = permanent field of the multipoles with
// the d / p scalings (dfieldEwaldReal), mode 1 = mutual field of the dipoles ud / up with the u scaling (ufieldEwaldReal).
// Same two-pass treatment of scaled pairs as above.  fd, fp are zeroed here.
extern "C" int ref_field_real(int mode, int n, long long npair, const int* pi, const int* pk, const double* R, const double* scale,
   const double* rpole, const double* pdamp, const double* pga, const double* ud, const double* up, double aewald, int ewald, double* fd, double* fp)
{
   std::memset(fd, 0, sizeof(double) * 3 * (size_t)n);
   std::memset(fp, 0, sizeof(double) * 3 * (size_t)n);
   for (long long p = 0; p < npair; ++p) {
      const int i = pi[p], k = pk[p];
      const real xr = R[3 * p], yr = R[3 * p + 1], zr = R[3 * p + 2];
      const real r2 = xr * xr + yr * yr + zr * zr;
      const double* s = scale + 4 * p;
      const real pdi = pdamp[i], pdk = pdamp[k], pg = pga[p];
      for (int pass = 0; pass < 2; ++pass) {
         real ds, ps, us, aw;
         bool ew;
         if (ewald && pass == 0)
            ds = ps = us = 1, aw = aewald, ew = true;
         else if (ewald)
            ds = s[1] - 1, ps = s[2] - 1, us = s[3] - 1, aw = 0, ew = false;
         else if (pass == 0)
            ds = s[1], ps = s[2], us = s[3], aw = 0, ew = false;
         else
            break;
         if (pass == 1 && (mode == 0 ? (ds == 0 && ps == 0) : us == 0))
            continue;
         real3 fid = make_real3(0, 0, 0), fip = make_real3(0, 0, 0), fkd = make_real3(0, 0, 0), fkp = make_real3(0, 0, 0);
         if (mode == 0) {
            const Atom A = load(rpole, i), B = load(rpole, k);
            if (ew)
               pair_dfield<EWALD>(r2, xr, yr, zr, ds, ps, MP(A), pdi, pg, MP(B), pdk, pg, aw, fid, fip, fkd, fkp);
            else
               pair_dfield<NON_EWALD>(r2, xr, yr, zr, ds, ps, MP(A), pdi, pg, MP(B), pdk, pg, aw, fid, fip, fkd, fkp);
         } else {
            const double *a = ud + 3 * i, *b = up + 3 * i, *c = ud + 3 * k, *d = up + 3 * k;
            if (ew)
               pair_ufield<EWALD>(r2, xr, yr, zr, us, (real)a[0], (real)a[1], (real)a[2], (real)b[0], (real)b[1], (real)b[2], pdi, pg, (real)c[0], (real)c[1],
                  (real)c[2], (real)d[0], (real)d[1], (real)d[2], pdk, pg, aw, fid, fip, fkd, fkp);
            else
               pair_ufield<NON_EWALD>(r2, xr, yr, zr, us, (real)a[0], (real)a[1], (real)a[2], (real)b[0], (real)b[1], (real)b[2], pdi, pg, (real)c[0],
                  (real)c[1], (real)c[2], (real)d[0], (real)d[1], (real)d[2], pdk, pg, aw, fid, fip, fkd, fkp);
         }
         fd[3 * i] += fid.x, fd[3 * i + 1] += fid.y, fd[3 * i + 2] += fid.z;
         fp[3 * i] += fip.x, fp[3 * i + 1] += fip.y, fp[3 * i + 2] += fip.z;
         fd[3 * k] += fkd.x, fd[3 * k + 1] += fkd.y, fd[3 * k + 2] += fkd.z;
         fp[3 * k] += fkp.x, fp[3 * k + 1] += fkp.y, fp[3 * k + 2] += fkp.z;
      }
   }
   return 0;
}

// Buffered 14-7 energy, force on the reduced sites and virial over a given pair list (pair_hal_v2, as ehal_acc does pair by
// pair): d[p] = x_i - x_k after the minimum-image shift, rv / eps the pair parameters (eps already times the pair's scale).
extern "C" int ref_hal_pairs(int n, long long npair, const int* pi, const int* pk, const double* d, const double* rv, const double* eps, double evcut,
   double evoff, double ghal, double dhal, double* ev, double* gred, double* vir9)
{
   *ev = 0;
   std::memset(gred, 0, sizeof(double) * 3 * (size_t)n);
   double v[9] = {0};
   for (long long p = 0; p < npair; ++p) {
      const double x = d[3 * p], y = d[3 * p + 1], z = d[3 * p + 2];
      const real r = std::sqrt(x * x + y * y + z * z);
      real e, de;
      pair_hal_v2<true, 1>(r, 1, (real)rv[p], (real)eps[p], (real)evcut, (real)evoff, 1, (real)ghal, (real)dhal, 5, (real)0.7, e, de);
      *ev += e;
      const double f = de / r, fx = f * x, fy = f * y, fz = f * z;
      const int i = pi[p], k = pk[p];
      gred[3 * i] += fx, gred[3 * i + 1] += fy, gred[3 * i + 2] += fz;
      gred[3 * k] -= fx, gred[3 * k + 1] -= fy, gred[3 * k + 2] -= fz;
      v[0] += x * fx, v[1] += x * fy, v[2] += x * fz, v[3] += y * fx, v[4] += y * fy, v[5] += y * fz, v[6] += z * fx, v[7] += z * fy, v[8] += z * fz;
   }
   std::memcpy(vir9, v, sizeof(v));
   return 0;
}
