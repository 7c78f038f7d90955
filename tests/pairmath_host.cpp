// CPU harness for the real-space pair math of the CUDA library -- TEST INFRASTRUCTURE.  Compiles the very header the kernels
// use (csrc/pairmath.cuh -> pairmath_body.inc, in float and in double) with g++ and walks a pair list in a plain loop, so the
// ERROR BUDGET of the mixed build can be measured against the float64 oracle without a GPU (tests/test_pairmath_host.py):
// which part of the force error comes from the coordinates the pair separation is formed from, which from evaluating the
// bonded-range (listed) pairs as "all scales 1" + "(scale - 1) correction" in float, and what float pair math costs by itself.
// The device intrinsics of the mixed build (__expf, rsqrtf, __fdividef) are replaced by their libm counterparts here, so the
// numbers are a lower bound of the GPU's by a fraction of an ulp per call.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#define __expf(x) expf(x)      // glibc declares __expf but does not export it
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v)
{
   unsigned long long o = *p;
   *p += v;
   return o;
}
static inline float atomicAdd(float* p, float v)
{
   float o = *p;
   *p += v;
   return o;
}
static inline double atomicAdd(double* p, double v)
{
   double o = *p;
   *p += v;
   return o;
}

#include "pairmath.cuh"

#include "pairmath_host_eval.inc"
namespace pm64 {
#include "pairmath_host_eval.inc"
}

namespace {      // (BoxD: apx_internal.h)
void invert3(const double* m, double* inv)
{
   double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
   double id = 1.0 / det;
   inv[0] = (m[4] * m[8] - m[5] * m[7]) * id, inv[1] = (m[2] * m[7] - m[1] * m[8]) * id, inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
   inv[3] = (m[5] * m[6] - m[3] * m[8]) * id, inv[4] = (m[0] * m[8] - m[2] * m[6]) * id, inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
   inv[6] = (m[3] * m[7] - m[4] * m[6]) * id, inv[7] = (m[1] * m[6] - m[0] * m[7]) * id, inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}
// fractional coordinates in [0,1), as wrap_pos (csrc/wrap.cuh) forms them
void frac_of(const BoxD& b, const double* x, double* f)
{
   for (int a = 0; a < 3; ++a) {
      double v = x[0] * b.r[3 * a] + x[1] * b.r[3 * a + 1] + x[2] * b.r[3 * a + 2];
      v -= floor(v);
      if (v >= 1.0)
         v = 0.0;
      f[a] = v;
   }
}
}

// mode bits: 1 = float pair math (else double); 2 = separation from 32-bit fractional coordinates (else from the wrapped
// Cartesian coordinates rounded to the math type, what posd holds); 4 = listed pairs evaluated once, with their true
// scales, in double (else: all-ones pass + (scale-1) correction, both in the math type).
// pairs: i < k, scale4 = (m, d, p, u); listed = any scale != 1.  Output: real-space gradient and torque [n][3] (f64).
extern "C" int pairmath_host_mplar(int mode, int n, long long npair, const int* pi, const int* pk, const double* scale4, const double* xyz,
   const double* lvec9, const double* rpole, const double* ud, const double* up, const double* pdamp, const double* thole, double aewald,
   int ewald, int mutual, double felec, double* grad, double* trq)
{
   const bool f32 = mode & 1, u32 = mode & 2, listed64 = mode & 4;
   BoxD b;
   memcpy(b.l, lvec9, sizeof(b.l));
   invert3(lvec9, b.r);
   // what the device holds per atom
   std::vector<double> fr(3 * (size_t)n);
   std::vector<float> wf(3 * (size_t)n);         // wrapped Cartesian, float (posd)
   std::vector<double> wd(3 * (size_t)n);        // wrapped Cartesian, double
   std::vector<uint32_t> qf(3 * (size_t)n);      // 32-bit fractional
   for (int i = 0; i < n; ++i) {
      double* f = &fr[3 * (size_t)i];
      frac_of(b, xyz + 3 * (size_t)i, f);
      for (int a = 0; a < 3; ++a) {
         double w = f[0] * b.l[3 * a] + f[1] * b.l[3 * a + 1] + f[2] * b.l[3 * a + 2];
         wd[3 * (size_t)i + a] = w;
         wf[3 * (size_t)i + a] = (float)w;
         qf[3 * (size_t)i + a] = (uint32_t)(unsigned long long)(f[a] * 4294967296.0);
      }
   }
   // float accumulators: 16 partial sums per atom (the 16 lanes of a row group), reduced in float at the end
   std::vector<float> ga(48 * (size_t)n, 0.0f), ta(48 * (size_t)n, 0.0f);
   std::vector<int> slot(n, 0);
   std::vector<double> gd(3 * (size_t)n, 0.0), td(3 * (size_t)n, 0.0);
   float lf[9], qs[9], qlo[9];
   for (int a = 0; a < 9; ++a) {
      lf[a] = (float)b.l[a];
      qs[a] = (float)(b.l[a] / 4294967296.0);
      qlo[a] = (float)(b.l[a] / 4294967296.0 - (double)qs[a]);      // Box::qlo (csrc/apx_api.cu: set_box)
   }
   float rf[9];
   for (int a = 0; a < 9; ++a)
      rf[a] = (float)b.r[a];
   for (long long p = 0; p < npair; ++p) {
      const int i = pi[p], k = pk[p];
      const double* sc = scale4 + 4 * p;
      const bool listed = sc[0] != 1.0 || sc[1] != 1.0 || sc[2] != 1.0 || sc[3] != 1.0;
      // ---- separation in double (oracle quality)
      double Rd[3];
      {
         double df[3];
         for (int a = 0; a < 3; ++a) {
            df[a] = fr[3 * (size_t)k + a] - fr[3 * (size_t)i + a];
            df[a] -= rint(df[a]);
         }
         for (int a = 0; a < 3; ++a)
            Rd[a] = df[0] * b.l[3 * a] + df[1] * b.l[3 * a + 1] + df[2] * b.l[3 * a + 2];
      }
      // ---- separation as the float kernels see it
      float Rf[3];
      if (u32) {
         float df[3];
         for (int a = 0; a < 3; ++a)
            df[a] = (float)(int32_t)(qf[3 * (size_t)k + a] - qf[3 * (size_t)i + a]);
         for (int a = 0; a < 3; ++a)
            Rf[a] = df[0] * qs[3 * a] + df[1] * qs[3 * a + 1] + df[2] * qs[3 * a + 2]
               + (df[0] * qlo[3 * a] + df[1] * qlo[3 * a + 1] + df[2] * qlo[3 * a + 2]);      // pair_delta (csrc/pairmath.cuh)
      } else {
         float d[3], f[3];
         for (int a = 0; a < 3; ++a)
            d[a] = wf[3 * (size_t)k + a] - wf[3 * (size_t)i + a];
         for (int a = 0; a < 3; ++a) {
            f[a] = d[0] * rf[3 * a] + d[1] * rf[3 * a + 1] + d[2] * rf[3 * a + 2];
            f[a] -= rintf(f[a]);
         }
         for (int a = 0; a < 3; ++a)
            Rf[a] = f[0] * lf[3 * a] + f[1] * lf[3 * a + 1] + f[2] * lf[3 * a + 2];
      }
      const double one4[4] = {1, 1, 1, 1};
      auto add64 = [&](int what, const double* s4, const double* Rx) {
         pm64::LabSite I = pm64::lab_site(rpole + 10 * (size_t)i, ud + 3 * (size_t)i, up + 3 * (size_t)i, pdamp[i], thole[i]);
         pm64::LabSite K = pm64::lab_site(rpole + 10 * (size_t)k, ud + 3 * (size_t)k, up + 3 * (size_t)k, pdamp[k], thole[k]);
         pm64::V3 g, ti, tk;
         pm64::lab_eval(what, pm64::v3(Rx[0], Rx[1], Rx[2]), I, K, s4, aewald, ewald != 0, mutual != 0, g, ti, tk);
         const double gg[3] = {g.x, g.y, g.z}, a[3] = {ti.x, ti.y, ti.z}, c[3] = {tk.x, tk.y, tk.z};
         for (int q = 0; q < 3; ++q) {
            gd[3 * (size_t)i + q] -= felec * gg[q], gd[3 * (size_t)k + q] += felec * gg[q];
            td[3 * (size_t)i + q] += felec * a[q], td[3 * (size_t)k + q] += felec * c[q];
         }
      };
      auto add32 = [&](int what, const double* s4, bool rows) {
         // the float kernels read multipoles / dipoles rounded to float
         LabSite I = lab_site(rpole + 10 * (size_t)i, ud + 3 * (size_t)i, up + 3 * (size_t)i, pdamp[i], thole[i]);
         LabSite K = lab_site(rpole + 10 * (size_t)k, ud + 3 * (size_t)k, up + 3 * (size_t)k, pdamp[k], thole[k]);
         V3 g, ti, tk;
         lab_eval(what, v3(Rf[0], Rf[1], Rf[2]), I, K, s4, (float)aewald, ewald != 0, mutual != 0, g, ti, tk);
         const float fe = (float)felec;
         const float gg[3] = {g.x, g.y, g.z}, a[3] = {ti.x, ti.y, ti.z}, c[3] = {tk.x, tk.y, tk.z};
         if (rows) {      // row pass: float partial sums per lane, the factor f applied once per atom at the end
            const int si = (slot[i]++) & 15, sk = (slot[k]++) & 15;
            for (int q = 0; q < 3; ++q) {
               ga[48 * (size_t)i + 3 * si + q] -= gg[q], ga[48 * (size_t)k + 3 * sk + q] += gg[q];
               ta[48 * (size_t)i + 3 * si + q] += a[q], ta[48 * (size_t)k + 3 * sk + q] += c[q];
            }
         } else {         // exclusion pass: fixed-point atomics per pair (exact sums of float values)
            for (int q = 0; q < 3; ++q) {
               gd[3 * (size_t)i + q] -= (double)(fe * gg[q]), gd[3 * (size_t)k + q] += (double)(fe * gg[q]);
               td[3 * (size_t)i + q] += (double)(fe * a[q]), td[3 * (size_t)k + q] += (double)(fe * c[q]);
            }
         }
      };
      if (!f32) {
         add64(LAB_TRUE, sc, Rd);
      } else if (listed && listed64) {
         add64(LAB_TRUE, sc, Rd);
      } else {
         add32(LAB_ONES, one4, true);
         if (listed)
            add32(LAB_CORRECTION, sc, false);
      }
   }
   const float fe = (float)felec;
   for (int i = 0; i < n; ++i)
      for (int q = 0; q < 3; ++q) {
         float g = 0, t = 0;
         for (int s = 0; s < 16; ++s)
            g += ga[48 * (size_t)i + 3 * s + q], t += ta[48 * (size_t)i + 3 * s + q];
         grad[3 * (size_t)i + q] = gd[3 * (size_t)i + q] + (double)(fe * g);
         trq[3 * (size_t)i + q] = td[3 * (size_t)i + q] + (double)(fe * t);
      }
   return 0;
}
