// oracle/_ref/libref_pme.so -- the REFERENCE'S OWN PME operators, executed on the CPU.  TEST INFRASTRUCTURE ONLY.
//
// oracle/Makefile compiles the reference's host PME translation unit src/acc/pme.cpp unmodified, where it lies under
// /root/reference (-DTINKER_DOUBLE_PRECISION; g++ ignores its OpenACC pragmas, so the loops run serially), and links it with
// this file, which is our code and does two things:
//   1. shim: defines the handful of process globals and runtime hooks that TU reads (n, x, y, z, the box vectors, rpole,
//      cmp, electric, dielec, g::q0, bufferSize, boxVolume, the darray allocation hooks, ~PME) -- 17 symbols, not the
//      Fortran module set the reference's front-ends need;
//   2. driver: a C ABI over gridMpole / gridUind / pmeConv / fphiMpole / fphiUind / fphiUind2 / rpoleToCmp / cmpToFmp /
//      cuindToFuind / fphiToCphi (src/acc/pme.cpp:168-190, 305-320, 700-730, 735-930) with the grid handed in and out, so
//      that the FFT between them (the reference uses FFTW, src/host/fft.cpp -- third-party arithmetic, any correct FFT is
//      interchangeable) is done by the caller.
// The B-spline moduli (bsmod1..3, computed by Fortran dftmod in the reference, src/pme.cpp:96-111) are passed in.
#define TINKER_EXTERN_DEFINITION_FILE 1
#include "ff/atom.h"
#include "ff/box.h"
#include "ff/elec.h"
#include "ff/energybuffer.h"
#include "ff/hippo/erepel.h"
#include "ff/modamoeba.h"
#include "ff/pme.h"
#include "tool/accasync.h"
#include "tool/darray.h"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace tinker {
// ---- shim: runtime hooks of the host build
size_t bufferSize() { return 1; }      // one accumulator slot: offset & (bufsize - 1) == 0
real boxVolume()
{
   return lvec1.x * (lvec2.y * lvec3.z - lvec2.z * lvec3.y) - lvec1.y * (lvec2.x * lvec3.z - lvec2.z * lvec3.x)
      + lvec1.z * (lvec2.x * lvec3.y - lvec2.y * lvec3.x);
}
void deviceMemoryZeroBytesAsync(void* dst, size_t nbytes, int) { std::memset(dst, 0, nbytes); }
void deviceMemoryAllocateBytes(void** pptr, size_t nbytes) { *pptr = std::malloc(nbytes); }
void deviceMemoryDeallocate(void* ptr) { std::free(ptr); }
void deviceMemoryCopyinBytesAsync(void* dst, const void* src, size_t nbytes, int) { std::memcpy(dst, src, nbytes); }
PME::~PME() {}      // the buffers belong to the vectors below

// ---- the operators of src/acc/pme.cpp
void pmeConv_acc(PMEUnit, EnergyBuffer, VirialBuffer);
void cmpToFmp_acc(PMEUnit, const real (*)[10], real (*)[10]);
void cuindToFuind_acc(PMEUnit, const real (*)[3], const real (*)[3], real (*)[3], real (*)[3]);
void fphiToCphi_acc(PMEUnit, const real (*)[20], real (*)[10]);
void gridMpole_acc(PMEUnit, real (*)[10]);
void gridUind_acc(PMEUnit, real (*)[3], real (*)[3]);
void fphiMpole_acc(PMEUnit, real (*)[20]);
void fphiUind_acc(PMEUnit, real (*)[10], real (*)[10], real (*)[20]);
void fphiUind2_acc(PMEUnit, real (*)[10], real (*)[10]);
void rpoleToCmp_acc();
// ---- src/acc/amoeba/rotpole.cpp, torque.cpp (compiled unmodified into this library as well)
void chkpole_acc();
void rotpole_acc();
void torque_acc(int vers, grad_prec* gx, grad_prec* gy, grad_prec* gz);
}

using namespace tinker;

namespace {
PMEUnit g_unit;
std::vector<real> g_x, g_y, g_z, g_qgrid, g_b1, g_b2, g_b3, g_rpole, g_cmp;
size_t g_k = 0;
}

extern "C" {
int ref_pme_open(int natoms, const double* xyz, const double* lvec9, const double* recip9, const int* nfft, int bsorder, double aewald,
   const double* bsmod1, const double* bsmod2, const double* bsmod3, double electric_, double dielec_)
{
   n = natoms;
   g_x.resize(n), g_y.resize(n), g_z.resize(n);
   for (int i = 0; i < n; ++i)
      g_x[i] = xyz[3 * i], g_y[i] = xyz[3 * i + 1], g_z[i] = xyz[3 * i + 2];
   x = g_x.data(), y = g_y.data(), z = g_z.data();
   lvec1 = make_real3(lvec9[0], lvec9[1], lvec9[2]), lvec2 = make_real3(lvec9[3], lvec9[4], lvec9[5]), lvec3 = make_real3(lvec9[6], lvec9[7], lvec9[8]);
   recipa = make_real3(recip9[0], recip9[1], recip9[2]), recipb = make_real3(recip9[3], recip9[4], recip9[5]);
   recipc = make_real3(recip9[6], recip9[7], recip9[8]);
   const double off = std::fabs(lvec9[1]) + std::fabs(lvec9[2]) + std::fabs(lvec9[3]) + std::fabs(lvec9[5]) + std::fabs(lvec9[6]) + std::fabs(lvec9[7]);
   box_shape = off < 1e-12 ? BoxShape::ORTHO : BoxShape::TRI;
   electric = electric_, dielec = dielec_;
   g::q0 = 0;
   g_unit = PMEUnit::open();
   PME& st = *g_unit;
   st.aewald = aewald, st.nfft1 = nfft[0], st.nfft2 = nfft[1], st.nfft3 = nfft[2], st.bsorder = bsorder;
   g_k = (size_t)nfft[0] * nfft[1] * nfft[2];
   g_qgrid.assign(2 * g_k, 0);
   g_b1.assign(bsmod1, bsmod1 + nfft[0]), g_b2.assign(bsmod2, bsmod2 + nfft[1]), g_b3.assign(bsmod3, bsmod3 + nfft[2]);
   st.qgrid = g_qgrid.data(), st.bsmod1 = g_b1.data(), st.bsmod2 = g_b2.data(), st.bsmod3 = g_b3.data();
   st.igrid = nullptr, st.thetai1 = st.thetai2 = st.thetai3 = nullptr;
   g_rpole.assign(10 * (size_t)n, 0), g_cmp.assign(10 * (size_t)n, 0);
   rpole = reinterpret_cast<real(*)[MPL_TOTAL]>(g_rpole.data());
   cmp = reinterpret_cast<real(*)[10]>(g_cmp.data());
   return 0;
}

long long ref_pme_grid_size(void) { return (long long)g_k; }
void ref_pme_qgrid_get(double* out) { std::memcpy(out, g_qgrid.data(), sizeof(double) * 2 * g_k); }      // [n3][n2][n1][re,im]
void ref_pme_qgrid_set(const double* in) { std::memcpy(g_qgrid.data(), in, sizeof(double) * 2 * g_k); }

void ref_pme_rpole_to_cmp(const double* rp, double* out)
{
   std::memcpy(g_rpole.data(), rp, sizeof(double) * 10 * (size_t)n);
   rpoleToCmp_acc();
   std::memcpy(out, g_cmp.data(), sizeof(double) * 10 * (size_t)n);
}
void ref_pme_cmp_to_fmp(const double* c, double* f) { cmpToFmp_acc(g_unit, reinterpret_cast<const real(*)[10]>(c), reinterpret_cast<real(*)[10]>(f)); }
void ref_pme_cuind_to_fuind(const double* ud, const double* up, double* fud, double* fup)
{
   cuindToFuind_acc(g_unit, reinterpret_cast<const real(*)[3]>(ud), reinterpret_cast<const real(*)[3]>(up), reinterpret_cast<real(*)[3]>(fud),
      reinterpret_cast<real(*)[3]>(fup));
}
void ref_pme_fphi_to_cphi(const double* f, double* c) { fphiToCphi_acc(g_unit, reinterpret_cast<const real(*)[20]>(f), reinterpret_cast<real(*)[10]>(c)); }
void ref_pme_grid_mpole(double* fmp) { gridMpole_acc(g_unit, reinterpret_cast<real(*)[10]>(fmp)); }
void ref_pme_grid_uind(double* fud, double* fup) { gridUind_acc(g_unit, reinterpret_cast<real(*)[3]>(fud), reinterpret_cast<real(*)[3]>(fup)); }
// the grid must hold the forward transform; on return it holds the product with the influence function.  e: reciprocal energy
// (0.5 f sum expterm |Q|^2), vir6: {xx, yx, zx, yy, zy, zz}
void ref_pme_conv(double* e, double* vir6)
{
   e_prec eb[1] = {0};
   v_prec vb[1][8] = {{0}};
   pmeConv_acc(g_unit, e ? eb : nullptr, vir6 ? vb : nullptr);
   if (e)
      *e = eb[0];
   if (vir6)
      for (int q = 0; q < 6; ++q)
         vir6[q] = vb[0][q];
}
void ref_pme_fphi_mpole(double* fphi) { fphiMpole_acc(g_unit, reinterpret_cast<real(*)[20]>(fphi)); }
void ref_pme_fphi_uind(double* f1, double* f2, double* fdp)
{
   fphiUind_acc(g_unit, reinterpret_cast<real(*)[10]>(f1), reinterpret_cast<real(*)[10]>(f2), reinterpret_cast<real(*)[20]>(fdp));
}
void ref_pme_fphi_uind2(double* f1, double* f2) { fphiUind2_acc(g_unit, reinterpret_cast<real(*)[10]>(f1), reinterpret_cast<real(*)[10]>(f2)); }
}

// ---- local frames: chkpole + rotpole (src/acc/amoeba/rotpole.cpp over include/seq/rotpole.h) and torque -> force on the frame
//      atoms with the torque virial (src/acc/amoeba/torque.cpp:20-395).  zaxis[i] = {z, x, y (signed, from ONE), polaxe}.
extern "C" {
static void frames_bind(int natoms, const double* xyz, const int* zax, std::vector<real>& px, std::vector<real>& py, std::vector<real>& pz)
{
   n = natoms;
   px.resize(n), py.resize(n), pz.resize(n);
   for (int i = 0; i < n; ++i)
      px[i] = xyz[3 * i], py[i] = xyz[3 * i + 1], pz[i] = xyz[3 * i + 2];
   x = px.data(), y = py.data(), z = pz.data();
   zaxis = reinterpret_cast<LocalFrame*>(const_cast<int*>(zax));
}

int ref_frames_rotpole(int natoms, const double* xyz, const int* zax, const double* pole_in, double* pole_chk, double* rpole_out)
{
   std::vector<real> px, py, pz, p(pole_in, pole_in + 10 * (size_t)natoms), rp(10 * (size_t)natoms, 0);
   frames_bind(natoms, xyz, zax, px, py, pz);
   pole = reinterpret_cast<real(*)[MPL_TOTAL]>(p.data());
   rpole = reinterpret_cast<real(*)[MPL_TOTAL]>(rp.data());
   chkpole_acc();
   rotpole_acc();
   std::memcpy(pole_chk, p.data(), sizeof(double) * p.size());
   std::memcpy(rpole_out, rp.data(), sizeof(double) * rp.size());
   pole = nullptr, rpole = nullptr;
   return 0;
}

int ref_frames_torque(int natoms, const double* xyz, const int* zax, const double* trq, double* grad, double* vir6)
{
   std::vector<real> px, py, pz, tx(natoms), ty(natoms), tz(natoms), gx(natoms, 0), gy(natoms, 0), gz(natoms, 0);
   frames_bind(natoms, xyz, zax, px, py, pz);
   for (int i = 0; i < natoms; ++i)
      tx[i] = trq[3 * i], ty[i] = trq[3 * i + 1], tz[i] = trq[3 * i + 2];
   trqx = tx.data(), trqy = ty.data(), trqz = tz.data();
   v_prec vb[1][8] = {{0}};
   vir_trq = vb;
   torque_acc(calc::grad | calc::virial, gx.data(), gy.data(), gz.data());
   for (int i = 0; i < natoms; ++i)
      grad[3 * i] = gx[i], grad[3 * i + 1] = gy[i], grad[3 * i + 2] = gz[i];
   for (int q = 0; q < 6; ++q)
      vir6[q] = vb[0][q];
   trqx = trqy = trqz = nullptr, vir_trq = nullptr;
   return 0;
}
}
