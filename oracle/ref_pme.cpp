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
#include "ff/modhippo.h"
#include "ff/nblist.h"
#include "ff/pme.h"
#include "ff/switch.h"
#include "math/parallelacc.h"
#include "tool/gpucard.h"
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

int gpuGridSize(int) { return 1; }
NBList::~NBList() {}
static real g_cutoff = 0;
real switchOff(Switch) { return g_cutoff; }
real switchCut(Switch) { return g_cutoff; }
template <class T>
void scaleArray_acc(T* dst, T scal, size_t nelem, int)
{
   for (size_t i = 0; i < nelem; ++i)
      dst[i] *= scal;
}
template void scaleArray_acc<double>(double*, double, size_t, int);

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
// ---- src/acc/hippo/empole.cpp, src/acc/amoeba/epolarewald.cpp: reciprocal energy / force / torque / virial assembly
void empoleChgpenEwaldRecip_acc(int vers, int use_cf);
void epolarEwaldRecipSelf_acc(int vers, const real (*uind)[3], const real (*uinp)[3]);
// ---- src/acc/amoeba/rotpole.cpp, torque.cpp (compiled unmodified into this library as well)
void chkpole_acc();
void rotpole_acc();
void torque_acc(int vers, grad_prec* gx, grad_prec* gy, grad_prec* gz);
}

// ---- front-end dispatchers of src/pme.cpp (TINKER_FCALL wrappers in the reference): forwarded to the *_acc functions; the FFT
//      (src/host/fft.cpp uses FFTW) is a callback into the caller, which transforms the grid in place
extern "C" typedef void (*ref_fft_fn)(int forward);
static ref_fft_fn g_fft = nullptr;
static tinker::real* g_cur = nullptr;      // the grid ref_pme_qgrid_get / _set address: that of the unit being transformed
namespace tinker {
void fftfront(PMEUnit pu) { real* keep = g_cur; g_cur = pu->qgrid, g_fft(1), g_cur = keep; }
void fftback(PMEUnit pu) { real* keep = g_cur; g_cur = pu->qgrid, g_fft(0), g_cur = keep; }
void pmeConv(PMEUnit pu) { pmeConv_acc(pu, nullptr, nullptr); }
void pmeConv(PMEUnit pu, VirialBuffer v) { pmeConv_acc(pu, nullptr, v); }
void pmeConv(PMEUnit pu, EnergyBuffer e) { pmeConv_acc(pu, e, nullptr); }
void pmeConv(PMEUnit pu, EnergyBuffer e, VirialBuffer v) { pmeConv_acc(pu, e, v); }
void cmpToFmp(PMEUnit pu, const real (*c)[10], real (*f)[10]) { cmpToFmp_acc(pu, c, f); }
void cuindToFuind(PMEUnit pu, const real (*a)[3], const real (*b)[3], real (*c)[3], real (*d)[3]) { cuindToFuind_acc(pu, a, b, c, d); }
void fphiToCphi(PMEUnit pu, const real (*f)[20], real (*c)[10]) { fphiToCphi_acc(pu, f, c); }
void gridMpole(PMEUnit pu, real (*f)[10]) { gridMpole_acc(pu, f); }
void gridUind(PMEUnit pu, real (*a)[3], real (*b)[3]) { gridUind_acc(pu, a, b); }
void fphiMpole(PMEUnit pu) { fphiMpole_acc(pu, fphi); }
void fphiUind(PMEUnit pu, real (*a)[10], real (*b)[10], real (*c)[20]) { fphiUind_acc(pu, a, b, c); }
void fphiUind2(PMEUnit pu, real (*a)[10], real (*b)[10]) { fphiUind2_acc(pu, a, b); }
}

using namespace tinker;

namespace {
PMEUnit g_unit, g_unit2;      // g_unit2: the second grid of the polarization virial (pvpme_unit, epolarewald.cpp:593)
std::vector<real> g_x, g_y, g_z, g_qgrid, g_qgrid2, g_b1, g_b2, g_b3, g_rpole, g_cmp;
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
   g_unit.deviceptrUpdate(st, 0);
   g_unit2 = PMEUnit::open();
   PME& s2 = *g_unit2;
   g_qgrid2.assign(2 * g_k, 0);
   s2 = st, s2.qgrid = g_qgrid2.data();
   g_unit2.deviceptrUpdate(s2, 0);
   g_cur = g_qgrid.data();
   g_rpole.assign(10 * (size_t)n, 0), g_cmp.assign(10 * (size_t)n, 0);
   rpole = reinterpret_cast<real(*)[MPL_TOTAL]>(g_rpole.data());
   cmp = reinterpret_cast<real(*)[10]>(g_cmp.data());
   return 0;
}

long long ref_pme_grid_size(void) { return (long long)g_k; }
void ref_pme_qgrid_get(double* out) { std::memcpy(out, g_cur, sizeof(double) * 2 * g_k); }      // [n3][n2][n1][re,im]
void ref_pme_qgrid_set(const double* in) { std::memcpy(g_cur, in, sizeof(double) * 2 * g_k); }

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

// ---- reciprocal-space energy / gradient / torque / virial of the permanent multipoles (empoleChgpenEwaldRecip_acc, AMOEBA branch:
//      use_cf = 0) and of the induced dipoles incl. the self term (epolarEwaldRecipSelf_acc), calc::v1.  ref_pme_open must have been
//      called; fft(forward) transforms the grid (ref_pme_qgrid_get / _set) in place, unnormalised in both directions.  The polar
//      call reuses cmp / fmp / cphi / fphi left by the multipole call, as the reference does (epolarrecip.cu:419-420).
extern "C" {
static std::vector<real> r_fmp, r_fphi, r_cphi, r_fuind, r_fuinp, r_fd1, r_fd2, r_cphidp, r_fphidp;
static v_prec r_vm[1][8];      // vir_m: the convolution virial of the multipole call, subtracted again by the polar call (epolarewald.cpp:505-510)

int ref_recip_mpole(ref_fft_fn fft, const double* rpole_in, double* e, double* grad, double* trq, double* vir6)
{
   g_fft = fft;
   const size_t N = (size_t)n;
   std::memcpy(g_rpole.data(), rpole_in, sizeof(double) * 10 * N);
   rpoleToCmp_acc();
   r_fmp.assign(10 * N, 0), r_fphi.assign(20 * N, 0), r_cphi.assign(10 * N, 0);
   fmp = reinterpret_cast<real(*)[10]>(r_fmp.data()), fphi = reinterpret_cast<real(*)[20]>(r_fphi.data());
   cphi = reinterpret_cast<real(*)[10]>(r_cphi.data());
   std::vector<real> gx(N, 0), gy(N, 0), gz(N, 0), tx(N, 0), ty(N, 0), tz(N, 0);
   demx = gx.data(), demy = gy.data(), demz = gz.data(), trqx = tx.data(), trqy = ty.data(), trqz = tz.data();
   e_prec eb[1] = {0};
   v_prec vb[1][8] = {{0}};
   std::memset(r_vm, 0, sizeof r_vm);
   em = eb, vir_em = vb, vir_m = r_vm, pot = nullptr;
   epme_unit = g_unit;
   empoleChgpenEwaldRecip_acc(calc::v1, 0);
   *e = eb[0];
   for (size_t i = 0; i < N; ++i) {
      grad[3 * i] = gx[i], grad[3 * i + 1] = gy[i], grad[3 * i + 2] = gz[i];
      trq[3 * i] = tx[i], trq[3 * i + 1] = ty[i], trq[3 * i + 2] = tz[i];
   }
   for (int q = 0; q < 6; ++q)
      vir6[q] = vb[0][q];
   demx = demy = demz = trqx = trqy = trqz = nullptr, em = nullptr, vir_em = nullptr, vir_m = nullptr;
   return 0;
}

int ref_recip_polar(ref_fft_fn fft, const double* ud, const double* up, double* e, double* grad, double* trq, double* vir6)
{
   g_fft = fft;
   const size_t N = (size_t)n;
   if (r_fmp.size() != 10 * N)
      return 2;      // ref_recip_mpole first
   r_fuind.assign(3 * N, 0), r_fuinp.assign(3 * N, 0), r_fd1.assign(10 * N, 0), r_fd2.assign(10 * N, 0);
   r_cphidp.assign(10 * N, 0), r_fphidp.assign(20 * N, 0);
   fuind = reinterpret_cast<real(*)[3]>(r_fuind.data()), fuinp = reinterpret_cast<real(*)[3]>(r_fuinp.data());
   fdip_phi1 = reinterpret_cast<real(*)[10]>(r_fd1.data()), fdip_phi2 = reinterpret_cast<real(*)[10]>(r_fd2.data());
   cphidp = reinterpret_cast<real(*)[10]>(r_cphidp.data()), fphidp = reinterpret_cast<real(*)[20]>(r_fphidp.data());
   std::vector<real> gx(N, 0), gy(N, 0), gz(N, 0), tx(N, 0), ty(N, 0), tz(N, 0);
   depx = gx.data(), depy = gy.data(), depz = gz.data(), trqx = tx.data(), trqy = ty.data(), trqz = tz.data();
   e_prec eb[1] = {0};
   v_prec vb[1][8] = {{0}};
   ep = eb, vir_ep = vb, vir_m = r_vm;
   ppme_unit = g_unit, pvpme_unit = g_unit2, epme_unit = g_unit;
   epolarEwaldRecipSelf_acc(calc::v1, reinterpret_cast<const real(*)[3]>(ud), reinterpret_cast<const real(*)[3]>(up));
   *e = eb[0];
   for (size_t i = 0; i < N; ++i) {
      grad[3 * i] = gx[i], grad[3 * i + 1] = gy[i], grad[3 * i + 2] = gz[i];
      trq[3 * i] = tx[i], trq[3 * i + 1] = ty[i], trq[3 * i + 2] = tz[i];
   }
   for (int q = 0; q < 6; ++q)
      vir6[q] = vb[0][q];
   depx = depy = depz = trqx = trqy = trqz = nullptr, ep = nullptr, vir_ep = nullptr, vir_m = nullptr;
   return 0;
}
}
