// oracle/_ref/libref_cuda.so -- the REFERENCE'S OWN CUDA implementation of the path, run on the GPU beside ours.
// COMPARATOR / TEST INFRASTRUCTURE ONLY: loaded by bench.py's comparator leg and by tests/, never by the product path.
//
// oracle/Makefile compiles, unmodified and where they lie under /root/reference (nothing is copied), the reference's CUDA
// translation units of the AMOEBA electrostatics path -- src/cu/amoeba/{pcg,field,precond,emplar,rotpole,torque,binding}.cu,
// src/cu/{pme,epolarrecip,spatial,induce,diagprecond,upredict,mathparallel,mathzero,cumod,cudalib}.cu, src/cu/hippo/empole.cu --
// its CUDA runtime layer src/cudart/{darray,error,fft,thrustcache}.cpp, src/cudalib.cpp, and the Fortran-free front-ends
// src/amoeba/{field,induce}.cpp, src/spatial.cpp, src/energybuffer.cpp, src/mod.cpp (all process globals), with the reference's
// own release flags (nvcc -O3 --use_fast_math, mixed precision, GPU_LANG=CUDA, separable compilation; CMakeLists.txt:553-575,
// src/cu/CMakeLists.txt).  This file is ours and does two things:
//   1. shim: the few symbols those TUs need from front-ends that read Fortran modules (src/pme.cpp dispatchers, switchOff,
//      use(Potent), useEwald, boxVolume, gpuGridSize, the polpot / polpcg / inform / polar / extfld module variables);
//   2. driver: a C ABI that loads one system (the arrays of our System blob, which follow the reference's own layout:
//      pole[n][10], zaxis{z,x,y,polaxe}, m/dp/u/mdpu exclusion lists with scales), builds the reference's spatial lists and PME
//      units exactly as nblistData / pmeData do (src/nblist.cpp:415-464, src/pme.cpp:60-117), and runs
//         mpoleInit -> induce(uind, uinp)                        (src/amoeba/mpole.cpp:28-58, induce.cpp:108)
//         mpoleInit -> emplar_cu(vers) -> torque_cu              (src/amoeba/emplar.cpp:10-30: the fused energy path)
//      returning dipoles / energies / gradient / virial and CUDA-event timings on the reference's stream (g::s0).
// What it is for: SURVEY section 8(d) "Reference CUDA build (the 1.5x comparator)": the reference executable cannot be linked
// here (Fortran), so this runs the reference's kernels on the same inputs on the same B200.
#include "ff/amoeba/empole.h"
#include "ff/amoeba/epolar.h"
#include "ff/amoeba/induce.h"
#include "ff/atom.h"
#include "ff/box.h"
#include "ff/elec.h"
#include "ff/energybuffer.h"
#include "ff/evdw.h"
#include "ff/modamoeba.h"
#include "ff/nblist.h"
#include "ff/pme.h"
#include "ff/potent.h"
#include "ff/spatial.h"
#include "ff/switch.h"
#include "math/pow2.h"
#include "tool/accasync.h"
#include "tool/cudalib.h"
#include "tool/darray.h"
#include "tool/error.h"
#include "tool/externfunc.h"
#include "tool/gpucard.h"
#include "tool/platform.h"
#include "tool/rcman.h"
#include "tool/thrustcache.h"
#include <tinker/detail/extfld.hh>
#include <tinker/detail/inform.hh>
#include <tinker/detail/polar.hh>
#include <tinker/detail/polpcg.hh>
#include <tinker/detail/polpot.hh>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

// ------------------------------------------------------------------------------------------------------------------------
// 1. shim
// ------------------------------------------------------------------------------------------------------------------------
namespace {
double s_ewald_cut = 7, s_usolve_cut = 4.5, s_vdw_cut = 12, s_vdw_taper = 10.8;
int s_politer = 100, s_pcgprec = 1, s_pcgguess = 1, s_debug = 0, s_use_exfld = 0;
double s_poleps = 1e-5, s_pcgpeek = 1, s_texfld[3] = {0, 0, 0};
double* s_polarity = nullptr;
std::string s_err;
}

namespace tinker {
// Fortran module variables (ext/interface/cpp/tinker/detail/*.hh declare them as references bound to gfortran symbols)
namespace polpot {
int& politer = s_politer;
double& poleps = s_poleps;
}
namespace polpcg {
double& pcgpeek = s_pcgpeek;
int& pcgprec = s_pcgprec;
int& pcgguess = s_pcgguess;
}
namespace inform {
int& debug = s_debug;
}
namespace polar {
double*& polarity = s_polarity;
}
namespace extfld {
double (&texfld)[3] = s_texfld;
int& use_exfld = s_use_exfld;
}

// front-end queries
bool use(Potent term) { return term == Potent::MPOLE || term == Potent::POLAR; }
bool useEwald() { return true; }
bool useDEwald() { return false; }
real switchOff(Switch mode)
{
   return mode == Switch::USOLVE ? (real)s_usolve_cut : mode == Switch::VDW ? (real)s_vdw_cut : (real)s_ewald_cut;
}
real switchCut(Switch mode) { return mode == Switch::VDW ? (real)s_vdw_taper : switchOff(mode); }
real boxVolume()
{
   return lvec1.x * (lvec2.y * lvec3.z - lvec2.z * lvec3.y) - lvec1.y * (lvec2.x * lvec3.z - lvec2.z * lvec3.x)
      + lvec1.z * (lvec2.x * lvec3.y - lvec2.y * lvec3.x);
}
void extfieldModifyDField(real (*)[3], real (*)[3]) {}      // no external field

// src/cudart/gpucard.cpp:341-358 (that TU resets the device and shells out to nvidia-smi at start-up; same formulas here)
static cudaDeviceProp s_prop;
int gpuGridSize(int nthreads_per_block)
{
   nthreads_per_block = std::min(nthreads_per_block, s_prop.maxThreadsPerBlock);
   int per_mp = std::min((s_prop.maxThreadsPerMultiProcessor + nthreads_per_block - 1) / nthreads_per_block, s_prop.maxBlocksPerMultiProcessor);
   return s_prop.multiProcessorCount * per_mp;
}
int gpuMaxNParallel(int) { return s_prop.multiProcessorCount * s_prop.maxThreadsPerMultiProcessor; }

void printError() {}
void printBacktrace(std::FILE*) {}
void throwExceptionMissingFunction(const char* fn, const char* file, int line)
{
   throw std::runtime_error(std::string("missing function ") + fn + " at " + file + ":" + std::to_string(line));
}

PME::~PME()
{
   darray::deallocate(bsmod1, bsmod2, bsmod3, qgrid);
   darray::deallocate(igrid, thetai1, thetai2, thetai3);
}

// the dispatchers of src/pme.cpp:221-351 (that TU also reads the Fortran pme / ewald modules): CUDA build -> *_cu
void bsplineFill_cu(PMEUnit, int);
void gridMpole_cu(PMEUnit, real (*)[10]);
void gridUind_cu(PMEUnit, real (*)[3], real (*)[3]);
void pmeConv_cu(PMEUnit, EnergyBuffer, VirialBuffer);
void fphiMpole_cu(PMEUnit, real (*)[20]);
void fphiUind_cu(PMEUnit, real (*)[10], real (*)[10], real (*)[20]);
void fphiUind2_cu(PMEUnit, real (*)[10], real (*)[10]);
void rpoleToCmp_cu();
void cmpToFmp_cu(PMEUnit, const real (*)[10], real (*)[10]);
void cuindToFuind_cu(PMEUnit, const real (*)[3], const real (*)[3], real (*)[3], real (*)[3]);
void fphiToCphi_cu(PMEUnit, const real (*)[20], real (*)[10]);
void bsplineFill(PMEUnit pu, int level) { bsplineFill_cu(pu, level); }
void gridMpole(PMEUnit pu, real (*f)[10]) { gridMpole_cu(pu, f); }      // the *_cu spreads zero the grid themselves (pme.cu:287-321)
void gridUind(PMEUnit pu, real (*a)[3], real (*b)[3]) { gridUind_cu(pu, a, b); }
void pmeConv(PMEUnit pu) { pmeConv_cu(pu, nullptr, nullptr); }
void pmeConv(PMEUnit pu, VirialBuffer v) { pmeConv_cu(pu, nullptr, v); }
void pmeConv(PMEUnit pu, EnergyBuffer e) { pmeConv_cu(pu, e, nullptr); }
void pmeConv(PMEUnit pu, EnergyBuffer e, VirialBuffer v) { pmeConv_cu(pu, e, v); }
void fphiMpole(PMEUnit pu) { fphiMpole_cu(pu, fphi); }
void fphiUind(PMEUnit pu, real (*a)[10], real (*b)[10], real (*c)[20]) { fphiUind_cu(pu, a, b, c); }
void fphiUind2(PMEUnit pu, real (*a)[10], real (*b)[10]) { fphiUind2_cu(pu, a, b); }
void rpoleToCmp() { rpoleToCmp_cu(); }
void cmpToFmp(PMEUnit pu, const real (*c)[10], real (*f)[10]) { cmpToFmp_cu(pu, c, f); }
void cuindToFuind(PMEUnit pu, const real (*a)[3], const real (*b)[3], real (*c)[3], real (*d)[3]) { cuindToFuind_cu(pu, a, b, c, d); }
void fphiToCphi(PMEUnit pu, const real (*f)[20], real (*c)[10]) { fphiToCphi_cu(pu, f, c); }

// the dispatchers of src/amoeba/empole.cpp:67-71, src/hippo/empole.cpp:69, src/amoeba/epolar.cpp:538-542, 651-655, mpole.cpp:8-26
void empoleChgpenEwaldRecip_cu(int, int);
void epolarEwaldRecipSelf_cu(int, const real (*)[3], const real (*)[3]);
void epolar0DotProd_cu(const real (*)[3], const real (*)[3]);
void emplar_cu(int);
void torque_cu(int, grad_prec*, grad_prec*, grad_prec*);
void chkpole_cu();
void rotpole_cu();
void mpoleDataBinding_cu(RcOp);
void epolarDataBinding_cu(RcOp);
void ehal_cu(int);
void ehalReduceXyz_cu();
void ehalResolveGradient_cu();
void ehalResolveGradient() { ehalResolveGradient_cu(); }      // src/evdw.cpp:573-577
void empoleEwaldRecip(int vers) { empoleChgpenEwaldRecip_cu(vers, 0); }
void epolarEwaldRecipSelf(int vers) { epolarEwaldRecipSelf_cu(vers, uind, uinp); }
void epolar0DotProd(const real (*u)[3], const real (*v)[3]) { epolar0DotProd_cu(u, v); }
}

using namespace tinker;

// ------------------------------------------------------------------------------------------------------------------------
// 2. driver
// ------------------------------------------------------------------------------------------------------------------------
namespace {
bool s_open = false;
int s_have_virial = 1;

// src/amoeba/mpole.cpp:28-58
void mpole_init(int vers)
{
   if (vers & calc::grad)
      darray::zero(g::q0, n, trqx, trqy, trqz);
   if (vers & calc::virial)
      darray::zero(g::q0, bufferSize(), vir_trq);
   chkpole_cu();
   rotpole_cu();
   rpoleToCmp();
   if (vir_m)
      darray::zero(g::q0, bufferSize(), vir_m);
   bool precompute_theta = (!TINKER_CU_THETA_ON_THE_FLY_GRID_MPOLE) || (!TINKER_CU_THETA_ON_THE_FLY_GRID_UIND);
   if (precompute_theta) {
      bsplineFill(epme_unit, 3);
      if (pvpme_unit.valid())
         bsplineFill(pvpme_unit, 2);
   }
}

// src/pme.cpp:60-117 (pmeOpAlloc + pmeOpCopyin; the B-spline moduli come from the caller instead of Fortran dftmod)
PMEUnit open_pme(double aewald, const int* nfft, int bsorder, const double* b1, const double* b2, const double* b3)
{
   PMEUnit u = PMEUnit::open();
   PME& st = *u;
   darray::allocate(nfft[0], &st.bsmod1);
   darray::allocate(nfft[1], &st.bsmod2);
   darray::allocate(nfft[2], &st.bsmod3);
   darray::allocate(2 * (size_t)nfft[0] * nfft[1] * nfft[2], &st.qgrid);
   darray::allocate(3 * (size_t)n, &st.igrid);
   darray::allocate((size_t)padded_n * bsorder * 4, &st.thetai1, &st.thetai2, &st.thetai3);
   st.aewald = aewald, st.nfft1 = nfft[0], st.nfft2 = nfft[1], st.nfft3 = nfft[2], st.bsorder = bsorder;
   darray::copyin(g::q0, nfft[0], st.bsmod1, b1);
   darray::copyin(g::q0, nfft[1], st.bsmod2, b2);
   darray::copyin(g::q0, nfft[2], st.bsmod3, b3);
   waitFor(g::q0);
   u.deviceptrUpdate(st, g::q0);
   waitFor(g::q0);
   return u;
}

template <class F>
int guarded(F&& f)
{
   try {
      f();
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
         s_err = cudaGetErrorString(e);
         return 2;
      }
      return 0;
   } catch (const std::exception& e) {
      s_err = e.what();
      return 1;
   } catch (...) {
      s_err = "unknown exception";
      return 1;
   }
}
}

extern "C" {
const char* refcu_last_error(void) { return s_err.c_str(); }

struct refcu_system {
   int n;
   const double* xyz;        // [n][3]
   const double* lvec9;      // rows = lvec1, lvec2, lvec3
   const double* recip9;     // rows = recipa, recipb, recipc
   const double* pole;       // [n][10], MPL_PME order, before chkpole
   const int* zaxis;         // [n][4] = {z, x, y (signed, from ONE), polaxe}
   const double* polarity;   // [n]
   const double* thole;      // [n]
   const double* pdamp;      // [n]
   const int* jpolar;        // [n], from zero
   int njpolar;
   const double* thlval;     // [njpolar][njpolar]
   int nmexclude;
   const int* mexclude;
   const double* mexclude_scale;
   int ndpexclude;
   const int* dpexclude;
   const double* dpexclude_scale;      // [ndp][2]
   int nuexclude;
   const int* uexclude;
   const double* uexclude_scale;
   int nmdpuexclude;
   const int* mdpuexclude;
   const double* mdpuexclude_scale;      // [nmdpu][4]
   double aewald;
   int nfft[3];
   int bsorder;
   const double* bsmod1;
   const double* bsmod2;
   const double* bsmod3;
   double ewald_cutoff, usolve_cutoff, list_buffer;
   double poleps;
   int politer, pcgprec, pcgguess;
   double pcgpeek, uaccel, electric, dielec;
};

// layout check for the ctypes mirror (tests/test_ref_cuda_host.py)
int refcu_sizeof_system(void) { return (int)sizeof(refcu_system); }
int refcu_offsetof_dielec(void) { return (int)offsetof(refcu_system, dielec); }

int refcu_open(const refcu_system* s)
{
   if (s_open) {
      s_err = "refcu_open: one system per process";
      return 3;
   }
   return guarded([&] {
      int dev = 0, ndev = 0;
      if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
         throw std::runtime_error("no CUDA device");
      always_check_rt(cudaGetDevice(&dev));
      always_check_rt(cudaGetDeviceProperties(&s_prop, dev));
      idevice = dev, ndevice = 1;
      pltfm_config = Platform::CUDA;
      rc_flag = calc::xyz | calc::energy | calc::grad | calc::virial;
      s_ewald_cut = s->ewald_cutoff, s_usolve_cut = s->usolve_cutoff;
      s_poleps = s->poleps, s_politer = s->politer, s_pcgprec = s->pcgprec, s_pcgguess = s->pcgguess, s_pcgpeek = s->pcgpeek;

      // atoms and buffers (src/atom.cpp:27-41)
      n = s->n;
      padded_n = (n + WARP_SIZE - 1) / WARP_SIZE * WARP_SIZE;
      nelem_buffer = pow2Ge(gpuMaxNParallel(idevice));
      cudalibData(RcOp::ALLOC);
      darray::allocate(n, &x, &y, &z);
      std::vector<double> cx(n), cy(n), cz(n);
      for (int i = 0; i < n; ++i)
         cx[i] = s->xyz[3 * i], cy[i] = s->xyz[3 * i + 1], cz[i] = s->xyz[3 * i + 2];
      darray::copyin(g::q0, n, x, cx.data());
      darray::copyin(g::q0, n, y, cy.data());
      darray::copyin(g::q0, n, z, cz.data());
      const double* L = s->lvec9;
      const double* R = s->recip9;
      lvec1 = make_real3(L[0], L[1], L[2]), lvec2 = make_real3(L[3], L[4], L[5]), lvec3 = make_real3(L[6], L[7], L[8]);
      recipa = make_real3(R[0], R[1], R[2]), recipb = make_real3(R[3], R[4], R[5]), recipc = make_real3(R[6], R[7], R[8]);
      const double off = std::fabs(L[1]) + std::fabs(L[2]) + std::fabs(L[3]) + std::fabs(L[5]) + std::fabs(L[6]) + std::fabs(L[7]);
      box_shape = off < 1e-12 ? BoxShape::ORTHO : BoxShape::TRI;
      electric = s->electric, dielec = s->dielec;

      // permanent multipoles (mpoleData, src/elec.cpp:54-146)
      darray::allocate(n, &zaxis, &pole, &rpole);
      darray::copyin(g::q0, n, zaxis, reinterpret_cast<const LocalFrame*>(s->zaxis));
      darray::copyin(g::q0, n, pole, s->pole);
      darray::allocate(n, &trqx, &trqy, &trqz);
      // without calc::analyz the terms share the electrostatic accumulators (src/amoeba/empole.cpp:37-39, epolar.cpp:405-407)
      darray::allocate(bufferSize(), &vir_trq, &em, &vir_em, &nem, &nep);
      darray::allocate(n, &demx, &demy, &demz);
      ep = em, vir_ep = vir_em, depx = demx, depy = demy, depz = demz;
      mpoleDataBinding_cu(RcOp::ALLOC);

      // polarization (epolarData, src/amoeba/epolar.cpp:395-511)
      njpolar = s->njpolar;
      darray::allocate(n, &polarity, &thole, &pdamp, &polarity_inv, &jpolar);
      darray::allocate((size_t)njpolar * njpolar, &thlval);
      darray::allocate(n, &udir, &udirp, &uind, &uinp, &ufld, &dufld);
      darray::allocate(n, &work01_, &work02_, &work03_, &work04_, &work05_);
      darray::allocate(n, &work06_, &work07_, &work08_, &work09_, &work10_);
      std::vector<double> pinv(n);
      for (int i = 0; i < n; ++i)
         pinv[i] = 1.0 / std::max(s->polarity[i], 1.0e-16);
      darray::copyin(g::q0, n, polarity, s->polarity);
      darray::copyin(g::q0, n, thole, s->thole);
      darray::copyin(g::q0, n, pdamp, s->pdamp);
      darray::copyin(g::q0, n, polarity_inv, pinv.data());
      darray::copyin(g::q0, n, jpolar, s->jpolar);
      darray::copyin(g::q0, (size_t)njpolar * njpolar, thlval, s->thlval);
      udiag = s->uaccel;
      polpred = UPred::NONE, maxualt = 0, nualt = 0;
      epolarDataBinding_cu(RcOp::ALLOC);

      // exclusion lists with their scale factors
      nmexclude = s->nmexclude, ndpexclude = s->ndpexclude, nuexclude = s->nuexclude, nmdpuexclude = s->nmdpuexclude;
      darray::allocate(std::max(nmexclude, 1), &mexclude, &mexclude_scale);
      darray::allocate(std::max(ndpexclude, 1), &dpexclude, &dpexclude_scale);
      darray::allocate(std::max(nuexclude, 1), &uexclude, &uexclude_scale);
      darray::allocate(std::max(nmdpuexclude, 1), &mdpuexclude, &mdpuexclude_scale);
      if (nmexclude)
         darray::copyin(g::q0, nmexclude, mexclude, s->mexclude), darray::copyin(g::q0, nmexclude, mexclude_scale, s->mexclude_scale);
      if (ndpexclude)
         darray::copyin(g::q0, ndpexclude, dpexclude, s->dpexclude), darray::copyin(g::q0, ndpexclude, dpexclude_scale, s->dpexclude_scale);
      if (nuexclude)
         darray::copyin(g::q0, nuexclude, uexclude, s->uexclude), darray::copyin(g::q0, nuexclude, uexclude_scale, s->uexclude_scale);
      if (nmdpuexclude)
         darray::copyin(g::q0, nmdpuexclude, mdpuexclude, s->mdpuexclude),
            darray::copyin(g::q0, nmdpuexclude, mdpuexclude_scale, s->mdpuexclude_scale);
      waitFor(g::q0);

      // PME units (pmeData, src/pme.cpp:176-215): electrostatics == polarization unit, a second grid for the polarization virial
      epme_unit = open_pme(s->aewald, s->nfft, s->bsorder, s->bsmod1, s->bsmod2, s->bsmod3);
      ppme_unit = epme_unit;
      pvpme_unit = open_pme(s->aewald, s->nfft, s->bsorder, s->bsmod1, s->bsmod2, s->bsmod3);
      darray::allocate(n, &cmp, &fmp, &cphi, &fphi);
      darray::allocate(n, &fuind, &fuinp, &fdip_phi1, &fdip_phi2, &cphidp, &fphidp);
      darray::allocate(bufferSize(), &vir_m);
      fftData(RcOp::ALLOC);
      fftData(RcOp::INIT);

      // spatial lists (nblistData, src/nblist.cpp:415-464): the 4-scale-set multipole list and the preconditioner list
      Spatial::dataAlloc(mspatial_v2_unit, n, s->ewald_cutoff, s->list_buffer, x, y, z, 4, nmdpuexclude, mdpuexclude, nmexclude, mexclude,
         ndpexclude, dpexclude, nuexclude, uexclude);
      Spatial::dataAlloc(uspatial_v2_unit, n, s->usolve_cutoff, s->list_buffer, x, y, z, 1, nuexclude, uexclude, 0, nullptr, 0, nullptr, 0, nullptr);
      ThrustCache::allocate();
      Spatial::dataInit(mspatial_v2_unit);
      Spatial::dataInit(uspatial_v2_unit);
      waitFor(g::q0);
      s_open = true;
   });
}

// New coordinates (same box); rebuild = 1 rebuilds both spatial lists, as nblistRefresh does when an atom moved too far.
int refcu_set_xyz(const double* xyz, int rebuild)
{
   return guarded([&] {
      std::vector<double> cx(n), cy(n), cz(n);
      for (int i = 0; i < n; ++i)
         cx[i] = xyz[3 * i], cy[i] = xyz[3 * i + 1], cz[i] = xyz[3 * i + 2];
      darray::copyin(g::q0, n, x, cx.data());
      darray::copyin(g::q0, n, y, cy.data());
      darray::copyin(g::q0, n, z, cz.data());
      if (rebuild) {
         Spatial::dataInit(mspatial_v2_unit);
         Spatial::dataInit(uspatial_v2_unit);
      } else {
         Spatial::dataUpdateSorted(mspatial_v2_unit);
         Spatial::dataUpdateSorted(uspatial_v2_unit);
      }
      waitFor(g::q0);
   });
}

// mpoleInit + induce(uind, uinp): the converged d / p induced dipoles [n][3] (electron-Angstrom) and the direct dipoles.
int refcu_induce(double* ud, double* up, double* udir_out, double* udirp_out)
{
   return guarded([&] {
      mpole_init(calc::v0);
      induce(uind, uinp);
      if (ud)
         darray::copyout(g::q0, n, ud, uind);
      if (up)
         darray::copyout(g::q0, n, up, uinp);
      if (udir_out)
         darray::copyout(g::q0, n, udir_out, udir);
      if (udirp_out)
         darray::copyout(g::q0, n, udirp_out, udirp);
      waitFor(g::q0);
   });
}

static void energy_step(int vers)
{
   // zeroEGV of the electrostatic accumulators (src/energy.cpp:319-340), then emplar() of src/amoeba/emplar.cpp:10-30
   if (vers & calc::energy)
      darray::zero(g::q0, bufferSize(), em);
   if (vers & calc::virial)
      darray::zero(g::q0, bufferSize(), vir_em);
   if (vers & calc::grad)
      darray::zero(g::q0, n, demx, demy, demz);
   mpole_init(vers);
   emplar_cu(vers);
   torque_cu(vers, demx, demy, demz);
}

// the reductions energy_core performs on the electrostatic accumulators (src/energy.cpp:345-444) and emplar() on vir_trq
static void energy_reduce(int vers, double* esum, double* vir9)
{
   if (vers & calc::energy) {
      energy_prec e = energyReduce(em);
      if (esum)
         *esum = e;
   }
   if (vers & calc::virial) {
      virial_prec v1[9], v3[9];
      virialReduce(v1, vir_em), virialReduce(v3, vir_trq);
      if (vir9)
         for (int q = 0; q < 9; ++q)
            vir9[q] = v1[q] + v3[q];
   }
}

// The fused electrostatics of energy(vers): vers = 0x10 energy | 0x20 grad | 0x40 virial (calc::v0 / v1 / v4 / v5 / v6).
// esum = E_mpole + E_polar, kcal/mol; grad [n][3] kcal/mol/A; vir9 row-major.  Any output pointer may be null.
int refcu_energy(int vers, double* esum, double* grad, double* vir9)
{
   return guarded([&] {
      energy_step(vers);
      energy_reduce(vers, esum, vir9);
      if ((vers & calc::grad) && grad) {
         std::vector<grad_prec> a(n);
         grad_prec* dm[3] = {demx, demy, demz};
         for (int c = 0; c < 3; ++c) {
            darray::copyout(g::q0, n, a.data(), dm[c]);
            waitFor(g::q0);
            for (int i = 0; i < n; ++i)
               grad[3 * i + c] = toFloatingPoint<double>(a[i]);
         }
      }
   });
}

// CUDA-event timings on the reference's stream: what = 0 mpoleInit + induce(), 1 the fused energy step with vers incl. its reductions, 2 a rebuild of
// both spatial lists.  ms[reps] receives one elapsed time per repetition (host gaps inside the call included, as for ours).
int refcu_time(int what, int vers, int warmup, int reps, float* ms)
{
   return guarded([&] {
      cudaEvent_t e0, e1;
      always_check_rt(cudaEventCreate(&e0));
      always_check_rt(cudaEventCreate(&e1));
      for (int r = -warmup; r < reps; ++r) {
         always_check_rt(cudaStreamSynchronize(g::s0));
         always_check_rt(cudaEventRecord(e0, g::s0));
         if (what == 0) {
            mpole_init(calc::v0);
            induce(uind, uinp);
         } else if (what == 1) {
            energy_step(vers);
            energy_reduce(vers, nullptr, nullptr);
         } else {
            Spatial::dataInit(mspatial_v2_unit);
            Spatial::dataInit(uspatial_v2_unit);
         }
         always_check_rt(cudaEventRecord(e1, g::s0));
         always_check_rt(cudaEventSynchronize(e1));
         float t = 0;
         always_check_rt(cudaEventElapsedTime(&t, e0, e1));
         if (r >= 0)
            ms[r] = t;
      }
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
   });
}

// ---- the buffered 14-7 van der Waals term of the same step (src/cu/ehal.cu, unmodified): optional, after refcu_open.
struct refcu_vdw {
   const int* ired;        // [n] atom each site is reduced toward, from zero
   const double* kred;     // [n]
   const int* jvdw;        // [n] class index, from zero
   int njvdw;
   const double* radmin;   // [njvdw][njvdw]
   const double* epsilon;  // [njvdw][njvdw]
   int nvexclude;
   const int* vexclude;    // [nvexclude][2]
   const double* vexclude_scale;
   double cutoff, taper, list_buffer;
};
static bool s_vdw_open = false;
int refcu_sizeof_vdw(void) { return (int)sizeof(refcu_vdw); }

int refcu_vdw_open(const refcu_vdw* v)
{
   if (!s_open || s_vdw_open) {
      s_err = "refcu_vdw_open: needs refcu_open first, once";
      return 3;
   }
   return guarded([&] {
      s_vdw_cut = v->cutoff, s_vdw_taper = v->taper;
      vdwtyp = Vdw::HAL;
      ghal = 0.12, dhal = 0.07;      // compile-time constants of ehal.cu:96-99
      vcouple = Vdw::DECOUPLE, vlam = 1, scexp = 5, scalpha = 0.7;
      njvdw = v->njvdw;
      nvexclude = v->nvexclude;
      darray::allocate(n, &ired, &kred, &jvdw, &mut, &xred, &yred, &zred, &gxred, &gyred, &gzred, &devx, &devy, &devz);
      darray::allocate((size_t)njvdw * njvdw, &radmin, &epsilon);
      darray::allocate(std::max(nvexclude, 1), &vexclude, &vexclude_scale);
      darray::allocate(bufferSize(), &ev, &vir_ev, &nev);
      darray::copyin(g::q0, n, ired, v->ired);
      darray::copyin(g::q0, n, kred, v->kred);
      darray::copyin(g::q0, n, jvdw, v->jvdw);
      darray::zero(g::q0, n, mut);
      darray::copyin(g::q0, (size_t)njvdw * njvdw, radmin, v->radmin);
      darray::copyin(g::q0, (size_t)njvdw * njvdw, epsilon, v->epsilon);
      if (nvexclude)
         darray::copyin(g::q0, nvexclude, vexclude, v->vexclude), darray::copyin(g::q0, nvexclude, vexclude_scale, v->vexclude_scale);
      waitFor(g::q0);
      // vlist on the reduced sites (src/nblist.cpp:366-374)
      Spatial::dataAlloc(vspatial_v2_unit, n, v->cutoff, v->list_buffer, xred, yred, zred, 1, nvexclude, vexclude, 0, nullptr, 0, nullptr, 0,
         nullptr);
      ehalReduceXyz_cu();
      Spatial::dataInit(vspatial_v2_unit);
      waitFor(g::q0);
      s_vdw_open = true;
   });
}

// what the vdW term costs per step: reduced sites (nblistRefresh does this every step, src/nblist.cpp:545-548), zeroed accumulators,
// ehal_cu(vers) incl. the gradient hand-back to the real atoms, reductions.  No long-range correction (a host constant).
static void ehal_step(int vers, double* e, double* vir9)
{
   if (vers & calc::energy)
      darray::zero(g::q0, bufferSize(), ev);
   if (vers & calc::virial)
      darray::zero(g::q0, bufferSize(), vir_ev);
   if (vers & calc::grad)
      darray::zero(g::q0, n, devx, devy, devz);
   ehalReduceXyz_cu();
   ehal_cu(vers);
   if (vers & calc::energy) {
      energy_prec e1 = energyReduce(ev);
      if (e)
         *e = e1;
   }
   if (vers & calc::virial) {
      virial_prec v1[9];
      virialReduce(v1, vir_ev);
      if (vir9)
         for (int q = 0; q < 9; ++q)
            vir9[q] = v1[q];
   }
}

int refcu_ehal(int vers, double* e, double* grad, double* vir9)
{
   if (!s_vdw_open) {
      s_err = "refcu_ehal: refcu_vdw_open first";
      return 3;
   }
   return guarded([&] {
      ehal_step(vers, e, vir9);
      if ((vers & calc::grad) && grad) {
         std::vector<grad_prec> a(n);
         grad_prec* dv[3] = {devx, devy, devz};
         for (int c = 0; c < 3; ++c) {
            darray::copyout(g::q0, n, a.data(), dv[c]);
            waitFor(g::q0);
            for (int i = 0; i < n; ++i)
               grad[3 * i + c] = toFloatingPoint<double>(a[i]);
         }
      }
   });
}

int refcu_time_ehal(int vers, int warmup, int reps, float* ms)
{
   if (!s_vdw_open) {
      s_err = "refcu_time_ehal: refcu_vdw_open first";
      return 3;
   }
   return guarded([&] {
      cudaEvent_t e0, e1;
      always_check_rt(cudaEventCreate(&e0));
      always_check_rt(cudaEventCreate(&e1));
      for (int r = -warmup; r < reps; ++r) {
         always_check_rt(cudaStreamSynchronize(g::s0));
         always_check_rt(cudaEventRecord(e0, g::s0));
         ehal_step(vers, nullptr, nullptr);
         always_check_rt(cudaEventRecord(e1, g::s0));
         always_check_rt(cudaEventSynchronize(e1));
         float t = 0;
         always_check_rt(cudaEventElapsedTime(&t, e0, e1));
         if (r >= 0)
            ms[r] = t;
      }
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
   });
}
}
