// oracle/_ref/libref_dropin.so -- the REFERENCE'S OWN FRONT-ENDS running on OUR kernels.  TEST INFRASTRUCTURE ONLY.
//
// Links (oracle/Makefile, target dropin): the reference's unmodified front-end translation units src/amoeba/field.cpp
// (dfield, ufield), src/amoeba/induce.cpp (induce, sparsePrecondApply, diagPrecond), src/amoeba/emplar.cpp (emplar),
// src/amoeba/mpole.cpp (mpoleInit, torque), its energy-buffer reductions src/energybuffer.cpp (energyReduce, virialReduce) with
// the generic reduction / zeroing kernels they dispatch to (src/cu/mathparallel.cu, mathzero.cu), its stream / scratch set-up
// src/cudalib.cpp + src/cu/cudalib.cu, its global-variable TU src/mod.cpp and its device-memory layer src/cudart/darray.cpp,
// error.cpp -- with integration/apx_adapter.cpp (the `*_cu` operator symbols forwarded to libapx) INSTEAD of the reference's
// CUDA kernels of the path, and with tinker-gpu_b200/libapx.so.  This file is scaffolding only: it
// stands in for the front-ends that read Fortran modules (src/pme.cpp dispatchers, use(), useEwald(), switchOff, the polpot
// / polpcg / inform / polar module variables) and offers a small C ABI so a test can call the reference's functions
//    tinker::induce(uind, uinp)      tinker::dfield(field, fieldp)      tinker::ufield(uind, uinp, field, fieldp)
//    tinker::emplar(vers) followed by the reference's own energyReduce(em) / virialReduce(vir_em) / gradient buffers
// and read back the reference's own device globals.  What it shows: the adapter satisfies the operator boundary the reference
// links against, symbol for symbol, and the answers that come out of the reference's front-ends are the library's.
#include "apx.h"
#include "ff/amoeba/induce.h"
#include "ff/atom.h"
#include "ff/box.h"
#include "ff/elec.h"
#include "ff/energy.h"
#include "ff/modamoeba.h"
#include "ff/pme.h"
#include "ff/potent.h"
#include "ff/switch.h"
#include "tool/accasync.h"
#include "tool/cudalib.h"
#include "tool/darray.h"
#include "tool/error.h"
#include "tool/platform.h"
#include "tool/rcman.h"
#include "ff/amoeba/emplar.h"
#include "ff/amoeba/empole.h"
#include "math/pow2.h"
#include "tool/gpucard.h"
#include <algorithm>
#include <type_traits>
#include <tinker/detail/inform.hh>
#include <tinker/detail/polar.hh>
#include <tinker/detail/polpcg.hh>
#include <tinker/detail/polpot.hh>
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>
#include <vector>

namespace {
double s_ewald_cut = 7, s_usolve_cut = 2.5;
int s_politer = 100, s_pcgprec = 1, s_pcgguess = 1, s_debug = 0;
double s_poleps = 1e-5, s_pcgpeek = 1;
double* s_polarity = nullptr;
std::string s_err;
bool s_open = false;
}

namespace tinker {
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
bool use(Potent term) { return term == Potent::MPOLE || term == Potent::POLAR; }
bool useEwald() { return true; }
real switchOff(Switch mode) { return mode == Switch::USOLVE ? (real)s_usolve_cut : (real)s_ewald_cut; }
void extfieldModifyDField(real (*)[3], real (*)[3]) {}
void printError() {}
void printBacktrace(std::FILE*) {}
void throwExceptionMissingFunction(const char* fn, const char* file, int line)
{
   throw std::runtime_error(std::string("missing function ") + fn + " at " + file + ":" + std::to_string(line));
}
PME::~PME() {}
void exfield(int, int) {}      // src/elec.cpp:821 (Fortran-reading TU): no external field
// src/cudart/gpucard.cpp:341-358 (that TU resets the device and shells out to nvidia-smi at start-up; same formulas here)
static cudaDeviceProp s_prop;
int gpuGridSize(int nthreads_per_block)
{
   nthreads_per_block = std::min(nthreads_per_block, s_prop.maxThreadsPerBlock);
   int per_mp = std::min((s_prop.maxThreadsPerMultiProcessor + nthreads_per_block - 1) / nthreads_per_block, s_prop.maxBlocksPerMultiProcessor);
   return s_prop.multiProcessorCount * per_mp;
}
int gpuMaxNParallel(int) { return s_prop.multiProcessorCount * s_prop.maxThreadsPerMultiProcessor; }

// the dispatchers of src/pme.cpp:221-351 (Fortran-reading TU): CUDA build -> *_cu, which the adapter defines
void gridMpole_cu(PMEUnit, real (*)[10]);
void gridUind_cu(PMEUnit, real (*)[3], real (*)[3]);
void pmeConv_cu(PMEUnit, EnergyBuffer, VirialBuffer);
void fphiMpole_cu(PMEUnit, real (*)[20]);
void fphiUind_cu(PMEUnit, real (*)[10], real (*)[10], real (*)[20]);
void fphiUind2_cu(PMEUnit, real (*)[10], real (*)[10]);
void cmpToFmp_cu(PMEUnit, const real (*)[10], real (*)[10]);
void cuindToFuind_cu(PMEUnit, const real (*)[3], const real (*)[3], real (*)[3], real (*)[3]);
void fphiToCphi_cu(PMEUnit, const real (*)[20], real (*)[10]);
void rpoleToCmp_cu();
void bsplineFill_cu(PMEUnit, int);
void rpoleToCmp() { rpoleToCmp_cu(); }
void bsplineFill(PMEUnit pu, int level) { bsplineFill_cu(pu, level); }
void gridMpole(PMEUnit pu, real (*f)[10]) { gridMpole_cu(pu, f); }
void gridUind(PMEUnit pu, real (*a)[3], real (*b)[3]) { gridUind_cu(pu, a, b); }
void pmeConv(PMEUnit pu) { pmeConv_cu(pu, nullptr, nullptr); }
void pmeConv(PMEUnit pu, VirialBuffer v) { pmeConv_cu(pu, nullptr, v); }
void fphiMpole(PMEUnit pu) { fphiMpole_cu(pu, fphi); }
void fphiUind(PMEUnit pu, real (*a)[10], real (*b)[10], real (*c)[20]) { fphiUind_cu(pu, a, b, c); }
void fphiUind2(PMEUnit pu, real (*a)[10], real (*b)[10]) { fphiUind2_cu(pu, a, b); }
void cmpToFmp(PMEUnit pu, const real (*c)[10], real (*f)[10]) { cmpToFmp_cu(pu, c, f); }
void cuindToFuind(PMEUnit pu, const real (*a)[3], const real (*b)[3], real (*c)[3], real (*d)[3]) { cuindToFuind_cu(pu, a, b, c, d); }
void fphiToCphi(PMEUnit pu, const real (*f)[20], real (*c)[10]) { fphiToCphi_cu(pu, f, c); }

// integration/apx_adapter.cpp
void apxAdapterCreate(const apx_system& s, int device);
void apxAdapterDestroy();
void apxAdapterRefreshPositions();
}

using namespace tinker;

namespace {
// a floating-point value in the element type of the reference's gradient / energy buffers of this build (fixed point in the
// mixed build, include/ff/precision.h:68-106)
template <class T>
T toBuf(double v)
{
   return std::is_same<T, fixed>::value ? static_cast<T>(static_cast<long long>(v * 0x100000000ull)) : static_cast<T>(v);
}
template <class F>
int guarded(F&& f)
{
   try {
      f();
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
const char* dropin_last_error(void) { return s_err.c_str(); }

// sys: the apx_system of include/apx.h (what apxCreateFromModules of INTEGRATION.md fills from the Fortran modules)
int dropin_open(const apx_system* sys, double list_buffer_of_usolve)
{
   if (s_open) {
      s_err = "dropin_open: one system per process";
      return 3;
   }
   return guarded([&] {
      int ndev = 0;
      if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
         throw std::runtime_error("no CUDA device");
      int dev = 0;
      always_check_rt(cudaGetDevice(&dev));
      always_check_rt(cudaGetDeviceProperties(&s_prop, dev));
      idevice = dev, ndevice = 1;
      pltfm_config = Platform::CUDA;
      rc_flag = calc::xyz | calc::energy | calc::grad | calc::virial;      // a DYNAMIC / TESTGRAD run: no calc::analyz
      n = sys->n;
      padded_n = (n + 31) / 32 * 32;
      nelem_buffer = pow2Ge(gpuMaxNParallel(idevice));      // src/atom.cpp:27-41
      cudalibData(RcOp::ALLOC);                              // the reference's own streams g::s0 / g::s1, queues and reduction scratch
      // coordinates as the reference holds them (include/ff/atom.h:39-45): the adapter reads these device arrays
      darray::allocate(n, &x, &y, &z);
      darray::allocate(n, &xpos, &ypos, &zpos);      // the integrator's coordinates (mdData, pos_prec)
      {
         std::vector<double> cx(n), cy(n), cz(n);
         for (int i = 0; i < n; ++i)
            cx[i] = sys->xyz[3 * i], cy[i] = sys->xyz[3 * i + 1], cz[i] = sys->xyz[3 * i + 2];
         darray::copyin(g::q0, n, x, cx.data());
         darray::copyin(g::q0, n, y, cy.data());
         darray::copyin(g::q0, n, z, cz.data());
         darray::copyin(g::q0, n, xpos, cx.data());
         darray::copyin(g::q0, n, ypos, cy.data());
         darray::copyin(g::q0, n, zpos, cz.data());
         waitFor(g::q0);
      }
      s_ewald_cut = sys->cutoff;
      s_usolve_cut = sys->usolve_cutoff > 0 ? sys->usolve_cutoff - list_buffer_of_usolve : 0;
      s_poleps = sys->poleps, s_politer = sys->politer, s_pcgprec = sys->pcgprec, s_pcgguess = sys->pcgguess, s_pcgpeek = sys->pcgpeek;
      polpred = UPred::NONE, maxualt = 0, nualt = 0;
      // the reference's device globals its front-ends read and write (epolarData / pmeData allocate them)
      darray::allocate(n, &uind, &uinp, &udir, &udirp);
      darray::allocate(n, &work01_, &work02_, &work03_, &work04_, &work05_);
      darray::allocate(n, &work06_, &work07_, &work08_, &work09_, &work10_);
      darray::allocate(n, &cmp, &fmp, &cphi, &fphi, &fuind, &fuinp, &fdip_phi1, &fdip_phi2);
      // the electrostatic accumulators of a non-analyze run (src/energy.cpp / egvData: one shared set for all electrostatic
      // terms): em = ep = eng_buf_elec, vir_em = vir_ep = vir_buf_elec, demx = depx = gx_elec; torque arrays and vir_trq
      darray::allocate(bufferSize(), &eng_buf_elec, &vir_buf_elec, &vir_trq);
      darray::allocate(n, &gx_elec, &gy_elec, &gz_elec, &trqx, &trqy, &trqz);
      em = ep = eng_buf_elec, vir_em = vir_ep = vir_buf_elec;
      demx = depx = gx_elec, demy = depy = gy_elec, demz = depz = gz_elec;
      vir_m = nullptr;
      epme_unit = PMEUnit::open();      // a handle only: the library owns the grids
      PME& st = *epme_unit;
      st.aewald = sys->aewald, st.nfft1 = sys->nfft[0], st.nfft2 = sys->nfft[1], st.nfft3 = sys->nfft[2], st.bsorder = sys->bsorder;
      st.bsmod1 = st.bsmod2 = st.bsmod3 = st.qgrid = nullptr, st.igrid = nullptr, st.thetai1 = st.thetai2 = st.thetai3 = nullptr;
      ppme_unit = epme_unit;
      apxAdapterCreate(*sys, 0);
      s_open = true;
   });
}

void dropin_close(void)
{
   apxAdapterDestroy();
   s_open = false;
}

static void out3(double* dst, const real (*dev)[3])
{
   if (dst)
      darray::copyout(g::q0, n, dst, dev);
}

// the reference's induce() front-end (src/amoeba/induce.cpp:108-113); the dipoles are read from the REFERENCE'S globals
int dropin_induce(double* ud, double* up, double* ud_dir, double* up_dir)
{
   return guarded([&] {
      induce(uind, uinp);
      out3(ud, uind), out3(up, uinp), out3(ud_dir, udir), out3(up_dir, udirp);
      waitFor(g::q0);
   });
}

// the reference's dfield() front-end (src/amoeba/field.cpp:17-63) into its work arrays
int dropin_dfield(double* field, double* fieldp)
{
   return guarded([&] {
      dfield(work01_, work02_);
      out3(field, work01_), out3(fieldp, work02_);
      waitFor(g::q0);
   });
}

// the reference's ufield() front-end (src/amoeba/field.cpp:78-117) of prescribed dipoles
int dropin_ufield(const double* ud, const double* up, double* field, double* fieldp)
{
   return guarded([&] {
      darray::copyin(g::q0, n, work09_, ud);
      darray::copyin(g::q0, n, work10_, up);
      ufield(work09_, work10_, work01_, work02_);
      out3(field, work01_), out3(fieldp, work02_);
      waitFor(g::q0);
   });
}

// the reference's sparsePrecondApply() front-end (src/amoeba/induce.cpp:21-25)
int dropin_precond(const double* rd, const double* rp, double* zd, double* zp)
{
   return guarded([&] {
      darray::copyin(g::q0, n, work03_, rd);
      darray::copyin(g::q0, n, work04_, rp);
      sparsePrecondApply(work03_, work04_, work05_, work06_);
      out3(zd, work05_), out3(zp, work06_);
      waitFor(g::q0);
   });
}

// The reference's emplar() front-end (src/amoeba/emplar.cpp:10-28: mpoleInit, emplar_cu, torque, virialReduce(vir_trq)) between
// the two halves of the reference's energy() (src/energy.cpp:333-446): the accumulators are zeroed as zeroEGV does, `preload`
// is first ADDED to every slot-0 / gradient entry (standing for the other electrostatic terms that share these buffers), then
// the energy, virial and gradient come out of the reference's own energyReduce / virialReduce and its gx_elec arrays.
// With an adapter that assigned instead of accumulating, the preload would be lost.
int dropin_emplar(int vers, double preload, double* e_elec, double* grad, double* vir9)
{
   return guarded([&] {
      darray::zero(g::q0, bufferSize(), eng_buf_elec, vir_buf_elec);
      darray::zero(g::q0, n, gx_elec, gy_elec, gz_elec);
      for (int q = 0; q < 9; ++q)
         virial_elec[q] = 0;
      if (preload != 0) {
         std::vector<grad_prec> a(n, toBuf<grad_prec>(preload));
         darray::copyin(g::q0, n, gx_elec, a.data());
         darray::copyin(g::q0, n, gy_elec, a.data());
         darray::copyin(g::q0, n, gz_elec, a.data());
         EnergyBufferTraits::type e0 = toBuf<EnergyBufferTraits::type>(preload);
         darray::copyin(g::q0, 1, eng_buf_elec, &e0);
         waitFor(g::q0);
      }
      apxAdapterRefreshPositions();      // copyPosToXyz + nblistRefresh from the reference's x / y / z
      emplar(vers);
      if (e_elec)
         *e_elec = (vers & calc::energy) ? energyReduce(eng_buf_elec) : 0.0;
      if (vir9) {
         virial_prec v[9] = {0};
         if (vers & calc::virial)
            virialReduce(v, vir_buf_elec);
         for (int q = 0; q < 9; ++q)
            vir9[q] = v[q] + virial_elec[q];      // + what the front-end reduced from vir_trq
      }
      if (grad && (vers & calc::grad)) {
         std::vector<grad_prec> a(n);
         grad_prec* dm[3] = {gx_elec, gy_elec, gz_elec};
         for (int c = 0; c < 3; ++c) {
            darray::copyout(g::q0, n, a.data(), dm[c]);
            waitFor(g::q0);
            for (int i = 0; i < n; ++i)
               grad[3 * i + c] = toFloatingPoint<double>(a[i]);
         }
      }
   });
}

// new coordinates into the reference's x / y / z (what mdPos / copyPosToXyz leave there)
int dropin_set_xyz(const double* xyz)
{
   return guarded([&] {
      std::vector<double> cx(n), cy(n), cz(n);
      for (int i = 0; i < n; ++i)
         cx[i] = xyz[3 * i], cy[i] = xyz[3 * i + 1], cz[i] = xyz[3 * i + 2];
      darray::copyin(g::q0, n, x, cx.data());
      darray::copyin(g::q0, n, y, cy.data());
      darray::copyin(g::q0, n, z, cz.data());
      darray::copyin(g::q0, n, xpos, cx.data());
      darray::copyin(g::q0, n, ypos, cy.data());
      darray::copyin(g::q0, n, zpos, cz.data());
      waitFor(g::q0);
      apxAdapterRefreshPositions();
   });
}

// device time of `reps` calls of the reference's induce() front-end on libapx, CUDA events on the reference's stream
int dropin_time_induce(int reps, double* ms_per_call)
{
   return guarded([&] {
      cudaEvent_t e0, e1;
      always_check_rt(cudaEventCreate(&e0));
      always_check_rt(cudaEventCreate(&e1));
      induce(uind, uinp);
      always_check_rt(cudaEventRecord(e0, g::s0));
      for (int r = 0; r < reps; ++r)
         induce(uind, uinp);
      always_check_rt(cudaEventRecord(e1, g::s0));
      always_check_rt(cudaEventSynchronize(e1));
      float ms = 0;
      always_check_rt(cudaEventElapsedTime(&ms, e0, e1));
      *ms_per_call = ms / reps;
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
   });
}
}
