// apx_adapter.cpp -- the ONE translation unit a Tinker-GPU maintainer adds (as src/cu/apx_adapter.cpp) to run the reference's
// unmodified front-ends on libapx: it defines the `*_cu` operator symbols that TINKER_FCALL2 resolves to in a GPU_LANG=CUDA
// build (include/tool/externfunc.h:26-39) by forwarding to the C ABI of include/apx.h, and the reference's own
// src/cu/amoeba/*.cu, src/cu/pme.cu, src/cu/epolarrecip.cu, src/cu/induce.cu are dropped from src/cu/cmakesrc.txt.
//
// It is compiled HERE against the reference's headers where they lie (oracle/Makefile, target dropin) and linked with the
// reference's unmodified front-ends src/amoeba/field.cpp and src/amoeba/induce.cpp into oracle/_ref/libref_dropin.so, so
// that tinker::induce(), tinker::dfield() and tinker::ufield() -- the reference's own functions -- run on our kernels
// (oracle/ref_dropin.cpp is the scaffolding that stands in for the Fortran-reading front-ends; tests/test_dropin_host.py,
// tests/test_zgpu_9_refcuda.py).  INTEGRATION.md has the symbol map.
//
// State convention: the reference keeps its arrays in process-global device pointers (include/ff/modamoeba.h); libapx keeps
// its own resident copies behind the context.  Where a front-end reads a global after the call (uind / uinp / udir / udirp,
// field / fieldp work arrays, the gradient and energy accumulators), the stub copies the library's result into it.
#include "apx.h"
#include "ff/amoeba/induce.h"
#include "ff/atom.h"
#include "ff/energy.h"
#include "ff/modamoeba.h"
#include "ff/pme.h"
#include "tool/darray.h"
#include "tool/error.h"
#include <vector>

namespace tinker {
static apx_ctx* g_apx = nullptr;

static void chk(int rc)
{
   if (rc)
      TINKER_THROW(apx_last_error());      // error convention of the boundary: include/tool/error.h:16-45
}

// epolarData(RcOp::ALLOC | RcOp::INIT) calls this once the Fortran modules are read (INTEGRATION.md shows the filling of
// apx_system from atoms:: / mpole:: / polar:: / polpot:: / ewald:: / pme::); RcOp::DEALLOC calls apxAdapterDestroy.
void apxAdapterCreate(const apx_system& s, int device)
{
   if (g_apx)
      apx_destroy(g_apx), g_apx = nullptr;
   chk(apx_create(&s, device, &g_apx));
}
void apxAdapterDestroy()
{
   if (g_apx)
      apx_destroy(g_apx), g_apx = nullptr;
}
apx_ctx* apxAdapterContext() { return g_apx; }

// copyPosToXyz + nblistRefresh (src/nblist.cpp:521): new coordinates, list check / rebuild inside the library
void apxAdapterSetPositions(const double* xyz) { chk(apx_set_positions(g_apx, xyz)); }

namespace {
struct Host3 {      // host double [n][3] staging of one of the reference's device real [n][3] arrays
   std::vector<double> v;
   Host3()
      : v(3 * (size_t)n)
   {}
   void from(const real (*dev)[3])
   {
      darray::copyout(g::q0, n, v.data(), dev);
      waitFor(g::q0);
   }
   void to(real (*dev)[3]) const { darray::copyin(g::q0, n, dev, v.data()); }
};
}

// ---- src/amoeba/mpole.cpp:8-26
void chkpole_cu() {}      // chkpole + rotpole + rpoleToCmp are one kernel inside the library (frames.cu)
void rotpole_cu() { chk(apx_mpole_init(g_apx)); }
void rpoleToCmp_cu() {}
void torque_cu(int, grad_prec*, grad_prec*, grad_prec*) {}      // inside apx_energy / apx_empole / apx_epolar

// ---- src/amoeba/field.cpp:8-117.  The front-end composes the reciprocal part from fine-grained PME operators and then calls
//      the real-space sweep; the library's operator contains all of it, so the last call of each sequence does the work and
//      the earlier ones have nothing left to do (the PME dispatchers of src/pme.cpp resolve to the empty stubs further down).
void dfieldEwaldRecipSelfP2_cu(real (*)[3]) {}
void dfieldEwaldReal_cu(real (*field)[3], real (*fieldp)[3])
{
   Host3 a, b;
   chk(apx_dfield(g_apx, a.v.data(), b.v.data()));
   a.to(field), b.to(fieldp);
}
void dfieldNonEwald_cu(real (*field)[3], real (*fieldp)[3]) { dfieldEwaldReal_cu(field, fieldp); }
void ufieldEwaldRecipSelfP1_cu(const real (*)[3], const real (*)[3], real (*)[3], real (*)[3]) {}
void ufieldEwaldReal_cu(const real (*ud)[3], const real (*up)[3], real (*field)[3], real (*fieldp)[3])
{
   Host3 u, p, a, b;
   u.from(ud), p.from(up);
   chk(apx_ufield(g_apx, u.v.data(), p.v.data(), a.v.data(), b.v.data()));
   a.to(field), b.to(fieldp);
}
void ufieldNonEwald_cu(const real (*ud)[3], const real (*up)[3], real (*field)[3], real (*fieldp)[3])
{
   ufieldEwaldReal_cu(ud, up, field, fieldp);
}

// ---- src/amoeba/induce.cpp:12-73
void sparsePrecondApply_cu(const real (*rsd)[3], const real (*rsdp)[3], real (*zrsd)[3], real (*zrsdp)[3])
{
   Host3 r, q, a, b;
   r.from(rsd), q.from(rsdp);
   chk(apx_precond(g_apx, r.v.data(), q.v.data(), a.v.data(), b.v.data()));
   a.to(zrsd), b.to(zrsdp);
}
void diagPrecond_cu(const real (*rsd)[3], const real (*rsdp)[3], real (*zrsd)[3], real (*zrsdp)[3])
{
   sparsePrecondApply_cu(rsd, rsdp, zrsd, zrsdp);      // apx_system.usolve_cutoff <= 0 selects the diagonal form
}
void ulspredSaveP1_cu(real (*)[3], real (*)[3], const real (*)[3], const real (*)[3]) {}      // history ring lives in apx_induce
void ulspredSum_cu(real (*)[3], real (*)[3]) {}
void induceMutualPcg1_cu(real (*ud)[3], real (*up)[3])
{
   chk(apx_induce(g_apx));
   Host3 a, b;
   chk(apx_get_uind(g_apx, a.v.data(), b.v.data()));
   a.to(ud), b.to(up);
   if (udir && udirp) {      // epolar0DotProd and the OPT / print paths read the direct dipoles
      chk(apx_get_udir(g_apx, a.v.data(), b.v.data()));
      a.to(udir), b.to(udirp);
   }
   waitFor(g::q0);
}

// ---- src/amoeba/emplar.cpp:9, empole.cpp:53-71, epolar.cpp:515-655: the library reduces on the device and hands the totals
//      back; the front-ends' own accumulators receive them (energy_em / energy_ep / virial_em, and the gradient as the
//      fixed-point or floating grad_prec the build uses)
static void store_gradient(grad_prec* gx, grad_prec* gy, grad_prec* gz)
{
   if (!gx)
      return;
   std::vector<double> g(3 * (size_t)n);
   chk(apx_get_gradient(g_apx, g.data()));
   std::vector<grad_prec> c[3];
   for (int k = 0; k < 3; ++k) {
      c[k].resize(n);
      for (int i = 0; i < n; ++i) {
#if TINKER_DETERMINISTIC_FORCE
         c[k][i] = static_cast<grad_prec>(static_cast<long long>(g[3 * (size_t)i + k] * 0x100000000ull));
#else
         c[k][i] = static_cast<grad_prec>(g[3 * (size_t)i + k]);
#endif
      }
   }
   darray::copyin(g::q0, n, gx, c[0].data());
   darray::copyin(g::q0, n, gy, c[1].data());
   darray::copyin(g::q0, n, gz, c[2].data());
   waitFor(g::q0);
}
static void run(int (*op)(apx_ctx*, int, apx_energy_result*), int vers, bool mpole, bool polar)
{
   apx_energy_result r;
   chk(op(g_apx, vers, &r));
   if (mpole)
      energy_em = r.em;
   if (polar)
      energy_ep = r.ep;
   if (vers & calc::virial)
      for (int i = 0; i < 9; ++i)
         (mpole ? virial_em : virial_ep)[i] = r.virial[i];
   if (vers & calc::grad)
      mpole ? store_gradient(demx, demy, demz) : store_gradient(depx, depy, depz);
}
void emplar_cu(int vers) { run(apx_energy, vers, true, true); }
void empoleEwaldRealSelf_cu(int vers) { run(apx_empole, vers, true, false); }
void empoleNonEwald_cu(int vers) { run(apx_empole, vers, true, false); }
void empoleChgpenEwaldRecip_cu(int, int) {}      // contained in apx_empole
void epolarEwaldReal_cu(int vers, const real (*)[3], const real (*)[3]) { run(apx_epolar, vers, false, true); }
void epolarNonEwald_cu(int vers, const real (*)[3], const real (*)[3]) { run(apx_epolar, vers, false, true); }
void epolarEwaldRecipSelf_cu(int, const real (*)[3], const real (*)[3]) {}      // contained in apx_epolar
void epolar0DotProd_cu(const real (*)[3], const real (*)[3]) {}                // contained in apx_epolar / apx_energy

// ---- src/pme.cpp:221-347: fine-grained PME operators -- fused inside the library's field / energy operators
void bsplineFill_cu(PMEUnit, int) {}
void gridMpole_cu(PMEUnit, real (*)[10]) {}
void gridUind_cu(PMEUnit, real (*)[3], real (*)[3]) {}
void pmeConv_cu(PMEUnit, EnergyBuffer, VirialBuffer) {}
void fphiMpole_cu(PMEUnit, real (*)[20]) {}
void fphiUind_cu(PMEUnit, real (*)[10], real (*)[10], real (*)[20]) {}
void fphiUind2_cu(PMEUnit, real (*)[10], real (*)[10]) {}
void cmpToFmp_cu(PMEUnit, const real (*)[10], real (*)[10]) {}
void cuindToFuind_cu(PMEUnit, const real (*)[3], const real (*)[3], real (*)[3], real (*)[3]) {}
void fphiToCphi_cu(PMEUnit, const real (*)[20], real (*)[10]) {}

// ---- src/cudart/fft.cpp:16-100 (dropped together with the kernels): the library owns its FFT plans and grids
void fftData(RcOp) {}
void fftfront(PMEUnit) {}
void fftback(PMEUnit) {}
}
