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
// field / fieldp work arrays) the stub has the library write it device to device; the gradient, energy and virial
// accumulators are ADDED to on the device (the *_dev entry points of include/apx.h) -- no stub stages through host memory.
#include "apx.h"
#include "ff/amoeba/induce.h"
#include "ff/atom.h"
#include "ff/energy.h"
#include "ff/modamoeba.h"
#include "ff/pme.h"
#include "tool/darray.h"
#include "tool/cudalib.h"
#include "tool/error.h"
#include <type_traits>

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
constexpr int EB = (int)sizeof(real);      // element size of the reference's real (*)[3] device arrays
// grad_prec / the energy and virial buffer element of this build (include/ff/precision.h:68-106, ff/energybuffer.h)
template <class T>
constexpr int devKind()
{
   return std::is_same<T, fixed>::value ? APX_DEV_FIXED : (sizeof(T) == 4 ? APX_DEV_F32 : APX_DEV_F64);
}
void* stream0() { return (void*)g::s0; }      // the stream the reference's front-ends enqueue on
}

// copyPosToXyz + nblistRefresh (src/nblist.cpp:521) from the reference's own device arrays, no host copy: the integrator's
// xpos / ypos / zpos (pos_prec: double also in the mixed build, include/ff/atom.h:39-45) when they exist -- the float copies
// x / y / z have lost 4e-6 A at 60 A, which alone costs 2e-5 kcal/mol/A of force accuracy -- else x / y / z
void apxAdapterRefreshPositions()
{
   if (xpos && ypos && zpos)
      chk(apx_set_positions_dev(g_apx, xpos, ypos, zpos, (int)sizeof(pos_prec), stream0()));
   else
      chk(apx_set_positions_dev(g_apx, x, y, z, EB, stream0()));
}

// ---- *DataBinding_cu (src/cu/amoeba/binding.cu:11-52, called from src/elec.cpp:53 and src/amoeba/epolar.cpp:24): the
//      reference mirrors its device pointers into __device__ globals (d::rpole, d::pdamp, ...) for its own kernels.  The
//      library keeps its state behind the context, so there is nothing to bind -- but the symbols must exist once
//      src/cu/amoeba/binding.cu is dropped from the build.
void mpoleDataBinding_cu(RcOp) {}
void epolarDataBinding_cu(RcOp) {}

// ---- src/amoeba/mpole.cpp:8-26
void chkpole_cu() {}      // chkpole + rotpole + rpoleToCmp are one kernel inside the library (frames.cu)
void rotpole_cu() { chk(apx_mpole_init(g_apx)); }
void rpoleToCmp_cu() {}
// torque(vers, demx, demy, demz) (src/amoeba/emplar.cpp:20): the library's energy operators already turned the torques
// into forces and booked their virial with the pair virial, so the reference's torque arrays and vir_trq stay as mpoleInit
// zeroed them and the front-end's virialReduce(vir_trq) adds nothing
void torque_cu(int, grad_prec*, grad_prec*, grad_prec*) {}

// ---- src/amoeba/field.cpp:8-117.  The front-end composes the reciprocal part from fine-grained PME operators and then calls
//      the real-space sweep; the library's operator contains all of it, so the last call of each sequence does the work and
//      the earlier ones have nothing left to do (the PME dispatchers of src/pme.cpp resolve to the empty stubs further down).
//      The field arrays are ASSIGNED here because the front-end's own sequence starts by zeroing them (darray::zero in
//      dfieldEwaldRecipSelfP2 / ufieldEwaldRecipSelfP1's callers, field.cpp:26,86) and the stubs that would have accumulated
//      the reciprocal part are empty.
void dfieldEwaldRecipSelfP2_cu(real (*)[3]) {}
void dfieldEwaldReal_cu(real (*field)[3], real (*fieldp)[3]) { chk(apx_dfield_dev(g_apx, field, fieldp, EB, stream0())); }
void dfieldNonEwald_cu(real (*field)[3], real (*fieldp)[3]) { dfieldEwaldReal_cu(field, fieldp); }
void ufieldEwaldRecipSelfP1_cu(const real (*)[3], const real (*)[3], real (*)[3], real (*)[3]) {}
void ufieldEwaldReal_cu(const real (*ud)[3], const real (*up)[3], real (*field)[3], real (*fieldp)[3])
{
   chk(apx_ufield_dev(g_apx, ud, up, field, fieldp, EB, stream0()));
}
void ufieldNonEwald_cu(const real (*ud)[3], const real (*up)[3], real (*field)[3], real (*fieldp)[3])
{
   ufieldEwaldReal_cu(ud, up, field, fieldp);
}

// ---- src/amoeba/induce.cpp:12-73
void sparsePrecondApply_cu(const real (*rsd)[3], const real (*rsdp)[3], real (*zrsd)[3], real (*zrsdp)[3])
{
   chk(apx_precond_dev(g_apx, rsd, rsdp, zrsd, zrsdp, EB, stream0()));
}
void diagPrecond_cu(const real (*rsd)[3], const real (*rsdp)[3], real (*zrsd)[3], real (*zrsdp)[3])
{
   sparsePrecondApply_cu(rsd, rsdp, zrsd, zrsdp);      // apx_system.usolve_cutoff <= 0 selects the diagonal form
}
void ulspredSaveP1_cu(real (*)[3], real (*)[3], const real (*)[3], const real (*)[3]) {}      // history ring lives in apx_induce
void ulspredSum_cu(real (*)[3], real (*)[3]) {}
void induceMutualPcg1_cu(real (*ud)[3], real (*up)[3])
{
   // epolar0DotProd and the OPT / print paths read the direct dipoles: handed back with the solution, device to device
   chk(apx_induce_dev(g_apx, ud, up, udir, udirp, EB, stream0()));
}

// ---- src/amoeba/emplar.cpp:9, empole.cpp:53-71, epolar.cpp:515-655.  Contract (SURVEY 8b "Ownership",
//      src/energy.cpp:333-446): the operator ADDS its energy to one slot of the term's energy buffer, its virial to one slot of
//      the virial buffer and its gradient to the term's gradient arrays, all on the device; energy() reduces the buffers
//      (energyReduce / virialReduce -> esum, vir) and sums gx_elec into gx.  In a non-analyze run em / ep alias eng_buf_elec,
//      vir_em / vir_ep alias vir_buf_elec and demx / depx alias gx_elec, so anything assigned instead of added would wipe the
//      other electrostatic terms.  The analyze-mode host scalars (energy_em ...) are NOT touched here: the reference's
//      front-ends fill them from the buffers themselves (empole.cpp:118-137, epolar.cpp:621-647).
static void hand_back(int vers, const apx_energy_result& r, double e, EnergyBuffer eb, VirialBuffer vb, grad_prec* gx_, grad_prec* gy_,
   grad_prec* gz_)
{
   if ((vers & calc::energy) && eb)
      chk(apx_add_scalars_dev(g_apx, eb, &e, 1, devKind<EnergyBufferTraits::type>(), stream0()));
   if ((vers & calc::virial) && vb) {
      // xx yx zx yy zy zz (src/energybuffer.cpp:81-94)
      const double v6[6] = {r.virial[0], r.virial[1], r.virial[2], r.virial[4], r.virial[5], r.virial[8]};
      chk(apx_add_scalars_dev(g_apx, &vb[0][0], v6, 6, devKind<VirialBufferTraits::type>(), stream0()));
   }
   if ((vers & calc::grad) && gx_)
      chk(apx_add_gradient_dev(g_apx, gx_, gy_, gz_, devKind<grad_prec>(), stream0()));
}
void emplar_cu(int vers)
{
   apx_energy_result r;
   chk(apx_energy(g_apx, vers, &r));
   // fused term: em and ep are one buffer in the only mode emplar runs in (src/energy.cpp:262-270: !analyz)
   hand_back(vers, r, r.em + r.ep, em, vir_em, demx, demy, demz);
}
void empoleEwaldRealSelf_cu(int vers)
{
   apx_energy_result r;
   chk(apx_empole(g_apx, vers, &r));
   hand_back(vers, r, r.em, em, vir_em, demx, demy, demz);
   if ((vers & calc::analyz) && nem)      // interaction count into slot 0 of the count buffer (countReduce sums it)
   {
      const double c1 = (double)r.nem;
      chk(apx_add_scalars_dev(g_apx, nem, &c1, 1, APX_DEV_I32, stream0()));
   }
}
void empoleNonEwald_cu(int vers) { empoleEwaldRealSelf_cu(vers); }
void empoleChgpenEwaldRecip_cu(int, int) {}      // contained in apx_empole
void epolarEwaldReal_cu(int vers, const real (*)[3], const real (*)[3])
{
   apx_energy_result r;
   chk(apx_epolar(g_apx, vers, &r));
   // without calc::analyz the front-end takes the energy from epolar0DotProd (epolar.cpp:651-655), which is part of apx_epolar:
   // handed back here either way, the dot-product stub below adds nothing
   hand_back(vers, r, r.ep, ep, vir_ep, depx, depy, depz);
   if ((vers & calc::analyz) && nep) {
      const double c1 = (double)r.nep;
      chk(apx_add_scalars_dev(g_apx, nep, &c1, 1, APX_DEV_I32, stream0()));
   }
}
void epolarNonEwald_cu(int vers, const real (*a)[3], const real (*b)[3]) { epolarEwaldReal_cu(vers, a, b); }
void epolarEwaldRecipSelf_cu(int, const real (*)[3], const real (*)[3]) {}      // contained in apx_epolar
void epolar0DotProd_cu(const real (*)[3], const real (*)[3]) {}                // contained in apx_epolar / apx_energy
void epolarPairwiseExtfield_cu(const real (*)[3]) {}                            // no external field in the library

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
