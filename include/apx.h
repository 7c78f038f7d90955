/* apx.h -- C ABI of the B200-native AMOEBA polarizable-electrostatics back end.
 *
 * Drop-in boundary.  The reference dispatches every operator of this
 * path through TINKER_FCALL2(acc, cu, F, args...) to a free function F_cu(...) that reads
 * process-global device arrays (include/tool/externfunc.h:10-77).  This library exposes
 * the same operators over an opaque context instead of globals; INTEGRATION.md shows
 * the ~200-line adapter TU that defines the reference's *_cu symbols on top of it.
 *
 * Each entry point cites the reference interface it replaces.  All pointers are plain
 * host pointers unless the name ends in _dev (device memory, see the block further down).  Arrays are in the CALLER's atom order
 * (the library keeps its own spatially sorted copies).  Return value: 0 on success,
 * non-zero on error with the message available from apx_last_error() -- the C-ABI
 * image of the reference's TINKER_THROW / FatalError (include/tool/error.h:16-45).
 *
 * There is no CPU fallback: every compute entry point fails if no CUDA device is usable.
 */
#ifndef APX_H
#define APX_H

#ifdef __cplusplus
extern "C" {
#endif

/* calc:: flags, include/tool/rcman.h:107-133 */
enum {
   APX_ENERGY = 0x010,
   APX_GRAD = 0x020,
   APX_VIRIAL = 0x040,
   APX_ANALYZ = 0x080,
   APX_V0 = 0x010,
   APX_V1 = 0x070,
   APX_V3 = 0x090,
   APX_V4 = 0x030,
   APX_V5 = 0x020,
   APX_V6 = 0x060
};

/* What mpoleData / epolarData / mdpuscaleData / pmeData upload
 * (src/elec.cpp:54-500, src/amoeba/epolar.cpp:25-511, src/pme.cpp:154-217). */
typedef struct apx_system {
   int n;
   const double* xyz;      /* [n][3]  atom.h x,y,z */
   double lvec[9];         /* box.h lvec1..3 (row major) */
   const double* pole;     /* [n][10] local-frame multipoles, MPL_PME order (mpole.h:6-15) */
   const int* zaxis;       /* [n][4]  LocalFrame {zaxis, xaxis, yaxis(signed, 1-based), polaxe} */
   const double* polarity; /* [n] */
   const double* thole;    /* [n] */
   const double* pdamp;    /* [n] */
   const int* jpolar;      /* [n] */
   int njpolar;
   const double* thlval;   /* [njpolar][njpolar] */
   int nmdpu;              /* fused exclusion list, pairs i<k */
   const int* mdpu_ik;     /* [nmdpu][2] */
   const double* mdpu_scale; /* [nmdpu][4]  m, d, p, u */
   int use_ewald, use_mpole, use_polar;
   int poltyp_mutual;      /* 1 MUTUAL, 0 DIRECT */
   double aewald;
   int nfft[3];
   int bsorder;            /* only 5 is built */
   double cutoff;          /* switchOff(EWALD) or switchOff(MPOLE) */
   double usolve_cutoff;   /* range the sparse preconditioner applies, <=0: diagonal */
   double list_buffer;
   double poleps;
   int politer;
   double uaccel;          /* udiag */
   int pcgprec, pcgguess;
   double pcgpeek;
   double electric, dielec;
   int polpred;            /* UPred (include/ff/amoeba/mpole.h:38): 0 NONE, 1 ASPC, 2 GEAR; 3 LSQR is rejected like
                              ulspredSum_cu does (src/cu/upredict.cu:204-206) */
} apx_system;

/* What evdwData(RcOp::ALLOC|INIT) uploads for the buffered 14-7 term, Vdw::HAL (src/evdw.cpp:62-470):
 * the "next" row after the electrostatics path. */
typedef struct apx_vdw {
   int n;
   const int* ired;          /* [n]  vdw::ired, 0-based: the atom a reduced hydrogen site hangs off (itself otherwise) */
   const double* kred;       /* [n]  vdw::kred */
   const int* jvdw;          /* [n]  compressed class index (src/evdw.cpp:176-188) */
   int njvdw;
   const double* radmin;     /* [njvdw][njvdw] */
   const double* epsilon;    /* [njvdw][njvdw] */
   int nvexclude;            /* pairs i<k whose vdW scale is not 1 (src/evdw.cpp:196-262) */
   const int* vexclude;      /* [nvexclude][2] */
   const double* vexclude_scale;
   double cutoff, taper;     /* switchOff / switchCut(Switch::VDW) */
   double ghal, dhal;
   double elrc_vol, vlrc_vol; /* long-range correction x volume (src/evdw.cpp:443-452) */
} apx_vdw;

/* What ebondData ... etortorData upload for the AMOEBA valence terms (src/bonded/ebond.cpp, eangle.cpp, estrbnd.cpp,
 * eurey.cpp, eopbend.cpp, etors.cpp, epitors.cpp, etortor.cpp): SURVEY.md 8f rank 3.  Atom indices are 0-based and
 * already resolved (the reference keeps indirections through iang / ibnd / ibitor).  Term order everywhere:
 * bond, angle, strbnd, urey, opbend, torsion, pitors, tortor. */
typedef struct apx_valence {
   int n;
   int nbond;
   const int* ibnd;          /* [nbond][2] */
   const double *bk, *bl;
   int nangle;
   const int* iang;          /* [nangle][4]; [3] = out-of-plane atom of an in-plane angle */
   const double *ak, *anat;  /* anat in degrees */
   const int* angtyp;        /* 0 HARMONIC, 1 IN-PLANE */
   int nstrbnd;
   const int* isb;           /* [nstrbnd][3] atoms a, b, c */
   const double* sbk;        /* [nstrbnd][2] */
   const double* sb_anat;    /* ideal angle of the parent angle */
   const double* sb_bl;      /* [nstrbnd][2] ideal a-b and c-b lengths */
   int nurey;
   const int* iury;          /* [nurey][3] */
   const double *uk, *ul;
   int nopbend;
   const int* iopb;          /* [nopbend][4] a, b (centre), c, d (out of plane) */
   const double* opbk;
   int opbtyp;               /* 0 W-D-C, 1 ALLINGER */
   int ntors;
   const int* itors;         /* [ntors][4] */
   const double* tors_v;     /* [ntors][6] amplitudes of folds 1..6 */
   const double* tors_phase; /* [ntors][6] phases in degrees */
   int npitors;
   const int* ipit;          /* [npitors][6] */
   const double* kpit;
   int ntortor;
   const int* itt;           /* [ntortor][5] atoms in table order */
   const int* tt_chk;        /* chirality probe atom (src/bonded/etortor.cpp:86-132) or -1 */
   const int* tt_grid;
   int ngrid;
   const int *tnx, *tny, *tt_off, *tt_xoff, *tt_yoff;
   const double *ttx, *tty, *tbf, *tbx, *tby, *tbxy;
   double consts[20];        /* bndunit cbnd qbnd angunit cang qang pang sang stbnunit ureyunit cury qury opbunit copb qopb popb
                                sopb torsunit ptorunit ttorunit */
   int use[8];               /* use_bond ... use_tortor */
} apx_valence;

typedef struct apx_valence_result {
   double e[8];              /* energy_eb, ea, eba, eub, eopb, et, ept, ett */
   int count[8];
   double esum;              /* energy_valence */
   double virial[9];         /* virial_valence */
} apx_valence_result;

typedef struct apx_ctx apx_ctx;

typedef struct apx_energy_result {
   double em, ep, esum;
   double virial[9];
   int nem, nep;
   int pcg_iterations;
   double pcg_eps;         /* final RMS residual in Debye */
   double ev;              /* energy_ev, 0 unless a vdW term is attached */
   int nev;
   double evalence;        /* energy_valence, 0 unless valence terms are attached (then part of esum and virial) */
   double eval_term[8];    /* eb, ea, eba, eub, eopb, et, ept, ett */
   int nval_term[8];
} apx_energy_result;

/* timing / counters of the most recent operator call, CUDA events on the library stream */
typedef struct apx_stats {
   float ms_induce, ms_energy, ms_list;
   float ms_ufield_real;   /* mean device time of the real-space ufield row kernel, last induce() */
   int pcg_iterations;
   int kernel_launches;    /* launches of this library's own kernels since apx_stats_reset */
   int list_rebuilds;
   long long nverlet;      /* directed entries of the Verlet rows (cutoff + buffer) at the last list build */
   long long npairs_m;     /* pairs inside the real-space cutoff at the last list build */
   long long npairs_u;     /* pairs inside the preconditioner range at the last list build */
   float ms_ehal;          /* device time of the vdW row kernel of the last evaluation */
   long long nverlet_vdw;  /* directed entries of the vdW Verlet rows */
   int energy_retries;     /* evaluations repeated because the solver's first, unawaited batch of iterations did not converge */
} apx_stats;

const char* apx_last_error(void);
const char* apx_version(void);            /* "apx <n> (float|double)" */
int apx_precision_bytes(void);            /* sizeof(real): 4 mixed build, 8 double build */

/* initialize()/finish(): src/rcman.cpp:20-80 (deviceData ALLOC|INIT / DEALLOC) */
int apx_create(const apx_system* sys, int device, apx_ctx** out);
void apx_destroy(apx_ctx* ctx);

/* ---- several GPUs of one node (no counterpart in the reference, which is single-GPU).
 * Every rank (one per GPU) creates a context over the SAME system and positions; the library splits
 * the box into z-slabs, exchanges halo dipoles and PME planes and reduces energies/forces, so every
 * entry point below is then COLLECTIVE: all ranks call it, in the same order, and all receive the
 * complete result.  transport "nccl": handle = the 128-byte id from apx_nccl_unique_id() of rank 0,
 * nccl_lib = path of the libnccl.so.2 the process uses (NULL: default search); the bulk exchanges (halo vectors, PME planes,
 * FFT transposes) and the solver's scalar all-reduces then run as the library's own kernels writing into the peers'
 * CUDA-IPC-mapped buffers over NVLink (APX_DIST_P2P=3, default; 2 = windowed push/pull, 0 = NCCL send/recv), NCCL keeps
 * the start-up handshake and the large force reduction.  transport "direct": the same peer-memory kernels with no NCCL at
 * all; handle = 16 bytes of job id shared by the ranks (processes of one node), nccl_lib = NULL.  transport "local":
 * handle = apx_local_hub_create(world); the ranks are host threads of one process sharing a GPU. */
int apx_create_dist(const apx_system* sys, int device, int rank, int world, const char* transport, const void* handle,
   const char* nccl_lib, apx_ctx** out);
int apx_nccl_unique_id(const char* nccl_lib, void* out128);
void* apx_local_hub_create(int world);
void apx_local_hub_destroy(void* hub);
int apx_get_dist_info(apx_ctx* ctx, int* info8 /* rank, world, a0, a1, halo atoms, planes, halo planes lo, hi */);
/* per-phase device time of the decomposed path since the previous call (ms): [0] halo exchanges of per-atom vectors,
 * [1] forward slab FFTs (plane reduction, 2-D FFTs, transpose, 1-D FFTs), [2] inverse slab FFTs, [3] scalar all-reduces,
 * [4..7] their call counts.  on: keep collecting.  Call after apx_synchronize. */
int apx_dist_profile(apx_ctx* ctx, int on, double* out8);
/* host-only self-test of the /dev/shm rendezvous that carries the CUDA IPC handles of transport "direct" at start-up: `rounds`
 * all-gathers of `bytes` bytes (round r sends mine[q] + r); out [world][bytes] = the last round.  0 on success. */
int apx_rendezvous_selftest(const void* job_id16, int rank, int world, const void* mine, int bytes, int rounds, void* out);
/* host-only: halo plan of `rank` from the sorted atoms' PME z-coordinates and the ranks' sorted ranges */
int apx_dist_plan(int n, const float* w3_sorted, const int* bounds, int world, int rank, double range_frac, int* send_idx,
   int* send_off, int* recv_idx, int* recv_off);

/* copyPosToXyz + nblistRefresh: src/nblist.cpp:521-531, spatialCheck_cu (src/cu/spatial.cu:943-960) */
int apx_set_positions(apx_ctx* ctx, const double* xyz);
/* box change (Monte-Carlo barostat): src/box.cpp boxSetCurrent */
int apx_set_box(apx_ctx* ctx, const double lvec[9]);

/* mpoleInit: chkpole_cu, rotpole_cu, rpoleToCmp_cu (src/amoeba/mpole.cpp:28-58) */
int apx_mpole_init(apx_ctx* ctx);
int apx_get_rpole(apx_ctx* ctx, double* rpole /* [n][10] */);

/* dfield(field, fieldp): src/amoeba/field.cpp:56-64 */
int apx_dfield(apx_ctx* ctx, double* field, double* fieldp);
/* ufield(uind, uinp, field, fieldp): src/amoeba/field.cpp:111-117 */
int apx_ufield(apx_ctx* ctx, const double* uind, const double* uinp, double* field, double* fieldp);
/* sparsePrecondApply / diagPrecond: src/amoeba/induce.cpp:12-25 */
int apx_precond(apx_ctx* ctx, const double* rsd, const double* rsdp, double* zrsd, double* zrsdp);

/* induce(uind, uinp) -> induceMutualPcg1_cu: src/amoeba/induce.cpp:108-113, src/cu/amoeba/pcg.cu:14 */
int apx_induce(apx_ctx* ctx);
int apx_get_uind(apx_ctx* ctx, double* uind, double* uinp);
/* induced-dipole predictor, ulspredSave / ulspredSum (src/amoeba/induce.cpp:27-69, src/cu/upredict.cu):
 * every apx_induce() stores its solution in a ring of maxualt (ASPC 16, GEAR 6) entries and, once the ring
 * is full, starts from the extrapolated dipoles instead of the direct guess.  apx_upred_set switches the
 * predictor kind and empties the ring (nualt = 0, as epolarData(RcOp::INIT) does, src/amoeba/epolar.cpp:446-447). */
int apx_upred_set(apx_ctx* ctx, int polpred);
int apx_upred_count(apx_ctx* ctx, int* nualt, int* maxualt);
int apx_get_udir(apx_ctx* ctx, double* udir, double* udirp);

/* energy(vers) restricted to the electrostatic terms: empole+epolar or fused emplar
 * (src/energy.cpp:87-114,262-270,319-448); includes torque() and the fixed-point reductions. */
int apx_energy(apx_ctx* ctx, int vers, apx_energy_result* out);
/* empole(vers), epolar(vers): src/amoeba/empole.cpp:81-138, src/amoeba/epolar.cpp:574-649 */
int apx_empole(apx_ctx* ctx, int vers, apx_energy_result* out);
int apx_epolar(apx_ctx* ctx, int vers, apx_energy_result* out);
/* evdwData(ALLOC|INIT): src/evdw.cpp:62-470.  Once attached, apx_energy() also evaluates the vdW term -- on its own
 * stream, beside the induced-dipole solver -- and esum / virial / gradient include it (energy(vers), src/energy.cpp:319-448). */
int apx_vdw_attach(apx_ctx* ctx, const apx_vdw* vdw);
/* evdw(vers) -> ehal_cu: src/evdw.cpp:472-530, src/cu/ehal.cu:125-160 (vdW alone: ev, nev, virial, gradient) */
int apx_evdw(apx_ctx* ctx, int vers, apx_energy_result* out);
/* ebondData ... etortorData (src/bonded/*.cpp).  Once attached, apx_energy() also evaluates the valence terms -- one
 * fused launch on its own stream (evalence_cu1, src/cu/evalence.cu) -- and esum / virial / gradient include them. */
int apx_valence_attach(apx_ctx* ctx, const apx_valence* val);
/* the valence terms alone: per-term energies and counts, virial_valence (energy(vers) with only bonded terms active,
 * test/bond.cpp ... test/tortor.cpp); gradient through apx_get_valence_gradient */
int apx_evalence(apx_ctx* ctx, int vers, apx_valence_result* out);
int apx_get_valence_gradient(apx_ctx* ctx, double* grad /* [n][3] */);

/* ---- DYNAMIC: velocity Verlet / r-RESPA + Bussi thermostat on the device (src/md/integrator.cpp:70-222,
 * src/md/propagator.cpp:170-187, src/mdpt.cpp:41-71).  Positions and velocities stay in HBM between steps. */
typedef struct apx_md_config {
   double dt;              /* outer time step in ps */
   int nrespa;             /* mdstuf::nrespa: 1 = velocity Verlet, >1 = r-RESPA with the valence terms on the inner level */
   int thermostat;         /* 0 none (NVE), 1 Bussi (ThermostatEnum::BUSSI) */
   double kelvin, tautemp; /* bath::kelvin, bath::tautemp (ps) */
   int nfree;              /* mdstuf::nfree; <= 0: 3n - 3 */
   unsigned long long seed;
} apx_md_config;
typedef struct apx_md_report {
   double epot, ekin, temp;   /* after the last step: potential, kinetic energy (kcal/mol), temperature (K) */
   double e_valence, e_nonbonded;
   double last_scale;         /* velocity scale the thermostat applied in the last step */
   int steps, pcg_iterations, list_rebuilds;
   long long total_steps;
   float ms_device;           /* device time of these steps (CUDA events on the library stream) */
} apx_md_report;
/* mdData + integrator kick-off: masses, starting velocities (NULL: at rest), gradients at the current positions */
int apx_md_init(apx_ctx* ctx, const double* mass, const double* vel, const apx_md_config* cfg);
/* nsteps x BasicIntegrator::dynamic(istep, dt) */
int apx_md_steps(apx_ctx* ctx, int nsteps, apx_md_report* out);
int apx_md_get_state(apx_ctx* ctx, double* xyz /* [n][3] or NULL */, double* vel /* [n][3] or NULL */);
/* positions / velocities from host buffers (a host-side driver that owns x and v; pinned memory makes the copies true DMA).
 * forces_valid != 0: xyz is what the last apx_md_steps left, the saved forces still apply; 0: they are recomputed. */
int apx_md_set_state(apx_ctx* ctx, const double* xyz, const double* vel, int forces_valid);

/* copyGradient: src/egvop.cpp:64-111 (fixed -> double, caller's order) */
int apx_get_gradient(apx_ctx* ctx, double* grad /* [n][3] */);

/* ---- device-pointer entry points: what the reference's *_cu operators are handed (csrc/devio.cu).
 * Arrays are DEVICE memory in the caller's atom order.  elem_bytes: 4 = float, 8 = double (the reference's `real`,
 * include/ff/precision.h).  `stream` is the caller's cudaStream_t (NULL = the legacy default stream; the reference's g::s0):
 * the call is ordered after the work already enqueued there and its results are visible to what is enqueued there next.
 * Nothing is staged through host memory; outputs that the reference ACCUMULATES into (gradient, energy and virial buffers:
 * SURVEY.md 8b "Ownership", src/energy.cpp:333-446) are added to, never assigned.  Single-GPU contexts only. */
enum { APX_DEV_FIXED = 0, APX_DEV_I32 = 1, APX_DEV_F32 = 4, APX_DEV_F64 = 8 }; /* 2^32 fixed point in unsigned long long (grad_prec / energy
                                                                 buffers of the mixed build, include/ff/precision.h:68-106) */
/* copyPosToXyz + nblistRefresh from the reference's x, y, z device arrays (include/ff/atom.h:39-45) */
int apx_set_positions_dev(apx_ctx* ctx, const void* x, const void* y, const void* z, int elem_bytes, void* stream);
/* dfieldEwaldRecipSelfP2_cu + dfieldEwaldReal_cu / dfieldNonEwald_cu: src/amoeba/field.cpp:8-63 */
int apx_dfield_dev(apx_ctx* ctx, void* field, void* fieldp, int elem_bytes, void* stream);
/* ufieldEwaldRecipSelfP1_cu + ufieldEwaldReal_cu / ufieldNonEwald_cu: src/amoeba/field.cpp:67-117 */
int apx_ufield_dev(apx_ctx* ctx, const void* uind, const void* uinp, void* field, void* fieldp, int elem_bytes, void* stream);
/* sparsePrecondApply_cu / diagPrecond_cu: src/amoeba/induce.cpp:12-25 */
int apx_precond_dev(apx_ctx* ctx, const void* rsd, const void* rsdp, void* zrsd, void* zrsdp, int elem_bytes, void* stream);
/* induceMutualPcg1_cu(uind, uinp): src/amoeba/induce.cpp:73; udir / udirp may be NULL */
int apx_induce_dev(apx_ctx* ctx, void* uind, void* uinp, void* udir, void* udirp, int elem_bytes, void* stream);
int apx_get_uind_dev(apx_ctx* ctx, void* uind, void* uinp, void* udir, void* udirp, int elem_bytes, void* stream);
/* g[i] += dE/dx_i of the last energy / empole / epolar / evdw call: what emplar_cu, epolar*_cu, ehal_cu do to
 * gx_elec / gx_vdw (src/amoeba/emplar.cpp:10-28, src/energy.cpp:443-444); kind = APX_DEV_* */
int apx_add_gradient_dev(apx_ctx* ctx, void* gx, void* gy, void* gz, int kind, void* stream);
/* dst[q] += vals[q], q < count <= 16: one slot of an energy buffer (count 1), of a virial buffer (count 6: xx yx zx yy zy
 * zz, src/energybuffer.cpp:81-94) or of a count buffer (APX_DEV_I32), which energyReduce / virialReduce / countReduce then
 * sum (src/energy.cpp:345-348,371-374) */
int apx_add_scalars_dev(apx_ctx* ctx, void* dst, const double* vals, int count, int kind, void* stream);

/* PME operators exposed for parity tests: gridMpole/gridUind + fftfront + pmeConv + fftback
 * + fphiMpole/fphiUind2 (src/pme.cpp:229-351).  Output in fractional coordinates. */
int apx_pme_mpole_fphi(apx_ctx* ctx, double* fphi /* [n][20] */);
int apx_pme_uind_fphi(apx_ctx* ctx, const double* uind, const double* uinp, double* fdip_phi1 /* [n][10] */,
   double* fdip_phi2 /* [n][10] */);

/* PME convolution alone: grid <- IFFT(influence * FFT(grid)), grid = [nfft3][nfft2][nfft1] complex
   (re,im interleaved), unnormalised like fftfront/pmeConv/fftback of src/pme.cpp:229-351.
   apx_set_native_fft(ctx, 0) forces cuFFT instead of the fused 64^3 kernels (fft64.cu). */
int apx_pme_convolve_grid(apx_ctx* ctx, const double* grid_in, double* grid_out);
int apx_set_native_fft(apx_ctx* ctx, int on);
/* Deterministic PME spreading: charge / dipole contributions are summed onto the grid as 2^32 fixed-point integers (64-bit
 * integer reductions, order independent) instead of float reductions; results are then bit-reproducible from run to run.
 * Off by default (the float vector reductions are ~3x cheaper; the reference's spreading, src/cu/pme.cu:14-255, is float
 * atomics as well).  APX_PME_FIXED=1 in the environment turns it on for every context. */
int apx_set_pme_fixed_point(apx_ctx* ctx, int on);

int apx_get_stats(apx_ctx* ctx, apx_stats* out);
int apx_stats_reset(apx_ctx* ctx);
/* cudaStream_t the library launches on (for callers that time with their own events) */
void* apx_stream(apx_ctx* ctx);
int apx_synchronize(apx_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
