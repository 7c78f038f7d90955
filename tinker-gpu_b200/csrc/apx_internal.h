// Internal declarations shared by the CUDA translation units of libapx.
// Layout of everything that lives in HBM is described in DESIGN.md §3.
#pragma once
#include "apx.h"
#include <cuda_runtime.h>
#include <cufft.h>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <functional>
#include <map>
#include <vector>

#ifdef APX_DOUBLE
typedef double real;
typedef double2 real2;
typedef double4 real4;
typedef cufftDoubleComplex cplx;
#define APX_PREC_NAME "double"
#else
typedef float real;
typedef float2 real2;
typedef float4 real4;
typedef cufftComplex cplx;
#define APX_PREC_NAME "float"
#endif

typedef unsigned long long fixed_t;   // 2^32 fixed point, as include/ff/precision.h:68-106
#define APX_FIXED_SCALE 4294967296.0  // 0x100000000

struct real3 {
   real x, y, z;
};

// Sorted per-step positions as the pair kernels read them.  Mixed build: 32-bit FRACTIONAL coordinates (resolution
// L / 2^32 = 1.4e-8 A at 62 A) + the Thole damping length as float bits: the separation of a pair is an integer
// subtraction that wraps (= the minimum image, exactly), converted to float AFTER the subtraction -- 2e-7 A for a 3 A pair
// where wrapped float coordinates lose 4e-6 A each (tests/test_pairmath_host.py measures what that
// is worth in force error).  Double build: wrapped Cartesian coordinates + pdamp, as posd.
#ifdef APX_DOUBLE
typedef real4 pos_t;
#else
typedef uint4 pos_t;
#endif

#define APX_WARP 32
#define APX_BLOCK 128

struct ApxError : std::runtime_error {
   using std::runtime_error::runtime_error;
};

void apx_throw(const char* file, int line, const std::string& msg);
#define APX_THROW(msg) apx_throw(__FILE__, __LINE__, (msg))
#define CUDA_CHECK(expr)                                                                           \
   do {                                                                                            \
      cudaError_t e__ = (expr);                                                                    \
      if (e__ != cudaSuccess)                                                                      \
         apx_throw(__FILE__, __LINE__, std::string(#expr) + ": " + cudaGetErrorString(e__));       \
   } while (0)
#define CUFFT_CHECK(expr)                                                                          \
   do {                                                                                            \
      cufftResult r__ = (expr);                                                                    \
      if (r__ != CUFFT_SUCCESS)                                                                    \
         apx_throw(__FILE__, __LINE__, std::string(#expr) + ": cufft error " + std::to_string(r__)); \
   } while (0)

template <class T>
struct DevBuf {
   T* p = nullptr;
   size_t cap = 0;
   void ensure(size_t n)
   {
      if (n <= cap)
         return;
      if (p)
         cudaFree(p);
      p = nullptr;
      size_t want = n + n / 8 + 64;
      CUDA_CHECK(cudaMalloc(&p, want * sizeof(T)));
      cap = want;
   }
   void release()
   {
      if (p)
         cudaFree(p);
      p = nullptr;
      cap = 0;
   }
   operator T*() const { return p; }
};

// Box: orthogonal fast path + general triclinic image via reciprocal vectors.
struct Box {
   real lx, ly, lz;         // orthogonal edge lengths
   real ilx, ily, ilz;
   real l[9];               // lvec rows
   real r[9];               // recip rows
   real q[9];               // l / 2^32: Cartesian separation from a difference of 32-bit fractional coordinates
   real qlo[9];             // l / 2^32 - q in double: the product (difference x l / 2^32) is formed from the two parts
   int orthogonal;
   real volume;
};

// The cell in double: wrapping of the caller's f64 coordinates into fractional coordinates must not see the float-rounded
// reciprocal vectors of Box (1/L in float is off by up to 6e-8 relative: at |x| = 15 A that moves an atom by 1e-6 A, and
// inconsistently with the listed pairs, which are formed from the f64 coordinates directly)
struct BoxD {
   double l[9], r[9];
};

// Directed per-atom neighbor rows in sorted order (rows.cu).  vnbr holds the Verlet rows
// (r <= cutoff + buffer at build time, both directions, k ascending); every step nbr receives,
// at the same row offsets, the atoms with r <= ucut first (count cntu) and then those with
// ucut < r <= cutoff (row length cnt).  No pair is shared between two rows, so the row kernels
// need no atomics and their sums are order-deterministic.
struct RowList {
   DevBuf<int> vstart;      // [n+1] row offsets
   DevBuf<int> vcnt;        // [n+1] Verlet row lengths (scan input)
   DevBuf<int> vnbr;        // Verlet rows
   DevBuf<int> nbr;         // per-step compacted rows
   DevBuf<int> cnt, cntu;   // [n]
   DevBuf<real4> sctr, sext;  // bounding boxes of super-blocks (32 blocks)
   DevBuf<unsigned long long> total;   // [2]: directed pairs within cutoff / ucut at the last compaction
   long long nverlet = 0;
   // one-pass rebuild (rows.cu): row lengths of the previous build in CALLER order give every row a padded slot, so the
   // search runs once (fill + count) and a bandwidth-bound copy packs the rows; overflow falls back to an exact fill
   DevBuf<int> prev_o;      // [n] row length at the last build, caller order
   DevBuf<int> capstart;    // [n+1] offsets of the padded slots
   DevBuf<int> vpad;        // padded rows
   DevBuf<int> oflow;       // [1]
   long long prev_total = 0;
   int have_prev = 0;
};

// Groups of 64 sorted atoms with their j-blocks and 16-bit slot rows (staged.cu)
struct GroupList {
   int ok = 0, ngrp = 0;
   DevBuf<int> vjb, nvjb;                // [ngrp][SG_VJB_CAP] Verlet j-blocks (ascending), their number
   DevBuf<int> ajb, najb;                // the blocks still holding a partner inside the cutoff this step
   DevBuf<unsigned short> vslot, nbr16;  // Verlet rows / per-step rows in slot form, same offsets as RowList::vnbr / nbr
   DevBuf<int> oflow;
};

struct PairExcl {           // exclusion pair in SORTED indices with (scale-1) factors
   int i, k;
   real m, d, p, u;
};

// ---- multi-GPU spatial decomposition (dist.cu).  Every rank keeps the whole atom
// set in the same sorted order ("replicated data") but owns the contiguous sorted range [a0,a1) of
// one z-slab and the matching planes of the PME grid ("partitioned work").
struct ApxComm;
struct DistState {
   int on = 0, rank = 0, world = 1;
   ApxComm* comm = nullptr;
   std::vector<int> bounds;              // [world+1] first sorted index of every rank's slab
   // halo of per-atom vectors: atoms of other ranks within list range of my slab (and vice versa)
   DevBuf<int> send_idx, recv_idx;       // concatenated by peer
   std::vector<int> send_off, recv_off;  // [world+1]
   DevBuf<real4> sendbuf, recvbuf;       // 2 real4 per atom (packed (d,p) pair)
   // PME slab decomposition: planes [z0, z0+pz) owned, hl / hu halo planes below / above
   int pz = 0, z0 = 0, hl = 0, hu = 0, py = 0, y0 = 0;
   cufftHandle plan2d = 0, plan1d = 0;
   int plans_ok = 0;
   DevBuf<cplx> tbuf, sbuf, tbuf2, hbuf;  // transposed grid [n3][py][n1], pack buffer, cross-virial copy, halo planes
   long long halo_atoms = 0;
   // per-phase device time of the decomposed path (apx_dist_profile): event pairs on the main stream around every halo
   // exchange (kind 0), forward slab FFT with its plane reduction and transpose (1), inverse (2), scalar all-reduce (3)
   int prof_on = 0, prof_used = 0;
   std::vector<cudaEvent_t> prof_ev;
   std::vector<int> prof_kind;
};

// ---- buffered 14-7 vdW term (ehal.cu)
struct VdwExcl {            // pair in SORTED indices whose scale is neither 0 nor 1: (scale - 1) correction
   int i, k;
   real s;
};
struct VdwState {
   int on = 0, nj = 0;
   DevBuf<int> ired_o, jvdw_o;           // caller order
   DevBuf<real> kred_o;
   DevBuf<real2> tab;                    // [nj*nj] {1/radmin, epsilon}
   DevBuf<int> exoff, exlist;            // CSR (caller indices) of partners with scale 0: never listed
   DevBuf<int> xs_ik;                    // pairs with another scale, caller order
   DevBuf<real> xs_sc;
   DevBuf<VdwExcl> xs_s;
   int nxs = 0;
   real exrange = 0;                     // no excluded pair is farther apart than this (bonded topology)
   DevBuf<real4> pred;                   // sorted reduced sites {x,y,z wrapped, jvdw bits}
   DevBuf<int> ired_s;                   // sorted slot of the parent atom
   DevBuf<real> kred_s;
   DevBuf<real4> ctr, ext;               // block boxes of the reduced sites
   RowList rows;
   DevBuf<fixed_t> vbuf;                 // [0] ev, [1..6] virial xx yx zx yy zy zz
   DevBuf<int> vcnt;                     // nev
   fixed_t h_vbuf[8] = {0};              // host copies, valid after the main stream is synchronised
   int h_vcnt[2] = {0, 0};
   real cut = 0, off = 0, ghal = 0, dhal = 0;
   double elrc_vol = 0, vlrc_vol = 0;
   cudaStream_t stream = nullptr;        // ehal runs here, beside induce() on the main stream
   cudaEvent_t ev_go = nullptr, ev_done = nullptr, t0 = nullptr, t1 = nullptr;
};

struct ValState;           // evalence.cu
struct MdState;            // md.cu

struct apx_ctx {
   int device = 0;
   VdwState vdw;
   ValState* val = nullptr;              // valence terms, when attached
   MdState* md = nullptr;                // integrator state, after apx_md_init
   DistState dist;
   int a0 = 0, a1 = 0;                   // owned sorted range (single GPU: [0, n))
   int zbase = 0, nzl = 0;               // local PME planes: global plane zg lives at (zg - zbase) mod nfft3, nzl planes held
   int qy0 = 0, qny = 0;                 // influence function / transformed grid cover k2 in [qy0, qy0+qny) (single GPU: all)
   cudaStream_t stream = nullptr;
   int n = 0, nblk = 0, npad = 0;
   int sm_count = 148;
   apx_system opt;                       // scalar options (pointers invalid after create)
   real f_elec = 0;
   Box box;

   // ---- caller-order static data
   DevBuf<double> xyz_d;                 // [n][3] positions as given (f64)
   DevBuf<double> xyz_ref;               // positions at the last list build
   DevBuf<int> zaxis;                    // [n][4]
   DevBuf<real> pole;                    // [n][10] local frame (chkpole may flip signs)
   DevBuf<real> polarity_o, thole_o, pdamp_o;
   DevBuf<int> jpolar_o;
   DevBuf<real> thlval;
   int thole_table = 0;                  // 1: per-pair lookup needed (polpair present)
   DevBuf<int> excl_ik;                  // [nx][2] caller order
   DevBuf<real> excl_sc;                 // [nx][4]
   DevBuf<double> excl_sc_d;             // [nx][4] the same scale factors in double, for the listed-pair pass of mplar.cu
   double recip_d[9] = {0};              // reciprocal cell vectors in double (rows), set with the box
   int nexcl = 0;
   int nexcl_u = 0;                      // exclusions whose u-scale != 1 (none in stock AMOEBA)

   // ---- sorted-order data (rebuilt with the list)
   DevBuf<int> perm, inv;                // perm[s] = caller index, inv[i] = sorted slot
   DevBuf<unsigned> sortkey, sortkey2;
   DevBuf<real> w3;                      // PME z-coordinate of every atom (caller order), slab decomposition only
   DevBuf<int> permtmp;
   DevBuf<char> cubtmp;
   size_t cubtmp_bytes = 0;
   DevBuf<real4> posd;                   // {x,y,z wrapped, pdamp}: list build / compaction, block boxes
   DevBuf<uint4> posq_buf;               // mixed build: pos_t array (32-bit fractional coordinates + pdamp bits)
   pos_t* posq = nullptr;                // what the pair kernels and the spline tables read (double build: = posd)
   DevBuf<real4> tpj;                    // {thole, polarity, 1/polarity, jpolar bits}
   DevBuf<real4> mp0, mp1;               // rpole {c,dx,dy,dz}, {qxx,qxy,qxz,qyy}
   DevBuf<real2> mp2;                    // {qyz,qzz}
   DevBuf<real4> blk_ctr, blk_ext;       // block bounding boxes
   DevBuf<PairExcl> excl_s;              // exclusions in sorted indices
   real list_cutoff = 0, list_buffer = 0;
   RowList rows;
   GroupList grp;                        // staged real-space operator (staged.cu)
   int staged_on = 0;                    // APX_STAGED=1: staged operator (measured slower than the rows, profiles/r02d_*; kept for A/B)
   // ---- stored-tensor real-space operator (tlist.cu): one real4 per directed pair inside the cutoff, written once per induce()
   DevBuf<real4> tl_T;                   // same offsets as rows.nbr: {B1 (LSB = sign of B2), sqrt|B2| R}
   DevBuf<real4> tl_P;                   // preconditioner tensors of the first cntu entries of every row
   int tlist_on = 1;                     // APX_TLIST=0: recompute the pair geometry in every operator application (field.cu)
   int tl_valid = 0;                     // tl_T belongs to the current positions and rows
   int tl_p_valid = 0;                   // ... and so does tl_P (written by the permanent-field rows only)
   int staged_cap = 64;                  // APX_STAGED_CAP: j-blocks of a group staged in shared memory (1.5 KB each)
   int staged_min_atoms = 0;             // APX_STAGED_MIN: systems smaller than this keep the row operator
   int list_valid = 0;
   DevBuf<int> flags;                    // device flags: [0] rebuild-needed, [1] pcg done, [2] iter ...
   int* flags_h = nullptr;               // pinned mirror

   // ---- fields / CG vectors, sorted order, [npad][3]
   DevBuf<real> field, fieldp, udir, udirp, uind, uinp;
   DevBuf<real> field_rs;                // real-space part of the permanent field while the PME part is computed beside it
   DevBuf<real> rsd, rsdp, zrsd, zrsdp, conj, conjp, vec, vecp;
   DevBuf<long long> qfix;      // deterministic spreading: fixed-point shadow of qgrid, two int64 per point (pme.cu)
   int pme_fixed = 0;
   DevBuf<real4> pk_p, pk_r, pk_z, pk_v, pk_f;   // packed (d,p) pairs, dp.cuh: direction, residual, M r, A p, real-space field
   DevBuf<real4> uf_rec;                 // [npad][3] interleaved neighbour records of the ufield rows (field.cu)
   int rows_onepass = 1;                 // APX_ROWS_ONEPASS: 1 = one-pass list rebuild with padded rows (default), 2 = the same with zero slack
                                         // (every grown row overflows: tests the fall-back), 0 = always count + fill
   int use_records = 1;                  // APX_NO_RECORDS=1: the separate-array kernel instead
   int uf_ctas = 8, uf_smem_kb = 0;      // residency of the ufield rows beside the PME chain (APX_UF_CTAS, APX_UF_SMEM)
   cudaStream_t stream2 = nullptr;       // real-space operator of an iteration runs here, beside the PME chain
   cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
   cudaEvent_t ev_dev_in = nullptr, ev_dev_out = nullptr;   // ordering of the _dev entry points against the caller's stream (devio.cu)
   DevBuf<double> scal;                  // PCG scalars (sum, sump, a, ap, sum1, sump1, epsd, epsp ...)
   double* scal_h = nullptr;             // pinned
   int last_iters = 6;

   // ---- induced-dipole predictor (pcg.cu): ring of packed (d,p) solutions in CALLER order, so a list
   //      rebuild (new sorted order) or a change of slab ownership does not disturb the history
   DevBuf<real4> upred_hist;             // [maxualt][n][2]
   int nualt = 0, maxualt = 0;

   // ---- PME
   int nfft1 = 0, nfft2 = 0, nfft3 = 0;
   cufftHandle plan = 0;
   int plan_ok = 0;
   int native_fft = 1;                   // use fft64.cu when the grid is 64^3 (mixed build); 0 = always cuFFT
   DevBuf<cplx> qgrid;
   DevBuf<real> qfac;                    // influence function (expterm of pmeConv)
   DevBuf<real> bsmod1, bsmod2, bsmod3;
   DevBuf<real4> theta;                  // [n][16] per-step spline tables + stencil origins (pme.cu)
   DevBuf<real> fphi;                    // [n][20] permanent, sorted order
   DevBuf<real> fmp;                     // [n][10]
   DevBuf<real> fphid, fphip, fphidp;    // [n][10],[n][10],[n][20]
   int mpole_pme_valid = 0;              // fphi/fmp correspond to current positions
   double recip_e_h = 0;

   // ---- energy / gradient accumulators (sorted order)
   DevBuf<fixed_t> gx, gy, gz;
   DevBuf<real> trq;                     // [npad][3] torques as real (input of the torque kernel)
   DevBuf<fixed_t> trqf;                 // [npad][3] fixed-point torque accumulators
   DevBuf<real4> mpx_a, mpx_b;           // scratch multipole heads for the cross virial
   DevBuf<cplx> qgrid2;
   DevBuf<fixed_t> ebuf;                 // [0]=em [1]=ep ; [2..7] vir  (fixed point)
   DevBuf<double> dbuf;                  // double accumulators (recip energies, virials)
   DevBuf<int> cnt;                      // nem, nep

   // ---- io staging
   DevBuf<double> io_a, io_b, io_c, io_d;
   double* pin_a = nullptr;              // pinned staging for host<->device in e2e paths
   size_t pin_bytes = 0;

   // ---- accumulator arenas: the buffers below live back to back so one memset clears them
   //      energy(): gx gy gz trqf ebuf dbuf cnt ;  induce(): scal flags
   DevBuf<char> arena_e, arena_p;
   size_t arena_e_bytes = 0, arena_p_bytes = 0;

   // ---- CUDA graphs of PCG iteration batches (pcg.cu), rebuilt when pointers / box / options change
   struct PcgGraph {
      int it0, nit;
      cudaGraphExec_t exec;
      int launches;
   };
   std::vector<PcgGraph> graphs;
   // the PCG loop as ONE graph: a conditional WHILE node whose body is one iteration; the preconditioner kernel ends the loop
   // from the device when the stopping rule is met (pcg.cu)
   cudaGraphExec_t loop_exec = nullptr;
   int loop_launches = 0;          // kernels of ours per iteration of the body
   int loop_warm = 0;
   int use_loop = 0;               // APX_LOOP=1: the WHILE-node loop instead of generic iteration batches (same median step,
                                   // 3 % slower in batches of steps on this driver: profiles/r02j)
   int pcg_n = 0, pcg_n_slack = 0; // iterations in the first batch of a solve (largest count seen recently)
   int uf_ext_iter = 0;            // capture of a batch: the operator launch being enqueued gets external timing events (slots 2,3)
   int uf_iter_timed = 0;          // the batch graph launched by this solve carries them
   int vdw_fork_vers = -1;         // >= 0: energy() wants apx_induce_impl to fork the vdW stream after its prologue (mplar.cu)
   std::function<void()> epi_tail;     // md.cu: closing half-kick + thermostat, enqueued inside the energy epilogue's IF body (mplar.cu)
   int induce_copy_pending = 0;    // an unawaited solver batch left its flags on the device: apx_induce_copy_out
   int epi_tail_ran = 0;           // the last epilogue launch contained (or replayed a graph that contains) epi_tail
   int epend_deferred = 0;         // energy_once stage 1 -> 2: the solve of the pending evaluation was deferred
   int induce_iter_launched = 0;   // iterations enqueued by the deferred first batch
   int induce_pending = 0, induce_pending_predict = 0;      // a deferred solve awaits apx_induce_finish
   int use_graph = 1;
   int capturing = 0;
   // ---- CUDA graphs of the fixed launch sequences around the solver (pcg.cu: apx_graph_begin/end): mpoleInit + zeroing,
   //      the permanent-field / first-residual prologue of induce(), the energy/force epilogue.  Keyed by what varies.
   struct StepGraph {
      cudaGraphExec_t exec = nullptr;
      int launches = 0, warm = 0;
      int conditional = 0;      // the region is the body of an IF node keyed on a device flag (apx_graph_begin with cond_flag)
   };
   int cond_nodes_ok = -1;               // conditional graph nodes on this driver: -1 not tried yet, 1 work, 0 unavailable
   cudaGraph_t cond_outer = nullptr;     // graph under construction that owns the IF node whose body is being captured
   std::map<int, StepGraph> step_graphs;
   int graph_key_open = -1, graph_launches_before = 0;
   unsigned char* red_h = nullptr;       // pinned: one copy brings every reduced scalar of a step to the host

   // ---- stats
   apx_stats stats;
   cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
   std::vector<cudaEvent_t> uf_ev;       // event pairs around the real-space ufield launches of one induce()
   int uf_used = 0;
   const int* skip = nullptr;            // device flag: kernels of speculative CG iterations return at once when set
   int mpole_inited = 0;
   int induced_valid = 0;
   int diag_skip = 0;                    // APX_DIAG_SKIP (diagnostics only): 1 = no real-space pair kernels in energy(), 2 = no reciprocal force kernels
   int md_forces_valid = 0;              // the accumulators hold the integrator's saved fast + slow forces at the current positions:
                                         // cleared by every public call that rewrites them (md.cu recomputes them when clear)
};

inline BoxD apx_box_d(const apx_ctx* c)
{
   BoxD b;
   for (int q = 0; q < 9; ++q)
      b.l[q] = c->opt.lvec[q], b.r[q] = c->recip_d[q];
   return b;
}

#define APX_COUNT_LAUNCH(ctx) ((ctx)->stats.kernel_launches++)

// ---- nblist.cu
void apx_list_refresh(apx_ctx* c, bool force, int known_moved = -1);
void apx_list_check_enqueue(apx_ctx* c, cudaStream_t st, double* seq = nullptr);
void apx_update_sorted_positions(apx_ctx* c);
// ---- dist.cu
void apx_induce_copy_out(apx_ctx* c);                              // flags of an unawaited solver batch -> pinned memory
void apx_pme_fixed_setup(apx_ctx* c);                              // shadow grid of the deterministic spreading mode
void apx_dist_after_sort(apx_ctx* c);                              // ownership bounds + halo plan (at list rebuild)
void apx_dist_halo(apx_ctx* c, real4* V, cudaStream_t st);         // V[halo atoms] <- owners' values
void apx_dist_allreduce_f64(apx_ctx* c, double* p, size_t n);
void apx_dist_allreduce_u64(apx_ctx* c, unsigned long long* p, size_t n);
void apx_dist_allreduce_i32(apx_ctx* c, int* p, size_t n);
void apx_dist_share_owned(apx_ctx* c, void* base, size_t bytes_per_atom);   // every rank receives the owned ranges of the others
void apx_dist_pme_setup(apx_ctx* c);
void apx_dist_pme_destroy(apx_ctx* c);
void apx_dist_fft_forward(apx_ctx* c, cplx* tb);                   // local grid (halo-reduced) -> transformed slab tb
void apx_dist_fft_inverse(apx_ctx* c, cplx* tb);                   // tb -> local grid including halo planes
void apx_dist_destroy(apx_ctx* c);
int apx_dist_prof_begin(apx_ctx* c, int kind, cudaStream_t st);      // -1 when profiling is off
void apx_dist_prof_end(apx_ctx* c, int slot, cudaStream_t st);
// ---- ehal.cu
void apx_vdw_attach_impl(apx_ctx* c, const apx_vdw* v);
void apx_vdw_refresh(apx_ctx* c, bool rebuilt);                    // reduced sites (every step), rows (at list rebuild)
void apx_vdw_launch(apx_ctx* c, int vers);                         // enqueue ehal on the vdW stream (forked from the main stream)
void apx_vdw_join(apx_ctx* c, bool copies = true);                 // main stream waits for it (+ its scalars -> pinned memory)
void apx_vdw_copy_out(apx_ctx* c);
void apx_vdw_collect(apx_ctx* c, int vers, apx_energy_result* r);  // after the main stream is synchronised: ev, nev, virial into r
void apx_vdw_destroy(apx_ctx* c);
void apx_block_boxes(apx_ctx* c, const real4* pos, real4* ctr, real4* ext);
// ---- evalence.cu
void apx_valence_attach_impl(apx_ctx* c, const apx_valence* v);
bool apx_valence_on(const apx_ctx* c);
void apx_valence_enqueue(apx_ctx* c, int vers, cudaStream_t st, bool zero_grad);   // zero + one fused launch on st
void apx_valence_launch(apx_ctx* c, int vers);                     // on its own stream, forked from the main stream
void apx_valence_join(apx_ctx* c);                                 // main stream waits, scalars go to pinned memory
void apx_valence_fetch(apx_ctx* c, cudaStream_t st);
void apx_valence_collect(apx_ctx* c, int vers, apx_valence_result* r);   // after the stream is synchronised
void apx_valence_set_in_total(apx_ctx* c, int on);
bool apx_valence_in_total(const apx_ctx* c);
fixed_t* apx_valence_grad_buffer(apx_ctx* c);                      // [3][n] caller order
void apx_valence_grad_out(apx_ctx* c, double* dev_out, bool accumulate);
void apx_valence_destroy(apx_ctx* c);
// ---- md.cu
void apx_md_init_impl(apx_ctx* c, const double* mass, const double* vel, const apx_md_config* cfg);
void apx_md_steps_impl(apx_ctx* c, int nsteps, apx_md_report* out);
void apx_md_get_state_impl(apx_ctx* c, double* xyz, double* vel);
void apx_md_set_state_impl(apx_ctx* c, const double* xyz, const double* vel, int forces_valid);
void apx_md_destroy(apx_ctx* c);
// ---- rows.cu
void apx_rows_build(apx_ctx* c);      // Verlet rows, after the spatial sort
void apx_rows_build_on(apx_ctx* c, RowList& L, const real4* pos, const real4* bctr, const real4* bext, real range, const int* exoff,
   const int* exlist, real exrange, bool want_compact);
void apx_rows_compact(apx_ctx* c, bool count);    // per-step compaction to r <= cutoff
// ---- staged.cu
bool apx_staged_usable(const apx_ctx* c);
void apx_group_build(apx_ctx* c);                                 // after apx_rows_build
void apx_rows_compact_grouped(apx_ctx* c, bool count);            // replaces apx_rows_compact when the groups are usable
void apx_ufield_staged(apx_ctx* c, cudaStream_t st, real4* F);    // F = real-space field of the records in c->uf_rec
// tlist.cu
bool apx_tlist_usable(const apx_ctx* c);
void apx_tlist_reserve(apx_ctx* c);                               // after apx_rows_build: buffers sized to the Verlet rows
void apx_tlist_build(apx_ctx* c, cudaStream_t st);                // tensors of the current rows and positions
void apx_ufield_tlist(apx_ctx* c, cudaStream_t st, const real4* U, real4* F);
// ---- frames.cu
void apx_rotpole(apx_ctx* c);
void apx_torque(apx_ctx* c, bool do_v);
// ---- pme.cu
void apx_pme_setup(apx_ctx* c);
void apx_pme_destroy(apx_ctx* c);
void apx_pme_fill_theta(apx_ctx* c);                             // after every change of posd
void apx_pme_mpole(apx_ctx* c, bool want_ev);                   // fills fmp, fphi (and recip E/virial in dbuf)
void apx_pme_zero_grid(apx_ctx* c);
void apx_pme_spread_dp(apx_ctx* c, const real4* U);              // grid += spread of a packed dipole pair
void apx_pme_convolve(apx_ctx* c);                               // FFT, influence function, inverse FFT
void apx_pme_gather_dp(apx_ctx* c, int epi, const real4* U, const real4* F, real* fd, real* fp, real4* OUT, double* slot,
   const int* itp = nullptr);
void apx_pme_uind_fphi(apx_ctx* c, const real* ud, const real* up, bool full20);
void apx_pme_cross_virial(apx_ctx* c, real4* mpa, real4* mpb, double* out6);
// ---- fft64.cu
bool apx_fft64_usable(const apx_ctx* c);
void apx_fft64_setup(apx_ctx* c);
void apx_fft64_convolve(apx_ctx* c);                             // grid <- IFFT(qfac * FFT(grid))
// ---- field.cu
void apx_dfield_real(apx_ctx* c, cudaStream_t st, real* fd, real* fp, bool assign);
void apx_ufield_real_dp(apx_ctx* c, cudaStream_t st, const real4* U, real4* F);    // F = real-space field of U (assigned)
struct PcgTest {          // convergence test fused into the preconditioner kernel of iteration `it` (it = 0: none)
   int it = 0, miniter = 0, politer = 0;
   const int* itp = nullptr;       // device-side loop (pcg.cu): the iteration number is *itp, `slot` is the slot of iteration 1
   unsigned long long cond = 0;    // conditional handle of the WHILE node the iterations run in: set to 0 on convergence
   real poleps = 0, debye = 0, pcgpeek = 0;
   const double* slot = nullptr;   // slot of iteration it (r.r in quantities 4,5)
   double* result = nullptr;       // [0] eps, [1] iteration
   int* flags = nullptr;
   real* ud = nullptr;
   real* up = nullptr;
};
void apx_precond_dp(apx_ctx* c, const real4* R, real4* Z, double* slot, const PcgTest* test = nullptr);   // Z = M R ; partial R.Z -> slot
// iteration-relative scalar slots: with itp the kernels of an iteration find their slot themselves, base + PCG_SLOT (*itp - 1)
#define PCG_SLOT_DOUBLES 96
#ifdef __CUDACC__
__device__ __forceinline__ double* pcg_slot_of(double* base, const int* itp)
{
   return itp ? base + (size_t)PCG_SLOT_DOUBLES * (*itp - 1) : base;
}
__device__ __forceinline__ const double* pcg_slot_of(const double* base, const int* itp)
{
   return itp ? base + (size_t)PCG_SLOT_DOUBLES * (*itp - 1) : base;
}
#endif
void apx_precond_apply(apx_ctx* c, const real* rd, const real* rp, real* zd, real* zp);
// ---- pcg.cu
bool apx_induce_impl(apx_ctx* c, bool defer = false);      // true: apx_induce_finish is pending (after the caller's synchronisation)
void apx_induce_resume(apx_ctx* c);                         // after a finish that returned false: more iterations, waiting for them
bool apx_induce_finish(apx_ctx* c);                         // false: not converged in the deferred batch, redo without defer
void apx_pcg_graphs_invalidate(apx_ctx* c);
// if (apx_graph_begin(c, key)) { enqueue the region on c->stream; apx_graph_end(c, key); }
// first call: runs eagerly (lazy allocations happen); second: captured, instantiated, launched; later: replayed.
// cond_flag != nullptr: the captured region becomes the body of a conditional IF node that runs only when *cond_flag != 0 at
// that point of the stream (a one-thread kernel ahead of the node reads the flag) -- work enqueued behind a solver that may
// not have converged yet.  apx_graph_is_conditional tells whether the replayed graph really has that form.
bool apx_graph_begin(apx_ctx* c, int key, const int* cond_flag = nullptr);
bool apx_graph_is_conditional(apx_ctx* c, int key);
void apx_graph_end(apx_ctx* c, int key);
void apx_upred_configure(apx_ctx* c, int polpred);             // sets opt.polpred, sizes and empties the ring
void apx_pack_dp(apx_ctx* c, const real* d, const real* p, real4* out);
void apx_unpack_dp(apx_ctx* c, const real4* in, real* d, real* p);
void apx_ufield_full(apx_ctx* c, const real* ud, const real* up, real* fd, real* fp);
// ---- mplar.cu
void apx_energy_impl(apx_ctx* c, int vers, bool do_m, bool do_p, apx_energy_result* out, bool do_vdw = false, bool do_val = false);
void apx_energy_impl_md(apx_ctx* c, int vers, apx_energy_result* out);      // slow level of the integrator: electrostatics + vdW
void apx_energy_md_enqueue(apx_ctx* c, int vers);
bool apx_energy_md_collect(apx_ctx* c, int vers, apx_energy_result* out);      // true: a solver-batch miss was repaired here
