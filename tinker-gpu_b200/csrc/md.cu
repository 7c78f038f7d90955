// Molecular dynamics on the device: velocity Verlet and r-RESPA with the valence terms on the inner level, optional
// Bussi thermostat -- SURVEY.md section 8f rank 3.  Restates BasicIntegrator::dynamic / RespaIntegrator::KickOff
// (src/md/integrator.cpp:70-170, 205-222), RespaDevice::velR0/velR1/velR2 + mdPos / mdVel / mdVel2
// (src/md/propagator.cpp:170-187, src/acc/mdpq.cpp:12-115), kinetic() and bussiThermostat (src/mdpt.cpp:41-71).
//
// B200 shape: positions and velocities never leave the device.  One outer step is
//    [graph]  kick(fast dt_a/2 + slow dt/2) + drift          1 launch
//             { valence gradient, kick(fast dt_a) + drift }   2 launches x (nrespa - 1)
//             valence energy + gradient                       1 launch
//    neighbour-list check, induce + electrostatics + vdW      the hot path (pcg.cu, mplar.cu, ehal.cu)
//    kick(fast dt_a/2 + slow dt/2) + kinetic energy           1 launch
//    Bussi rescale (scale factor drawn on the device)         1 launch
// The kick/drift kernels read the fast gradient from the valence accumulator (caller order) and the slow one from the
// sorted fixed-point accumulators of the electrostatics path through inv[], and clear the valence accumulator for the
// next evaluation, so no separate copy / zero passes exist (the reference copies gx -> gx1/gx2 three arrays at a time).
// All per-atom passes are HBM streams of 24 (x) + 24 (v) + 24..48 (g) + 8 (1/m) B per atom: ~3 MB at dhfr2, L2-resident.
#include "apx_internal.h"
#include <chrono>
#include <cmath>
#include <cstring>

struct MdState {
   int on = 0, n = 0, nrespa = 1, thermostat = 0, nfree = 0;
   double dt = 0, kelvin = 0, tautemp = 0;
   unsigned long long seed = 0, step = 0;
   DevBuf<double> vel, massinv, mass;
   DevBuf<double> sc;      // [0] sum m v^2 of this step, [1] eksum (kcal/mol) after the thermostat, [2] temperature, [3] last scale
   double* sc_h = nullptr;
   cudaEvent_t t0 = nullptr, t1 = nullptr;
   int graph_has_check = 0;      // the captured step graph contains the neighbour-list test
   int seq_host = 0;             // sequence number the next published list test will carry (device copy: sc[7])
};

namespace {
constexpr double EKCAL = 418.4;                // units::ekcal, tinker/source/units.f:90
constexpr double GASCONST = 1.9872042586e-3;   // units::gasconst, units.f:84

__device__ __forceinline__ double fx2d(fixed_t v) { return (double)(long long)v * (1.0 / APX_FIXED_SCALE); }

// v += -ekcal/m (g_fast cf + g_slow cs); optionally x += dta v; the fast accumulator is cleared for its next evaluation.
// KIN: accumulate sum m v^2 (after the kick) into ksum.
template <bool DRIFT, bool KIN>
__global__ void k_md_kick(int n, double cf, double cs, double dta, const double* __restrict__ massinv, const double* __restrict__ mass,
   fixed_t* __restrict__ vg, const fixed_t* __restrict__ gx, const fixed_t* __restrict__ gy, const fixed_t* __restrict__ gz,
   const int* __restrict__ inv, double* __restrict__ vel, double* __restrict__ xyz, double* __restrict__ ksum)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   double mv2 = 0;
   if (i < n) {
      const double coef = -EKCAL * massinv[i];
      double g[3] = {0, 0, 0};
      if (vg && cf != 0) {
         g[0] = cf * fx2d(vg[i]), g[1] = cf * fx2d(vg[(size_t)n + i]), g[2] = cf * fx2d(vg[2 * (size_t)n + i]);
      }
      if (gx && cs != 0) {
         const int s = inv[i];
         g[0] += cs * fx2d(gx[s]), g[1] += cs * fx2d(gy[s]), g[2] += cs * fx2d(gz[s]);
      }
      if (vg && DRIFT)
         vg[i] = 0, vg[(size_t)n + i] = 0, vg[2 * (size_t)n + i] = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
         double v = vel[3 * (size_t)i + k] + coef * g[k];
         vel[3 * (size_t)i + k] = v;
         if (DRIFT)
            xyz[3 * (size_t)i + k] += dta * v;
         mv2 += v * v;
      }
      mv2 *= mass[i];
   }
   if (KIN) {
      for (int o = 16; o > 0; o >>= 1)
         mv2 += __shfl_down_sync(0xffffffffu, mv2, o);
      __shared__ double sh[8];
      if ((threadIdx.x & 31) == 0)
         sh[threadIdx.x >> 5] = mv2;
      __syncthreads();
      if (threadIdx.x == 0) {
         double t = 0;
         for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
            t += sh[w];
         atomicAdd(ksum, t);
      }
   }
}

// counter-based generator: the scale factor of step `ctr` is a pure function of (seed, ctr), so every block can draw it
__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
   z += 0x9e3779b97f4a7c15ull;
   z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
   z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
   return z ^ (z >> 31);
}
struct Rng {
   unsigned long long key, ctr;
   __device__ double uniform()      // (0,1)
   {
      unsigned long long r = mix64(key ^ mix64(ctr++));
      return ((double)(r >> 11) + 0.5) * (1.0 / 9007199254740992.0);
   }
   __device__ double normal()
   {
      const double u1 = uniform(), u2 = uniform();
      return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
   }
   __device__ double gamma(double a)      // Marsaglia-Tsang, a >= 1
   {
      const double d = a - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * d);
      for (int it = 0; it < 64; ++it) {
         const double x = normal();
         double v = 1.0 + cc * x;
         if (v <= 0)
            continue;
         v = v * v * v;
         const double u = uniform();
         if (log(u) < 0.5 * x * x + d - d * v + d * log(v))
            return d * v;
      }
      return d;
   }
};

// bussiThermostat (src/mdpt.cpp:41-71): c = exp(-dt/tau), d = (1-c)(T0/T)/nfree, scale^2 = c + (s + r^2) d + 2 r sqrt(c d)
// with r ~ N(0,1), s ~ chi^2(nfree-1).  thermostat == 0: only the kinetic energy / temperature are published.
// The step number that keys the random stream lives on the device (sc[6], advanced by the last CTA when `advance` is set): the
// kernel can then sit in a captured graph, behind work the host has not waited for.
__global__ void k_md_thermo(int n, int thermostat, int nfree, double dt, double tautemp, double kelvin, unsigned long long seed,
   int advance, double* __restrict__ vel, double* __restrict__ sc)
{
   __shared__ double sh_scale;
   const unsigned long long step = advance ? (unsigned long long)sc[6] + 1ull : 0ull;
   if (threadIdx.x == 0) {
      const double eksum = 0.5 * sc[0] / EKCAL;
      double temp = 2.0 * eksum / ((double)nfree * GASCONST);
      double scale = 1.0;
      if (thermostat == 1) {
         if (temp == 0)
            temp = 0.1;
         Rng g;
         g.key = mix64(seed), g.ctr = step << 8;
         const double c = exp(-dt / tautemp);
         const double d = (1.0 - c) * (kelvin / temp) / (double)nfree;
         const double r = g.normal();
         const double s = nfree > 1 ? 2.0 * g.gamma(0.5 * (double)(nfree - 1)) : 0.0;
         double s2 = c + (s + r * r) * d + 2.0 * r * sqrt(c * d);
         scale = sqrt(s2 > 0 ? s2 : 0.0);
         if (r + sqrt(c / d) < 0)
            scale = -scale;
      }
      sh_scale = scale;
      if (blockIdx.x == 0) {
         sc[1] = eksum * scale * scale;
         sc[2] = 2.0 * sc[1] / ((double)nfree * GASCONST);
         sc[3] = scale;
      }
   }
   __syncthreads();
   const double scale = sh_scale;
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n && scale != 1.0) {
      vel[3 * (size_t)i] *= scale, vel[3 * (size_t)i + 1] *= scale, vel[3 * (size_t)i + 2] *= scale;
   }
   if (advance) {
      // every CTA has read sc[6] before it takes its ticket
      __syncthreads();
      if (threadIdx.x == 0) {
         const double t = atomicAdd(&sc[5], 1.0);
         if (t == (double)(gridDim.x - 1)) {
            sc[5] = 0;
            sc[6] = (double)step;
         }
      }
   }
}

__global__ void k_md_zero1(double* p) { *p = 0; }
}      // namespace

static void md_kick(apx_ctx* c, bool drift, bool kin, double cf, double cs, double dta)
{
   MdState& M = *c->md;
   const int n = M.n, g = (n + 255) / 256;
   fixed_t* vg = apx_valence_on(c) ? apx_valence_grad_buffer(c) : nullptr;
   if (drift)
      k_md_kick<true, false><<<g, 256, 0, c->stream>>>(n, cf, cs, dta, M.massinv, M.mass, vg, c->gx, c->gy, c->gz, c->inv, M.vel, c->xyz_d, M.sc);
   else if (kin)
      k_md_kick<false, true><<<g, 256, 0, c->stream>>>(n, cf, cs, dta, M.massinv, M.mass, vg, c->gx, c->gy, c->gz, c->inv, M.vel, c->xyz_d, M.sc);
   else
      k_md_kick<false, false><<<g, 256, 0, c->stream>>>(n, cf, cs, dta, M.massinv, M.mass, vg, c->gx, c->gy, c->gz, c->inv, M.vel, c->xyz_d, M.sc);
   APX_COUNT_LAUNCH(c);
}

static void md_positions_changed(apx_ctx* c, int known_moved)
{
   c->mpole_inited = 0;
   c->mpole_pme_valid = 0;
   c->induced_valid = 0;
   apx_list_refresh(c, false, known_moved);
}

// kick-off (RespaIntegrator::KickOff, src/md/integrator.cpp:205-222): fast gradient into the valence accumulator, slow gradient
// into gx/gy/gz at the current positions.  The reference keeps private gx1/gx2 copies; here the saved forces live in the shared
// accumulators, so any public call that rewrites them (energy, evdw, empole, epolar, evalence, set_positions, set_box) clears
// md_forces_valid and the next apx_md_steps starts by recomputing them.
static void md_kickoff_forces(apx_ctx* c)
{
   cudaStream_t st = c->stream;
   if (!c->list_valid)
      apx_list_refresh(c, true);
   if (apx_valence_on(c)) {
      apx_valence_enqueue(c, APX_ENERGY | APX_GRAD, st, true);
      apx_valence_fetch(c, st);
   }
   apx_energy_result r;
   apx_energy_impl_md(c, APX_V4, &r);
   c->md_forces_valid = 1;
}

void apx_md_init_impl(apx_ctx* c, const double* mass, const double* vel, const apx_md_config* cfg)
{
   if (c->dist.on)
      APX_THROW("apx_md_init: the integrator is built for single-GPU contexts");
   if (!(cfg->dt > 0) || cfg->nrespa < 1)
      APX_THROW("apx_md_init: dt must be positive and nrespa >= 1");
   if (cfg->thermostat != 0 && cfg->thermostat != 1)
      APX_THROW("apx_md_init: thermostat must be 0 (none) or 1 (Bussi)");
   if (cfg->thermostat == 1 && !(cfg->tautemp > 0 && cfg->kelvin > 0))
      APX_THROW("apx_md_init: the Bussi thermostat needs kelvin > 0 and tautemp > 0");
   if (cfg->nrespa > 1 && !apx_valence_on(c))
      APX_THROW("apx_md_init: r-RESPA needs the valence terms on its inner level (apx_valence_attach)");
   if (!c->md)
      c->md = new MdState;
   MdState& M = *c->md;
   const int n = c->n;
   M.n = n, M.nrespa = cfg->nrespa, M.thermostat = cfg->thermostat, M.dt = cfg->dt, M.kelvin = cfg->kelvin, M.tautemp = cfg->tautemp;
   M.nfree = cfg->nfree > 0 ? cfg->nfree : 3 * n - 3;
   M.seed = cfg->seed, M.step = 0;
   std::vector<double> mi(n), v(3 * (size_t)n, 0.0);
   for (int i = 0; i < n; ++i) {
      if (mass[i] < 0)
         APX_THROW("apx_md_init: negative mass");
      mi[i] = mass[i] > 0 ? 1.0 / mass[i] : 0.0;
   }
   if (vel)
      memcpy(v.data(), vel, sizeof(double) * 3 * (size_t)n);
   M.vel.ensure(3 * (size_t)n), M.massinv.ensure(n), M.mass.ensure(n), M.sc.ensure(8);
   if (!M.sc_h)
      CUDA_CHECK(cudaMallocHost(&M.sc_h, sizeof(double) * 8));
   if (!M.t0) {
      CUDA_CHECK(cudaEventCreate(&M.t0));
      CUDA_CHECK(cudaEventCreate(&M.t1));
   }
   cudaStream_t st = c->stream;
   // the step graphs bake dt / dt_a into their kick and drift nodes: a second init must not replay the old coefficients
   for (auto it = c->step_graphs.begin(); it != c->step_graphs.end();) {
      if ((it->first >= 0x3000 && it->first < 0x4000) || (it->first >= 0x6000 && it->first < 0x7000) || (it->first & 0x10000)) {
         if (it->second.exec)
            cudaGraphExecDestroy(it->second.exec);
         it = c->step_graphs.erase(it);
      } else
         ++it;
   }
   CUDA_CHECK(cudaMemcpyAsync(M.vel.p, v.data(), sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, st));
   CUDA_CHECK(cudaMemcpyAsync(M.massinv.p, mi.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
   CUDA_CHECK(cudaMemcpyAsync(M.mass.p, mass, sizeof(double) * n, cudaMemcpyHostToDevice, st));
   CUDA_CHECK(cudaMemsetAsync(M.sc.p, 0, sizeof(double) * 8, st));
   M.seq_host = 0;
   c->flags_h[7] = 0;
   CUDA_CHECK(cudaStreamSynchronize(st));
   md_kickoff_forces(c);
   // kinetic energy of the starting velocities
   md_kick(c, false, true, 0.0, 0.0, 0.0);
   k_md_thermo<<<(n + 255) / 256, 256, 0, st>>>(n, 0, M.nfree, M.dt, 1.0, 1.0, M.seed, 0, M.vel, M.sc);
   APX_COUNT_LAUNCH(c);
   CUDA_CHECK(cudaStreamSynchronize(st));
   M.on = 1;
}

void apx_md_steps_impl(apx_ctx* c, int nsteps, apx_md_report* out)
{
   if (!c->md || !c->md->on)
      APX_THROW("apx_md_steps before apx_md_init");
   MdState& M = *c->md;
   cudaStream_t st = c->stream;
   const int n = M.n, nr = M.nrespa;
   const double dt = M.dt, dta = dt / nr;
   const bool val = apx_valence_on(c);
   const int rebuilds0 = c->stats.list_rebuilds;
   apx_energy_result r;
   memset(&r, 0, sizeof(r));
   apx_valence_result vr;
   memset(&vr, 0, sizeof(vr));
   if (!c->md_forces_valid && nsteps > 0)
      md_kickoff_forces(c);
   cudaEventRecord(M.t0, st);
   bool report_enqueued = false;
   for (int s = 0; s < nsteps; ++s) {
      bool check_async = false;
      if (apx_graph_begin(c, 0x3000 + nr)) {
         k_md_zero1<<<1, 1, 0, st>>>(M.sc.p);
         APX_COUNT_LAUNCH(c);
         md_kick(c, true, false, 0.5 * dta, 0.5 * dt, dta);                 // velR1 + pos(dta)
         for (int f = 1; f < nr; ++f) {
            apx_valence_enqueue(c, APX_GRAD, st, false);                    // energy(grad, RESPA_FAST)
            md_kick(c, true, false, dta, 0.0, dta);                         // velR0(dta) + pos(dta)
         }
         // the neighbour-list test of the new positions on the second stream, BESIDE the last valence evaluation: its answer
         // is in host memory before that kernel ends, so the host round trip that decides "refresh or rebuild" costs the GPU
         // nothing (it used to idle ~15 us per step waiting for it)
         check_async = c->list_valid != 0;
         if (check_async) {
            CUDA_CHECK(cudaEventRecord(c->ev_fork, st));
            CUDA_CHECK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
            apx_list_check_enqueue(c, c->stream2, M.sc.p + 7);
            CUDA_CHECK(cudaEventRecord(c->ev_join, c->stream2));
         }
         if (val)
            apx_valence_enqueue(c, APX_ENERGY | APX_GRAD, st, false);       // fast force at the new positions
         if (check_async)
            CUDA_CHECK(cudaStreamWaitEvent(st, c->ev_join, 0));
         apx_graph_end(c, 0x3000 + nr);
         M.graph_has_check = check_async ? 1 : 0;
      } else
         check_async = M.graph_has_check != 0;
      int moved = -1;
      if (check_async) {
         // spin on the sequence number the side branch publishes into pinned memory (k_list_publish)
         const int want = ++M.seq_host;
         volatile int* fh = c->flags_h;
         const auto t_begin = std::chrono::steady_clock::now();
         long spins = 0;
         while (fh[7] != want) {
            if ((++spins & 0xfff) == 0 && std::chrono::steady_clock::now() - t_begin > std::chrono::seconds(20)) {
               CUDA_CHECK(cudaStreamSynchronize(st));
               if (fh[7] != want)
                  APX_THROW("apx_md_steps: the neighbour-list test never reported back");
            }
         }
         moved = fh[6] != 0 ? 1 : 0;
      }
      md_positions_changed(c, moved);                                       // copyPosToXyz(true): list check / rebuild
      if (val)
         apx_valence_fetch(c, st);
      // slow force (induce + emplar + ehal), closing half-kick, thermostat.  Once the graphs exist nothing between them waits
      // for the GPU: the energy evaluation is only enqueued, the kick and the thermostat follow it as the body of an IF node
      // keyed on the solver's convergence flag (like the energy epilogue, mplar.cu), ONE synchronisation ends the step.  A
      // solver batch that was too short leaves both IF nodes idle; the collect step finishes the solve and the epilogue, and the
      // tail is launched again.
      const int tkey = 0x6000 + nr;
      const int* cflag = (c->opt.use_polar && c->opt.poltyp_mutual) ? c->flags.p + 1 : nullptr;
      auto tail = [&]() {
         if (apx_graph_begin(c, tkey, cflag)) {
            md_kick(c, false, true, 0.5 * dta, 0.5 * dt, 0.0);              // velR2 + sum m v^2
            k_md_thermo<<<(n + 255) / 256, 256, 0, st>>>(n, M.thermostat, M.nfree, dt, M.tautemp > 0 ? M.tautemp : 1.0, M.kelvin, M.seed,
               1, M.vel, M.sc);
            APX_COUNT_LAUNCH(c);
            apx_graph_end(c, tkey);
         }
      };
      if (c->cond_nodes_ok == 1 && apx_graph_is_conditional(c, tkey)) {
         // the kick and the thermostat ride inside the epilogue's IF body when the evaluation can take them (mplar.cu: epi_tail);
         // c->epi_tail_ran says whether it did -- while that merged graph is warmed up and captured, and wherever conditional
         // nodes are missing, the separate tail graph follows as before
         static const int merge = getenv("APX_MD_MERGE_TAIL") ? atoi(getenv("APX_MD_MERGE_TAIL")) : 1;
         if (merge)
            c->epi_tail = [&]() {
               md_kick(c, false, true, 0.5 * dta, 0.5 * dt, 0.0);
               k_md_thermo<<<(n + 255) / 256, 256, 0, st>>>(n, M.thermostat, M.nfree, dt, M.tautemp > 0 ? M.tautemp : 1.0, M.kelvin, M.seed,
                  1, M.vel, M.sc);
               APX_COUNT_LAUNCH(c);
            };
         struct Clear {
            apx_ctx* c;
            ~Clear() { c->epi_tail = nullptr; }
         } clear{c};
         apx_energy_md_enqueue(c, APX_V4);
         const bool merged = c->epi_tail_ran != 0;
         if (!merged)
            tail();
         // last step of the call: the report's copy rides in front of the same synchronisation (a one-step call, which is
         // what a host-side driver and the bench's per-step timing make, then has ONE host round trip instead of two)
         if (s == nsteps - 1) {
            cudaEventRecord(M.t1, st);
            CUDA_CHECK(cudaMemcpyAsync(M.sc_h, M.sc.p, sizeof(double) * 8, cudaMemcpyDeviceToHost, st));
            report_enqueued = true;
         }
         CUDA_CHECK(cudaStreamSynchronize(st));
         if (apx_energy_md_collect(c, APX_V4, &r)) {
            report_enqueued = false;      // (the solve was finished after that copy: take the report again)
            if (!c->epi_tail_ran)
               tail();      // (its IF node did not fire behind the unconverged batch)
         }
      } else {
         apx_energy_impl_md(c, APX_V4, &r);
         tail();
      }
      c->md_forces_valid = 1;
      M.step++;
   }
   if (!report_enqueued) {
      cudaEventRecord(M.t1, st);
      CUDA_CHECK(cudaMemcpyAsync(M.sc_h, M.sc.p, sizeof(double) * 8, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
   }
   if (val)
      apx_valence_collect(c, APX_ENERGY | APX_GRAD, &vr);
   if (out) {
      memset(out, 0, sizeof(*out));
      out->steps = nsteps;
      out->e_valence = vr.esum;
      out->e_nonbonded = r.esum;
      out->epot = vr.esum + r.esum;
      out->ekin = M.sc_h[1];
      out->temp = M.sc_h[2];
      out->last_scale = M.sc_h[3];
      out->pcg_iterations = r.pcg_iterations;
      out->list_rebuilds = c->stats.list_rebuilds - rebuilds0;
      out->total_steps = (long long)M.step;
      cudaEventElapsedTime(&out->ms_device, M.t0, M.t1);
   }
}

void apx_md_get_state_impl(apx_ctx* c, double* xyz, double* vel)
{
   if (!c->md || !c->md->on)
      APX_THROW("apx_md_get_state before apx_md_init");
   const size_t b = sizeof(double) * 3 * (size_t)c->n;
   if (xyz)
      CUDA_CHECK(cudaMemcpyAsync(xyz, c->xyz_d.p, b, cudaMemcpyDeviceToHost, c->stream));
   if (vel)
      CUDA_CHECK(cudaMemcpyAsync(vel, c->md->vel.p, b, cudaMemcpyDeviceToHost, c->stream));
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

// integrator state from host buffers (a host-side driver that keeps x and v itself: the e2e leg of bench.py).  forces_valid != 0:
// the positions are the ones the last apx_md_steps call left (handed back unchanged), so the saved forces still apply;
// otherwise they are recomputed at the start of the next apx_md_steps.
void apx_md_set_state_impl(apx_ctx* c, const double* xyz, const double* vel, int forces_valid)
{
   if (!c->md || !c->md->on)
      APX_THROW("apx_md_set_state before apx_md_init");
   const size_t b = sizeof(double) * 3 * (size_t)c->n;
   if (xyz)
      CUDA_CHECK(cudaMemcpyAsync(c->xyz_d.p, xyz, b, cudaMemcpyHostToDevice, c->stream));
   if (vel)
      CUDA_CHECK(cudaMemcpyAsync(c->md->vel.p, vel, b, cudaMemcpyHostToDevice, c->stream));
   if (xyz && !forces_valid) {
      c->md_forces_valid = 0;
      c->mpole_inited = 0, c->mpole_pme_valid = 0, c->induced_valid = 0;
      apx_list_refresh(c, false);
   }
}

void apx_md_destroy(apx_ctx* c)
{
   if (!c->md)
      return;
   MdState& M = *c->md;
   M.vel.release(), M.massinv.release(), M.mass.release(), M.sc.release();
   if (M.sc_h)
      cudaFreeHost(M.sc_h);
   if (M.t0) {
      cudaEventDestroy(M.t0);
      cudaEventDestroy(M.t1);
   }
   delete c->md;
   c->md = nullptr;
}
