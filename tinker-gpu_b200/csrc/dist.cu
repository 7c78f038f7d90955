// Multi-GPU spatial decomposition of the AMOEBA electrostatics path (the reference is
// single-GPU, so nothing here has a counterpart in it).
//
//   * z-slabs: GPU g owns the atoms whose PME grid coordinate w3 lies in [g/G, (g+1)/G) -- one
//     contiguous range [a0,a1) of the common sorted order (nblist.cu puts the slab in the top bits
//     of the sort key) -- and the nfft3/G planes of the PME grid those atoms sit on.  Per-atom
//     static data (positions, rotated multipoles, polarizabilities) is replicated: it is O(N)
//     streaming work per step and needs no exchange.  Everything that costs -- neighbor rows,
//     pair kernels, spreading, gathering, FFT, solver vectors -- is done for owned atoms/planes only.
//   * halo exchange per operator application: the packed (d,p) vectors of the atoms within
//     cutoff+buffer of a neighbour's slab go to that neighbour.  Both sides derive the same index
//     lists from the replicated coordinates, so no index traffic is needed; sorted indices are global,
//     so the direct transport writes every halo atom straight into the same slot of the peer's array.
//   * slab-decomposed 3-D FFT: halo planes of the spread grid are summed into their owners, 2-D
//     FFTs over the owned planes, an all-to-all transpose to [k3][k2 local][k1], 1-D FFTs along z,
//     the influence function, and the way back; the potential's halo planes are then returned for
//     the gather.
//   * scalars of the solver: one small all-reduce per dot-product pair (3 per iteration).
//
// Transports (ApxComm): (1) one process per GPU -- NCCL for the start-up handshake and the large force
// reduction, and for everything else the library's own peer-memory kernels over NVLink: `DirectComm`
// (default; exchange buffers registered through CUDA IPC, ONE kernel per exchange that writes into the
// peers' memory, flag-based scalar all-reduce), `P2pComm` (windowed push / pull, APX_DIST_P2P=2 or 1) or
// plain NCCL send/recv (APX_DIST_P2P=0); (2) transport "direct": DirectComm with no NCCL at all (IPC
// handles through a /dev/shm rendezvous) -- used to test the process-per-rank path on a one-GPU box;
// (3) an in-process transport for several ranks driven by host threads on ONE GPU, which lets the whole
// decomposition be parity-tested against the single-GPU path without any inter-process machinery.
#include "apx_internal.h"
#include "dp.cuh"
#include <nccl.h>
#include <dlfcn.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>

// ------------------------------------------------------------------------------------------------
// transports
// ------------------------------------------------------------------------------------------------
struct ApxComm {
   struct Op {
      int peer;
      void* ptr;
      size_t bytes;
   };
   int rank = 0, world = 1;
   virtual ~ApxComm() {}
   // dtype: 0 = f64, 1 = u64, 2 = i32 ; in place, sum
   virtual void allreduce(void* p, size_t n, int dtype, cudaStream_t st) = 0;
   // messages between the same pair of ranks are matched in the order they are listed
   virtual void exchange(const std::vector<Op>& sends, const std::vector<Op>& recvs, cudaStream_t st) = 0;
};

namespace {
// ---- NCCL, bound with dlsym so that libapx carries no link-time dependency on a particular libnccl
struct NcclApi {
   void* handle = nullptr;
   ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
   ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
   ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
   ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*GroupStart)() = nullptr;
   ncclResult_t (*GroupEnd)() = nullptr;
   const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl_api(const char* path)
{
   static NcclApi api;
   static std::mutex m;
   std::lock_guard<std::mutex> lk(m);
   if (api.handle)
      return &api;
   const char* names[] = {path && path[0] ? path : "libnccl.so.2", "libnccl.so.2", "libnccl.so"};
   for (const char* nm : names) {
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle)
         break;
   }
   if (!api.handle)
      APX_THROW(std::string("cannot load NCCL: ") + dlerror());
#define BIND(field, sym)                                                                                                 \
   api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym));                                            \
   if (!api.field)                                                                                                       \
      APX_THROW(std::string("NCCL symbol missing: ") + sym)
   BIND(GetUniqueId, "ncclGetUniqueId");
   BIND(CommInitRank, "ncclCommInitRank");
   BIND(CommDestroy, "ncclCommDestroy");
   BIND(AllReduce, "ncclAllReduce");
   BIND(AllGather, "ncclAllGather");
   BIND(Send, "ncclSend");
   BIND(Recv, "ncclRecv");
   BIND(GroupStart, "ncclGroupStart");
   BIND(GroupEnd, "ncclGroupEnd");
   BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
   return &api;
}

#define NCCL_CHECK(api, expr)                                                                                            \
   do {                                                                                                                  \
      ncclResult_t r__ = (expr);                                                                                         \
      if (r__ != ncclSuccess)                                                                                            \
         apx_throw(__FILE__, __LINE__, std::string(#expr) + ": " + (api)->GetErrorString(r__));                          \
   } while (0)

// ---- start-up exchange of small host blobs between the ranks of one node without NCCL (transport "direct"): every
// rank drops its blob into /dev/shm under a name derived from the job id and polls for the others'.  Round k files
// are removed by their owner once round k+1 has been read (every rank has then finished reading round k).
struct FileRendezvous {
   std::string prefix;
   int rank = 0, world = 1;
   unsigned round = 0;
   std::string name(int r, unsigned rd) const { return prefix + "_" + std::to_string(rd) + "_" + std::to_string(r); }
   void gather(const void* mine, size_t bytes, void* all)
   {
      ++round;
      const std::string fin = name(rank, round), tmp = fin + ".tmp";
      FILE* f = fopen(tmp.c_str(), "wb");
      if (!f || fwrite(mine, 1, bytes, f) != bytes)
         APX_THROW("rendezvous: cannot write " + tmp);
      fclose(f);
      if (rename(tmp.c_str(), fin.c_str()) != 0)
         APX_THROW("rendezvous: cannot publish " + fin);
      const auto t0 = std::chrono::steady_clock::now();
      for (int r = 0; r < world; ++r) {
         char* dst = static_cast<char*>(all) + (size_t)r * bytes;
         if (r == rank) {
            memcpy(dst, mine, bytes);
            continue;
         }
         const std::string fn = name(r, round);
         for (;;) {
            FILE* g = fopen(fn.c_str(), "rb");
            if (g) {
               const size_t got = fread(dst, 1, bytes, g);
               fclose(g);
               if (got == bytes)
                  break;
            }
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120))
               APX_THROW("rendezvous: rank " + std::to_string(r) + " did not show up within 120 s (" + fn + ")");
            std::this_thread::sleep_for(std::chrono::microseconds(200));
         }
      }
      if (round > 1)
         remove(name(rank, round - 1).c_str());
   }
};

struct NcclComm : ApxComm {
   NcclApi* api = nullptr;
   ncclComm_t comm = nullptr;
   FileRendezvous* rdv = nullptr;      // set instead of `comm` by the NCCL-free transport
   ~NcclComm() override
   {
      if (comm)
         api->CommDestroy(comm);
      delete rdv;
   }
   void need_nccl(const char* what) const
   {
      if (!comm)
         APX_THROW(std::string("transport \"direct\" has no NCCL communicator: ") + what);
   }
   // every rank's blob of `bytes` host bytes, concatenated by rank (start-up only: synchronises the host)
   void gather_bytes(const void* mine, size_t bytes, void* all)
   {
      if (rdv) {
         rdv->gather(mine, bytes, all);
         return;
      }
      char* dev = nullptr;
      CUDA_CHECK(cudaMalloc(&dev, bytes * (world + 1)));
      CUDA_CHECK(cudaMemcpy(dev + bytes * world, mine, bytes, cudaMemcpyHostToDevice));
      NCCL_CHECK(api, api->AllGather(dev + bytes * world, dev, bytes, ncclChar, comm, nullptr));
      CUDA_CHECK(cudaStreamSynchronize(nullptr));
      CUDA_CHECK(cudaMemcpy(all, dev, bytes * world, cudaMemcpyDeviceToHost));
      cudaFree(dev);
   }
   void allreduce(void* p, size_t n, int dtype, cudaStream_t st) override
   {
      need_nccl("all-reduce");
      ncclDataType_t t = dtype == 0 ? ncclFloat64 : (dtype == 1 ? ncclUint64 : ncclInt32);
      NCCL_CHECK(api, api->AllReduce(p, p, n, t, ncclSum, comm, st));
   }
   void exchange(const std::vector<Op>& sends, const std::vector<Op>& recvs, cudaStream_t st) override
   {
      need_nccl("send/recv");
      NCCL_CHECK(api, api->GroupStart());
      for (const Op& o : sends)
         NCCL_CHECK(api, api->Send(o.ptr, o.bytes, ncclChar, o.peer, comm, st));
      for (const Op& o : recvs)
         NCCL_CHECK(api, api->Recv(o.ptr, o.bytes, ncclChar, o.peer, comm, st));
      NCCL_CHECK(api, api->GroupEnd());
   }
};

// ---- peer-memory transport for the bulk exchanges (one process per GPU, NVLink / NVSwitch)
//
// NCCL's grouped send/recv moved the 36 MB transpose blocks of the 1 M-atom box at ~300 GB/s and small
// halo messages at ~55 GB/s (profiles/r01g_trace_water1m_n2.txt): four such exchanges per CG operator sit
// on the critical path.  Here every rank exposes a receive window through CUDA IPC; a sender copies
// straight into the receiver's window over NVLink (cudaMemcpyAsync to the mapped peer pointer: copy
// engines, no SMs), then raises a sequence flag in the receiver's memory; the receiver's stream waits
// on the flag with a one-thread kernel, moves the data from the window to its destination and
// acknowledges, so that the window can be reused two exchanges later (two slots per sender).
// Everything is stream ordered; the host never blocks.  Small all-reduces stay on NCCL, which also
// carries the IPC handles at start-up.  APX_DIST_P2P=0 forces NCCL send/recv, 1 these copy-engine windows, 2 the fused kernels below.
__global__ void k_flag_wait(const volatile unsigned* flag, unsigned want)
{
   // sequence numbers only grow; unsigned difference handles wrap-around
   while ((int)(*flag - want) < 0)
      __nanosleep(200);
   __threadfence_system();
}
__global__ void k_flag_set(volatile unsigned* flag, unsigned value)
{
   __threadfence_system();
   *flag = value;
}

// ---- one kernel per direction for ALL peers of an exchange (APX_DIST_P2P=2)
// k_xfer moves a table of messages with the SMs: a CTA first waits (thread 0 spins) until the flag of its message's
// peer has reached `want`, copies its share of the message with 16-byte loads/stores -- to the peer's window over
// NVLink when pushing, out of the local window when pulling -- and the last CTA of every peer raises that peer's flag
// (ready flag in the receiver's memory after a push, acknowledge flag in the sender's memory after a pull).
#define XFER_MAX_MSG 40
struct XferMsg {
   const char* src;
   char* dst;
   size_t bytes;
   int peer;            // index into the per-peer tables
   int cta0, nctas;     // CTAs [cta0, cta0 + nctas) move this message
};
struct XferTable {
   XferMsg m[XFER_MAX_MSG];
   int nmsg;
   const volatile unsigned* wait_flag[16];    // per peer: spin until >= wait_val (nullptr: no wait)
   unsigned wait_val[16];
   volatile unsigned* set_flag[16];           // per peer: written by the last CTA serving that peer
   unsigned set_val;
   int peer_ctas[16];                         // CTAs serving each peer
   unsigned* counters;                        // [16] arrival counters, self-resetting
};
__global__ void __launch_bounds__(256) k_xfer(const __grid_constant__ XferTable T)
{
   __shared__ int s_msg;
   if (threadIdx.x == 0) {
      int k = 0;
      while (k < T.nmsg - 1 && (int)blockIdx.x >= T.m[k].cta0 + T.m[k].nctas)
         ++k;
      s_msg = k;
      const int p = T.m[k].peer;
      if (T.wait_flag[p]) {
         while ((int)(*T.wait_flag[p] - T.wait_val[p]) < 0)
            __nanosleep(100);
         __threadfence_system();
      }
   }
   __syncthreads();
   const XferMsg M = T.m[s_msg];
   const int part = (int)blockIdx.x - M.cta0;
   // 16-byte lanes when both ends allow it (windows are 256-byte aligned; user arrays of 3-float atoms may not be),
   // 4-byte words otherwise; loads bypass L1 (the window was written by another GPU while this kernel was waiting)
   const bool a16 = (((size_t)M.src | (size_t)M.dst) & 15) == 0;
   const bool a4 = (((size_t)M.src | (size_t)M.dst | M.bytes) & 3) == 0;
   if (a16) {
      const size_t n16 = M.bytes / 16;
      const size_t per = (n16 + M.nctas - 1) / M.nctas;
      const size_t b0 = per * part, b1 = b0 + per < n16 ? b0 + per : n16;
      const uint4* s = reinterpret_cast<const uint4*>(M.src);
      uint4* d = reinterpret_cast<uint4*>(M.dst);
      size_t q = b0 + threadIdx.x;
      for (; q + 3 * blockDim.x < b1; q += 4 * blockDim.x) {      // four independent 16-byte transfers in flight per thread
         const uint4 v0 = __ldcg(s + q), v1 = __ldcg(s + q + blockDim.x), v2 = __ldcg(s + q + 2 * blockDim.x),
                     v3 = __ldcg(s + q + 3 * blockDim.x);
         d[q] = v0, d[q + blockDim.x] = v1, d[q + 2 * blockDim.x] = v2, d[q + 3 * blockDim.x] = v3;
      }
      for (; q < b1; q += blockDim.x)
         d[q] = __ldcg(s + q);
      if (part == M.nctas - 1)
         for (size_t r = n16 * 16 + threadIdx.x; r < M.bytes; r += blockDim.x)
            M.dst[r] = __ldcg(M.src + r);
   } else if (a4) {
      const size_t n4 = M.bytes / 4;
      const size_t per = (n4 + M.nctas - 1) / M.nctas;
      const size_t b0 = per * part, b1 = b0 + per < n4 ? b0 + per : n4;
      const unsigned* s = reinterpret_cast<const unsigned*>(M.src);
      unsigned* d = reinterpret_cast<unsigned*>(M.dst);
      for (size_t q = b0 + threadIdx.x; q < b1; q += blockDim.x)
         d[q] = __ldcg(s + q);
   } else {
      const size_t per = (M.bytes + M.nctas - 1) / M.nctas;
      const size_t b0 = per * part, b1 = b0 + per < M.bytes ? b0 + per : M.bytes;
      for (size_t q = b0 + threadIdx.x; q < b1; q += blockDim.x)
         M.dst[q] = __ldcg(M.src + q);
   }
   __threadfence_system();
   __syncthreads();
   if (threadIdx.x == 0) {
      const int p = M.peer;
      const unsigned old = atomicInc(&T.counters[p], (unsigned)T.peer_ctas[p] - 1);
      if (old == (unsigned)T.peer_ctas[p] - 1) {
         __threadfence_system();
         *T.set_flag[p] = T.set_val;
      }
   }
}

struct P2pComm : NcclComm {
   int fused = 0;                           // 1: k_xfer push/pull kernels instead of copy-engine copies + flag kernels
   unsigned* counters = nullptr;            // [32] device arrival counters (push: 0..15, pull: 16..31)
   size_t window = 0;                       // bytes one sender may put into one slot
   char* win = nullptr;                     // my receive windows: [sender][slot][window]
   unsigned* flg = nullptr;                 // my flags: ready[sender][slot], then ack[receiver][slot]
   std::vector<char*> peer_win;             // peers' windows / flags mapped into this process
   std::vector<unsigned*> peer_flg;
   unsigned seq = 0;
   bool ok = false;
   static size_t al(size_t b) { return (b + 255) / 256 * 256; }
   unsigned* ready_of(unsigned* base, int sender, int slot) const { return base + (sender * 2 + slot); }
   unsigned* ack_of(unsigned* base, int receiver, int slot) const { return base + 2 * world + (receiver * 2 + slot); }
   char* slot_of(char* base, int sender, int slot) const { return base + ((size_t)sender * 2 + slot) * window; }

   void setup(size_t window_bytes)
   {
      struct Handles {
         cudaIpcMemHandle_t w, f;
      };
      window = al(window_bytes);
      CUDA_CHECK(cudaMalloc(&win, (size_t)world * 2 * window));
      CUDA_CHECK(cudaMalloc(&flg, sizeof(unsigned) * 4 * world));
      CUDA_CHECK(cudaMemset(flg, 0, sizeof(unsigned) * 4 * world));
      CUDA_CHECK(cudaMalloc(&counters, sizeof(unsigned) * 32));
      CUDA_CHECK(cudaMemset(counters, 0, sizeof(unsigned) * 32));
      Handles mine;
      CUDA_CHECK(cudaIpcGetMemHandle(&mine.w, win));
      CUDA_CHECK(cudaIpcGetMemHandle(&mine.f, flg));
      std::vector<Handles> all(world);
      gather_bytes(&mine, sizeof(Handles), all.data());
      peer_win.assign(world, nullptr);
      peer_flg.assign(world, nullptr);
      for (int r = 0; r < world; ++r) {
         if (r == rank) {
            peer_win[r] = win;
            peer_flg[r] = flg;
            continue;
         }
         void* a = nullptr;
         void* b = nullptr;
         CUDA_CHECK(cudaIpcOpenMemHandle(&a, all[r].w, cudaIpcMemLazyEnablePeerAccess));
         CUDA_CHECK(cudaIpcOpenMemHandle(&b, all[r].f, cudaIpcMemLazyEnablePeerAccess));
         peer_win[r] = static_cast<char*>(a);
         peer_flg[r] = static_cast<unsigned*>(b);
      }
      ok = true;
   }
   ~P2pComm() override
   {
      if (!ok)
         return;
      cudaDeviceSynchronize();
      for (int r = 0; r < world; ++r)
         if (r != rank) {
            if (peer_win[r]) cudaIpcCloseMemHandle(peer_win[r]);
            if (peer_flg[r]) cudaIpcCloseMemHandle(peer_flg[r]);
         }
      // every peer must have unmapped before the owner frees: the communicator / rendezvous is still alive here
      try {
         char one = 0;
         std::vector<char> all(world);
         gather_bytes(&one, 1, all.data());
      } catch (...) {
      }
      cudaFree(win);
      cudaFree(flg);
      cudaFree(counters);
   }
   void exchange(const std::vector<Op>& sends, const std::vector<Op>& recvs, cudaStream_t st) override
   {
      if (!ok) {
         NcclComm::exchange(sends, recvs, st);
         return;
      }
      // per-peer totals decide the path of that pair; both ends see the same byte counts
      std::vector<size_t> tot_s(world, 0), tot_r(world, 0);
      for (const Op& o : sends)
         tot_s[o.peer] += al(o.bytes);
      for (const Op& o : recvs)
         tot_r[o.peer] += al(o.bytes);
      std::vector<Op> ns, nr;
      for (const Op& o : sends)
         if (tot_s[o.peer] > window)
            ns.push_back(o);
      for (const Op& o : recvs)
         if (tot_r[o.peer] > window)
            nr.push_back(o);
      ++seq;
      const int slot = (int)(seq & 1u);
      if (fused && sends.size() <= XFER_MAX_MSG && recvs.size() <= XFER_MAX_MSG) {
         exchange_fused(sends, recvs, tot_s, tot_r, ns, nr, slot, st);
         return;
      }
      // 1. my data into the peers' windows
      for (int p = 0; p < world; ++p) {
         if (p == rank || tot_s[p] == 0 || tot_s[p] > window)
            continue;
         if (seq > 2)      // the peer has emptied this slot (what I wrote two exchanges ago)
            k_flag_wait<<<1, 1, 0, st>>>(ack_of(flg, p, slot), last_used[p][slot]);
         size_t off = 0;
         char* dst = slot_of(peer_win[p], rank, slot);
         for (const Op& o : sends)
            if (o.peer == p) {
               if (o.bytes)
                  CUDA_CHECK(cudaMemcpyAsync(dst + off, o.ptr, o.bytes, cudaMemcpyDefault, st));
               off += al(o.bytes);
            }
         k_flag_set<<<1, 1, 0, st>>>(ready_of(peer_flg[p], rank, slot), seq);
         last_used[p][slot] = seq;
      }
      // 2. oversized pairs through NCCL
      if (!ns.empty() || !nr.empty())
         NcclComm::exchange(ns, nr, st);
      // 3. the peers' data out of my windows
      for (int p = 0; p < world; ++p) {
         if (p == rank || tot_r[p] == 0 || tot_r[p] > window)
            continue;
         k_flag_wait<<<1, 1, 0, st>>>(ready_of(flg, p, slot), seq);
         size_t off = 0;
         const char* src = slot_of(win, p, slot);
         for (const Op& o : recvs)
            if (o.peer == p) {
               if (o.bytes)
                  CUDA_CHECK(cudaMemcpyAsync(o.ptr, src + off, o.bytes, cudaMemcpyDeviceToDevice, st));
               off += al(o.bytes);
            }
         k_flag_set<<<1, 1, 0, st>>>(ack_of(peer_flg[p], rank, slot), seq);
      }
   }
   unsigned last_used[16][2] = {};

   int xfer_cap = 128;
   int ctas_for(size_t bytes) const { return (int)std::max<size_t>(1, std::min<size_t>((size_t)xfer_cap, bytes / (64 * 1024))); }

   void exchange_fused(const std::vector<Op>& sends, const std::vector<Op>& recvs, const std::vector<size_t>& tot_s,
      const std::vector<size_t>& tot_r, const std::vector<Op>& ns, const std::vector<Op>& nr, int slot, cudaStream_t st)
   {
      // push: everything I send, all peers in one kernel
      XferTable P;
      memset(&P, 0, sizeof(P));
      int grid = 0;
      std::vector<size_t> off(world, 0);
      for (const Op& o : sends) {
         const int p = o.peer;
         if (p == rank || tot_s[p] > window)
            continue;
         XferMsg& M = P.m[P.nmsg++];
         M.src = static_cast<const char*>(o.ptr);
         M.dst = slot_of(peer_win[p], rank, slot) + off[p];
         M.bytes = o.bytes;
         M.peer = p;
         M.cta0 = grid;
         M.nctas = ctas_for(o.bytes);
         grid += M.nctas;
         P.peer_ctas[p] += M.nctas;
         off[p] += al(o.bytes);
      }
      if (P.nmsg) {
         for (int p = 0; p < world; ++p) {
            if (!P.peer_ctas[p])
               continue;
            P.wait_flag[p] = ack_of(flg, p, slot);      // the peer emptied this slot (my write two exchanges ago)
            P.wait_val[p] = last_used[p][slot];
            P.set_flag[p] = ready_of(peer_flg[p], rank, slot);
            last_used[p][slot] = seq;
         }
         P.set_val = seq;
         P.counters = counters;
         k_xfer<<<grid, 256, 0, st>>>(P);
      }
      if (!ns.empty() || !nr.empty())
         NcclComm::exchange(ns, nr, st);
      // pull: everything I receive, out of my windows
      XferTable Q;
      memset(&Q, 0, sizeof(Q));
      grid = 0;
      std::fill(off.begin(), off.end(), 0);
      for (const Op& o : recvs) {
         const int p = o.peer;
         if (p == rank || tot_r[p] > window)
            continue;
         XferMsg& M = Q.m[Q.nmsg++];
         M.src = slot_of(win, p, slot) + off[p];
         M.dst = static_cast<char*>(o.ptr);
         M.bytes = o.bytes;
         M.peer = p;
         M.cta0 = grid;
         M.nctas = ctas_for(o.bytes);
         grid += M.nctas;
         Q.peer_ctas[p] += M.nctas;
         off[p] += al(o.bytes);
      }
      if (Q.nmsg) {
         for (int p = 0; p < world; ++p) {
            if (!Q.peer_ctas[p])
               continue;
            Q.wait_flag[p] = ready_of(flg, p, slot);
            Q.wait_val[p] = seq;
            Q.set_flag[p] = ack_of(peer_flg[p], rank, slot);
         }
         Q.set_val = seq;
         Q.counters = counters + 16;
         k_xfer<<<grid, 256, 0, st>>>(Q);
      }
   }
};

// ---- direct transport (APX_DIST_P2P=3, the default for 2..8 GPUs; also transport "direct", which has no NCCL at all)
//
// The windowed exchange above costs, per transpose of the slab FFT, a pack kernel, a push into the peer's window, a pull
// out of my window and an unpack kernel: the 36 MB block crosses HBM five times and NVLink once, in four launches with
// two cross-GPU flag waits (profiles/r02m_trace_water1m_n2.txt: 144 us of a 1550 us CG iteration per transpose, 24 % of
// the step in k_xfer).  Here the buffers an exchange lands in (PME planes, transposed slab, halo planes, the packed
// per-atom vectors) are REGISTERED: every rank maps the peers' allocations through CUDA IPC once, and ONE kernel per
// exchange writes every message straight to its final place in the peer's memory over NVLink -- strided chunk by chunk
// for the transposes (the transpose IS the address arithmetic of the copy), atom by atom for the per-atom halos (sorted
// indices are global, so the sender's index list is the receiver's).  Protocol, sequence numbers only grow:
//   arrive[from]  block 0 of my exchange kernel k tells every rank that will write to me "my stream has reached exchange
//                 k": every earlier kernel of mine has finished with the destination buffers, they may be overwritten;
//   (copy)        a CTA waits for its peer's arrive >= k, copies its share of the message, and the last CTA serving a
//                 peer raises
//   ready[from]   in the peer's memory; block 0 ends by waiting for ready >= k from every rank that writes to me, so the
//                 kernel that follows on my stream sees the data.
// No window, no second copy, no acknowledge round trip; self-addressed blocks are moved by the same kernel.  Small
// all-reduces (the solver's scalars) use the same flags: one CTA stores its values into a slot of every peer, raises the
// peer's flag, waits for the peers' and sums the slots in rank order -- every rank gets the same bits.
#define DX_MAX_MSG 24
#define DX_FLAG_ARRIVE 0
#define DX_FLAG_READY 16
#define DX_FLAG_AR 32
#define DX_FLAG_WORDS 64
#define DAR_MAX 512              // elements of one small all-reduce
struct DxMsg {
   const char* src;
   char* dst;
   unsigned long long src_stride, dst_stride;      // bytes between consecutive chunks
   unsigned chunk16;                               // kind 0: 16-byte units per chunk
   unsigned nchunks;                               // kind 0: chunks; kind 1: atoms
   const int* idx;                                 // kind 1: items of chunk16 units at item index idx[j] on both sides
   int peer, cta0, nctas, kind;
};
struct DxTable {
   DxMsg m[DX_MAX_MSG];
   int nmsg, rank, world;
   unsigned seq;
   unsigned recv_mask;                             // ranks that write to me in this exchange
   volatile unsigned* peer_flags[16];              // peers' flag blocks, mapped
   volatile unsigned* my_flags;
   int peer_ctas[16];
   unsigned* counters;                             // [16] arrival counters, self-resetting
   unsigned long long* trace;                      // APX_DX_TRACE: {entry, peers arrived (latest), copies done (latest), all data in} in ns
};

__device__ __forceinline__ unsigned long long dx_now()
{
   unsigned long long t;
   asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
   return t;
}
// a peer that never shows up must not hang the GPU: after 30 s the kernel traps and the host call fails
__device__ __forceinline__ void dx_wait(const volatile unsigned* f, unsigned want)
{
   if ((int)(*f - want) >= 0)
      return;
   const unsigned long long t0 = dx_now();
   while ((int)(*f - want) < 0) {
      __nanosleep(64);
      if (dx_now() - t0 > 30000000000ull)
         __trap();
   }
}

__global__ void __launch_bounds__(256) k_dxchg(const __grid_constant__ DxTable T)
{
   __shared__ int s_msg;
   const int tid = threadIdx.x;
   const bool gate = blockIdx.x == 0 && tid < T.world && ((T.recv_mask >> tid) & 1u);
   if (gate)
      T.peer_flags[tid][DX_FLAG_ARRIVE + T.rank] = T.seq;
   if (T.trace && blockIdx.x == 0 && tid == 0)
      T.trace[0] = dx_now();
   if (tid == 0) {
      int k = 0;
      while (k < T.nmsg - 1 && (int)blockIdx.x >= T.m[k].cta0 + T.m[k].nctas)
         ++k;
      s_msg = k;
      const int p = T.m[k].peer;
      if (T.nmsg > 0 && p != T.rank) {
         dx_wait(T.my_flags + DX_FLAG_ARRIVE + p, T.seq);
         __threadfence_system();
         if (T.trace)
            atomicMax(T.trace + 1, dx_now());
      }
   }
   __syncthreads();
   if (T.nmsg > 0) {
      const DxMsg& M = T.m[s_msg];
      const int part = (int)blockIdx.x - M.cta0;
      if (M.kind == 0) {
         // (messages are far below 64 GB: 32-bit unit counts keep the chunk / offset split a 32-bit division)
         const unsigned n16 = M.chunk16 * M.nchunks;
         const unsigned per = (n16 + M.nctas - 1) / M.nctas;
         const unsigned b0 = per * part, b1 = b0 + per < n16 ? b0 + per : n16;
         const unsigned c16 = M.chunk16;
         auto src_of = [&](unsigned q) {
            const unsigned ch = q / c16;
            return reinterpret_cast<const uint4*>(M.src + ch * M.src_stride) + (q - ch * c16);
         };
         auto dst_of = [&](unsigned q) {
            const unsigned ch = q / c16;
            return reinterpret_cast<uint4*>(M.dst + ch * M.dst_stride) + (q - ch * c16);
         };
         unsigned q = b0 + tid;
         const unsigned bd = blockDim.x;
         for (; q + 3 * bd < b1; q += 4 * bd) {      // four independent 16-byte transfers in flight per thread
            const uint4 v0 = __ldcg(src_of(q)), v1 = __ldcg(src_of(q + bd)), v2 = __ldcg(src_of(q + 2 * bd)), v3 = __ldcg(src_of(q + 3 * bd));
            *dst_of(q) = v0, *dst_of(q + bd) = v1, *dst_of(q + 2 * bd) = v2, *dst_of(q + 3 * bd) = v3;
         }
         for (; q < b1; q += bd)
            *dst_of(q) = __ldcg(src_of(q));
      } else {
         const unsigned c16 = M.chunk16;               // 16-byte units per atom: 2 (mixed build) or 4 (double build)
         const unsigned n2 = c16 * M.nchunks;
         const unsigned per = (n2 + M.nctas - 1) / M.nctas;
         const unsigned b0 = per * part, b1 = b0 + per < n2 ? b0 + per : n2;
         for (unsigned q = b0 + tid; q < b1; q += blockDim.x) {
            const unsigned a = q / c16;
            const size_t o = (size_t)c16 * M.idx[a] + (q - a * c16);
            reinterpret_cast<uint4*>(M.dst)[o] = __ldcg(reinterpret_cast<const uint4*>(M.src) + o);
         }
      }
      __threadfence_system();
      __syncthreads();
      if (T.trace && tid == 0)
         atomicMax(T.trace + 2, dx_now());
      if (tid == 0 && M.peer != T.rank) {
         const int p = M.peer;
         const unsigned old = atomicInc(&T.counters[p], (unsigned)T.peer_ctas[p] - 1);
         if (old == (unsigned)T.peer_ctas[p] - 1) {
            __threadfence_system();
            T.peer_flags[p][DX_FLAG_READY + T.rank] = T.seq;
         }
      }
   }
   if (gate) {
      dx_wait(T.my_flags + DX_FLAG_READY + tid, T.seq);
      __threadfence_system();
      if (T.trace)
         atomicMax(T.trace + 3, dx_now());
   }
}

// ---- the same exchange with the bulk-copy engine (TMA) doing the moving: one warp per CTA, whose lane 0 streams 16 KB tiles
// global -> shared (cp.async.bulk + mbarrier) -> global in the peer's memory (cp.async.bulk.global.shared), four tiles in flight.
// Measured (tools/probe/p2p_bw.cu, profiles/r02s_p2p_bw.txt): 16 such CTAs already move 18 MB at the 525 GB/s the 256-thread
// copy loop reaches with 148; in situ the wide CTAs of k_dxchg had to queue behind the resident CTAs of the real-space operator
// on stream2 (APX_DX_TRACE: the last CTA of a grid-halo exchange started 30 us after the first), a 32-thread CTA with no
// registers to speak of is scheduled at once and the copy engine does not compete for issue slots.
#define DXB_TILE 16384
#define DXB_STAGES 4
__device__ __forceinline__ unsigned dx_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(32) k_dxchg_bulk(const __grid_constant__ DxTable T)
{
   extern __shared__ __align__(128) char dxb_sm[];
   __shared__ unsigned long long bar[DXB_STAGES];
   const int tid = threadIdx.x;
   const bool gate = blockIdx.x == 0 && tid < T.world && ((T.recv_mask >> tid) & 1u);
   if (gate)
      T.peer_flags[tid][DX_FLAG_ARRIVE + T.rank] = T.seq;
   if (T.trace && blockIdx.x == 0 && tid == 0)
      T.trace[0] = dx_now();
   if (tid == 0 && T.nmsg > 0) {
      int k = 0;
      while (k < T.nmsg - 1 && (int)blockIdx.x >= T.m[k].cta0 + T.m[k].nctas)
         ++k;
      const DxMsg& M = T.m[k];
      const int p = M.peer;
      for (int s = 0; s < DXB_STAGES; ++s)
         asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dx_smem(&bar[s])));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      if (p != T.rank) {
         dx_wait(T.my_flags + DX_FLAG_ARRIVE + p, T.seq);
         __threadfence_system();
         if (T.trace)
            atomicMax(T.trace + 1, dx_now());
      }
      // tiles of this message: chunk c, tile j of the chunk (the last one of a chunk may be short)
      const unsigned cbytes = M.chunk16 * 16u;
      const unsigned tpc = (cbytes + DXB_TILE - 1) / DXB_TILE;
      const unsigned ntile = tpc * M.nchunks;
      const unsigned part = blockIdx.x - M.cta0, stride = M.nctas;
      auto tile_of = [&](unsigned t, const char*& sp, char*& dp, unsigned& nb) {
         const unsigned ch = t / tpc, j = t - ch * tpc;
         const unsigned off = j * DXB_TILE;
         nb = cbytes - off < DXB_TILE ? cbytes - off : DXB_TILE;
         sp = M.src + ch * M.src_stride + off;
         dp = M.dst + ch * M.dst_stride + off;
      };
      auto load = [&](unsigned t, int s) {
         const char* sp;
         char* dp;
         unsigned nb;
         tile_of(t, sp, dp, nb);
         asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dx_smem(&bar[s])), "r"(nb) : "memory");
         asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dx_smem(dxb_sm + s * DXB_TILE)),
            "l"(sp), "r"(nb), "r"(dx_smem(&bar[s])) : "memory");
      };
      unsigned phase = 0;      // bit s: parity the next wait on stage s expects
      unsigned issued = part;
      for (int s = 0; s < DXB_STAGES && issued < ntile; ++s, issued += stride)
         load(issued, s);
      int s = 0;
      for (unsigned t = part; t < ntile; t += stride) {
         unsigned done = 0;
         while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done)
                         : "r"(dx_smem(&bar[s])), "r"((phase >> s) & 1u) : "memory");
         phase ^= 1u << s;
         const char* sp;
         char* dp;
         unsigned nb;
         tile_of(t, sp, dp, nb);
         asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dp), "r"(dx_smem(dxb_sm + s * DXB_TILE)), "r"(nb) : "memory");
         asm volatile("cp.async.bulk.commit_group;" ::: "memory");
         if (issued < ntile) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the store has read the stage: refill it
            load(issued, s);
            issued += stride;
         }
         s = s + 1 == DXB_STAGES ? 0 : s + 1;
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      __threadfence_system();
      if (T.trace)
         atomicMax(T.trace + 2, dx_now());
      if (p != T.rank) {
         const unsigned old = atomicInc(&T.counters[p], (unsigned)T.peer_ctas[p] - 1);
         if (old == (unsigned)T.peer_ctas[p] - 1) {
            __threadfence_system();
            T.peer_flags[p][DX_FLAG_READY + T.rank] = T.seq;
         }
      }
   }
   __syncwarp();
   if (gate) {
      dx_wait(T.my_flags + DX_FLAG_READY + tid, T.seq);
      __threadfence_system();
      if (T.trace)
         atomicMax(T.trace + 3, dx_now());
   }
}

struct DarTable {
   char* peer_stage[16];                 // peers' staging areas, mapped: [parity][from][DAR_MAX] 8-byte slots
   const char* my_stage;
   volatile unsigned* peer_flags[16];
   volatile unsigned* my_flags;
   int rank, world;
   unsigned seq;
};
template <class V>
__global__ void __launch_bounds__(DAR_MAX) k_dar(V* __restrict__ data, int n, const __grid_constant__ DarTable A)
{
   const int tid = threadIdx.x;
   const size_t par = (size_t)(A.seq & 1u) * 16;
   V mine = 0;
   if (tid < n) {
      mine = data[tid];
      for (int p = 0; p < A.world; ++p)
         if (p != A.rank)
            reinterpret_cast<V*>(A.peer_stage[p] + (par + A.rank) * DAR_MAX * 8)[tid] = mine;
   }
   __threadfence_system();
   __syncthreads();
   if (tid < A.world && tid != A.rank) {
      A.peer_flags[tid][DX_FLAG_AR + A.rank] = A.seq;
      dx_wait(A.my_flags + DX_FLAG_AR + tid, A.seq);
      __threadfence_system();
   }
   __syncthreads();
   if (tid < n) {
      V s = 0;
      for (int r = 0; r < A.world; ++r)      // rank order: the same bits on every rank
         s += r == A.rank ? mine : reinterpret_cast<const volatile V*>(A.my_stage + (par + r) * DAR_MAX * 8)[tid];
      data[tid] = s;
   }
}

struct DirectComm : P2pComm {
   struct Region {
      char* base = nullptr;
      size_t size = 0;
      char* peer[16] = {};
   };
   struct Mapping {
      cudaIpcMemHandle_t h;
      char* p;
      int used;
   };
   std::vector<Region> regions;
   std::vector<Mapping> maps[16];
   std::vector<void*> wanted;           // pointers (anywhere inside their allocations) the next sync registers
   bool dirty = true;
   char* dmem = nullptr;                // my flag block + all-reduce staging, written by the peers
   char* peer_dmem[16] = {};
   unsigned* dcounters = nullptr;
   unsigned dseq = 0, arseq = 0;
   unsigned long long* trace = nullptr;      // APX_DX_TRACE=1: 4 timestamps per exchange, ring of TRACE_N
   static constexpr int TRACE_N = 4096;
   std::vector<int> trace_tag;
   int bulk = 1;                        // APX_DX_BULK: strided messages through the bulk-copy engine (k_dxchg_bulk)
   int bulk_ctas = 64;                  // APX_DX_BULK_CTAS: one-warp CTAs of one bulk exchange
   int total_ctas = 296;                // CTAs of one exchange kernel, shared out by bytes (measured at 2 GPUs: 296 16.6 ms, 592 16.9, 1184 17.2)
   bool dok = false;
   static constexpr size_t FLAG_BYTES = 256;
   static constexpr size_t STAGE_BYTES = (size_t)2 * 16 * DAR_MAX * 8;

   typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
   range_fn addr_range = nullptr;

   void dsetup()
   {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qr;
      if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn)
         APX_THROW("direct transport: cuMemGetAddressRange is not available");
      addr_range = reinterpret_cast<range_fn>(fn);
      CUDA_CHECK(cudaMalloc(&dmem, FLAG_BYTES + STAGE_BYTES));
      CUDA_CHECK(cudaMemset(dmem, 0, FLAG_BYTES + STAGE_BYTES));
      CUDA_CHECK(cudaMalloc(&dcounters, sizeof(unsigned) * 16));
      CUDA_CHECK(cudaMemset(dcounters, 0, sizeof(unsigned) * 16));
      CUDA_CHECK(cudaDeviceSynchronize());
      cudaIpcMemHandle_t mine;
      CUDA_CHECK(cudaIpcGetMemHandle(&mine, dmem));
      std::vector<cudaIpcMemHandle_t> all(world);
      gather_bytes(&mine, sizeof(mine), all.data());
      for (int r = 0; r < world; ++r) {
         if (r == rank) {
            peer_dmem[r] = dmem;
            continue;
         }
         void* a = nullptr;
         CUDA_CHECK(cudaIpcOpenMemHandle(&a, all[r], cudaIpcMemLazyEnablePeerAccess));
         peer_dmem[r] = static_cast<char*>(a);
      }
      if (const char* e = getenv("APX_DX_CTAS"))
         total_ctas = std::max(8, std::min(4096, atoi(e)));
      if (const char* e = getenv("APX_DX_BULK"))
         bulk = atoi(e);
      if (const char* e = getenv("APX_DX_BULK_CTAS"))
         bulk_ctas = std::max(2, std::min(1024, atoi(e)));
      CUDA_CHECK(cudaFuncSetAttribute(k_dxchg_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, DXB_STAGES * DXB_TILE));
      if (getenv("APX_DX_TRACE") && atoi(getenv("APX_DX_TRACE"))) {
         CUDA_CHECK(cudaMalloc(&trace, sizeof(unsigned long long) * 4 * TRACE_N));
         CUDA_CHECK(cudaMemset(trace, 0, sizeof(unsigned long long) * 4 * TRACE_N));
         trace_tag.assign(TRACE_N, -1);
      }
      dok = true;
   }
   ~DirectComm() override
   {
      if (!dok)
         return;
      cudaDeviceSynchronize();
      if (trace)
         dump_trace();
      for (int r = 0; r < world; ++r) {
         if (r == rank)
            continue;
         if (peer_dmem[r])
            cudaIpcCloseMemHandle(peer_dmem[r]);
         for (Mapping& m : maps[r])
            cudaIpcCloseMemHandle(m.p);
      }
      try {      // (owners free only after every peer has unmapped)
         char one = 0;
         std::vector<char> all(world);
         gather_bytes(&one, 1, all.data());
      } catch (...) {
      }
      cudaFree(dmem);
      cudaFree(dcounters);
   }

   // per kind of exchange (tag): mean us from kernel entry to "every peer has arrived", to "my copies are done", to "all data in"
   void dump_trace()
   {
      std::vector<unsigned long long> h(4 * (size_t)TRACE_N);
      if (cudaMemcpy(h.data(), trace, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost) != cudaSuccess)
         return;
      static const char* names[] = {"per-atom halo", "grid halo (forward)", "transpose (forward)", "transpose (inverse)", "grid halo (inverse)"};
      for (int tag = 0; tag < 5; ++tag) {
         double a = 0, b = 0, d = 0;
         int m = 0;
         const int last = (int)std::min<unsigned>(dseq, TRACE_N);
         for (int k = last / 2; k < last; ++k) {      // second half of the record: warm
            if (trace_tag[k] != tag || !h[4 * k] || !h[4 * k + 3])
               continue;
            const double t0 = (double)h[4 * k];
            a += (h[4 * k + 1] ? (double)h[4 * k + 1] - t0 : 0), b += (double)h[4 * k + 2] - t0, d += (double)h[4 * k + 3] - t0;
            ++m;
         }
         if (m)
            fprintf(stderr, "[apx dx trace] rank %d %-20s n=%4d  peers arrived +%6.1f us  copies done +%6.1f us  all data in +%6.1f us\n", rank,
               names[tag], m, 1e-3 * a / m, 1e-3 * b / m, 1e-3 * d / m);
      }
      cudaFree(trace);
      trace = nullptr;
   }

   // collective: (re)register the allocations behind `wanted`; mappings of allocations no region uses any more are closed
   void sync_regions()
   {
      struct Rec {
         cudaIpcMemHandle_t h;
         unsigned long long size;
         int valid;
      };
      const int nr = (int)wanted.size();
      std::vector<Rec> mine(nr), all((size_t)nr * world);
      regions.assign(nr, Region());
      for (int k = 0; k < nr; ++k) {
         memset(&mine[k], 0, sizeof(Rec));
         if (!wanted[k])
            continue;
         unsigned long long base = 0;
         size_t size = 0;
         if (addr_range(&base, &size, (unsigned long long)(uintptr_t)wanted[k]) != 0)
            APX_THROW("direct transport: pointer does not belong to a device allocation");
         regions[k].base = reinterpret_cast<char*>((uintptr_t)base);
         regions[k].size = size;
         CUDA_CHECK(cudaIpcGetMemHandle(&mine[k].h, regions[k].base));
         mine[k].size = size;
         mine[k].valid = 1;
      }
      CUDA_CHECK(cudaDeviceSynchronize());      // nothing of mine is in flight towards a mapping that is about to be closed
      if (nr)
         gather_bytes(mine.data(), sizeof(Rec) * nr, all.data());
      for (int r = 0; r < world; ++r) {
         for (Mapping& m : maps[r])
            m.used = 0;
         for (int k = 0; k < nr; ++k) {
            if (r == rank) {
               regions[k].peer[r] = regions[k].base;
               continue;
            }
            const Rec& R = all[(size_t)r * nr + k];
            if (!R.valid)
               continue;
            Mapping* hit = nullptr;
            for (Mapping& m : maps[r])
               if (memcmp(&m.h, &R.h, sizeof(R.h)) == 0)
                  hit = &m;
            if (!hit) {
               void* a = nullptr;
               CUDA_CHECK(cudaIpcOpenMemHandle(&a, R.h, cudaIpcMemLazyEnablePeerAccess));
               maps[r].push_back({R.h, static_cast<char*>(a), 0});
               hit = &maps[r].back();
            }
            hit->used = 1;
            regions[k].peer[r] = hit->p;
         }
         for (size_t q = 0; q < maps[r].size();)
            if (!maps[r][q].used) {
               cudaIpcCloseMemHandle(maps[r][q].p);
               maps[r].erase(maps[r].begin() + q);
            } else
               ++q;
      }
      dirty = false;
   }
   // the address, in `peer`'s memory, of what `mine` is in my memory (same offset in the peer's registered allocation)
   char* remote(const void* mine, int peer) const
   {
      const char* q = static_cast<const char*>(mine);
      for (const Region& R : regions)
         if (R.base && q >= R.base && q < R.base + R.size) {
            if (!R.peer[peer])
               break;
            return R.peer[peer] + (q - R.base);
         }
      APX_THROW("direct transport: destination buffer is not registered");
      return nullptr;
   }

   struct Msg {            // one message of an exchange, in terms of MY buffers: dst is translated to the peer's
      int peer;
      const void* src;
      void* dst;
      size_t chunk_bytes, nchunks, src_stride, dst_stride;
      const int* idx = nullptr;      // per-atom scatter (nchunks atoms of chunk_bytes each) when set
   };
   void xchg(const std::vector<Msg>& msgs, unsigned recv_mask, cudaStream_t st, int tag = 0)
   {
      if (dirty)
         sync_regions();
      ++dseq;
      if (msgs.empty() && !recv_mask)
         return;
      if ((int)msgs.size() > DX_MAX_MSG)
         APX_THROW("direct transport: too many messages in one exchange");
      DxTable T;
      memset(&T, 0, sizeof(T));
      size_t total = 0;
      bool strided = !msgs.empty();
      for (const Msg& m : msgs) {
         total += m.chunk_bytes * m.nchunks;
         strided = strided && !m.idx;
      }
      const bool use_bulk = bulk && strided;
      const size_t budget = use_bulk ? (size_t)bulk_ctas : (size_t)total_ctas;
      int grid = 0;
      for (const Msg& m : msgs) {
         DxMsg& M = T.m[T.nmsg++];
         const size_t bytes = m.chunk_bytes * m.nchunks;
         M.src = static_cast<const char*>(m.src);
         M.dst = m.peer == rank ? static_cast<char*>(m.dst) : remote(m.dst, m.peer);
         M.src_stride = m.src_stride, M.dst_stride = m.dst_stride;
         if (!m.idx && bytes / 16 >= 0xffffffffull)
            APX_THROW("direct transport: message too large");
         if (!m.idx && (m.chunk_bytes % 16 || m.src_stride % 16 || m.dst_stride % 16 || ((uintptr_t)M.src | (uintptr_t)M.dst) % 16))
            APX_THROW("direct transport: messages must be 16-byte aligned");
         M.chunk16 = (unsigned)(m.chunk_bytes / 16);
         M.nchunks = (unsigned)m.nchunks;
         M.idx = m.idx;
         M.kind = m.idx ? 1 : 0;
         M.peer = m.peer;
         M.cta0 = grid;
         const size_t gran = use_bulk ? (size_t)DXB_TILE * DXB_STAGES : 16384;
         M.nctas = (int)std::max<size_t>(1, std::min<size_t>(budget * bytes / std::max<size_t>(total, 1), bytes / gran + 1));
         grid += M.nctas;
         T.peer_ctas[m.peer] += M.nctas;
      }
      T.rank = rank, T.world = world, T.seq = dseq, T.recv_mask = recv_mask & ~(1u << rank);
      for (int p = 0; p < world; ++p)
         T.peer_flags[p] = reinterpret_cast<volatile unsigned*>(peer_dmem[p]);
      T.my_flags = reinterpret_cast<volatile unsigned*>(dmem);
      T.counters = dcounters;
      if (trace && dseq <= (unsigned)TRACE_N) {
         T.trace = trace + 4 * (size_t)(dseq - 1);
         trace_tag[dseq - 1] = tag;
      }
      if (use_bulk)
         k_dxchg_bulk<<<std::max(grid, 1), 32, DXB_STAGES * DXB_TILE, st>>>(T);
      else
         k_dxchg<<<std::max(grid, 1), 256, 0, st>>>(T);
   }

   void allreduce(void* p, size_t n, int dtype, cudaStream_t st) override
   {
      const size_t es = dtype == 2 ? 4 : 8;
      if (n > DAR_MAX && comm) {
         NcclComm::allreduce(p, n, dtype, st);
         return;
      }
      for (size_t o = 0; o < n; o += DAR_MAX) {      // (more than one pass only without NCCL: tests on small systems)
         const int m = (int)std::min<size_t>(DAR_MAX, n - o);
         DarTable A;
         memset(&A, 0, sizeof(A));
         for (int r = 0; r < world; ++r) {
            A.peer_stage[r] = peer_dmem[r] + FLAG_BYTES;
            A.peer_flags[r] = reinterpret_cast<volatile unsigned*>(peer_dmem[r]);
         }
         A.my_stage = dmem + FLAG_BYTES;
         A.my_flags = reinterpret_cast<volatile unsigned*>(dmem);
         A.rank = rank, A.world = world, A.seq = ++arseq;
         char* q = static_cast<char*>(p) + o * es;
         if (dtype == 0)
            k_dar<double><<<1, DAR_MAX, 0, st>>>((double*)q, m, A);
         else if (dtype == 1)
            k_dar<unsigned long long><<<1, DAR_MAX, 0, st>>>((unsigned long long*)q, m, A);
         else
            k_dar<int><<<1, DAR_MAX, 0, st>>>((int*)q, m, A);
      }
   }
};

// ---- in-process transport: `world` ranks = host threads of one process on one device
struct LocalHub {
   int world = 1;
   std::mutex m;
   std::condition_variable cv;
   int count = 0;
   long gen = 0;
   bool broken = false;
   std::vector<std::vector<ApxComm::Op>> box;     // [src * world + dst]
   std::vector<cudaEvent_t> ready, done;          // per rank
   std::vector<void*> stage;
   std::vector<size_t> stage_cap;
   explicit LocalHub(int w) : world(w), box((size_t)w * w), ready(w, nullptr), done(w, nullptr), stage(w, nullptr), stage_cap(w, 0) {}
   void barrier()
   {
      std::unique_lock<std::mutex> lk(m);
      if (broken)
         APX_THROW("local transport: a peer rank failed");
      long g = gen;
      if (++count == world) {
         count = 0;
         ++gen;
         cv.notify_all();
         return;
      }
      if (!cv.wait_for(lk, std::chrono::seconds(120), [&] { return gen != g || broken; })) {
         broken = true;
         cv.notify_all();
         APX_THROW("local transport: ranks did not reach the same collective within 120 s");
      }
      if (broken)
         APX_THROW("local transport: a peer rank failed");
   }
};

struct StagePtrs {
   const void* p[16];
};
template <class T>
__global__ void k_sum_stages(size_t n, int world, StagePtrs sp, T* __restrict__ out)
{
   size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n)
      return;
   T v = 0;
   for (int r = 0; r < world; ++r)        // fixed order: every rank gets the same bits
      v += static_cast<const T*>(sp.p[r])[i];
   out[i] = v;
}

struct LocalComm : ApxComm {
   LocalHub* hub = nullptr;
   void allreduce(void* p, size_t n, int dtype, cudaStream_t st) override
   {
      const size_t bytes = n * (dtype == 2 ? 4 : 8);
      if (world > 16)
         APX_THROW("local transport supports at most 16 ranks");
      if (hub->stage_cap[rank] < bytes) {
         if (hub->stage[rank])
            cudaFree(hub->stage[rank]);
         CUDA_CHECK(cudaMalloc(&hub->stage[rank], bytes + bytes / 4 + 256));
         hub->stage_cap[rank] = bytes + bytes / 4 + 256;
      }
      CUDA_CHECK(cudaMemcpyAsync(hub->stage[rank], p, bytes, cudaMemcpyDeviceToDevice, st));
      CUDA_CHECK(cudaEventRecord(hub->ready[rank], st));
      hub->barrier();
      StagePtrs sp;
      for (int r = 0; r < world; ++r) {
         sp.p[r] = hub->stage[r];
         if (r != rank)
            CUDA_CHECK(cudaStreamWaitEvent(st, hub->ready[r], 0));
      }
      const unsigned g = (unsigned)((n + 255) / 256);
      if (dtype == 0)
         k_sum_stages<double><<<g, 256, 0, st>>>(n, world, sp, (double*)p);
      else if (dtype == 1)
         k_sum_stages<unsigned long long><<<g, 256, 0, st>>>(n, world, sp, (unsigned long long*)p);
      else
         k_sum_stages<int><<<g, 256, 0, st>>>(n, world, sp, (int*)p);
      CUDA_CHECK(cudaEventRecord(hub->done[rank], st));
      hub->barrier();
      for (int r = 0; r < world; ++r)
         if (r != rank)
            CUDA_CHECK(cudaStreamWaitEvent(st, hub->done[r], 0));
   }
   void exchange(const std::vector<Op>& sends, const std::vector<Op>& recvs, cudaStream_t st) override
   {
      for (int d = 0; d < world; ++d)
         hub->box[(size_t)rank * world + d].clear();
      for (const Op& o : sends)
         hub->box[(size_t)rank * world + o.peer].push_back(o);
      CUDA_CHECK(cudaEventRecord(hub->ready[rank], st));
      hub->barrier();
      std::vector<int> taken(world, 0);
      std::vector<char> waited(world, 0);
      for (const Op& o : recvs) {
         auto& b = hub->box[(size_t)o.peer * world + rank];
         if (taken[o.peer] >= (int)b.size() || b[taken[o.peer]].bytes != o.bytes)
            APX_THROW("local transport: unmatched message");
         if (!waited[o.peer]) {
            CUDA_CHECK(cudaStreamWaitEvent(st, hub->ready[o.peer], 0));
            waited[o.peer] = 1;
         }
         CUDA_CHECK(cudaMemcpyAsync(o.ptr, b[taken[o.peer]].ptr, o.bytes, cudaMemcpyDeviceToDevice, st));
         taken[o.peer]++;
      }
      CUDA_CHECK(cudaEventRecord(hub->done[rank], st));
      hub->barrier();
      // my send buffers may be rewritten only after the receivers have copied them
      std::fill(waited.begin(), waited.end(), 0);
      for (const Op& o : sends)
         if (!waited[o.peer]) {
            CUDA_CHECK(cudaStreamWaitEvent(st, hub->done[o.peer], 0));
            waited[o.peer] = 1;
         }
   }
};

// self-addressed messages never reach a transport
void comm_exchange(apx_ctx* c, const std::vector<ApxComm::Op>& sends, const std::vector<ApxComm::Op>& recvs, cudaStream_t st)
{
   ApxComm* cm = c->dist.comm;
   std::vector<ApxComm::Op> s2, r2, ss, rs;
   for (auto& o : sends)
      (o.peer == cm->rank ? ss : s2).push_back(o);
   for (auto& o : recvs)
      (o.peer == cm->rank ? rs : r2).push_back(o);
   if (ss.size() != rs.size())
      APX_THROW("self messages do not pair up");
   for (size_t k = 0; k < ss.size(); ++k)
      if (ss[k].bytes)
         CUDA_CHECK(cudaMemcpyAsync(rs[k].ptr, ss[k].ptr, ss[k].bytes, cudaMemcpyDeviceToDevice, st));
   cm->exchange(s2, r2, st);
}

double frac_dist_to_slab(double w, int r, int world)
{
   // periodic distance (in units of the cell's third fractional coordinate) from w to [r/G, (r+1)/G)
   const double lo = (double)r / world, hi = (double)(r + 1) / world;
   if (w >= lo && w < hi)
      return 0.0;
   auto pd = [](double x) {
      x = fabs(x);
      x -= floor(x);
      return std::min(x, 1.0 - x);
   };
   return std::min(pd(w - lo), pd(w - hi));
}
} // namespace

// Halo plan shared by every rank: for sorted atom s with grid coordinate w3s[s], owned by the rank whose
// bounds contain s, rank r needs it when its distance to r's slab is <= range_frac.  Fills, for
// `rank`, the atoms it sends to every peer and the atoms it receives from every peer (sorted
// indices, ascending, concatenated by peer; *_off has world+1 entries).  Pure host code.
static void plan_halo(int n, const float* w3s, const int* bounds, int world, int rank, double range_frac, std::vector<int>& send_idx,
   std::vector<int>& send_off, std::vector<int>& recv_idx, std::vector<int>& recv_off)
{
   std::vector<std::vector<int>> snd(world), rcv(world);
   for (int owner = 0; owner < world; ++owner)
      for (int s = bounds[owner]; s < bounds[owner + 1]; ++s) {
         const double w = w3s[s];
         if (owner == rank) {
            for (int r = 0; r < world; ++r)
               if (r != rank && frac_dist_to_slab(w, r, world) <= range_frac)
                  snd[r].push_back(s);
         } else if (frac_dist_to_slab(w, rank, world) <= range_frac) {
            rcv[owner].push_back(s);
         }
      }
   send_idx.clear(), recv_idx.clear();
   send_off.assign(world + 1, 0), recv_off.assign(world + 1, 0);
   for (int r = 0; r < world; ++r) {
      send_idx.insert(send_idx.end(), snd[r].begin(), snd[r].end());
      recv_idx.insert(recv_idx.end(), rcv[r].begin(), rcv[r].end());
      send_off[r + 1] = (int)send_idx.size();
      recv_off[r + 1] = (int)recv_idx.size();
   }
}

namespace {
__global__ void k_halo_pack(int m, const int* __restrict__ idx, const real4* __restrict__ V, real4* __restrict__ buf)
{
   int j = blockIdx.x * blockDim.x + threadIdx.x;
   if (j >= 2 * m)
      return;
   buf[j] = V[2 * (size_t)idx[j >> 1] + (j & 1)];
}
__global__ void k_halo_unpack(int m, const int* __restrict__ idx, const real4* __restrict__ buf, real4* __restrict__ V)
{
   int j = blockIdx.x * blockDim.x + threadIdx.x;
   if (j >= 2 * m)
      return;
   V[2 * (size_t)idx[j >> 1] + (j & 1)] = buf[j];
}

__global__ void k_grid_add(size_t m, const cplx* __restrict__ src, cplx* __restrict__ dst)
{
   size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= m)
      return;
   cplx a = dst[i], b = src[i];
   a.x += b.x;
   a.y += b.y;
   dst[i] = a;
}

// planes [z][y][x] of this rank -> blocks [r][z][y local to r][x] for the transpose, and back
__global__ void k_transpose_pack(int pz, int n2, int n1, int py, const cplx* __restrict__ planes, cplx* __restrict__ buf)
{
   size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   size_t tot = (size_t)pz * n2 * n1;
   if (i >= tot)
      return;
   int x = (int)(i % n1);
   int y = (int)((i / n1) % n2);
   int z = (int)(i / ((size_t)n1 * n2));
   int r = y / py, yl = y - r * py;
   buf[(((size_t)r * pz + z) * py + yl) * n1 + x] = planes[i];
}
__global__ void k_transpose_unpack(int pz, int n2, int n1, int py, const cplx* __restrict__ buf, cplx* __restrict__ planes)
{
   size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   size_t tot = (size_t)pz * n2 * n1;
   if (i >= tot)
      return;
   int x = (int)(i % n1);
   int y = (int)((i / n1) % n2);
   int z = (int)(i / ((size_t)n1 * n2));
   int r = y / py, yl = y - r * py;
   planes[i] = buf[(((size_t)r * pz + z) * py + yl) * n1 + x];
}

// both halo-plane sums of the forward transform in one launch
__global__ void k_grid_add2(size_t m1, const cplx* __restrict__ s1, cplx* __restrict__ d1, size_t m2, const cplx* __restrict__ s2,
   cplx* __restrict__ d2)
{
   size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   const cplx* s = s1;
   cplx* d = d1;
   if (i >= m1) {
      i -= m1;
      if (i >= m2)
         return;
      s = s2, d = d2;
   }
   cplx a = d[i], b = s[i];
   a.x += b.x;
   a.y += b.y;
   d[i] = a;
}

// the direct transport of this context, with the buffers its exchanges land in registered (a collective re-registration
// follows whenever one of them has moved: sizes, and with them reallocations, are the same on every rank)
DirectComm* direct_of(apx_ctx* c)
{
   DirectComm* dc = dynamic_cast<DirectComm*>(c->dist.comm);
   if (!dc || !dc->dok)
      return nullptr;
   void* w[] = {c->qgrid.p, c->dist.tbuf.p, c->dist.hbuf.p, c->pk_p.p, c->pk_r.p};
   const size_t nw = sizeof(w) / sizeof(w[0]);
   if (dc->wanted.size() != nw || !std::equal(w, w + nw, dc->wanted.begin())) {
      dc->wanted.assign(w, w + nw);
      dc->dirty = true;
   }
   return dc;
}

inline void exec_fft(cufftHandle plan, cplx* p, int dir)
{
#ifdef APX_DOUBLE
   CUFFT_CHECK(cufftExecZ2Z(plan, p, p, dir));
#else
   CUFFT_CHECK(cufftExecC2C(plan, p, p, dir));
#endif
}
} // namespace

// ------------------------------------------------------------------------------------------------
// ownership + halo plan, at every list rebuild (after the sort)
// ------------------------------------------------------------------------------------------------
void apx_dist_after_sort(apx_ctx* c)
{
   DistState& D = c->dist;
   const int n = c->n, G = D.world;
   std::vector<unsigned> keys(n);
   std::vector<int> perm(n);
   std::vector<real> w3(n);
   CUDA_CHECK(cudaMemcpyAsync(keys.data(), c->sortkey2.p, sizeof(unsigned) * n, cudaMemcpyDeviceToHost, c->stream));
   CUDA_CHECK(cudaMemcpyAsync(perm.data(), c->perm.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
   CUDA_CHECK(cudaMemcpyAsync(w3.data(), c->w3.p, sizeof(real) * n, cudaMemcpyDeviceToHost, c->stream));
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
   D.bounds.assign(G + 1, -1);
   D.bounds[G] = n;
   for (int s = n - 1; s >= 0; --s)
      D.bounds[keys[s] >> 27] = s;           // first index of every non-empty slab
   for (int g = G - 1; g >= 0; --g)
      if (D.bounds[g] < 0)
         D.bounds[g] = D.bounds[g + 1];
   D.bounds[0] = 0;
   c->a0 = D.bounds[D.rank];
   c->a1 = D.bounds[D.rank + 1];
   std::vector<float> w3s(n);
   for (int s = 0; s < n; ++s)
      w3s[s] = (float)w3[perm[s]];
   // list range in units of the third fractional coordinate: planes of constant w3 are 1/|recip_c| apart
   const double r3 = sqrt((double)c->box.r[6] * c->box.r[6] + (double)c->box.r[7] * c->box.r[7] + (double)c->box.r[8] * c->box.r[8]);
   const double range_frac = ((double)c->list_cutoff + (double)c->list_buffer) * r3 * (1.0 + 1e-6) + 1e-7;
   std::vector<int> si, ri;
   plan_halo(n, w3s.data(), D.bounds.data(), G, D.rank, range_frac, si, D.send_off, ri, D.recv_off);
   D.send_idx.ensure(si.size() + 1);
   D.recv_idx.ensure(ri.size() + 1);
   if (!si.empty())
      CUDA_CHECK(cudaMemcpyAsync(D.send_idx.p, si.data(), sizeof(int) * si.size(), cudaMemcpyHostToDevice, c->stream));
   if (!ri.empty())
      CUDA_CHECK(cudaMemcpyAsync(D.recv_idx.p, ri.data(), sizeof(int) * ri.size(), cudaMemcpyHostToDevice, c->stream));
   CUDA_CHECK(cudaStreamSynchronize(c->stream));
   D.sendbuf.ensure(2 * si.size() + 2);
   D.recvbuf.ensure(2 * ri.size() + 2);
   D.halo_atoms = (long long)ri.size();
}

int apx_dist_prof_begin(apx_ctx* c, int kind, cudaStream_t st)
{
   DistState& D = c->dist;
   if (!D.prof_on)
      return -1;
   if (D.prof_used + 2 > (int)D.prof_ev.size()) {
      const size_t old = D.prof_ev.size();
      D.prof_ev.resize(old + 256);
      D.prof_kind.resize(old + 256);
      for (size_t k = old; k < D.prof_ev.size(); ++k)
         CUDA_CHECK(cudaEventCreate(&D.prof_ev[k]));
   }
   const int slot = D.prof_used;
   D.prof_used += 2;
   D.prof_kind[slot] = kind;
   cudaEventRecord(D.prof_ev[slot], st);
   return slot;
}
void apx_dist_prof_end(apx_ctx* c, int slot, cudaStream_t st)
{
   if (slot >= 0)
      cudaEventRecord(c->dist.prof_ev[slot + 1], st);
}

// V[halo atoms] <- the owners' values; V is a packed (d,p) array (2 real4 per atom, dp.cuh)
void apx_dist_halo(apx_ctx* c, real4* V, cudaStream_t st)
{
   DistState& D = c->dist;
   const int G = D.world;
   const int pslot = apx_dist_prof_begin(c, 0, st);
   const int ns = D.send_off[G], nr = D.recv_off[G];
   if (DirectComm* dc = direct_of(c)) {
      // every halo atom goes straight from my V to the same slot of the neighbour's V
      std::vector<DirectComm::Msg> msgs;
      unsigned recv_mask = 0;
      for (int r = 0; r < G; ++r) {
         if (D.send_off[r + 1] > D.send_off[r]) {
            DirectComm::Msg m{r, V, V, 2 * sizeof(real4), (size_t)(D.send_off[r + 1] - D.send_off[r]), 0, 0};
            m.idx = D.send_idx.p + D.send_off[r];
            msgs.push_back(m);
         }
         if (D.recv_off[r + 1] > D.recv_off[r])
            recv_mask |= 1u << r;
      }
      dc->xchg(msgs, recv_mask, st);
      APX_COUNT_LAUNCH(c);
      apx_dist_prof_end(c, pslot, st);
      return;
   }
   if (ns > 0) {
      k_halo_pack<<<(2 * ns + 255) / 256, 256, 0, st>>>(ns, D.send_idx, V, D.sendbuf);
      APX_COUNT_LAUNCH(c);
   }
   std::vector<ApxComm::Op> sends, recvs;
   for (int r = 0; r < G; ++r) {
      if (D.send_off[r + 1] > D.send_off[r])
         sends.push_back({r, D.sendbuf.p + 2 * (size_t)D.send_off[r], sizeof(real4) * 2 * (size_t)(D.send_off[r + 1] - D.send_off[r])});
      if (D.recv_off[r + 1] > D.recv_off[r])
         recvs.push_back({r, D.recvbuf.p + 2 * (size_t)D.recv_off[r], sizeof(real4) * 2 * (size_t)(D.recv_off[r + 1] - D.recv_off[r])});
   }
   comm_exchange(c, sends, recvs, st);
   if (nr > 0) {
      k_halo_unpack<<<(2 * nr + 255) / 256, 256, 0, st>>>(nr, D.recv_idx, D.recvbuf, V);
      APX_COUNT_LAUNCH(c);
   }
   apx_dist_prof_end(c, pslot, st);
}

void apx_dist_allreduce_f64(apx_ctx* c, double* p, size_t n)
{
   const int pslot = apx_dist_prof_begin(c, 3, c->stream);
   c->dist.comm->allreduce(p, n, 0, c->stream);
   apx_dist_prof_end(c, pslot, c->stream);
}
void apx_dist_allreduce_u64(apx_ctx* c, unsigned long long* p, size_t n) { c->dist.comm->allreduce(p, n, 1, c->stream); }
void apx_dist_allreduce_i32(apx_ctx* c, int* p, size_t n) { c->dist.comm->allreduce(p, n, 2, c->stream); }

// every rank receives the owned ranges of the others (sorted-order arrays with a fixed stride per atom)
void apx_dist_share_owned(apx_ctx* c, void* base, size_t bpa)
{
   DistState& D = c->dist;
   std::vector<ApxComm::Op> sends, recvs;
   char* b = static_cast<char*>(base);
   for (int r = 0; r < D.world; ++r) {
      if (r == D.rank)
         continue;
      if (c->a1 > c->a0)
         sends.push_back({r, b + bpa * (size_t)c->a0, bpa * (size_t)(c->a1 - c->a0)});
      if (D.bounds[r + 1] > D.bounds[r])
         recvs.push_back({r, b + bpa * (size_t)D.bounds[r], bpa * (size_t)(D.bounds[r + 1] - D.bounds[r])});
   }
   comm_exchange(c, sends, recvs, c->stream);
}

// ------------------------------------------------------------------------------------------------
// slab-decomposed PME grid
// ------------------------------------------------------------------------------------------------
void apx_dist_pme_setup(apx_ctx* c)
{
   DistState& D = c->dist;
   const int G = D.world, n1 = c->nfft1, n2 = c->nfft2, n3 = c->nfft3;
   if (n3 % G || n2 % G)
      APX_THROW("slab PME: nfft2 and nfft3 must be multiples of the number of GPUs");
   D.pz = n3 / G;
   D.z0 = D.rank * D.pz;
   D.py = n2 / G;
   D.y0 = D.rank * D.py;
   // an atom's stencil covers planes ii-4 .. ii with ii the plane of its w3; between list rebuilds it
   // may drift buffer/2, i.e. m planes, in either direction
   const double r3 = sqrt((double)c->box.r[6] * c->box.r[6] + (double)c->box.r[7] * c->box.r[7] + (double)c->box.r[8] * c->box.r[8]);
   const int m = (int)ceil(0.5 * c->opt.list_buffer * n3 * r3) + 1;
   D.hl = 4 + m;
   D.hu = m;
   if (D.hl > D.pz || D.pz + D.hl + D.hu > n3)
      APX_THROW("slab PME: slabs of " + std::to_string(D.pz) + " planes are thinner than the " + std::to_string(D.hl)
         + "-plane halo; use fewer GPUs or a finer grid");
   c->zbase = ((D.z0 - D.hl) % n3 + n3) % n3;
   c->nzl = D.pz + D.hl + D.hu;
   c->qy0 = D.y0;
   c->qny = D.py;
   if (D.plans_ok) {
      cufftDestroy(D.plan2d);
      cufftDestroy(D.plan1d);
      D.plans_ok = 0;
   }
#ifdef APX_DOUBLE
   const cufftType ty = CUFFT_Z2Z;
#else
   const cufftType ty = CUFFT_C2C;
#endif
   int dims2[2] = {n2, n1};
   CUFFT_CHECK(cufftPlanMany(&D.plan2d, 2, dims2, nullptr, 1, n2 * n1, nullptr, 1, n2 * n1, ty, D.pz));
   int dims1[1] = {n3};
   int emb[1] = {n3};
   CUFFT_CHECK(cufftPlanMany(&D.plan1d, 1, dims1, emb, D.py * n1, 1, emb, D.py * n1, 1, ty, D.py * n1));
   CUFFT_CHECK(cufftSetStream(D.plan2d, c->stream));
   CUFFT_CHECK(cufftSetStream(D.plan1d, c->stream));
   D.plans_ok = 1;
   const size_t slab = (size_t)n3 * D.py * n1;
   D.tbuf.ensure(slab);
   D.sbuf.ensure(slab);
   D.hbuf.ensure((size_t)(D.hl + D.hu) * n2 * n1);
}

void apx_dist_pme_destroy(apx_ctx* c)
{
   DistState& D = c->dist;
   if (D.plans_ok) {
      cufftDestroy(D.plan2d);
      cufftDestroy(D.plan1d);
      D.plans_ok = 0;
   }
}

// local grid (spread contributions of my atoms on my planes + halo planes) -> tb = [k3][k2 local][k1]
void apx_dist_fft_forward(apx_ctx* c, cplx* tb)
{
   DistState& D = c->dist;
   const int G = D.world, n1 = c->nfft1, n2 = c->nfft2;
   const size_t plane = (size_t)n1 * n2;
   const int prev = (D.rank + G - 1) % G, next = (D.rank + 1) % G;
   cudaStream_t st = c->stream;
   cplx* g = c->qgrid.p;
   const int pslot = apx_dist_prof_begin(c, 1, st);
   if (DirectComm* dc = direct_of(c)) {
      // 1. my halo planes land in the neighbours' hbuf and are summed into the planes they belong to
      const size_t pb = plane * sizeof(cplx);
      std::vector<DirectComm::Msg> h = {{prev, g, D.hbuf.p, D.hl * pb, 1, 0, 0},
         {next, g + (size_t)(D.hl + D.pz) * plane, D.hbuf.p + D.hl * plane, D.hu * pb, 1, 0, 0}};
      dc->xchg(h, (1u << prev) | (1u << next), st, 1);
      const size_t m1 = D.hl * plane, m2 = D.hu * plane;
      k_grid_add2<<<(unsigned)((m1 + m2 + 255) / 256), 256, 0, st>>>(m1, D.hbuf.p, g + (size_t)D.pz * plane, m2, D.hbuf.p + D.hl * plane,
         g + (size_t)D.hl * plane);
      cplx* mine = g + (size_t)D.hl * plane;
      // 2. 2-D transforms of my planes; 3. the transpose is the address arithmetic of ONE copy kernel: rows y of rank r's
      // share of every plane go to [k3 = my planes][y local to r][x] of r's slab; 4. 1-D transforms along z
      exec_fft(D.plan2d, mine, CUFFT_FORWARD);
      const size_t chunk = (size_t)D.py * n1 * sizeof(cplx), blk = (size_t)D.pz * D.py * n1;
      std::vector<DirectComm::Msg> t;
      for (int q = 0; q < G; ++q) {
         const int r = (D.rank + q) % G;      // my own block first, then the peers in ring order: no two ranks start on the same target
         t.push_back({r, mine + (size_t)r * D.py * n1, tb + (size_t)D.rank * blk, chunk, (size_t)D.pz, pb, chunk});
      }
      dc->xchg(t, (1u << G) - 1u, st, 2);
      exec_fft(D.plan1d, tb, CUFFT_FORWARD);
      c->stats.kernel_launches += 6;
      apx_dist_prof_end(c, pslot, st);
      return;
   }
   // 1. halo planes go to the slabs they belong to and are summed there
   {
      std::vector<ApxComm::Op> sends = {{prev, g, D.hl * plane * sizeof(cplx)}, {next, g + (size_t)(D.hl + D.pz) * plane, D.hu * plane * sizeof(cplx)}};
      std::vector<ApxComm::Op> recvs = {{next, D.hbuf.p, D.hl * plane * sizeof(cplx)}, {prev, D.hbuf.p + D.hl * plane, D.hu * plane * sizeof(cplx)}};
      comm_exchange(c, sends, recvs, st);
      size_t m1 = D.hl * plane, m2 = D.hu * plane;
      k_grid_add<<<(unsigned)((m1 + 255) / 256), 256, 0, st>>>(m1, D.hbuf.p, g + (size_t)D.pz * plane);         // my top hl planes
      k_grid_add<<<(unsigned)((m2 + 255) / 256), 256, 0, st>>>(m2, D.hbuf.p + D.hl * plane, g + (size_t)D.hl * plane);   // my bottom hu planes
   }
   cplx* mine = g + (size_t)D.hl * plane;
   // 2. 2-D transforms of my planes, 3. transpose, 4. 1-D transforms along z
   exec_fft(D.plan2d, mine, CUFFT_FORWARD);
   const size_t tot = (size_t)D.pz * plane, blk = (size_t)D.pz * D.py * n1;
   k_transpose_pack<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(D.pz, n2, n1, D.py, mine, D.sbuf);
   {
      std::vector<ApxComm::Op> sends, recvs;
      for (int r = 0; r < G; ++r) {
         sends.push_back({r, D.sbuf.p + r * blk, blk * sizeof(cplx)});
         recvs.push_back({r, tb + r * blk, blk * sizeof(cplx)});
      }
      comm_exchange(c, sends, recvs, st);
   }
   exec_fft(D.plan1d, tb, CUFFT_FORWARD);
   c->stats.kernel_launches += 3;
   apx_dist_prof_end(c, pslot, st);
}

// tb -> potential on my planes and on the halo planes my atoms' stencils reach
void apx_dist_fft_inverse(apx_ctx* c, cplx* tb)
{
   DistState& D = c->dist;
   const int G = D.world, n1 = c->nfft1, n2 = c->nfft2;
   const size_t plane = (size_t)n1 * n2;
   const int prev = (D.rank + G - 1) % G, next = (D.rank + 1) % G;
   cudaStream_t st = c->stream;
   cplx* g = c->qgrid.p;
   cplx* mine = g + (size_t)D.hl * plane;
   const int pslot = apx_dist_prof_begin(c, 2, st);
   if (DirectComm* dc = direct_of(c)) {
      exec_fft(D.plan1d, tb, CUFFT_INVERSE);
      const size_t pb = plane * sizeof(cplx);
      const size_t chunk = (size_t)D.py * n1 * sizeof(cplx), blk = (size_t)D.pz * D.py * n1;
      std::vector<DirectComm::Msg> t;
      for (int q = 0; q < G; ++q) {
         const int r = (D.rank + q) % G;
         t.push_back({r, tb + (size_t)r * blk, mine + (size_t)D.rank * D.py * n1, chunk, (size_t)D.pz, chunk, pb});
      }
      dc->xchg(t, (1u << G) - 1u, st, 3);
      exec_fft(D.plan2d, mine, CUFFT_INVERSE);
      // my top hl planes are the low halo of the next slab, my bottom hu planes the high halo of the previous one
      std::vector<DirectComm::Msg> h = {{next, g + (size_t)D.pz * plane, g, D.hl * pb, 1, 0, 0},
         {prev, mine, g + (size_t)(D.hl + D.pz) * plane, D.hu * pb, 1, 0, 0}};
      dc->xchg(h, (1u << prev) | (1u << next), st, 4);
      c->stats.kernel_launches += 5;
      apx_dist_prof_end(c, pslot, st);
      return;
   }
   exec_fft(D.plan1d, tb, CUFFT_INVERSE);
   const size_t tot = (size_t)D.pz * plane, blk = (size_t)D.pz * D.py * n1;
   {
      std::vector<ApxComm::Op> sends, recvs;
      for (int r = 0; r < G; ++r) {
         sends.push_back({r, tb + r * blk, blk * sizeof(cplx)});
         recvs.push_back({r, D.sbuf.p + r * blk, blk * sizeof(cplx)});
      }
      comm_exchange(c, sends, recvs, st);
   }
   k_transpose_unpack<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(D.pz, n2, n1, D.py, D.sbuf, mine);
   exec_fft(D.plan2d, mine, CUFFT_INVERSE);
   {
      // my top hl planes are the low halo of the next slab, my bottom hu planes the high halo of the previous one
      std::vector<ApxComm::Op> sends = {{next, g + (size_t)D.pz * plane, D.hl * plane * sizeof(cplx)}, {prev, mine, D.hu * plane * sizeof(cplx)}};
      std::vector<ApxComm::Op> recvs = {{prev, g, D.hl * plane * sizeof(cplx)}, {next, g + (size_t)(D.hl + D.pz) * plane, D.hu * plane * sizeof(cplx)}};
      comm_exchange(c, sends, recvs, st);
   }
   c->stats.kernel_launches += 1;
   apx_dist_prof_end(c, pslot, st);
}

void apx_dist_destroy(apx_ctx* c)
{
   DistState& D = c->dist;
   for (auto& e : D.prof_ev)
      cudaEventDestroy(e);
   D.prof_ev.clear();
   apx_dist_pme_destroy(c);
   delete D.comm;
   D.comm = nullptr;
   D.send_idx.release(), D.recv_idx.release(), D.sendbuf.release(), D.recvbuf.release();
   D.tbuf.release(), D.sbuf.release(), D.tbuf2.release(), D.hbuf.release();
}

// ------------------------------------------------------------------------------------------------
// C ABI (declared in include/apx.h)
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_dist_err;
const char* apx_dist_error() { return g_dist_err.c_str(); }

ApxComm* apx_make_nccl_comm(int rank, int world, const char* lib, const void* unique_id)
{
   NcclApi* api = nccl_api(lib);
   // measured on the 1 M-atom box (profiles/r01p..., r01s..., r01v..x): 2 GPUs 22.3 ms over NCCL, 21.5 with copy-engine
   // windows (mode 1), 21.3 with the fused push/pull kernels (mode 2); 4 GPUs 13.0 (NCCL), 15.3 (mode 1: the per-peer copies
   // of one exchange serialise on one stream), 13.0 (mode 2).  Default: mode 2 at 2 GPUs, NCCL beyond (8 GPUs not yet run).
   // Round 2: mode 3 (direct transport, one kernel per exchange writing into the peers' registered buffers) is the default.
   int p2p = 3;
   if (const char* e = getenv("APX_DIST_P2P"))
      p2p = atoi(e);
   DirectComm* dcm = p2p >= 3 && world <= 16 ? new DirectComm() : nullptr;
   P2pComm* cm = dcm ? dcm : new P2pComm();
   cm->api = api;
   cm->rank = rank;
   cm->world = world;
   ncclUniqueId id;
   memcpy(&id, unique_id, sizeof(id));
   NCCL_CHECK(api, api->CommInitRank(&cm->comm, world, id, rank));
   if (p2p && world <= 16) {
      cm->fused = p2p >= 2 ? 1 : 0;
      if (const char* e = getenv("APX_DIST_XFER_CTAS"))
         cm->xfer_cap = std::max(1, std::min(1024, atoi(e)));
      size_t mb = 64;
      if (const char* e = getenv("APX_DIST_WINDOW_MB"))
         mb = (size_t)std::max(1, atoi(e));
      cm->setup(mb << 20);
      if (dcm)
         dcm->dsetup();
   }
   return cm;
}

// transport "direct": peer memory only, no NCCL anywhere -- ranks are processes of one node (one per GPU, or several sharing
// a GPU in tests), the 128-byte job id names the /dev/shm rendezvous through which the IPC handles travel at start-up
ApxComm* apx_make_direct_comm(int rank, int world, const void* job_id)
{
   if (world > 16)
      APX_THROW("direct transport supports at most 16 ranks");
   DirectComm* cm = new DirectComm();
   cm->rank = rank;
   cm->world = world;
   cm->fused = 1;
   cm->rdv = new FileRendezvous();
   cm->rdv->rank = rank, cm->rdv->world = world;
   static const char hex[] = "0123456789abcdef";
   std::string key;
   const unsigned char* b = static_cast<const unsigned char*>(job_id);
   for (int k = 0; k < 16; ++k)
      key += hex[b[k] >> 4], key += hex[b[k] & 15];
   cm->rdv->prefix = "/dev/shm/apx_" + key;
   size_t mb = 16;
   if (const char* e = getenv("APX_DIST_WINDOW_MB"))
      mb = (size_t)std::max(1, atoi(e));
   cm->setup(mb << 20);
   cm->dsetup();
   return cm;
}

ApxComm* apx_make_local_comm(int rank, int world, void* hub_)
{
   LocalHub* hub = static_cast<LocalHub*>(hub_);
   if (!hub || hub->world != world)
      APX_THROW("local transport: hub was created for a different number of ranks");
   LocalComm* cm = new LocalComm();
   cm->hub = hub;
   cm->rank = rank;
   cm->world = world;
   CUDA_CHECK(cudaEventCreateWithFlags(&hub->ready[rank], cudaEventDisableTiming));
   CUDA_CHECK(cudaEventCreateWithFlags(&hub->done[rank], cudaEventDisableTiming));
   return cm;
}

extern "C" {
#pragma GCC visibility push(default)
int apx_nccl_unique_id(const char* lib, void* out128)
{
   try {
      NcclApi* api = nccl_api(lib);
      ncclUniqueId id;
      NCCL_CHECK(api, api->GetUniqueId(&id));
      static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
      memcpy(out128, &id, sizeof(id));
   } catch (const std::exception& e) {
      g_dist_err = e.what();
      return 1;
   }
   return 0;
}

// per-phase device time of the decomposed path: on = 1 starts collecting (and empties the record), the second call
// returns, in ms summed over everything since: [0] halo exchanges of per-atom vectors, [1] forward slab FFTs (plane
// reduction + 2-D FFT + transpose + 1-D FFT), [2] inverse slab FFTs, [3] scalar all-reduces, [4..7] their call counts.
// The caller must have synchronised the context.
int apx_dist_profile(apx_ctx* c, int on, double* out8)
{
   DistState& D = c->dist;
   if (out8) {
      for (int q = 0; q < 8; ++q)
         out8[q] = 0;
      for (int k = 0; k + 1 < D.prof_used; k += 2) {
         float ms = 0;
         if (cudaEventElapsedTime(&ms, D.prof_ev[k], D.prof_ev[k + 1]) == cudaSuccess) {
            out8[D.prof_kind[k]] += ms;
            out8[4 + D.prof_kind[k]] += 1;
         }
      }
   }
   D.prof_used = 0;
   D.prof_on = on ? 1 : 0;
   return 0;
}

void* apx_local_hub_create(int world) { return world >= 1 && world <= 16 ? new LocalHub(world) : nullptr; }

void apx_local_hub_destroy(void* h)
{
   LocalHub* hub = static_cast<LocalHub*>(h);
   if (!hub)
      return;
   for (void* p : hub->stage)
      if (p)
         cudaFree(p);
   for (auto e : hub->ready)
      if (e)
         cudaEventDestroy(e);
   for (auto e : hub->done)
      if (e)
         cudaEventDestroy(e);
   delete hub;
}

// host-only: the start-up rendezvous of transport "direct" (no GPU needed; tests/test_dist_plan.py runs it between two processes):
// `rounds` gathers of `bytes` bytes each; out receives the blobs of the last round, concatenated by rank
int apx_rendezvous_selftest(const void* job_id16, int rank, int world, const void* mine, int bytes, int rounds, void* out)
{
   try {
      FileRendezvous rdv;
      rdv.rank = rank, rdv.world = world;
      static const char hex[] = "0123456789abcdef";
      std::string key;
      const unsigned char* b = static_cast<const unsigned char*>(job_id16);
      for (int k = 0; k < 16; ++k)
         key += hex[b[k] >> 4], key += hex[b[k] & 15];
      rdv.prefix = "/dev/shm/apx_" + key;
      std::vector<char> blob((size_t)bytes);
      for (int r = 0; r < rounds; ++r) {
         for (int q = 0; q < bytes; ++q)
            blob[q] = (char)(static_cast<const char*>(mine)[q] + r);      // every round carries different bytes
         rdv.gather(blob.data(), (size_t)bytes, out);
      }
   } catch (const std::exception& e) {
      g_dist_err = e.what();
      return 1;
   }
   return 0;
}

// host-only: the halo plan of `rank` (no GPU needed; CPU tests run it under gloo with 2 ranks)
int apx_dist_plan(int n, const float* w3_sorted, const int* bounds, int world, int rank, double range_frac, int* send_idx,
   int* send_off, int* recv_idx, int* recv_off)
{
   std::vector<int> si, so, ri, ro;
   plan_halo(n, w3_sorted, bounds, world, rank, range_frac, si, so, ri, ro);
   std::copy(si.begin(), si.end(), send_idx);
   std::copy(ri.begin(), ri.end(), recv_idx);
   std::copy(so.begin(), so.end(), send_off);
   std::copy(ro.begin(), ro.end(), recv_off);
   return 0;
}
#pragma GCC visibility pop
}
