// dmd_cuda.cu -- libdmdb200.so: CUDA backend (sm_100a) of the C ABI in include/dmdb200.h.
//
// Kernels (one warp per replica; 28 replicas per CTA; the two 28x28 tables every prediction reads and the hot constants
// are staged in shared memory):
//   dmd_event_loop_kernel    the persistent event loop, main.F90:484-1258 (dmdb_run, engine 1: warp per replica); its
//                            first CTAs are the list-rebuild service (nbor.f + events.f for the other CTAs' replicas)
//   dmd_block_loop_kernel    the same loop with one CTA per replica, state in shared memory, batched
//                            conservative commit of independent events (dmdb_run, engine 2; dmd_block.h)
//   dmd_grid_loop_kernel     the same rounds spread over every SM for ONE large system (dmdb_run, engine 3; dmd_grid.h)
//   dmd_bulk_*_kernel        grid-parallel, one thread per bead, any system size: run start (init, fix-up of
//                            main.F90:249-321 through the cell grid), cell_add.f, nbor.f, events.f, group minima
//                            (dmdb_set_state*, dmdb_nbor, dmdb_predict_all)
//   dmd_sync_positions_kernel main.F90:1288-1295
//   dmd_energy_kernel        energy.f
//   dmd_retemp_kernel        replica-exchange temperature change on resident state (new functionality)
//   dmd_evcode_kernel        ev_code(i,j) read-back for the parity tests
// There is no CPU fallback: be::init fails when no CUDA device is present.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <chrono>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "dmd_types.h"

// The engine source is compiled once per lane count (dmd_warp.h), each time into its own namespace:
//   dmd::w32 (inline: plain dmd:: names)  32 lanes per replica -- CTA-per-replica and whole-GPU engines, bulk kernels
//   dmd::evl                              DMD_EVL_W lanes per replica -- the warp-per-replica event loop (engine 1)
#ifndef DMD_EVL_W
#define DMD_EVL_W 8  // four replicas per hardware warp (measured on B200, 48-peptide box, tools/prof_run.py: 32 lanes per
                     // replica 2.15e8, 16: 2.55e8, 8: 2.77e8 events/s)
#endif
#define DMD_W 32
#define DMD_VARIANT_BEGIN inline namespace w32 {
#define DMD_VARIANT_END }
#include "dmd_block.h"
#include "dmd_grid.h"
#undef DMD_W
#undef DMD_VARIANT_BEGIN
#undef DMD_VARIANT_END
#if DMD_EVL_W != 32
#define DMD_W DMD_EVL_W
#define DMD_VARIANT_BEGIN namespace evl {
#define DMD_VARIANT_END }
#if DMD_EVL_W == 8  // four replicas per warp: 112 cascade queues per CTA must fit the static shared memory
#undef DMD_CQ_CAP
#define DMD_CQ_CAP 56
#endif
#include "dmd_engine.h"
#include "dmd_lockstep.h"  // the hot path of 32 / DMD_EVL_W replicas per hardware warp, executed in lockstep
#undef DMD_W
#undef DMD_VARIANT_BEGIN
#undef DMD_VARIANT_END
#else
namespace dmd {
namespace evl = w32;
}
#endif
#define DMD_W 32

#define CUDA_OK(x)                                                                                   \
  do {                                                                                               \
    cudaError_t e_ = (x);                                                                            \
    if (e_ != cudaSuccess) throw std::runtime_error(std::string(#x) + ": " + cudaGetErrorString(e_)); \
  } while (0)

#include "dmd_exchange.h"

namespace dmd {

#ifndef DMD_WPC
#define DMD_WPC 28  // one 28-warp CTA per SM (72 registers per thread) shares ONE shared-memory copy of the tables.
                    // Measured on B200 (48-peptide box), final round-1 build with the service split: 24 warps
                    // 1.67e8, 28: 1.93e8, 32 (64 registers): 1.90e8 events/s
#endif
constexpr int WARPS_PER_CTA = DMD_WPC;
constexpr int EVL_RPW = 32 / DMD_EVL_W;            // replicas per hardware warp of the event loop
constexpr int EVL_RPC = WARPS_PER_CTA * EVL_RPW;   // ... and per CTA
constexpr int BULK_THREADS = 256;  // CTA size of the thread-per-bead bulk kernels
#ifndef DMD_MIN_CTAS
#define DMD_MIN_CTAS 1
#endif

// read-only constants of the hot loop, copied into shared memory once per CTA
struct SmemConsts {
  HotTables tab;
  HotConst hot;
  double bl[6 * HOT_MAX_RES];
};
template <class ST>
__device__ __forceinline__ ST stage_consts_t(const DevArrays& d, SmemConsts* smem) {
  const double* src = reinterpret_cast<const double*>(d.tables);
  double* dst = reinterpret_cast<double*>(&smem->tab);
  for (int k = threadIdx.x; k < (int)(sizeof(HotTables) / 8); k += blockDim.x) dst[k] = src[k];  // the prefix of PairTables
  src = reinterpret_cast<const double*>(d.hot);
  dst = reinterpret_cast<double*>(&smem->hot);
  for (int k = threadIdx.x; k < (int)(sizeof(HotConst) / 8); k += blockDim.x) dst[k] = src[k];
  const bool bl_fits = d.nres <= HOT_MAX_RES;
  if (bl_fits)
    for (int k = threadIdx.x; k < 6 * d.nres; k += blockDim.x) smem->bl[k] = d.bl[k];
  __syncthreads();
  ST st;
  st.tab = &smem->tab;
  st.hot = &smem->hot;
  st.bl = bl_fits ? smem->bl : d.bl;
  return st;
}
__device__ __forceinline__ Staged stage_consts(const DevArrays& d, SmemConsts* smem) { return stage_consts_t<Staged>(d, smem); }

__device__ __forceinline__ int32_t* warp_queue() {
  __shared__ int32_t s_cq[WARPS_PER_CTA][CQ_CAP];
  return s_cq[threadIdx.x >> 5];
}

__device__ __forceinline__ int replica_of_warp(int r0, int nrep) {
  int w = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
  return w < nrep ? r0 + w : -1;
}

// ---- list-rebuild service (svc_request in dmd_engine.h is the requesting side).
// A ROLE of the event-loop kernel (its first n_srv CTAs): one kernel, so nothing depends on two kernels running side
// by side (a second-stream service kernel with its own shared memory measured the same and was dropped).  The cell grid of the rebuild aliases the CTA's static shared memory (tables +
// cascade queues, which a service CTA does not use); bead records and tables are read through L1/L2.
struct alignas(16) EvlSmem {
  SmemConsts consts;
  int32_t cq[EVL_RPC][evl::CQ_CAP];
};
#ifndef DMD_SVC_GROUPS
#define DMD_SVC_GROUPS 2  // a service CTA works as this many independent groups of warps, one replica each: with one
                          // bead per thread a 1344-bead replica leaves a third of 896 threads idle in its second round
#endif
#ifndef DMD_CHAINWISE_MAX_NEAR
#define DMD_CHAINWISE_MAX_NEAR 6  // chain-wise list rebuild while a chain has at most this many chains within reach (mean)
#endif
__device__ __forceinline__ void svc_group_sync(int grp, int gsz) { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(gsz) : "memory"); }

__device__ __noinline__ void svc_serve_in_kernel(DevArrays d, int r0, int nrep, int n_srv, unsigned char* scratch, size_t scratch_bytes) {
  __shared__ int s_pick[DMD_SVC_GROUPS], s_state[DMD_SVC_GROUPS];
  const evl::Staged tab = evl::staged_global(d);
  const int gsz = ((int)blockDim.x / 32 / DMD_SVC_GROUPS) * 32;  // threads per group (whole warps)
  const int grp = (int)threadIdx.x / gsz;
  if (grp >= DMD_SVC_GROUPS) return;  // left-over warps
  const int tid = (int)threadIdx.x - grp * gsz, nt = gsz;
  scratch_bytes /= DMD_SVC_GROUPS;
  scratch += (size_t)grp * scratch_bytes;
  const bool grid_fits = ((size_t)d.ncc3 + 2 * (size_t)d.n_beads) * 4 <= scratch_bytes;
  int32_t* const s_heads = reinterpret_cast<int32_t*>(scratch);
  int32_t* const s_cnext = s_heads + d.ncc3;
  uint32_t* const s_cpk = reinterpret_cast<uint32_t*>(s_cnext + d.n_beads);
  // Which request a free group takes: the OLDEST one.  Requesters queue their replica index in a ticket ring
  // (svc_request, dmd_types.h: SVC_Q_*); thread 0 of a free group claims the next ticket and reads its slot.  One pool,
  // first come first served: no scan of the request words, no collisions, nobody is passed over.  (Measured before:
  // every group scanning for the lowest pending index -- collisions, mean wait 560 us; scanning from a fixed point per
  // group -- the far end of the array starved, 0.15 % of the requests waited more than the 40 ms after which the warp
  // rebuilds in place; a stretch per group -- no pooling, mean wait 1100 us.)
  unsigned long long* const rq = d.svc_ctl;
  while (true) {
    if (tid == 0) {
      int state = 0, pick = 0;
      const unsigned long long hd = evl::svc_ld_relaxed64(&rq[SVC_Q_HEAD]), tl = evl::svc_ld_relaxed64(&rq[SVC_Q_TAIL]);
      if (hd < tl) {
        if (atomicCAS(&rq[SVC_Q_HEAD], hd, hd + 1ull) == hd) {  // ticket hd is this group's
          const unsigned long long* const slot = &rq[SVC_Q_RING + hd % rq[SVC_Q_CAP]];
          unsigned long long e = evl::svc_ld_acquire64(slot);
          const long long t0 = clock64();  // (the slot is written right after the ticket was taken: a short wait)
          while ((e >> 24) < hd + 1ull && clock64() - t0 < evl::SVC_TIMEOUT) e = evl::svc_ld_acquire64(slot);
          if ((e >> 24) == hd + 1ull) {
            pick = (int)(e & 0xffffffull);
            state = evl::svc_cas_acq_rel(d.svc_flag + pick, 1, 2) == 1 ? 1 : 0;  // 0: taken back by its warp meanwhile
          }
          // else: a later lap of the ring has overwritten the slot -- more than SVC_Q_CAP tickets were outstanding,
          // which takes a service so far behind that the tickets of requests taken back pile up -- or it was never
          // written.  The ticket is skipped; if its request is still pending, its warp takes it back after SVC_PATIENCE
        }
      } else {
        const unsigned long long done = *(volatile unsigned long long*)&d.svc_ctl[0];
        state = done >= (unsigned long long)nrep ? 2 : 0;
        if (state == 0) __nanosleep(500);
      }
      s_pick[grp] = pick;
      s_state[grp] = state;
    }
    svc_group_sync(grp, gsz);
    const int pick = s_pick[grp], state = s_state[grp];
    svc_group_sync(grp, gsz);  // (thread 0 writes both again at the top of the loop)
    if (state == 2) return;
    if (state == 0) continue;
    const long long t0 = clock64();
    const int rid = pick;  // (the queue carries indices into the request words of the whole replica set)
    evl::Rep q;
    evl::rep_bind(q, d, tab, nullptr, rid);  // scalars as saved by the requesting warp (tfalse = 0, new interval_max)
    q.error = 0;
    // small systems: chain-wise candidate search (dmd_engine.h); its scratch -- packed cell coordinates, bounding
    // spheres, near-chain masks -- lives in the group's share of the CTA's static shared memory
    const int nch = d.n_chains;
    const size_t cw_bytes = (size_t)d.n_beads * 4 + (size_t)nch * (sizeof(evl::ChainBound) + 8) + 32;
    bool cell_walk = true;
    // Small systems (SysConst.chainwise): two candidate searches that avoid the linked-list walk, chosen per rebuild by
    // how many chains come within reach of one another (measured on B200, 48-peptide boxes, per rebuild and group of 14
    // warps: dilute box  chain-wise 112 us, sorted grid 257 us, cell walk ~190 us; aggregated box  sorted grid 842 us,
    // chain-wise ~1700 us, cell walk ~1700 us).  Scratch = the group's share of the CTA's static shared memory.
    const size_t sg_end_words = ((size_t)d.ncc3 + 2 + 1) & ~(size_t)1, sg_sorted_words = ((size_t)d.n_beads + 1) & ~(size_t)1;
    const size_t sg_bytes = (sg_end_words + sg_sorted_words) * 2 + (size_t)nt * 4;
    if (d.chainwise && cw_bytes <= scratch_bytes) {
      evl::ChainBound* const cbound = reinterpret_cast<evl::ChainBound*>(scratch);
      unsigned* const near = reinterpret_cast<unsigned*>(cbound + nch);
      uint32_t* const cpk = near + 2 * nch;
      for (int k = tid; k < 2 * nch; k += nt) near[k] = 0u;
      if (tid == 0) s_pick[grp] = 0;  // (free between two picks) number of near chain pairs
      evl::chainwise_cells(q, cpk, tid, nt);
      evl::chainwise_bounds(q, cbound, nch, tid >> 5, nt >> 5);
      svc_group_sync(grp, gsz);
      evl::chainwise_near(q, cbound, near, nch, tid, nt);
      svc_group_sync(grp, gsz);
      for (int k = tid; k < 2 * nch; k += nt) atomicAdd(&s_pick[grp], __popc(near[k]));
      svc_group_sync(grp, gsz);
      const bool sparse = s_pick[grp] <= DMD_CHAINWISE_MAX_NEAR * nch;
      const long long t1 = clock64();
      if (sparse) {  // dilute: a chain has a few chains within reach
        evl::chainwise_lists(q, cpk, near, tid, nt);
        cell_walk = false;
      } else if (d.n_beads < 65536 && sg_bytes <= scratch_bytes) {  // aggregated: every chain is near most others
        svc_group_sync(grp, gsz);  // (the scratch changes hands)
        evl::SortedGrid sg;
        sg.end = reinterpret_cast<uint16_t*>(scratch);
        sg.sorted = sg.end + sg_end_words;
        sg.tot = reinterpret_cast<unsigned*>(sg.sorted + sg_sorted_words);
        evl::sorted_grid_build(q, sg, tid, nt, [&]() { svc_group_sync(grp, gsz); });
        evl::sorted_grid_lists(q, sg, tid >> 5, nt >> 5);
        cell_walk = false;
      }
      svc_group_sync(grp, gsz);
      if (tid == 0 && !cell_walk) {
        atomicAdd(&d.svc_ctl[5], (unsigned long long)(t1 - t0));
        atomicAdd(&d.svc_ctl[6], (unsigned long long)(clock64() - t1));
      }
    }
    if (cell_walk) {
      if (grid_fits) {
        for (int k = tid; k < d.ncc3; k += nt) s_heads[k] = -1;
        q.cellhead = s_heads;
        q.cnext = s_cnext;
        q.cpk = s_cpk;
        svc_group_sync(grp, gsz);
      }
      evl::cell_build(q, tid, nt);
      svc_group_sync(grp, gsz);
      evl::nbor_build(q, tid, nt);
      if (!grid_fits) {
        svc_group_sync(grp, gsz);
        evl::cell_clear(q, tid, nt);
      }
      svc_group_sync(grp, gsz);  // predict_all_flat reads the rows other threads wrote
    }
    evl::predict_all_flat(q, tid >> 5, nt >> 5);  // events.f:23-107; the requester refreshes the group minima
    if (q.error && evl::Warp::lane() == 0 && atomicCAS(&q.sc->error, 0, q.error) == 0) q.sc->error_info = q.error_info;
    __threadfence();
    svc_group_sync(grp, gsz);
    if (tid == 0) {
      evl::svc_st_release(d.svc_flag + rid, 0);
      atomicAdd(&d.svc_ctl[4], (unsigned long long)(clock64() - t0));
      atomicAdd(&d.svc_ctl[1], 1ull);
    }
  }
}

// the first n_srv CTAs are the list-rebuild service, the others run the event loop
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, DMD_MIN_CTAS) dmd_event_loop_kernel(DevArrays d, int r0, int nrep, long long n_events, int flags, int n_srv) {
  __shared__ EvlSmem sm;
  if ((int)blockIdx.x < n_srv) {
    svc_serve_in_kernel(d, r0, nrep, n_srv, reinterpret_cast<unsigned char*>(&sm), sizeof(EvlSmem));
    return;
  }
  const evl::Staged tab = stage_consts_t<evl::Staged>(d, &sm.consts);
  const int wl = (int)threadIdx.x / DMD_EVL_W;  // this replica's slot in the CTA (DMD_EVL_W lanes each)
  const int w = ((int)blockIdx.x - n_srv) * EVL_RPC + wl;
#if DMD_EVL_W < 32
  const bool pair = __all_sync(0xffffffffu, w < nrep);  // every lane group of this hardware warp holds a replica
#endif
  if (w >= nrep) return;
  const int rid = r0 + w;
  evl::Rep r;
  evl::rep_bind(r, d, tab, sm.cq[wl], rid);
  if (n_srv > 0) {
    r.svc = d.svc_flag + rid;
    r.svc_ctl = d.svc_ctl;
    r.svc_id = rid;
  }
#if defined(DMD_PHASE_PROF)
#if DMD_EVL_W >= 16
  __shared__ unsigned long long s_prof[EVL_RPC][16];
  r.prof = s_prof[wl];
#else  // no room beside the cascade queues: the accumulators sit in global memory (they stay in L1 / L2)
  __shared__ unsigned long long* s_prof_base;
  if (threadIdx.x == 0) s_prof_base = evl::g_phase_acc + (size_t)blockIdx.x * EVL_RPC * 16;
  __syncthreads();
  r.prof = s_prof_base + wl * 16;
#endif
  if (evl::Warp::lane() == 0) {
    for (int k = 0; k < 15; k++) r.prof[k] = 0;
    r.prof[15] = (unsigned long long)clock64();
  }
  evl::Warp::sync();
#endif
#if DMD_EVL_W < 32
  if (pair) evl::lk_run_events(r, n_events, (flags & 1) != 0);  // the replicas of the warp in lockstep
  else if (r.error == 0) evl::run_events(r, n_events, (flags & 1) != 0);
#else
  if (r.error == 0) evl::run_events(r, n_events, (flags & 1) != 0);
#endif
#if defined(DMD_PHASE_PROF)
  evl::Warp::sync();
  if (evl::Warp::lane() == 0)
    for (int k = 0; k < 15; k++) atomicAdd(&evl::g_phase_cyc[k], r.prof[k]);
  if ((threadIdx.x & 31) == 0) {  // when the CTA's last warp finished, and on which SM it ran
    unsigned long long gt;
    unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    atomicMax(&evl::g_cta_end[blockIdx.x], gt);
    evl::g_cta_sm[blockIdx.x] = smid;
  }
#endif
  evl::rep_save(r);
  if (n_srv > 0 && evl::Warp::lane() == 0) atomicAdd(&d.svc_ctl[0], 1ull);  // the service CTAs leave when all warps are done
}

// ---- CTA-per-replica engine -----------------------------------------------------------------------------------
// dynamic shared memory: BlkShared | PairTables | cascade queues | claim[N+3] | (rec[N] | cal[stride] | er34[2N])
struct BlkLayout {
  size_t tab, cq, claim, meta, nup, ndn, rec, cal, er34, total;
  bool topo;   // topology words and list lengths resident in shared memory
  bool state;  // ... and bead records / calendar / er34 too
};
__host__ __device__ inline size_t blk_align(size_t x) { return (x + 127) & ~(size_t)127; }
__host__ __device__ inline BlkLayout blk_layout(int N, int cal_stride, size_t limit) {
  BlkLayout L;
  size_t off = blk_align(sizeof(BlkShared));
  L.tab = off; off += blk_align(sizeof(SmemConsts));
  L.cq = off; off += blk_align((size_t)BK_MAXW * CQ_CAP * 4);
  L.claim = off; off += blk_align(((size_t)N + 3) * 4);
  L.total = off;  // the part that must fit
  L.meta = off; off += blk_align((size_t)N * 4);
  L.nup = off; off += blk_align((size_t)N * 2);
  L.ndn = off; off += blk_align((size_t)N * 2);
  L.topo = off <= limit;
  if (L.topo) L.total = off;
  L.rec = off; off += blk_align((size_t)N * sizeof(BeadRec));
  L.cal = off; off += blk_align((size_t)cal_stride * sizeof(CalEnt));
  L.er34 = off; off += blk_align((size_t)N * 8);
  L.state = L.topo && off <= limit;
  if (L.state) L.total = off;
  return L;
}

__global__ void __launch_bounds__(BK_MAXW * 32, 1) dmd_block_loop_kernel(DevArrays d, int r0, int nrep, long long n_events,
                                                                         unsigned smem_limit, int flags) {
  extern __shared__ __align__(128) unsigned char blk_smem[];
  const int N = d.sys->N;
  const BlkLayout L = blk_layout(N, d.cal_stride, smem_limit);
  BlkShared& S = *reinterpret_cast<BlkShared*>(blk_smem);
  SmemConsts* sconst = reinterpret_cast<SmemConsts*>(blk_smem + L.tab);
  int32_t* cq = reinterpret_cast<int32_t*>(blk_smem + L.cq);
  uint32_t* claim = reinterpret_cast<uint32_t*>(blk_smem + L.claim);
  const Staged tab = stage_consts(d, sconst);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5, tid = threadIdx.x, nt = blockDim.x;
  for (int rr = blockIdx.x; rr < nrep; rr += gridDim.x) {
    const int rid = r0 + rr;
    Rep r;
    rep_bind(r, d, tab, cq + w * CQ_CAP, rid);
    BeadRec* const grec = r.rec;
    CalEnt* const gcal = r.cal;
    int32_t* const ger34 = r.er34;
    uint16_t* const gnup = r.nup;
    uint16_t* const gndn = r.ndn;
    if (L.topo) {  // topology words and list lengths
      uint32_t* smeta = reinterpret_cast<uint32_t*>(blk_smem + L.meta);
      uint16_t* snup = reinterpret_cast<uint16_t*>(blk_smem + L.nup);
      uint16_t* sndn = reinterpret_cast<uint16_t*>(blk_smem + L.ndn);
      for (int k = tid; k < N; k += nt) {
        smeta[k] = d.meta[k];
        snup[k] = gnup[k];
        sndn[k] = gndn[k];
      }
      r.c.meta = smeta;
      r.nup = snup;
      r.ndn = sndn;
    }
    if (L.state) {  // stage the replica into shared memory (16-byte vector copies)
      uint4* srec = reinterpret_cast<uint4*>(blk_smem + L.rec);
      uint4* scal = reinterpret_cast<uint4*>(blk_smem + L.cal);
      int2* ser = reinterpret_cast<int2*>(blk_smem + L.er34);
      const uint4* g4 = reinterpret_cast<const uint4*>(grec);
      for (int k = tid; k < N * 4; k += nt) srec[k] = g4[k];
      const uint4* c4 = reinterpret_cast<const uint4*>(gcal);
      for (int k = tid; k < d.cal_stride; k += nt) scal[k] = c4[k];
      const int2* e2 = reinterpret_cast<const int2*>(ger34);
      for (int k = tid; k < N; k += nt) ser[k] = e2[k];
      r.rec = reinterpret_cast<BeadRec*>(srec);
      r.cal = reinterpret_cast<CalEnt*>(scal);
      r.er34 = reinterpret_cast<int32_t*>(ser);
    }
    for (int k = tid; k < N + 3; k += nt) claim[k] = CLAIM_FREE;
    if (w != 0) r.n_pair_pred = r.n_nbr_visits = 0;
    if (tid == 0) {
      S.coll = r.coll;
      S.target = r.coll + n_events;
      S.stop_at_output = flags & 1;
      S.window = r.interval * 0.02;
      S.error = r.error;
      S.error_info = r.error_info;
      S.st_rounds = S.st_exec = S.st_rollback = S.st_conflict = S.st_cold = 0;
      S.n_pair_pred = S.n_nbr_visits = 0;
      S.n_cand = 0;
      S.tlast = -1.0;
      for (int q = 0; q < 8; q++) S.cyc[q] = 0;
      for (int q = 0; q < 32; q++) S.nevents[q] = 0;
    }
    __syncthreads();
    blk_run(S, r, claim, w, nw);
    harvest_counters(r);
    if (w != 0 && Warp::lane() == 0) {
      blk_atomic_add64(&S.n_pair_pred, r.n_pair_pred);
      blk_atomic_add64(&S.n_nbr_visits, r.n_nbr_visits);
    }
    __syncthreads();
    if (L.topo) {
      for (int k = tid; k < N; k += nt) {  // list lengths may have changed (rebuilds)
        gnup[k] = r.nup[k];
        gndn[k] = r.ndn[k];
      }
      r.nup = gnup;
      r.ndn = gndn;
      r.c.meta = d.meta;
    }
    if (L.state) {
      uint4* g4 = reinterpret_cast<uint4*>(grec);
      const uint4* srec = reinterpret_cast<const uint4*>(blk_smem + L.rec);
      for (int k = tid; k < N * 4; k += nt) g4[k] = srec[k];
      uint4* c4 = reinterpret_cast<uint4*>(gcal);
      const uint4* scal = reinterpret_cast<const uint4*>(blk_smem + L.cal);
      for (int k = tid; k < d.cal_stride; k += nt) c4[k] = scal[k];
      int2* e2 = reinterpret_cast<int2*>(ger34);
      const int2* ser = reinterpret_cast<const int2*>(blk_smem + L.er34);
      for (int k = tid; k < N; k += nt) e2[k] = ser[k];
      r.rec = grec;
      r.cal = gcal;
      r.er34 = ger34;
    }
    __syncthreads();
    if (w == 0) {
      r.n_pair_pred += S.n_pair_pred;
      r.n_nbr_visits += S.n_nbr_visits;
      if (S.error && !r.error) {
        r.error = S.error;
        r.error_info = S.error_info;
      }
      rep_save(r);
      rebuild_all_groups(r);  // the warp engine's group minima, so that both engines can alternate on one handle
      if (Warp::lane() == 0) {
        for (int q = 0; q < 32; q++) r.sc->nevents[q] += S.nevents[q];
        long long* st = d.blkstat + (size_t)rid * 16;
        st[0] += S.st_rounds; st[1] += S.st_exec; st[2] += S.st_rollback; st[3] += S.st_conflict; st[4] += S.st_cold;
        for (int q = 0; q < 8; q++) st[8 + q] += S.cyc[q];
      }
    }
    __syncthreads();
  }
}

// ---- the interval pseudo-event (main.F90:1126-1187, interval_event_cold in dmd_engine.h) for a large system,
// spread over the grid: begin (one thread: event time -> scalars), advance (thread per bead: shift the calendar,
// move the beads, largest displacement), decide (one thread: displ.f:37-46, main.F90:1150-1157), and -- when the
// lists have to be rebuilt -- wrap + the bulk cell / nbor / predict kernels.  Same arithmetic as the serial code.
__global__ void dmd_interval_begin_kernel(DevArrays d, int rid) {
  Rep r;
  rep_bind(r, d, staged_global(d), nullptr, rid);
  RepScalars& q = *r.sc;
  const double tf = r.cal[r.N + 1].t;
  q.tfalse = tf;            // main.F90:639-643
  q.coll = q.coll + 1;
  q.t = q.t + tf;           // :1129
  q.interval_max = q.interval_max - tf;
  *reinterpret_cast<unsigned long long*>(&q.pad[2]) = 0ull;  // largest displacement (bits of a non-negative double)
}
__global__ void __launch_bounds__(BULK_THREADS) dmd_interval_advance_kernel(DevArrays d, int rid) {
  Rep r;
  rep_bind(r, d, staged_global(d), nullptr, rid);
  const int k = blockIdx.x * BULK_THREADS + threadIdx.x;
  const double tf = r.tfalse;
  if (k < r.N + 3) r.cal[k].t = r.cal[k].t - tf;  // :1133-1135
  double moved = 0.0;
  if (k < r.N) {  // :1140-1144 + displ.f:20-33
    BeadRec* p = &r.rec[k];
    double x = p->x + p->vx * tf, y = p->y + p->vy * tf, z = p->z + p->vz * tf;
    p->x = x; p->y = y; p->z = z;
    double a = r.oldr[3 * k] - x, b = r.oldr[3 * k + 1] - y, cc = r.oldr[3 * k + 2] - z;
    double dis = a * a + b * b + cc * cc;
    moved = dis / r.c.sys->hdelr;
  }
  moved = warp_max(moved);
  if (Warp::lane() == 0 && moved > 0.0)
    atomicMax(reinterpret_cast<unsigned long long*>(&r.sc->pad[2]), (unsigned long long)__double_as_longlong(moved));
}
__global__ void dmd_interval_decide_kernel(DevArrays d, int rid) {
  Rep r;
  rep_bind(r, d, staged_global(d), nullptr, rid);
  const double moved_far = __longlong_as_double((long long)*reinterpret_cast<unsigned long long*>(&r.sc->pad[2]));
  r.tfalse = 0.0;
  bool update = false;
  if (moved_far >= 0.1) {  // displ.f:37-46
    update = true;
    if (moved_far >= 1.25 * 1.25) {
      r.t_fact = r.t_fact / 1.01;
      r.interval = r.t_fact / dmd_sqrt(r.setemp);
    }
  }
  const bool rebuild = update || r.interval > r.interval_max;  // :1150-1179
  if (rebuild) {
    if (!update) {
      r.sc->nforcedupdate += 1;
      r.n_forced = r.n_forced * 1.01;
    }
    r.interval_max = r.interval * r.n_forced;
    r.sc->nupdates += 1;
  }
  r.cal[r.N + 1].t = r.interval * 0.999;  // :1181 (the bulk prediction only writes bead entries)
  r.old_tfalse = 0.0;
  if (r.n_log < r.c.sys->log_cap) {
    EventLogRec e;
    e.t = r.t;
    e.i = r.N + 2;
    e.j = 0;
    e.type = -2;
    e.evcode = 0;
    r.log[r.n_log] = e;
    r.n_log++;
  }
  RepScalars& q = *r.sc;  // (one thread: store the scalars directly)
  q.tfalse = r.tfalse; q.old_tfalse = r.old_tfalse; q.interval = r.interval; q.t_fact = r.t_fact;
  q.interval_max = r.interval_max; q.n_forced = r.n_forced; q.n_log = r.n_log;
  q.pad[1] = rebuild ? 1 : 0;
}
__global__ void __launch_bounds__(BULK_THREADS) dmd_interval_wrap_kernel(DevArrays d, int rid) {
  Rep r;
  rep_bind(r, d, staged_global(d), nullptr, rid);
  const int k = blockIdx.x * BULK_THREADS + threadIdx.x;
  if (k >= r.N) return;
  BeadRec* p = &r.rec[k];
  double x = p->x - dmd_round(p->x), y = p->y - dmd_round(p->y), z = p->z - dmd_round(p->z);
  p->x = x; p->y = y; p->z = z;
  r.oldr[3 * k] = x; r.oldr[3 * k + 1] = y; r.oldr[3 * k + 2] = z;
}

// ---- whole-GPU engine for one large system (dmd_grid.h): cooperative launch, one CTA per SM
__global__ void __launch_bounds__(BK_MAXW * 32, 1) dmd_grid_loop_kernel(DevArrays d, int rid, GridShared* S, uint32_t* claim) {
  __shared__ SmemConsts sconst;
  __shared__ int32_t s_cq[BK_MAXW][CQ_CAP];
  const Staged tab = stage_consts(d, &sconst);
  const int w = threadIdx.x >> 5;
  const int gw = blockIdx.x * BK_MAXW + w, ngw = gridDim.x * BK_MAXW;
  Rep r;
  rep_bind(r, d, tab, s_cq[w], rid);
  const int64_t pp0 = r.n_pair_pred, nv0 = r.n_nbr_visits;
  if (gw == 0) r.n_log = S->n_log;
  grid_run(*S, r, claim, gw, ngw);
  harvest_counters(r);
  if (Warp::lane() == 0 && gw != 0) {  // work counters of the other warps
    atomicAdd((unsigned long long*)&S->nevents[30], (unsigned long long)(r.n_pair_pred - pp0));
    atomicAdd((unsigned long long*)&S->nevents[31], (unsigned long long)(r.n_nbr_visits - nv0));
  }
  if (gw == 0) {
    if (S->error && !r.error) {
      r.error = S->error;
      r.error_info = S->error_info;
    }
    rep_save(r);
    if (Warp::lane() == 0) {
      for (int q = 0; q < 30; q++) r.sc->nevents[q] += S->nevents[q];
      long long* st = d.blkstat + (size_t)rid * 16;
      st[0] += S->st_rounds; st[1] += S->st_exec; st[2] += S->st_rollback; st[3] += S->st_conflict;
      st[8] += S->cyc[0]; st[9] += S->cyc[1]; st[10] += S->cyc[2]; st[11] += S->cyc[3]; st[12] += S->cyc[4]; st[13] += S->cyc[5];
    }
  }
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) dmd_sync_positions_kernel(DevArrays d, int r0, int nrep) {
  int rid = replica_of_warp(r0, nrep);
  if (rid < 0) return;
  Rep r;
  rep_bind(r, d, staged_global(d), warp_queue(), rid);
  sync_positions(r);
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) dmd_energy_kernel(DevArrays d, int r0, int nrep, OutRec* eout) {
  __shared__ SmemConsts sconst;
  const Staged tab = stage_consts(d, &sconst);
  int rid = replica_of_warp(r0, nrep);
  if (rid < 0) return;
  Rep r;
  rep_bind(r, d, tab, warp_queue(), rid);
  OutRec o;
  energy_of(r, o);
  if (Warp::lane() == 0) eout[rid] = o;
}

// beta-sheet observables (fibril_list_assign.f definitions) of every replica, one warp each, from resident state
constexpr int SHEET_WARPS = 4;
constexpr int SHEET_MAX_CHAINS = 96;
__global__ void __launch_bounds__(SHEET_WARPS * 32) dmd_sheet_kernel(DevArrays d, int nrep, int32_t* out) {
  __shared__ uint8_t s_hb[SHEET_WARPS][SHEET_MAX_CHAINS * SHEET_MAX_CHAINS];
  __shared__ int32_t s_lab[SHEET_WARPS][SHEET_MAX_CHAINS];
  const int w = threadIdx.x >> 5;
  const int rid = blockIdx.x * SHEET_WARPS + w;
  if (rid >= nrep) return;
  Rep r;
  rep_bind(r, d, staged_global(d), nullptr, rid);
  sheet_observables(r, out + 8 * (size_t)rid, s_hb[w], s_lab[w]);
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) dmd_retemp_kernel(DevArrays d, int r0, int nrep, const double* tstar_new) {
  __shared__ SmemConsts sconst;
  const Staged tab = stage_consts(d, &sconst);
  int rid = replica_of_warp(r0, nrep);
  if (rid < 0) return;
  const double tn = tstar_new[rid];
  if (!(tn > 0.0)) return;  // <= 0: leave this replica untouched
  Rep r;
  rep_bind(r, d, tab, warp_queue(), rid);
  retemp(r, tn);
  rep_save(r);
}

// ---- grid-parallel bulk kernels: ONE THREAD PER BEAD, blockIdx.y = replica.  Any system size and replica count
// (a 10^6-bead box is one replica spread over the whole GPU; 2368 small replicas are 2368 rows of CTAs).
// They read the tables through global memory (L1/L2 resident) instead of staging 31 KB per small CTA.
__device__ __forceinline__ bool bulk_bind(Rep& r, const DevArrays& d, int r0, int& k) {
  rep_bind(r, d, staged_global(d), nullptr, r0 + blockIdx.y);
  k = blockIdx.x * BULK_THREADS + threadIdx.x;
  return k < r.N;
}
__device__ __forceinline__ void bulk_error(Rep& r, int code, int info) {  // first error of the replica wins
  if (atomicCAS(&r.sc->error, 0, code) == 0) r.sc->error_info = info;
}

// clears the error word and the fix-up hit counter of each replica before a run start
__global__ void dmd_bulk_reset_kernel(DevArrays d, int r0, int nrep) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nrep) return;
  RepScalars& q = d.scal[r0 + k];
  q.error = 0;
  q.error_info = 0;
  q.pad[0] = 0;
}

// run start, step 1 (inputinfo.f:89-91, main.F90:143-156, 205-234, 408-423): records, scalars, calendar
__global__ void __launch_bounds__(BULK_THREADS) dmd_bulk_init_kernel(DevArrays d, int r0, const double* sv, size_t sv_stride,
                                                                      const int32_t* bp, size_t bp_stride,
                                                                      const double* tstar, unsigned long long seed0) {
  Rep r;
  int k;
  bulk_bind(r, d, r0, k);
  const int rid = r0 + blockIdx.y;
  const size_t c = blockIdx.y;
  init_scalars(r, tstar[rid], seed0 + (unsigned long long)rid);  // identical in every thread; lane 0 of each warp
                                                                 // stores the same tallies
  const double tghost = first_ghost_time(r);
  if (k < r.N && !init_bead(r, k, sv + c * sv_stride, bp ? bp + c * bp_stride : nullptr)) bulk_error(r, DMD_E_BAD_INPUT, k);
  if (k < d.cal_stride) r.cal[k] = init_cal_entry(r, k, tghost);
  if (k == 0) {  // the one copy of the scalars that is stored
    RepScalars& q = *r.sc;
    q.t = r.t; q.tfalse = r.tfalse; q.old_tfalse = r.old_tfalse; q.setemp = r.setemp; q.interval = r.interval;
    q.t_fact = r.t_fact; q.interval_max = r.interval_max; q.n_forced = r.n_forced; q.avegtime = r.avegtime;
    q.coll = r.coll; q.rng_ctr = r.ctr; q.n_pair_pred = 0; q.n_nbr_visits = 0;
    q.n_log = 0; q.n_out = 0;
  }
}

// cell_add.f:12-28
__global__ void __launch_bounds__(BULK_THREADS) dmd_bulk_cell_kernel(DevArrays d, int r0) {
  Rep r;
  int k;
  if (!bulk_bind(r, d, r0, k)) return;
  cell_build(r, k, 1 << 30);
}
__global__ void __launch_bounds__(BULK_THREADS) dmd_bulk_clear_kernel(DevArrays d, int r0) {
  Rep r;
  int k;
  if (!bulk_bind(r, d, r0, k)) return;
  cell_clear(r, k, 1 << 30);
}

// run start, step 2 (main.F90:249-321): every N / C bead looks through its cell stencil for the pairs the literal
// double loop acts on (the N-C well, 4.5 A, is shorter than the stencil reach) and appends them to scratch
__global__ void __launch_bounds__(BULK_THREADS) dmd_bulk_hits_kernel(DevArrays d, int r0) {
  Rep r;
  int k;
  if (!bulk_bind(r, d, r0, k)) return;
  const uint32_t mk = r.c.meta[k];
  const int cls = meta_cls(mk);
  if ((cls != 1 && cls != 2) || r.cnext[k] == -2) return;
  const BeadRec a = r.rec[k];
  const int ck = r.c.chain[k];
  int32_t* hits = reinterpret_cast<int32_t*>(r.up);
  const int hit_cap = (int)(((size_t)r.N * r.cap) / 4);
  int32_t* counter = &r.sc->pad[0];
  auto emit = [&](int kj) {
    const int pos = atomicAdd(counter, 1);
    if (pos < hit_cap) {
      hits[2 * pos] = k;
      hits[2 * pos + 1] = kj;
    }
  };
  if (a.bptnr > k) emit(a.bptnr);  // bonded pairs act whatever their distance
  stencil_visit(r, k, [&](int kj) {
    if (kj <= k || kj == a.bptnr) return;
    const int cj = meta_cls(r.c.meta[kj]);
    if (cj != 1 && cj != 2) return;
    if (fixup_pair_hit(r, k, a, mk, ck, kj)) emit(kj);
  });
}
// ... and one warp per replica replays them in the loop's order
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) dmd_fixup_kernel(DevArrays d, int r0, int nrep) {
  int rid = replica_of_warp(r0, nrep);
  if (rid < 0) return;
  Rep r;
  rep_bind(r, d, staged_global(d), nullptr, rid);
  int32_t* hits = reinterpret_cast<int32_t*>(r.up);
  const int hit_cap = (int)(((size_t)r.N * r.cap) / 4);
  int nh = r.sc->pad[0];
  if (nh > hit_cap) {
    if (Warp::lane() == 0) bulk_error(r, DMD_E_NBR_CAP, nh);
    nh = 0;
  }
  fixup_replay(r, hits, nh, hit_cap);
}

// nbor.f:33-137
__global__ void __launch_bounds__(BULK_THREADS) dmd_bulk_nbor_kernel(DevArrays d, int r0) {
  Rep r;
  int k;
  const bool in = bulk_bind(r, d, r0, k);
  r.error = 0;
  nbor_build(r, in ? k : r.N, 1 << 30);  // every thread takes part in the warp votes inside
  if (r.error && Warp::lane() == 0) bulk_error(r, r.error, r.error_info);
}

// events.f:23-107
__global__ void __launch_bounds__(BULK_THREADS) dmd_bulk_predict_kernel(DevArrays d, int r0) {
  Rep r;
  int k;
  if (!bulk_bind(r, d, r0, k)) return;
  redo_lane(r, k);
}
// per-group calendar minima of the warp-per-replica engine: one warp per group of 32 entries
__global__ void __launch_bounds__(BULK_THREADS) dmd_bulk_groups_kernel(DevArrays d, int r0) {
  Rep r;
  rep_bind(r, d, staged_global(d), nullptr, r0 + blockIdx.y);
  const int g = blockIdx.x * (BULK_THREADS / 32) + (threadIdx.x >> 5);
  if (g < r.G) group_min_update(r, g);
}

// device state -> the reference's sv(6,N) and bptnr(N) (1-based), for dmdb_get_state_all
__global__ void dmd_pack_kernel(DevArrays d, double* sv, int32_t* bp) {
  const size_t n = (size_t)d.n_replicas * d.n_beads;
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
    const BeadRec b = d.rec[k];
    double* o = sv + 6 * k;
    o[0] = b.x; o[1] = b.y; o[2] = b.z; o[3] = b.vx; o[4] = b.vy; o[5] = b.vz;
    bp[k] = b.bptnr + 1;
  }
}

__global__ void dmd_evcode_kernel(DevArrays d, int rid, int n_pairs, int32_t* buf) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_pairs) return;
  const int N = d.sys->N;
  int i = buf[k] - 1, j = buf[n_pairs + k] - 1;
  const BeadRec* rec = d.rec + (size_t)rid * N;
  int sc = static_code(*d.sys, d.meta[i], d.chain[i], i, d.meta[j], d.chain[j], j);
  buf[2 * n_pairs + k] = overlay_code(sc, i, rec[i], j, rec[j]);
}


// ---- replica exchange on the device (dmd_exchange.h): (E_pot, T*) of the local replicas, the decision for every
// ladder of the gathered set, and the selection of the local replicas whose temperature changes
__global__ void dmd_xch_pack_kernel(DevArrays d, const OutRec* eout, const double* tstar, double* local) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= d.n_replicas) return;
  local[2 * r] = eout[r].ered - 0.5 * eout[r].sumvel;  // e_int of main.F90:358
  local[2 * r + 1] = tstar[r];  // the T* the host set (setemp / 12 would not give back the same bits)
}
__global__ void dmd_xch_decide_kernel(const double* et, double* tnew, int n_ladders, int L, int world, int R, long long step,
                                      unsigned long long seed, XchCounts* cnt) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  int att = 0, acc = 0;
  if (l < n_ladders) xch_decide_ladder(et, tnew, l, L, world, R, step, seed, att, acc);
  att = __reduce_add_sync(0xffffffffu, att);
  acc = __reduce_add_sync(0xffffffffu, acc);
  if ((threadIdx.x & 31) == 0 && att) {
    atomicAdd(&cnt->attempted, att);
    atomicAdd(&cnt->accepted, acc);
  }
}
__global__ void dmd_xch_init_kernel(const double* et, double* tnew, int M, XchCounts* cnt, int n_ladders) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < M) tnew[g] = et[2 * g + 1];  // replicas outside every ladder keep their temperature
  if (g == 0) {
    cnt->attempted = cnt->accepted = cnt->changed_local = 0;
    cnt->ladders = n_ladders;
  }
}
__global__ void dmd_xch_select_kernel(const double* et, const double* tnew, int rank, int R, double* tsel, double* tcur, XchCounts* cnt) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  int ch = 0;
  if (r < R) {
    const int g = rank * R + r;
    const double tn = tnew[g], to = et[2 * g + 1];
    ch = tn != to;
    tsel[r] = ch ? tn : 0.0;  // dmd_retemp_kernel: <= 0 leaves the replica untouched
    tcur[r] = tn;
  }
  ch = __reduce_add_sync(0xffffffffu, ch);
  if ((threadIdx.x & 31) == 0 && ch) atomicAdd(&cnt->changed_local, ch);
}

}  // namespace dmd

namespace be {

// One CUDA device per process (one process per GPU, like every multi-GPU path of this library): the stream, the timing
// events and the whole-GPU engine's workspace below belong to that device.  A handle on a second device is refused
// (init), and every entry point re-selects the device first (bind), so a host that changes the current device between
// calls -- torch.cuda.set_device, another library -- cannot send our copies and launches elsewhere.
static cudaStream_t g_stream = nullptr;
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
static int g_device = -1;
inline void bind() {
  if (g_device >= 0) cudaSetDevice(g_device);
}

inline bool init(int device, std::string& err) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    err = std::string("libdmdb200 needs a CUDA device (sm_100a); none found: ") + cudaGetErrorString(e);
    return false;
  }
  if (device < 0 || device >= n) {
    err = "CUDA device ordinal out of range";
    return false;
  }
  if (g_device >= 0 && device != g_device) {
    err = "this process already drives CUDA device " + std::to_string(g_device) + ": libdmdb200 uses one device per process "
          "(start one process per GPU)";
    return false;
  }
  if ((e = cudaSetDevice(device)) != cudaSuccess) {
    err = cudaGetErrorString(e);
    return false;
  }
  g_device = device;
  if (!g_stream) {
    cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking);
    cudaEventCreate(&g_ev0);
    cudaEventCreate(&g_ev1);
  }
  return true;
}
inline void* alloc(size_t n) {
  bind();
  void* p = nullptr;
  CUDA_OK(cudaMalloc(&p, n ? n : 1));
  return p;
}
inline void release(void* p) { cudaFree(p); }
inline void h2d(void* d, const void* h, size_t n) {
  bind();
  CUDA_OK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, g_stream));
  CUDA_OK(cudaStreamSynchronize(g_stream));
}
inline void d2h(void* h, const void* d, size_t n) {
  bind();
  CUDA_OK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, g_stream));
  CUDA_OK(cudaStreamSynchronize(g_stream));
}
inline void zero(void* d, size_t n) {
  bind();
  CUDA_OK(cudaMemsetAsync(d, 0, n, g_stream));
}
inline void fill_i32(int32_t* d, int v, size_t n) {
  if (v == -1) CUDA_OK(cudaMemsetAsync(d, 0xff, n * 4, g_stream));
  else if (v == 0) CUDA_OK(cudaMemsetAsync(d, 0, n * 4, g_stream));
  else throw std::runtime_error("fill_i32: unsupported value");
}

static size_t g_smem_optin = 0;
inline size_t smem_optin() {
  if (!g_smem_optin) {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    g_smem_optin = (size_t)v;
  }
  return g_smem_optin;
}
// the CTA-per-replica engine needs at least its fixed part (tables, claims, undo logs) in shared memory
inline bool block_engine_fits(const dmd::SysConst& s) {
  return dmd::blk_layout(s.N, s.ngroups * 32, smem_optin()).total <= smem_optin();
}

// Service split of the event-loop kernel (measured on B200, 48-peptide box, 148 CTAs in all, final round-1 build with
// two service groups per CTA, 131 us per rebuild and CTA: 14 service CTAs 1.73e8 events/s, 16: 1.80e8 -- the service
// saturates and warps take their requests back --, 18: 2.11e8, 20: see DESIGN.md, 22: 2.08e8; without the service
// 1.5e8): about one service CTA per 6.4 event-loop CTAs, on the safe side of the cliff
inline int sm_count() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}
// one service CTA per 6.4 event-loop CTAs with one or two replicas per warp, per 5.7 with four (an event-loop CTA then
// processes ~1.5 x the events and asks for as many more rebuilds): 126 + 22 CTAs on a 148-SM B200.  Measured on the
// headline workload (tools/prof_run.py, 40 000 events per replica after 60 000): 16: 2.70e8, 18: 2.99e8, 20: 3.21e8,
// 22: 3.19e8, 24: 3.13e8, 26: 3.12e8, 28: 3.06e8 events/s; the ladder workload of bench.py --ladder (every replica
// restarts its lists at every exchange) 20: 2.97e8, 24: 3.01e8.  Below ~20 the service saturates and the warps queue.
// The count must be EVEN: the two SMs of a TPC get consecutive CTAs, an odd count puts one event-loop CTA next to a
// service CTA, that CTA runs ~9 % slower than the others (it shares the TPC's instruction cache with the rebuild code)
// and the launch ends when the last CTA does (measured: 23: 2.78e8, 24: 3.09e8, 25: 2.80e8, 26: 3.05e8).
inline int default_service_ctas(int worker_ctas) {
  if (worker_ctas < 32) return 0;
  return ((dmd::EVL_RPW >= 4 ? (worker_ctas * 22 + 64) / 128 : (worker_ctas * 10 + 32) / 64) + 1) & ~1;
}
inline void device_fill(int& replicas, int& service) {
  const int sms = sm_count();
  int w = sms;
  while (w > 1 && w + default_service_ctas(w) > sms) w--;
  replicas = w * dmd::EVL_RPC;
  service = default_service_ctas(w);
}

inline dim3 bulk_grid(const dmd::DevArrays& d, int nrep, int n) { return dim3((n + dmd::BULK_THREADS - 1) / dmd::BULK_THREADS, nrep); }
inline void launch_nbor(const dmd::DevArrays& d, int r0, int nrep) {  // = nbor(): cell_add.f + nbor.f
  using namespace dmd;
  const dim3 g = bulk_grid(d, nrep, d.n_beads);
  dmd_bulk_cell_kernel<<<g, BULK_THREADS, 0, g_stream>>>(d, r0);
  dmd_bulk_nbor_kernel<<<g, BULK_THREADS, 0, g_stream>>>(d, r0);
  dmd_bulk_clear_kernel<<<g, BULK_THREADS, 0, g_stream>>>(d, r0);
}
inline void launch_predict_all(const dmd::DevArrays& d, int r0, int nrep) {  // = events()
  using namespace dmd;
  dmd_bulk_predict_kernel<<<bulk_grid(d, nrep, d.n_beads), BULK_THREADS, 0, g_stream>>>(d, r0);
  const int groups = d.cal_stride / 32;
  dmd_bulk_groups_kernel<<<dim3((groups + BULK_THREADS / 32 - 1) / (BULK_THREADS / 32), nrep), BULK_THREADS, 0, g_stream>>>(d, r0);
}
inline void run_init(const dmd::DevArrays& d, int r0, int nrep, const double* sv, size_t sv_stride, const int32_t* bp,
                     size_t bp_stride, const double* tstar, unsigned long long seed0) {
  using namespace dmd;
  bind();
  const int n = d.cal_stride > d.n_beads ? d.cal_stride : d.n_beads;
  dmd_bulk_reset_kernel<<<(nrep + 255) / 256, 256, 0, g_stream>>>(d, r0, nrep);
  dmd_bulk_init_kernel<<<bulk_grid(d, nrep, n), BULK_THREADS, 0, g_stream>>>(d, r0, sv, sv_stride, bp, bp_stride, tstar, seed0);
  const dim3 g = bulk_grid(d, nrep, d.n_beads);
  dmd_bulk_cell_kernel<<<g, BULK_THREADS, 0, g_stream>>>(d, r0);
  dmd_bulk_hits_kernel<<<g, BULK_THREADS, 0, g_stream>>>(d, r0);
  dmd_bulk_clear_kernel<<<g, BULK_THREADS, 0, g_stream>>>(d, r0);
  dmd_fixup_kernel<<<(nrep + WARPS_PER_CTA - 1) / WARPS_PER_CTA, WARPS_PER_CTA * 32, 0, g_stream>>>(d, r0, nrep);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(g_stream));
}
inline int sheet_max_chains() { return dmd::SHEET_MAX_CHAINS; }
inline void run_sheets(const dmd::DevArrays& d, int32_t* out_dev) {
  using namespace dmd;
  bind();
  dmd_sheet_kernel<<<(d.n_replicas + SHEET_WARPS - 1) / SHEET_WARPS, SHEET_WARPS * 32, 0, g_stream>>>(d, d.n_replicas, out_dev);
  CUDA_OK(cudaGetLastError());
}
inline void run_pack(const dmd::DevArrays& d, double* sv, int32_t* bp) {
  bind();
  dmd::dmd_pack_kernel<<<148 * 8, 256, 0, g_stream>>>(d, sv, bp);
  CUDA_OK(cudaGetLastError());
}

inline bool grid_engine_available() { return true; }
// engine 3: one large system on the whole GPU.  The grid kernel commits batches of plain pair events until the head
// of the calendar is something else (ghost, interval, output, H-bond event); that one entry is processed by the
// warp-per-replica engine (after refreshing its group minima) and the grid kernel is relaunched.
static dmd::GridShared* g_grid = nullptr;
static uint32_t* g_grid_claim = nullptr;
static size_t g_grid_claim_n = 0;
inline int run_grid(const dmd::DevArrays& d, int r0, int nrep, long long n_events) {
  using namespace dmd;
  int dev = 0, sms = 0, coop = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  if (!coop) throw std::runtime_error("device does not support cooperative launch");
  if (!g_grid) g_grid = (GridShared*)alloc(sizeof(GridShared));
  const size_t nclaim = (size_t)d.n_beads + 3;
  if (g_grid_claim_n < nclaim) {
    if (g_grid_claim) release(g_grid_claim);
    g_grid_claim = (uint32_t*)alloc(nclaim * 4);
    g_grid_claim_n = nclaim;
  }
  CUDA_OK(cudaMemsetAsync(g_grid_claim, 0xff, nclaim * 4, g_stream));
  if (sms * BK_MAXW > GW) sms = GW / BK_MAXW;
  int launches = 0;
  const size_t header = offsetof(GridShared, cand_t);
  for (int rid = r0; rid < r0 + nrep; rid++) {
    RepScalars sc;
    d2h(&sc, d.scal + rid, sizeof(sc));
    if (sc.error) continue;
    const long long target = sc.coll + n_events;
    static GridShared hdr;  // only its header part is used on the host
    std::memset(&hdr, 0, header);
    hdr.window = sc.interval * 0.02;
    while (sc.coll < target && !sc.error) {
      hdr.tlast = -1.0;
      hdr.coll = sc.coll;
      hdr.target = target;
      hdr.tmin_bits = ~0ull;
      hdr.barrier = 0;
      hdr.n_cand = 0;
      hdr.first_cold = 0x7fffffff;
      hdr.first_lost = 0x7fffffff;
      hdr.status = 0;
      hdr.error = 0;
      hdr.n_log = sc.n_log;
      hdr.st_rounds = hdr.st_exec = hdr.st_rollback = hdr.st_conflict = 0;
      std::memset(hdr.nevents, 0, sizeof(hdr.nevents));
      std::memset(hdr.cyc, 0, sizeof(hdr.cyc));
      CUDA_OK(cudaMemcpyAsync(g_grid, &hdr, header, cudaMemcpyHostToDevice, g_stream));
      DevArrays dd = d;
      GridShared* S = g_grid;
      uint32_t* claim = g_grid_claim;
      void* args[] = {&dd, &rid, &S, &claim};
      const auto tk0 = std::chrono::steady_clock::now();
      CUDA_OK(cudaLaunchCooperativeKernel((void*)dmd_grid_loop_kernel, dim3(sms), dim3(BK_MAXW * 32), args, 0, g_stream));
      launches++;
      CUDA_OK(cudaMemcpyAsync(&hdr, g_grid, header, cudaMemcpyDeviceToHost, g_stream));
      CUDA_OK(cudaStreamSynchronize(g_stream));
      if (getenv("DMDB_DEBUG")) {
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tk0).count();
        fprintf(stderr, "grid launch: %.3f ms, status %d, coll %lld, rounds %lld\n", ms, hdr.status, hdr.coll, hdr.st_rounds);
      }
      d2h(&sc, d.scal + rid, sizeof(sc));
      if (hdr.nevents[30] || hdr.nevents[31]) {  // work counters of the warps other than the master
        sc.n_pair_pred += hdr.nevents[30];
        sc.n_nbr_visits += hdr.nevents[31];
        h2d(d.scal + rid, &sc, sizeof(sc));
      }
      if (hdr.status == 1 && !sc.error && sc.coll < target && hdr.head_owner == d.n_beads + 1) {
        // the interval pseudo-event, grid-wide
        const dim3 gb = bulk_grid(d, 1, d.n_beads + 3);
        dmd_interval_begin_kernel<<<1, 1, 0, g_stream>>>(d, rid);
        dmd_interval_advance_kernel<<<gb.x, BULK_THREADS, 0, g_stream>>>(d, rid);
        dmd_interval_decide_kernel<<<1, 1, 0, g_stream>>>(d, rid);
        launches += 3;
        CUDA_OK(cudaGetLastError());
        d2h(&sc, d.scal + rid, sizeof(sc));
        if (sc.pad[1]) {  // lists + calendar rebuilt
          dmd_interval_wrap_kernel<<<gb.x, BULK_THREADS, 0, g_stream>>>(d, rid);
          launch_nbor(d, rid, 1);
          dmd_bulk_predict_kernel<<<bulk_grid(d, 1, d.n_beads), BULK_THREADS, 0, g_stream>>>(d, rid);
          launches += 5;
          CUDA_OK(cudaGetLastError());
          d2h(&sc, d.scal + rid, sizeof(sc));
        }
      } else if (hdr.status == 1 && !sc.error && sc.coll < target) {  // one calendar entry for the serial engine
        dmd_bulk_groups_kernel<<<dim3((d.cal_stride / 32 + BULK_THREADS / 32 - 1) / (BULK_THREADS / 32), 1), BULK_THREADS, 0, g_stream>>>(d, rid);
        const auto tc0 = std::chrono::steady_clock::now();
        dmd_event_loop_kernel<<<1, WARPS_PER_CTA * 32, 0, g_stream>>>(d, rid, 1, 1, 0, 0);
        launches += 2;
        CUDA_OK(cudaGetLastError());
        d2h(&sc, d.scal + rid, sizeof(sc));
        if (getenv("DMDB_DEBUG"))
          fprintf(stderr, "cold event: %.3f ms\n",
                  std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count());
      } else if (hdr.status == 2) {
        break;
      }
    }
  }
  return launches;
}

// one launcher per device operation; times the kernel with CUDA events on the launching stream
inline void run_op(const dmd::DevArrays& d, int op, int r0, int nrep, long long arg, int32_t* ibuf, dmd::OutRec* eout,
                   double* ms, int* launches, int flags = 0) {
  using namespace dmd;
  const int block = WARPS_PER_CTA * 32;
  const int grid = (nrep + WARPS_PER_CTA - 1) / WARPS_PER_CTA;  // 32-lane kernels: one replica per hardware warp
  const int grid_evl = (nrep + EVL_RPC - 1) / EVL_RPC;          // event loop: DMD_EVL_W lanes per replica
  int nl = 1;
  bind();
  CUDA_OK(cudaEventRecord(g_ev0, g_stream));
  switch (op) {
    case 0: launch_nbor(d, r0, nrep); launch_predict_all(d, r0, nrep); nl = 5; break;
    case 1: launch_nbor(d, r0, nrep); nl = 3; break;
    case 2: launch_predict_all(d, r0, nrep); nl = 2; break;
    case 3: {
      static bool carve_done = false;  // tuning knob: shared-memory carve-out (percent) of the event-loop kernel
      if (!carve_done) {
        carve_done = true;
        if (const char* cv = getenv("DMDB_CARVEOUT"))
          CUDA_OK(cudaFuncSetAttribute(dmd_event_loop_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(cv)));
      }
      // list-rebuild service CTAs (svc_serve_in_kernel)
      int n_srv = 0;
      {
        const int sms = sm_count();
        const int asked = ((flags >> 8) & 0xffff) - 1;  // dmdb_set_service_ctas; -1 = automatic
        const char* sv = getenv("DMDB_SVC");          // tuning override
        if (sv || asked >= 0) {
          const int want = sv ? atoi(sv) : asked;
          n_srv = want < sms - 1 ? want : sms - 1;
        } else if (grid_evl <= sms) {  // one wave: as many service CTAs as fit beside the event-loop CTAs
          const int want = default_service_ctas(grid_evl);
          n_srv = (want < sms - grid_evl ? want : sms - grid_evl) & ~1;  // even: see default_service_ctas
          if (3 * n_srv < want) n_srv = 0;  // too few would only make the warps wait (measured: 4 of 22 is a loss)
        } else {  // several waves: event-loop CTAs take turns on the SMs the service CTAs leave free, if that is faster
          int w = sms, srv = 0;
          device_fill(w, srv);
          w /= EVL_RPC;
          // per event-loop SM: 1.37e6 events/s with the service, 0.93e6 without
          if (srv > 0 && ((grid_evl + w - 1) / w) / 1.37 < ((grid_evl + sms - 1) / sms) / 0.93) n_srv = srv;
        }
        if (n_srv < 0) n_srv = 0;
      }
      flags &= 0xff;
      if (n_srv > 0) CUDA_OK(cudaMemsetAsync(d.svc_ctl, 0, SVC_CTL_WORDS * 8, g_stream));
#if defined(DMD_PHASE_PROF)
      {
        unsigned long long z[16] = {0};
        CUDA_OK(cudaMemcpyToSymbolAsync(evl::g_phase_cyc, z, sizeof(z), 0, cudaMemcpyHostToDevice, g_stream));
      }
#endif
      dmd_event_loop_kernel<<<grid_evl + n_srv, block, 0, g_stream>>>(d, r0, nrep, arg, flags, n_srv);
#if defined(DMD_PHASE_PROF)
      if (getenv("DMDB_DEBUG")) {
        unsigned long long c[16];
        CUDA_OK(cudaStreamSynchronize(g_stream));
        CUDA_OK(cudaMemcpyFromSymbol(c, evl::g_phase_cyc, sizeof(c)));
        const double ev = (double)nrep * (double)arg;
        static const char* nm[12] = {"flush", "pop", "event dynamics", "main pass", "cascade prep", "cascade pass", "interval/output",
                                     "loop", "H-bond event", "ghost event", "wait for partner", "wait for service"};
        double tot = 0;
        for (int k = 0; k < 12; k++) tot += (double)c[k];
        for (int k = 0; k < 12; k++) fprintf(stderr, "phase %-16s %9.1f cycles/event  %5.1f %%\n", nm[k], c[k] / ev, 100.0 * c[k] / tot);
        fprintf(stderr, "phase total         %9.1f cycles/event\n", tot / ev);
        {  // spread of the event-loop CTAs' finishing times (the launch ends with the last one)
          static unsigned long long e[512];
          static unsigned sm[512];
          CUDA_OK(cudaMemcpyFromSymbol(e, evl::g_cta_end, sizeof(e)));
          CUDA_OK(cudaMemcpyFromSymbol(sm, evl::g_cta_sm, sizeof(sm)));
          unsigned long long lo = ~0ull, hi = 0;
          for (int b = n_srv; b < grid_evl + n_srv && b < 512; b++) {
            if (e[b] < lo) lo = e[b];
            if (e[b] > hi) hi = e[b];
          }
          fprintf(stderr, "event-loop CTAs finish within %.3f ms of one another; (ms before the last one : SM)", (hi - lo) * 1e-6);
          for (int b = n_srv; b < grid_evl + n_srv && b < 512; b++) fprintf(stderr, " %.1f:%u", (hi - e[b]) * 1e-6, sm[b]);
          fprintf(stderr, "\n");
          unsigned long long z[512] = {0};
          CUDA_OK(cudaMemcpyToSymbol(evl::g_cta_end, z, sizeof(z)));
        }
      }
#endif
      if (n_srv > 0 && getenv("DMDB_DEBUG")) {
        unsigned long long ctl[SVC_CTL_WORDS];
        d2h(ctl, d.svc_ctl, sizeof(ctl));
        fprintf(stderr, "service: %d CTAs, %llu worker warps done, %llu rebuilds served, %.1f us each (cells+bounds %.1f, lists %.1f), "
                "%.1f ms busy per CTA; %llu requests taken back, mean wait %.1f us\n", n_srv, ctl[0], ctl[1],
                ctl[1] ? (double)ctl[4] / ctl[1] / 1.9e3 : 0.0, ctl[1] ? (double)ctl[5] / ctl[1] / 1.9e3 : 0.0,
                ctl[1] ? (double)ctl[6] / ctl[1] / 1.9e3 : 0.0, (double)ctl[4] / n_srv / 1.9e6, ctl[2],
                ctl[1] + ctl[2] ? (double)ctl[3] / (ctl[1] + ctl[2]) / 1.9e3 : 0.0);
      }
      break;
    }
    case 4: dmd_sync_positions_kernel<<<grid, block, 0, g_stream>>>(d, r0, nrep); break;
    case 5: dmd_energy_kernel<<<grid, block, 0, g_stream>>>(d, r0, nrep, eout); break;
    case 6: dmd_evcode_kernel<<<((int)arg + 127) / 128, 128, 0, g_stream>>>(d, r0, (int)arg, ibuf); break;
    case 7: dmd_retemp_kernel<<<grid, block, 0, g_stream>>>(d, r0, nrep, (const double*)ibuf); break;
    case 8: {
      int dev = 0, sms = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const BlkLayout L = blk_layout(d.n_beads, d.cal_stride, smem_optin());  // host-side copies: d.sys is a device pointer
      if (L.total > smem_optin()) throw std::runtime_error("system too large for the CTA-per-replica engine");
      CUDA_OK(cudaFuncSetAttribute(dmd_block_loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
      const int g = nrep < sms ? nrep : sms;
      dmd_block_loop_kernel<<<g, BK_MAXW * 32, L.total, g_stream>>>(d, r0, nrep, arg, (unsigned)smem_optin(), flags);
      break;
    }
    case 9: nl = run_grid(d, r0, nrep, arg); break;
    default: throw std::runtime_error("unknown device op");
  }
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(g_ev1, g_stream));
  CUDA_OK(cudaStreamSynchronize(g_stream));
  float t = 0;
  CUDA_OK(cudaEventElapsedTime(&t, g_ev0, g_ev1));
  if (ms) *ms = t;
  if (launches) *launches = nl;
}


// ---- NCCL, bound at run time (no link-time dependency: a host that never exchanges needs no NCCL).  When the host
// process already has NCCL loaded (torch.distributed, or a Fortran/C++ host linked against it), dlopen by soname
// returns that very instance, so communicators created by the host can be passed in.
namespace nccl {
struct UniqueId {
  char internal[128];
};
typedef int (*GetUniqueId_t)(UniqueId*);
typedef int (*CommInitRank_t)(void**, int, UniqueId, int);
typedef int (*CommDestroy_t)(void*);
typedef int (*AllGather_t)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*GetErrorString_t)(int);
typedef int (*CommQuery_t)(void*, int*);
static void* lib = nullptr;
static GetUniqueId_t GetUniqueId = nullptr;
static CommInitRank_t CommInitRank = nullptr;
static CommDestroy_t CommDestroy = nullptr;
static AllGather_t AllGather = nullptr;
static GetErrorString_t GetErrorString = nullptr;
static CommQuery_t CommCount = nullptr, CommUserRank = nullptr;
constexpr int kDouble = 8;  // ncclFloat64
inline void load() {
  if (lib) return;
  const char* names[] = {getenv("DMDB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n) continue;
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) throw std::runtime_error("NCCL not found (libnccl.so.2; set DMDB_NCCL_LIB): replica exchange across GPUs needs it");
  GetUniqueId = (GetUniqueId_t)dlsym(lib, "ncclGetUniqueId");
  CommInitRank = (CommInitRank_t)dlsym(lib, "ncclCommInitRank");
  CommDestroy = (CommDestroy_t)dlsym(lib, "ncclCommDestroy");
  AllGather = (AllGather_t)dlsym(lib, "ncclAllGather");
  GetErrorString = (GetErrorString_t)dlsym(lib, "ncclGetErrorString");
  CommCount = (CommQuery_t)dlsym(lib, "ncclCommCount");
  CommUserRank = (CommQuery_t)dlsym(lib, "ncclCommUserRank");
  if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllGather) throw std::runtime_error("NCCL symbols missing");
}
inline void ok(int rc, const char* what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + ": " + (GetErrorString ? GetErrorString(rc) : "NCCL error"));
}
}  // namespace nccl
inline void nccl_unique_id(char out[128]) {
  nccl::load();
  nccl::UniqueId id;
  nccl::ok(nccl::GetUniqueId(&id), "ncclGetUniqueId");
  std::memcpy(out, id.internal, 128);
}
inline void* nccl_comm_init(const char id128[128], int world, int rank) {
  nccl::load();
  nccl::UniqueId id;
  std::memcpy(id.internal, id128, 128);
  void* comm = nullptr;
  nccl::ok(nccl::CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
  return comm;
}
inline void nccl_comm_geometry(void* comm, int& world, int& rank) {  // of a communicator the host created itself
  nccl::load();
  if (!nccl::CommCount || !nccl::CommUserRank) throw std::runtime_error("NCCL symbols missing");
  nccl::ok(nccl::CommCount(comm, &world), "ncclCommCount");
  nccl::ok(nccl::CommUserRank(comm, &rank), "ncclCommUserRank");
}
inline void nccl_comm_destroy(void* comm) {
  if (comm && nccl::CommDestroy) nccl::CommDestroy(comm);
}

// One exchange step on the library's stream: energy kernel -> (E_pot, T*) -> all-gather (NCCL; or the host's own gather
// passed in `gathered_host`) -> decision kernel -> retemp kernel for the local replicas whose temperature changed.
// xb: device scratch [2R local | 2M gathered | M new temperatures | R selected | R current | XchCounts].
inline void exchange(const dmd::DevArrays& d, dmd::OutRec* eout, double* xb, void* comm, const double* gathered_host, int world,
                     int rank, long long step, unsigned long long seed, int L, dmd::XchCounts* counts_out, double* tstar_out,
                     double* ms, int* launches) {
  using namespace dmd;
  bind();
  const int R = d.n_replicas, M = world * R, n_ladders = M / L;
  double* local = xb;
  double* all = world > 1 ? xb + 2 * (size_t)R : local;
  double* tnew = xb + 2 * (size_t)R + 2 * (size_t)M;
  double* tsel = tnew + M;
  double* tcur = tsel + R;
  XchCounts* cnt = reinterpret_cast<XchCounts*>(tcur + R);
  const int block = WARPS_PER_CTA * 32, grid = (R + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  CUDA_OK(cudaEventRecord(g_ev0, g_stream));
  int nl = 0;
  if (!gathered_host) {
    CUDA_OK(cudaMemcpyAsync(tcur, tstar_out, sizeof(double) * R, cudaMemcpyHostToDevice, g_stream));  // current T* (in/out)
    dmd_energy_kernel<<<grid, block, 0, g_stream>>>(d, 0, R, eout);
    dmd_xch_pack_kernel<<<(R + 255) / 256, 256, 0, g_stream>>>(d, eout, tcur, local);
    nl += 2;
    if (world > 1) {
      if (!comm) throw std::runtime_error("dmdb_exchange: more than one rank needs an NCCL communicator (dmdb_comm_init)");
      nccl::load();
      nccl::ok(nccl::AllGather(local, all, 2 * (size_t)R, nccl::kDouble, comm, g_stream), "ncclAllGather");
    }
  } else {
    CUDA_OK(cudaMemcpyAsync(all, gathered_host, sizeof(double) * 2 * (size_t)M, cudaMemcpyHostToDevice, g_stream));
  }
  dmd_xch_init_kernel<<<(M + 255) / 256, 256, 0, g_stream>>>(all, tnew, M, cnt, n_ladders);
  if (n_ladders > 0) dmd_xch_decide_kernel<<<(n_ladders + 127) / 128, 128, 0, g_stream>>>(all, tnew, n_ladders, L, world, R, step, seed, cnt);
  dmd_xch_select_kernel<<<(R + 255) / 256, 256, 0, g_stream>>>(all, tnew, rank, R, tsel, tcur, cnt);
  dmd_retemp_kernel<<<grid, block, 0, g_stream>>>(d, 0, R, tsel);
  nl += 4;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(g_ev1, g_stream));
  CUDA_OK(cudaMemcpyAsync(counts_out, cnt, sizeof(XchCounts), cudaMemcpyDeviceToHost, g_stream));
  CUDA_OK(cudaMemcpyAsync(tstar_out, tcur, sizeof(double) * R, cudaMemcpyDeviceToHost, g_stream));
  CUDA_OK(cudaStreamSynchronize(g_stream));
  float t = 0;
  CUDA_OK(cudaEventElapsedTime(&t, g_ev0, g_ev1));
  if (ms) *ms = t;
  if (launches) *launches = nl;
}

}  // namespace be

#include "dmd_capi_impl.h"
