// trace_lib.cpp -- TEST SCAFFOLDING, never shipped and never loaded by the product path.
//
// Compiles the engine source (csrc/dmd_engine.h) with DMD_HOST_TRACE: a 1-lane "warp" executed on the CPU,
// behind the same C ABI, so that the event-loop LOGIC (calendar, cascades, bookkeeping, rebuilds) can be
// diffed against the oracle in the GPU-less build container (tests/test_engine_logic_hosttrace.py).
// It exercises none of the 32-lane collectives; the real parity tests are the `-m gpu` ones.
#define DMD_HOST_TRACE 1
#define DMD_W 1  // one lane per replica (dmd_warp.h)
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include <pthread.h>

#include <thread>
#include <vector>

#include "../../parallel_dmd_for_biomolecules_b200/csrc/dmd_block.h"  // includes dmd_engine.h (no include guard: once)
#include "../../parallel_dmd_for_biomolecules_b200/csrc/dmd_exchange.h"
#include "../../parallel_dmd_for_biomolecules_b200/csrc/dmd_types.h"

// The CTA-per-replica engine (dmd_block.h) is emulated with one host thread per (1-lane) virtual warp; its
// __syncthreads() becomes a pthread barrier, so claims, validation and rollback run with real concurrency.
namespace dmd {
static thread_local pthread_barrier_t* g_blk_barrier = nullptr;
void blk_sync() { pthread_barrier_wait(g_blk_barrier); }

static int trace_block_warps() {
  const char* e = std::getenv("DMDB_TRACE_WARPS");
  int n = e ? std::atoi(e) : 8;
  return n < 1 ? 1 : (n > BK_MAXW ? BK_MAXW : n);
}

static void trace_run_block(const DevArrays& d, int rid, long long n_events, int flags) {
  const int nw = trace_block_warps();
  const int N = d.sys->N;
  BlkShared* S = new BlkShared();
  std::memset(S, 0, sizeof(*S));
  std::vector<uint32_t> claim(N + 3, CLAIM_FREE);
  std::vector<int32_t> cq((size_t)nw * CQ_CAP);
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, nullptr, nw);
  {
    Rep r0;
    rep_bind(r0, d, staged_global(d), cq.data(), rid);
    S->coll = r0.coll;
    S->target = r0.coll + n_events;
    S->stop_at_output = flags & 1;
    S->window = r0.interval * 0.02;
    S->tlast = -1.0;
    S->error = r0.error;
    S->error_info = r0.error_info;
  }
  std::vector<std::thread> th;
  for (int w = 0; w < nw; w++)
    th.emplace_back([&, w]() {
      g_blk_barrier = &bar;
      Rep r;
      rep_bind(r, d, staged_global(d), cq.data() + (size_t)w * CQ_CAP, rid);
      if (w != 0) r.n_pair_pred = r.n_nbr_visits = 0;
      blk_run(*S, r, claim.data(), w, nw);
      if (w != 0) {
        harvest_counters(r);
        blk_atomic_add64(&S->n_pair_pred, r.n_pair_pred);
        blk_atomic_add64(&S->n_nbr_visits, r.n_nbr_visits);
      }
      blk_sync();
      if (w == 0) {
        r.n_pair_pred += S->n_pair_pred;
        r.n_nbr_visits += S->n_nbr_visits;
        if (S->error && !r.error) {
          r.error = S->error;
          r.error_info = S->error_info;
        }
        rep_save(r);
        rebuild_all_groups(r);
        for (int q = 0; q < 32; q++) r.sc->nevents[q] += S->nevents[q];
        long long* st = d.blkstat + (size_t)rid * 16;
        st[0] += S->st_rounds; st[1] += S->st_exec; st[2] += S->st_rollback; st[3] += S->st_conflict; st[4] += S->st_cold;
      }
    });
  for (auto& t : th) t.join();
  pthread_barrier_destroy(&bar);
  delete S;
}
}  // namespace dmd

namespace be {
inline bool init(int, std::string&) { return true; }
inline void* alloc(size_t n) { return std::calloc(n ? n : 1, 1); }
inline void release(void* p) { std::free(p); }
inline void h2d(void* d, const void* h, size_t n) { std::memcpy(d, h, n); }
inline void d2h(void* h, const void* d, size_t n) { std::memcpy(h, d, n); }
inline void zero(void* d, size_t n) { std::memset(d, 0, n); }
inline void fill_i32(int32_t* d, int v, size_t n) {
  for (size_t k = 0; k < n; k++) d[k] = v;
}
inline bool block_engine_fits(const dmd::SysConst&) { return true; }
inline bool grid_engine_available() { return false; }  // device only
inline void device_fill(int& replicas, int& service) { replicas = 1; service = 0; }
inline void run_init(const dmd::DevArrays& d, int r0, int nrep, const double* sv, size_t sv_stride, const int32_t* bp,
                     size_t bp_stride, const double* tstar, unsigned long long seed0) {
  using namespace dmd;
  for (int rid = r0; rid < r0 + nrep; rid++) {
    Rep r;
    int32_t cq[CQ_CAP];
    rep_bind(r, d, staged_global(d), cq, rid);
    const size_t k = (size_t)(rid - r0);
    init_replica(r, sv + k * sv_stride, bp ? bp + k * bp_stride : nullptr, tstar[rid], seed0 + (unsigned long long)rid,
                 d.nc_beads, d.n_nc, d.cal_stride);
    rep_save(r);
  }
}
inline int sheet_max_chains() { return 1 << 20; }
inline void run_sheets(const dmd::DevArrays& d, int32_t* out) {
  using namespace dmd;
  const int nc = d.sys->N / d.sys->numbeads[0];
  std::vector<uint8_t> hb((size_t)nc * nc);
  std::vector<int32_t> lab(nc);
  for (int rid = 0; rid < d.n_replicas; rid++) {
    Rep r;
    rep_bind(r, d, staged_global(d), nullptr, rid);
    sheet_observables(r, out + 8 * (size_t)rid, hb.data(), lab.data());
  }
}
inline void run_pack(const dmd::DevArrays& d, double* sv, int32_t* bp) {
  const size_t n = (size_t)d.n_replicas * d.n_beads;
  for (size_t k = 0; k < n; k++) {
    const dmd::BeadRec b = d.rec[k];
    double* o = sv + 6 * k;
    o[0] = b.x; o[1] = b.y; o[2] = b.z; o[3] = b.vx; o[4] = b.vy; o[5] = b.vz;
    bp[k] = b.bptnr + 1;
  }
}
inline void run_op(const dmd::DevArrays& d, int op, int r0, int nrep, long long arg, int32_t* ibuf, dmd::OutRec* eout,
                   double* ms, int* launches, int flags = 0) {
  using namespace dmd;
  if (op == 6) {
    const int N = d.sys->N, n_pairs = (int)arg;
    const BeadRec* rec = d.rec + (size_t)r0 * N;
    for (int k = 0; k < n_pairs; k++) {
      int i = ibuf[k] - 1, j = ibuf[n_pairs + k] - 1;
      int sc = static_code(*d.sys, d.meta[i], d.chain[i], i, d.meta[j], d.chain[j], j);
      ibuf[2 * n_pairs + k] = overlay_code(sc, i, rec[i], j, rec[j]);
    }
  } else {
    for (int rid = r0; rid < r0 + nrep; rid++) {
      Rep r;
      int32_t cq[CQ_CAP];
      rep_bind(r, d, staged_global(d), cq, rid);
      switch (op) {
        case 0:
          nbor(r);
          predict_all(r);
          rep_save(r);
          break;
        case 1: nbor(r); rep_save(r); break;
        case 2: predict_all(r); rep_save(r); break;
        case 3: if (r.error == 0) run_events(r, arg, (flags & 1) != 0); rep_save(r); break;
        case 8: if (r.error == 0) trace_run_block(d, rid, arg, flags); break;
        case 4: sync_positions(r); break;
        case 5: { OutRec o; energy_of(r, o); eout[rid] = o; } break;
        case 7: { double tn = ((const double*)ibuf)[rid]; if (tn > 0.0) { retemp(r, tn); rep_save(r); } } break;
        default: throw std::runtime_error("unknown op");
      }
    }
  }
  if (ms) *ms = 0;
  if (launches) *launches = 0;
}
// replica exchange: the same decision code (dmd_exchange.h) on the CPU; more than one rank needs the host's own gather
inline void nccl_unique_id(char*) { throw std::runtime_error("the host-trace library has no NCCL"); }
inline void* nccl_comm_init(const char*, int, int) { throw std::runtime_error("the host-trace library has no NCCL"); }
inline void nccl_comm_geometry(void*, int&, int&) { throw std::runtime_error("the host-trace library has no NCCL"); }
inline void nccl_comm_destroy(void*) {}
inline void exchange(const dmd::DevArrays& d, dmd::OutRec* eout, double* xb, void*, const double* gathered_host, int world,
                     int rank, long long step, unsigned long long seed, int L, dmd::XchCounts* counts_out, double* tstar_out,
                     double* ms, int* launches) {
  using namespace dmd;
  const int R = d.n_replicas, M = world * R, n_ladders = M / L;
  std::vector<double> all(2 * (size_t)M), tnew(M);
  if (gathered_host) {
    std::memcpy(all.data(), gathered_host, sizeof(double) * 2 * (size_t)M);
  } else {
    if (world > 1) throw std::runtime_error("dmdb_exchange: more than one rank needs the gathered (E_pot, T*) array here");
    run_op(d, 5, 0, R, 0, nullptr, eout, nullptr, nullptr);
    for (int r = 0; r < R; r++) {
      all[2 * r] = eout[r].ered - 0.5 * eout[r].sumvel;
      all[2 * r + 1] = d.scal[r].setemp / 12.0;
    }
  }
  XchCounts c;
  c.attempted = c.accepted = c.changed_local = 0;
  c.ladders = n_ladders;
  for (int g = 0; g < M; g++) tnew[g] = all[2 * g + 1];
  for (int l = 0; l < n_ladders; l++) xch_decide_ladder(all.data(), tnew.data(), l, L, world, R, step, seed, c.attempted, c.accepted);
  double* tsel = xb;
  for (int r = 0; r < R; r++) {
    const int g = rank * R + r;
    const bool ch = tnew[g] != all[2 * g + 1];
    tsel[r] = ch ? tnew[g] : 0.0;
    tstar_out[r] = tnew[g];
    c.changed_local += ch ? 1 : 0;
  }
  run_op(d, 7, 0, R, 0, (int32_t*)tsel, nullptr, nullptr, nullptr);
  *counts_out = c;
  if (ms) *ms = 0;
  if (launches) *launches = 0;
}
}  // namespace be

#include "../../parallel_dmd_for_biomolecules_b200/csrc/dmd_capi_impl.h"
