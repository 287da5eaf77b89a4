// dmd_capi_impl.h -- the C ABI of include/dmdb200.h, written once against a tiny backend namespace `be`
// (allocation, copies, and one launcher per device operation).  Included by
//   dmd_cuda.cu                      -> libdmdb200.so (the product: CUDA backend, sm_100a kernels)
//   tests/host_trace/trace_lib.cpp   -> test-only 1-lane CPU trace of the same engine source (never shipped)
// The including file must define namespace be { init, alloc, release, h2d, d2h, fill_i32, run_op, ... }.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dmdb200.h"
#include "dmd_host.h"
#include "dmd_types.h"

namespace dmd {
enum Op { OP_START = 0, OP_NBOR, OP_PREDICT_ALL, OP_RUN, OP_SYNC_POS, OP_ENERGY, OP_EVCODE, OP_RETEMP, OP_RUN_BLOCK, OP_RUN_GRID };
}

static thread_local std::string g_create_error;

struct dmdb_handle {
  dmd::HostModel model;
  dmd::DevArrays d;
  std::vector<void*> allocs;
  std::vector<double> tstar;  // per replica
  std::vector<char> loaded;   // per replica: dmdb_set_state done
  dmd::OutRec* eout = nullptr;   // device: n_replicas
  int32_t* pair_buf = nullptr;   // device scratch for dmdb_get_evcode
  double* temp_buf = nullptr;    // device: n_replicas temperatures (run start, dmdb_apply_temperatures)
  double* stage_sv = nullptr;    // device staging: n_replicas x N x 6 raw sv (dmdb_set_state*, dmdb_get_state_all)
  int32_t* stage_bp = nullptr;   // device staging: n_replicas x N bptnr
  size_t pair_cap = 0;
  std::string err;
  double last_ms = 0;
  int last_launches = 0;
  int service_ctas = -1;  // dmdb_set_service_ctas
  // replica exchange (dmdb_exchange): NCCL communicator of dmdb_comm_init (owned) and device scratch
  void* comm = nullptr;
  int world = 1, rank = 0;
  double* xbuf = nullptr;
  size_t xbuf_doubles = 0;
  bool synced = false;  // dmdb_sync_positions has advanced every replica to true positions and nothing has run since
  std::vector<dmd::RepScalars> sc_cache;  // the per-replica scalars as last downloaded (one D2H serves the error
                                          // check and the tallies of a run)
};

namespace {

template <class T>
T* dalloc(dmdb_handle* h, size_t n) {
  void* p = be::alloc(n * sizeof(T));
  h->allocs.push_back(p);
  return (T*)p;
}

int fail(dmdb_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  else g_create_error = msg;
  return code;
}

#define DMDB_TRY(h, ...)                                   \
  try {                                                    \
    __VA_ARGS__                                            \
  } catch (const std::exception& e) {                      \
    return fail(h, DMDB_ERR_CUDA, e.what());               \
  }

int check_replica(dmdb_handle* h, int replica, bool need_loaded) {
  if (!h) return DMDB_ERR_ARG;
  if (replica < 0 || replica >= h->d.n_replicas) return fail(h, DMDB_ERR_ARG, "replica index out of range");
  if (need_loaded && !h->loaded[replica]) return fail(h, DMDB_ERR_STATE, "dmdb_set_state has not been called for this replica");
  return 0;
}

int all_loaded(dmdb_handle* h) {
  if (!h) return DMDB_ERR_ARG;
  for (char c : h->loaded)
    if (!c) return fail(h, DMDB_ERR_STATE, "dmdb_set_state has not been called for every replica");
  return 0;
}

const char* device_error_text(int e) {
  switch (e) {
    case dmd::DMD_E_NBR_CAP: return "neighbour list capacity exceeded (raise dmdb_params.nbr_capacity)";
    case dmd::DMD_E_CAL_EMPTY: return "event calendar empty";
    case dmd::DMD_E_NEG_TIME: return "negative event time";
    case dmd::DMD_E_GRID: return "bead outside the cell grid";
    case dmd::DMD_E_BAD_INPUT: return "bptnr entry out of range";
  }
  return "unknown device error";
}

// after a device op: surface device-side error words (downloads the per-replica scalars into h->sc_cache)
int check_device_errors(dmdb_handle* h) {
  const int R = h->d.n_replicas;
  std::vector<dmd::RepScalars>& sc = h->sc_cache;
  sc.resize(R);
  be::d2h(sc.data(), h->d.scal, sizeof(dmd::RepScalars) * R);
  for (int r = 0; r < R; r++)
    if (sc[r].error) {
      char buf[256];
      snprintf(buf, sizeof buf, "replica %d: %s (info %d)", r, device_error_text(sc[r].error), sc[r].error_info);
      int code = sc[r].error == dmd::DMD_E_NBR_CAP ? DMDB_ERR_CAPACITY
                 : (sc[r].error == dmd::DMD_E_BAD_INPUT ? DMDB_ERR_ARG : DMDB_ERR_PHYSICS);
      return fail(h, code, buf);
    }
  return 0;
}

}  // namespace

extern "C" {

int dmdb_create(const dmdb_params* p, const dmdb_topology* topo, const dmdb_tables* tab, dmdb_handle** out) {
  if (!p || !topo || !tab || !out) return fail(nullptr, DMDB_ERR_ARG, "null argument");
  if (p->n_replicas < 1) return fail(nullptr, DMDB_ERR_ARG, "n_replicas must be >= 1");
  if (p->engine < 0 || p->engine > 3) return fail(nullptr, DMDB_ERR_ARG, "engine must be 0 (auto), 1 (warp), 2 (block) or 3 (grid)");
  std::string err;
  if (!be::init(p->device, err)) return fail(nullptr, DMDB_ERR_NO_DEVICE, err);
  std::unique_ptr<dmdb_handle> h(new dmdb_handle());
  try {
    dmd::build_model(*p, *topo, *tab, h->model);
  } catch (const std::exception& e) {
    return fail(nullptr, DMDB_ERR_ARG, e.what());
  }
  try {
    const dmd::SysConst& s = h->model.sys;
    const size_t R = (size_t)p->n_replicas, N = (size_t)s.N;
    dmd::DevArrays& d = h->d;
    std::memset(&d, 0, sizeof(d));
    d.n_replicas = (int)R;
    d.cal_stride = s.ngroups * 32;
    d.n_beads = s.N;
    d.cap = s.cap;
    d.ngroups = s.ngroups;
    d.log_cap = s.log_cap;
    d.out_cap = s.out_cap;
    {
      const int ncc0 = (s.ncr + 1) >> 1;
      d.ncc3 = ncc0 * ncc0 * ncc0;
      d.n_chains = s.nch[0] + (s.n_species == 2 ? s.nch[1] : 0);
      d.chainwise = s.chainwise;
    }
    dmd::SysConst* dsys = dalloc<dmd::SysConst>(h.get(), 1);
    be::h2d(dsys, &s, sizeof(s));
    d.sys = dsys;
    dmd::PairTables* dtab = dalloc<dmd::PairTables>(h.get(), 1);
    be::h2d(dtab, &h->model.tab, sizeof(dmd::PairTables));
    d.tables = dtab;
    dmd::HotConst* dhot = dalloc<dmd::HotConst>(h.get(), 1);
    be::h2d(dhot, &h->model.hot, sizeof(dmd::HotConst));
    d.hot = dhot;
    double* dbl = dalloc<double>(h.get(), h->model.bl.size());
    be::h2d(dbl, h->model.bl.data(), h->model.bl.size() * sizeof(double));
    d.bl = dbl;
    d.nres = h->model.hot.nres;
    int32_t* dnc = dalloc<int32_t>(h.get(), h->model.nc_beads.size());
    be::h2d(dnc, h->model.nc_beads.data(), h->model.nc_beads.size() * 4);
    d.nc_beads = dnc;
    d.n_nc = (int)h->model.nc_beads.size();
    uint32_t* dmeta = dalloc<uint32_t>(h.get(), N);
    be::h2d(dmeta, h->model.meta.data(), N * 4);
    d.meta = dmeta;
    int32_t* dchain = dalloc<int32_t>(h.get(), N);
    be::h2d(dchain, h->model.chain.data(), N * 4);
    d.chain = dchain;
    if (!h->model.sctab.empty()) {
      uint8_t* dsct = dalloc<uint8_t>(h.get(), h->model.sctab.size());
      be::h2d(dsct, h->model.sctab.data(), h->model.sctab.size());
      d.sctab = dsct;
    }
    d.rec = dalloc<dmd::BeadRec>(h.get(), R * N);
    d.cal = dalloc<dmd::CalEnt>(h.get(), R * d.cal_stride);
    d.er34 = dalloc<int32_t>(h.get(), R * 2 * N);
    d.up = dalloc<uint32_t>(h.get(), R * N * s.cap);
    d.dn = dalloc<uint32_t>(h.get(), R * N * s.cap);
    d.nup = dalloc<uint16_t>(h.get(), R * N);
    d.ndn = dalloc<uint16_t>(h.get(), R * N);
    d.oldr = dalloc<double>(h.get(), R * 3 * N);
    const size_t ncc = (size_t)((s.ncr + 1) >> 1), nc3 = ncc * ncc * ncc;
    d.cellhead = dalloc<int32_t>(h.get(), R * nc3);
    be::fill_i32(d.cellhead, -1, R * nc3);
    d.cpk = dalloc<uint32_t>(h.get(), R * N);
    d.cnext = dalloc<int32_t>(h.get(), R * N);
    d.cellof = dalloc<int32_t>(h.get(), R * N);
    d.tmin1 = dalloc<double>(h.get(), R * s.ngroups);
    d.scal = dalloc<dmd::RepScalars>(h.get(), R);
    d.log = dalloc<dmd::EventLogRec>(h.get(), R * (size_t)std::max(s.log_cap, 1));
    d.out = dalloc<dmd::OutRec>(h.get(), R * (size_t)s.out_cap);
    d.blkstat = dalloc<long long>(h.get(), R * 16);
    be::zero(d.blkstat, R * 16 * sizeof(long long));
    d.svc_flag = dalloc<int32_t>(h.get(), R);
    be::zero(d.svc_flag, R * 4);
    {
      const size_t qcap = 2 * (size_t)R + 64, words = dmd::SVC_Q_RING + qcap;
      d.svc_ctl = dalloc<unsigned long long>(h.get(), words);
      be::zero(d.svc_ctl, words * 8);
      const unsigned long long cap_word = qcap;
      be::h2d(d.svc_ctl + dmd::SVC_Q_CAP, &cap_word, 8);
    }
    h->eout = dalloc<dmd::OutRec>(h.get(), R);
    h->temp_buf = dalloc<double>(h.get(), R);
    be::zero(d.nup, R * N * 2);
    be::zero(d.ndn, R * N * 2);
    h->tstar.assign(R, p->tstar);
    h->loaded.assign(R, 0);
    std::vector<dmd::RepScalars> sc(R);
    std::memset(sc.data(), 0, sizeof(dmd::RepScalars) * R);
    be::h2d(d.scal, sc.data(), sizeof(dmd::RepScalars) * R);
  } catch (const std::exception& e) {
    for (void* q : h->allocs) be::release(q);
    return fail(nullptr, DMDB_ERR_CUDA, e.what());
  }
  *out = h.release();
  return DMDB_OK;
}

void dmdb_destroy(dmdb_handle* h) {
  if (!h) return;
  for (void* q : h->allocs) be::release(q);
  if (h->pair_buf) be::release(h->pair_buf);
  if (h->xbuf) be::release(h->xbuf);
  if (h->comm) be::nccl_comm_destroy(h->comm);
  delete h;
}

const char* dmdb_last_error(const dmdb_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }
int dmdb_num_beads(const dmdb_handle* h) { return h ? h->model.sys.N : -1; }
int dmdb_num_cells(const dmdb_handle* h) { return h ? h->model.sys.num_cell : -1; }

// Run start of replicas [r0, r1): ONE host->device copy of the raw sv (and bptnr) into staging, then the device
// builds every per-replica array (init_replica in dmd_engine.h).  sv_stride / bp_stride = 0 gives every replica
// the same configuration (different RNG streams).
static void upload_replicas(dmdb_handle* h, int r0, int r1, const double* sv, size_t sv_stride, const int32_t* bptnr,
                            size_t bp_stride) {
  const dmd::SysConst& s = h->model.sys;
  const size_t N = (size_t)s.N, n = (size_t)(r1 - r0), R = (size_t)h->d.n_replicas;
  if (!h->stage_sv) {
    h->stage_sv = dalloc<double>(h, R * N * 6);
    h->stage_bp = dalloc<int32_t>(h, R * N);
  }
  const size_t n_cfg = sv_stride ? n : 1;
  be::h2d(h->stage_sv, sv, n_cfg * N * 6 * sizeof(double));
  if (bptnr) be::h2d(h->stage_bp, bptnr, (bp_stride ? n : 1) * N * sizeof(int32_t));
  be::h2d(h->temp_buf, h->tstar.data(), R * sizeof(double));
  be::run_init(h->d, r0, (int)n, h->stage_sv, sv_stride, bptnr ? h->stage_bp : nullptr, bp_stride, h->temp_buf,
               h->model.params.seed);
  for (int r = r0; r < r1; r++) h->loaded[r] = 1;
  h->synced = false;
}

int dmdb_set_state(dmdb_handle* h, int replica, const double* sv, const int32_t* bptnr) {
  if (!h || !sv) return fail(h, DMDB_ERR_ARG, "null argument");
  if (replica != -1) {
    int rc = check_replica(h, replica, false);
    if (rc) return rc;
  }
  try {
    const int r0 = replica < 0 ? 0 : replica, r1 = replica < 0 ? h->d.n_replicas : replica + 1;
    upload_replicas(h, r0, r1, sv, 0, bptnr, 0);
    be::run_op(h->d, dmd::OP_START, r0, r1 - r0, 0, nullptr, nullptr, &h->last_ms, &h->last_launches);
  } catch (const std::exception& e) {
    return fail(h, DMDB_ERR_ARG, e.what());
  }
  return check_device_errors(h);
}

int dmdb_set_state_all(dmdb_handle* h, const double* sv_all, const int32_t* bptnr_all) {
  if (!h || !sv_all) return fail(h, DMDB_ERR_ARG, "null argument");
  try {
    const size_t N = (size_t)h->model.sys.N;
    upload_replicas(h, 0, h->d.n_replicas, sv_all, 6 * N, bptnr_all, N);
    be::run_op(h->d, dmd::OP_START, 0, h->d.n_replicas, 0, nullptr, nullptr, &h->last_ms, &h->last_launches);
  } catch (const std::exception& e) {
    return fail(h, DMDB_ERR_ARG, e.what());
  }
  return check_device_errors(h);
}

int dmdb_get_state(dmdb_handle* h, int replica, double* sv, int32_t* bptnr, int32_t* identity, int32_t* extra_repuls,
                   double* t, double* tfalse, int64_t* coll);

int dmdb_set_temperature(dmdb_handle* h, int replica, double tstar) {
  if (!h || !(tstar > 0)) return fail(h, DMDB_ERR_ARG, "bad argument");
  const int r0 = replica < 0 ? 0 : replica, r1 = replica < 0 ? h->d.n_replicas : replica + 1;
  if (replica >= 0) {
    int rc = check_replica(h, replica, true);
    if (rc) return rc;
  }
  DMDB_TRY(h, {
    // what a new `./dmd < temp_0xx` run does -- true positions (main.F90:1288-1295) -> restart files -> start-up path
    // (inputinfo.f:76-101, main.F90:205-424) -- with the "files" staying on the device: the records are packed into
    // the staging arrays the upload path uses (sv(6,N) + bptnr, the exact doubles config.f would write) and the run
    // start reads them from there.  No host round trip.
    const size_t N = (size_t)h->model.sys.N, R = (size_t)h->d.n_replicas;
    if (!h->synced) be::run_op(h->d, dmd::OP_SYNC_POS, r0, r1 - r0, 0, nullptr, nullptr, &h->last_ms, &h->last_launches);
    if (!h->stage_sv) {
      h->stage_sv = dalloc<double>(h, R * N * 6);
      h->stage_bp = dalloc<int32_t>(h, R * N);
    }
    be::run_pack(h->d, h->stage_sv, h->stage_bp);
    for (int r = r0; r < r1; r++) h->tstar[r] = tstar;
    be::h2d(h->temp_buf, h->tstar.data(), R * sizeof(double));
    be::run_init(h->d, r0, r1 - r0, h->stage_sv + (size_t)r0 * N * 6, N * 6, h->stage_bp + (size_t)r0 * N, N, h->temp_buf,
                 h->model.params.seed);
    be::run_op(h->d, dmd::OP_START, r0, r1 - r0, 0, nullptr, nullptr, &h->last_ms, &h->last_launches);
    if (r1 - r0 == h->d.n_replicas) h->synced = false;
  })
  return check_device_errors(h);
}

int dmdb_nbor(dmdb_handle* h) {
  int rc = all_loaded(h);
  if (rc) return rc;
  DMDB_TRY(h, be::run_op(h->d, dmd::OP_NBOR, 0, h->d.n_replicas, 0, nullptr, nullptr, &h->last_ms, &h->last_launches);)
  return check_device_errors(h);
}

int dmdb_predict_all(dmdb_handle* h) {
  int rc = all_loaded(h);
  if (rc) return rc;
  DMDB_TRY(h, be::run_op(h->d, dmd::OP_PREDICT_ALL, 0, h->d.n_replicas, 0, nullptr, nullptr, &h->last_ms,
                         &h->last_launches);)
  return check_device_errors(h);
}

int dmdb_get_replica_stats(dmdb_handle* h, int replica, dmdb_stats* s);

int dmdb_set_service_ctas(dmdb_handle* h, int n) {
  if (!h) return DMDB_ERR_ARG;
  if (n < -1 || n > 0xfffe) return fail(h, DMDB_ERR_ARG, "service CTAs: -1 (automatic), 0 (none) or a positive count");
  h->service_ctas = n;
  return DMDB_OK;
}

int dmdb_device_fill(int device, int32_t* n_replicas, int32_t* n_service_ctas) {
  if (!n_replicas) return fail(nullptr, DMDB_ERR_ARG, "null argument");
  std::string err;
  if (!be::init(device, err)) return fail(nullptr, DMDB_ERR_NO_DEVICE, err);
  int workers = 0, service = 0;
  be::device_fill(workers, service);
  *n_replicas = workers;
  if (n_service_ctas) *n_service_ctas = service;
  return DMDB_OK;
}

// dmdb_sync_positions leaves the beads at their true positions with the event clock unchanged (main.F90:1288-1295 is
// the end of a run): the next thing must be a restart
static const char* const SYNCED_MSG = "positions were advanced by dmdb_sync_positions: restart the run first (dmdb_set_state* or "
                                      "dmdb_set_temperature)";

static void stats_from_scalars(const dmdb_handle* h, const std::vector<dmd::RepScalars>& sc, int replica, dmdb_stats* s) {
  const int R = (int)sc.size();
  std::memset(s, 0, sizeof(*s));
  for (int r = (replica < 0 ? 0 : replica); r < (replica < 0 ? R : replica + 1); r++) {
    s->events += sc[r].coll;
    for (int k = 0; k < 32; k++) {
      s->nevents[k] += sc[r].nevents[k];
      s->pair_events += sc[r].nevents[k];
    }
    s->ghosts += sc[r].numghosts;
    s->updates += sc[r].nupdates - sc[r].nforcedupdate;
    s->forced_updates += sc[r].nforcedupdate;
    s->pair_predictions += sc[r].n_pair_pred;
    s->nbr_visits += sc[r].n_nbr_visits;
  }
  s->device_ms = h->last_ms;
  s->kernel_launches = h->last_launches;
}

static int run_impl(dmdb_handle* h, int64_t n_events, dmdb_stats* stats, int flags);
int dmdb_run(dmdb_handle* h, int64_t n_events, dmdb_stats* stats) { return run_impl(h, n_events, stats, 0); }
int dmdb_run_until_output(dmdb_handle* h, int64_t max_events, dmdb_stats* stats) { return run_impl(h, max_events, stats, 1); }

static int run_impl(dmdb_handle* h, int64_t n_events, dmdb_stats* stats, int flags) {
  int rc = all_loaded(h);
  if (rc) return rc;
  if (n_events < 0) return fail(h, DMDB_ERR_ARG, "n_events must be >= 0");
  // engine choice: one warp per replica fills the GPU from ~1000 replicas on; below that the CTA-per-replica
  // engine (batched conservative commit, state in shared memory) is an order of magnitude faster per trajectory
  int engine = h->model.params.engine;
  // a single system of >= 8192 beads gets the whole GPU per round (engine 3): measured 8.0e5 vs 5.1e5 events/s at
  // 12 288 beads, 5.5e6 at 10^6 beads; below that its kernel relaunches at pseudo-events cost more than it gains
  if (engine == 0) engine = h->d.n_replicas >= 1184 ? 1 : ((h->d.n_replicas == 1 && h->model.sys.N >= 8192 && !flags) ? 3 : 2);
  if (engine == 3 && (flags || !be::grid_engine_available())) engine = 2;  // (run_until_output: engines 1 and 2)
  if (engine == 2 && !be::block_engine_fits(h->model.sys)) engine = 1;
  const int op = engine == 3 ? dmd::OP_RUN_GRID : (engine == 2 ? dmd::OP_RUN_BLOCK : dmd::OP_RUN);
  const int svc_bits = op == dmd::OP_RUN ? ((h->service_ctas + 1) & 0xffff) << 8 : 0;  // run_op: bits 8-23 = 1 + service CTAs (0 = automatic)
  if (h->synced) return fail(h, DMDB_ERR_STATE, SYNCED_MSG);
  DMDB_TRY(h, be::run_op(h->d, op, 0, h->d.n_replicas, n_events, nullptr, nullptr, &h->last_ms, &h->last_launches, flags | svc_bits);)
  rc = check_device_errors(h);
  if (rc) return rc;
  if (stats) stats_from_scalars(h, h->sc_cache, -1, stats);  // the scalars just downloaded: no second copy
  return DMDB_OK;
}

int dmdb_sync_positions(dmdb_handle* h) {
  int rc = all_loaded(h);
  if (rc) return rc;
  if (h->synced) return DMDB_OK;  // already at true positions: advancing twice would move the beads again
  DMDB_TRY(h, be::run_op(h->d, dmd::OP_SYNC_POS, 0, h->d.n_replicas, 0, nullptr, nullptr, &h->last_ms,
                         &h->last_launches);)
  h->synced = true;
  return DMDB_OK;
}

int dmdb_get_cells(dmdb_handle* h, int replica, int32_t* cell_of_bead) {
  int rc = check_replica(h, replica, true);
  if (rc) return rc;
  const size_t N = (size_t)h->model.sys.N;
  DMDB_TRY(h, be::d2h(cell_of_bead, h->d.cellof + (size_t)replica * N, N * 4);)
  return DMDB_OK;
}

int dmdb_get_nbors(dmdb_handle* h, int replica, int down, int32_t* offsets, int32_t* nb) {
  int rc = check_replica(h, replica, true);
  if (rc) return rc;
  const dmd::SysConst& s = h->model.sys;
  const size_t N = (size_t)s.N, cap = (size_t)s.cap;
  DMDB_TRY(h, {
    std::vector<uint16_t> cnt(N);
    be::d2h(cnt.data(), (down ? h->d.ndn : h->d.nup) + (size_t)replica * N, N * 2);
    std::vector<uint32_t> lst;
    if (nb) {
      lst.resize(N * cap);
      be::d2h(lst.data(), (down ? h->d.dn : h->d.up) + (size_t)replica * N * cap, N * cap * 4);
    }
    int n = 0;
    for (size_t i = 0; i < N; i++) {
      offsets[i] = n;
      if (nb) {
        for (int k = 0; k < cnt[i]; k++) nb[n + k] = (int32_t)(lst[i * cap + k] & dmd::NB_MASK) + 1;
        std::sort(nb + n, nb + n + cnt[i]);
      }
      n += cnt[i];
    }
    offsets[N] = n;
  })
  return DMDB_OK;
}

int dmdb_get_calendar(dmdb_handle* h, int replica, double* tim, int32_t* nptnr, int32_t* coltype) {
  int rc = check_replica(h, replica, true);
  if (rc) return rc;
  const size_t N = (size_t)h->model.sys.N;
  DMDB_TRY(h, {
    std::vector<dmd::CalEnt> cal(N + 3);
    be::d2h(cal.data(), h->d.cal + (size_t)replica * h->d.cal_stride, (N + 3) * sizeof(dmd::CalEnt));
    for (size_t k = 0; k < N + 3; k++) {
      tim[k] = cal[k].t;
      coltype[k] = (int8_t)(cal[k].type & 0xff);
      nptnr[k] = cal[k].ptnr >= 0 ? cal[k].ptnr + 1 : cal[k].ptnr;  // 1-based; -1 none, -2 pseudo (main.F90:231-234)
    }
  })
  return DMDB_OK;
}

int dmdb_get_state(dmdb_handle* h, int replica, double* sv, int32_t* bptnr, int32_t* identity, int32_t* extra_repuls,
                   double* t, double* tfalse, int64_t* coll) {
  int rc = check_replica(h, replica, true);
  if (rc) return rc;
  const size_t N = (size_t)h->model.sys.N;
  DMDB_TRY(h, {
    std::vector<dmd::BeadRec> rec(N);
    be::d2h(rec.data(), h->d.rec + (size_t)replica * N, N * sizeof(dmd::BeadRec));
    std::vector<int32_t> er34;
    if (extra_repuls) {
      er34.resize(2 * N);
      be::d2h(er34.data(), h->d.er34 + (size_t)replica * 2 * N, 2 * N * 4);
    }
    for (size_t k = 0; k < N; k++) {
      const dmd::BeadRec& b = rec[k];
      if (sv) {
        double* o = sv + 6 * k;
        o[0] = b.x; o[1] = b.y; o[2] = b.z; o[3] = b.vx; o[4] = b.vy; o[5] = b.vz;
      }
      if (bptnr) bptnr[k] = b.bptnr + 1;
      if (identity) identity[k] = b.ident;
      if (extra_repuls) {
        extra_repuls[0 * N + k] = b.er1 + 1;
        extra_repuls[1 * N + k] = b.er2 + 1;
        extra_repuls[2 * N + k] = er34[2 * k] + 1;
        extra_repuls[3 * N + k] = er34[2 * k + 1] + 1;
      }
    }
    if (t || tfalse || coll) {
      dmd::RepScalars sc;
      be::d2h(&sc, h->d.scal + replica, sizeof(sc));
      if (t) *t = sc.t;
      if (tfalse) *tfalse = sc.tfalse;
      if (coll) *coll = sc.coll;
    }
  })
  return DMDB_OK;
}

int dmdb_get_state_all(dmdb_handle* h, double* sv_all, int32_t* bptnr_all) {
  int rc = all_loaded(h);
  if (rc) return rc;
  const size_t N = (size_t)h->model.sys.N, R = (size_t)h->d.n_replicas;
  DMDB_TRY(h, {
    if (!h->stage_sv) {
      h->stage_sv = dalloc<double>(h, R * N * 6);
      h->stage_bp = dalloc<int32_t>(h, R * N);
    }
    be::run_pack(h->d, h->stage_sv, h->stage_bp);  // records -> sv(6,N) + bptnr on the device
    if (sv_all) be::d2h(sv_all, h->stage_sv, R * N * 6 * sizeof(double));
    if (bptnr_all) be::d2h(bptnr_all, h->stage_bp, R * N * sizeof(int32_t));
  })
  return DMDB_OK;
}

int dmdb_apply_temperatures(dmdb_handle* h, const double* tstar_new) {
  int rc = all_loaded(h);
  if (rc) return rc;
  if (!tstar_new) return fail(h, DMDB_ERR_ARG, "null argument");
  if (h->synced) return fail(h, DMDB_ERR_STATE, SYNCED_MSG);
  const int R = h->d.n_replicas;
  DMDB_TRY(h, {
    std::vector<double> tn(R);
    bool any = false;
    for (int r = 0; r < R; r++) {
      tn[r] = (tstar_new[r] > 0 && tstar_new[r] != h->tstar[r]) ? tstar_new[r] : 0.0;
      any = any || tn[r] > 0;
    }
    if (any) {
      be::h2d(h->temp_buf, tn.data(), sizeof(double) * R);
      be::run_op(h->d, dmd::OP_RETEMP, 0, R, 0, (int32_t*)h->temp_buf, nullptr, &h->last_ms, &h->last_launches);
      for (int r = 0; r < R; r++)
        if (tn[r] > 0) h->tstar[r] = tn[r];
    }
  })
  return check_device_errors(h);
}

// ---- replica exchange (new functionality)
int dmdb_nccl_unique_id(char id[128]) {
  if (!id) return fail(nullptr, DMDB_ERR_ARG, "null argument");
  try {
    be::nccl_unique_id(id);
  } catch (const std::exception& e) {
    return fail(nullptr, DMDB_ERR_CUDA, e.what());
  }
  return DMDB_OK;
}

int dmdb_comm_init(dmdb_handle* h, const char id[128], int world, int rank) {
  if (!h || !id) return DMDB_ERR_ARG;
  if (world < 1 || rank < 0 || rank >= world) return fail(h, DMDB_ERR_ARG, "dmdb_comm_init: need 0 <= rank < world");
  DMDB_TRY(h, {
    if (h->comm) be::nccl_comm_destroy(h->comm);
    h->comm = nullptr;
    if (world > 1) h->comm = be::nccl_comm_init(id, world, rank);
    h->world = world;
    h->rank = rank;
  })
  return DMDB_OK;
}

static int exchange_impl(dmdb_handle* h, void* comm, const double* gathered, int world, int rank, int64_t step, uint64_t seed,
                         int32_t ladder_size, dmdb_exchange_stats* out) {
  int rc = all_loaded(h);
  if (rc) return rc;
  if (h->synced) return fail(h, DMDB_ERR_STATE, SYNCED_MSG);
  const int R = h->d.n_replicas;
  if (world < 1 || rank < 0 || rank >= world) return fail(h, DMDB_ERR_ARG, "exchange: need 0 <= rank < world");
  const long long M = (long long)world * R;
  if (ladder_size <= 0) ladder_size = (int32_t)(M < dmd::XCH_MAX_LADDER ? M : dmd::XCH_MAX_LADDER);
  if (ladder_size < 2 || ladder_size > dmd::XCH_MAX_LADDER) return fail(h, DMDB_ERR_ARG, "exchange: ladder size must be 2..32");
  DMDB_TRY(h, {
    const size_t need = 2 * (size_t)R + 3 * (size_t)M + 2 * (size_t)R + 8;
    if (h->xbuf_doubles < need) {
      if (h->xbuf) be::release(h->xbuf);
      h->xbuf = (double*)be::alloc(need * sizeof(double));
      h->xbuf_doubles = need;
    }
    dmd::XchCounts c;
    std::memset(&c, 0, sizeof(c));
    int launches = 0;
    double ms = 0;
    be::exchange(h->d, h->eout, h->xbuf, comm, gathered, world, rank, (long long)step, (unsigned long long)seed, ladder_size, &c,
                 h->tstar.data(), &ms, &launches);
    h->last_ms = ms;
    h->last_launches = launches;
    if (out) {
      out->ladders = c.ladders; out->attempted = c.attempted; out->accepted = c.accepted; out->changed_local = c.changed_local;
      out->device_ms = ms; out->kernel_launches = launches; out->reserved = 0;
    }
  })
  return check_device_errors(h);
}

int dmdb_exchange(dmdb_handle* h, void* nccl_comm, int64_t step, uint64_t seed, int32_t ladder_size, dmdb_exchange_stats* out) {
  if (!h) return DMDB_ERR_ARG;
  int world = h->world, rank = h->rank;
  void* comm = h->comm;
  if (nccl_comm) {
    try {
      be::nccl_comm_geometry(nccl_comm, world, rank);
    } catch (const std::exception& e) {
      return fail(h, DMDB_ERR_CUDA, e.what());
    }
    comm = nccl_comm;
  }
  return exchange_impl(h, comm, nullptr, world, rank, step, seed, ladder_size, out);
}

int dmdb_exchange_gathered(dmdb_handle* h, const double* gathered, int world, int rank, int64_t step, uint64_t seed,
                           int32_t ladder_size, dmdb_exchange_stats* out) {
  if (!h || !gathered) return DMDB_ERR_ARG;
  return exchange_impl(h, nullptr, gathered, world, rank, step, seed, ladder_size, out);
}

int dmdb_get_evcode(dmdb_handle* h, int replica, int n_pairs, const int32_t* i, const int32_t* j, int32_t* code) {
  int rc = check_replica(h, replica, true);
  if (rc) return rc;
  if (n_pairs <= 0) return DMDB_OK;
  const int N = h->model.sys.N;
  for (int k = 0; k < n_pairs; k++)
    if (i[k] < 1 || i[k] > N || j[k] < 1 || j[k] > N || i[k] == j[k]) return fail(h, DMDB_ERR_ARG, "bad bead pair");
  DMDB_TRY(h, {
    if ((size_t)n_pairs > h->pair_cap) {
      if (h->pair_buf) be::release(h->pair_buf);
      h->pair_buf = (int32_t*)be::alloc((size_t)n_pairs * 3 * 4);
      h->pair_cap = (size_t)n_pairs;
    }
    be::h2d(h->pair_buf, i, (size_t)n_pairs * 4);
    be::h2d(h->pair_buf + n_pairs, j, (size_t)n_pairs * 4);
    be::run_op(h->d, dmd::OP_EVCODE, replica, 1, n_pairs, h->pair_buf, nullptr, &h->last_ms, &h->last_launches);
    be::d2h(code, h->pair_buf + 2 * (size_t)n_pairs, (size_t)n_pairs * 4);
  })
  return DMDB_OK;
}

int dmdb_energy_of(dmdb_handle* h, int replica, dmdb_energy* e) {
  int rc = check_replica(h, replica, true);
  if (rc) return rc;
  DMDB_TRY(h, {
    be::run_op(h->d, dmd::OP_ENERGY, replica, 1, 0, nullptr, h->eout, &h->last_ms, &h->last_launches);
    dmd::OutRec o;
    be::d2h(&o, h->eout + replica, sizeof(o));
    e->ered = o.ered; e->tred = o.tred; e->sumvel = o.sumvel; e->ehh_ii = o.ehh_ii; e->ehh_ij = o.ehh_ij;
    e->hb_alpha = o.hb_alpha; e->hb_ii = o.hb_ii; e->hb_ij = o.hb_ij; e->reserved = 0;
  })
  return DMDB_OK;
}

int dmdb_sheet_observables(dmdb_handle* h, int32_t* out) {
  int rc = all_loaded(h);
  if (rc) return rc;
  if (!out) return fail(h, DMDB_ERR_ARG, "null argument");
  const dmd::SysConst& s = h->model.sys;
  if (s.n_species == 2 && s.nch[1] > 0 && (s.numbeads[1] != s.numbeads[0] || s.chnln[1] != s.chnln[0]))
    return fail(h, DMDB_ERR_ARG, "sheet observables: fibril_list_assign.f handles one peptide species (equal chains)");
  if (s.N / s.numbeads[0] > be::sheet_max_chains()) return fail(h, DMDB_ERR_CAPACITY, "sheet observables: too many chains");
  const size_t R = (size_t)h->d.n_replicas;
  DMDB_TRY(h, {
    int32_t* dev = (int32_t*)be::alloc(R * 8 * sizeof(int32_t));
    try {
      be::run_sheets(h->d, dev);
      be::d2h(out, dev, R * 8 * sizeof(int32_t));
    } catch (...) {
      be::release(dev);
      throw;
    }
    be::release(dev);
  })
  return DMDB_OK;
}

int dmdb_potential_energies(dmdb_handle* h, double* epot, double* tstar) {
  int rc = all_loaded(h);
  if (rc) return rc;
  const int R = h->d.n_replicas;
  DMDB_TRY(h, {
    be::run_op(h->d, dmd::OP_ENERGY, 0, R, 0, nullptr, h->eout, &h->last_ms, &h->last_launches);
    std::vector<dmd::OutRec> o(R);
    be::d2h(o.data(), h->eout, sizeof(dmd::OutRec) * R);
    for (int r = 0; r < R; r++) {
      if (epot) epot[r] = o[r].ered - 0.5 * o[r].sumvel;  // e_int of main.F90:358
      if (tstar) tstar[r] = h->tstar[r];
    }
  })
  return DMDB_OK;
}

int dmdb_get_event_log(dmdb_handle* h, int replica, int64_t first, int64_t n, dmdb_event* out, int64_t* n_out) {
  int rc = check_replica(h, replica, true);
  if (rc) return rc;
  static_assert(sizeof(dmdb_event) == sizeof(dmd::EventLogRec), "event record layout");
  if (first < 0 || n < 0 || (n > 0 && !out)) return fail(h, DMDB_ERR_ARG, "event log: first and n must be >= 0");
  DMDB_TRY(h, {
    dmd::RepScalars sc;
    be::d2h(&sc, h->d.scal + replica, sizeof(sc));
    int64_t avail = sc.n_log, m = 0;
    if (first < avail) {
      m = std::min<int64_t>(n, avail - first);
      be::d2h(out, h->d.log + (size_t)replica * std::max(h->model.sys.log_cap, 1) + first, (size_t)m * sizeof(dmdb_event));
    }
    if (n_out) *n_out = m;
  })
  return DMDB_OK;
}

int dmdb_get_batch_stats(dmdb_handle* h, int replica, int64_t out[16]) {
  if (!h || !out) return DMDB_ERR_ARG;
  if (replica >= 0) {
    int rc = check_replica(h, replica, false);
    if (rc) return rc;
  }
  DMDB_TRY(h, {
    const int R = h->d.n_replicas;
    std::vector<long long> st((size_t)R * 16);
    be::d2h(st.data(), h->d.blkstat, st.size() * sizeof(long long));
    for (int k = 0; k < 16; k++) out[k] = 0;
    for (int r = (replica < 0 ? 0 : replica); r < (replica < 0 ? R : replica + 1); r++)
      for (int k = 0; k < 16; k++) out[k] += st[(size_t)r * 16 + k];
  })
  return DMDB_OK;
}

int dmdb_get_replica_stats(dmdb_handle* h, int replica, dmdb_stats* s) {
  if (!h || !s) return DMDB_ERR_ARG;
  if (replica >= 0) {
    int rc = check_replica(h, replica, false);
    if (rc) return rc;
  }
  DMDB_TRY(h, {
    const int R = h->d.n_replicas;
    h->sc_cache.resize(R);
    be::d2h(h->sc_cache.data(), h->d.scal, sizeof(dmd::RepScalars) * R);
    stats_from_scalars(h, h->sc_cache, replica, s);
  })
  return DMDB_OK;
}

}  // extern "C"
