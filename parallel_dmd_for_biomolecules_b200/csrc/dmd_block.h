// dmd_block.h -- the CTA-per-replica engine: batched conservative commit of independent events.
//
// One CTA (BK_MAXW warps) owns one replica whose bead records, calendar and auxiliary lists are resident in
// SHARED MEMORY (N = 1344: 119 KB).  This is the intra-trajectory parallelism of BASELINE.json's north_star item
// (4): the reference ships the events of the first calendar bucket to MPI workers speculatively and abandons the
// rest of the batch when an applied event invalidates it (main.F90:563-636, 955-993).  Here a round is
//
//   scan     block arg-min over the calendar (replaces add_tbin.f / del_tbin.f + main.F90:496-545)
//   select   every entry within `window` of the minimum becomes a candidate; candidates are ranked by
//            (time, bead index) -- the serial processing order
//   claim    candidate q (one warp each) stamps its rank on every bead of its FOOTPRINT
//            {i, j} + up/down neighbours + auxiliary partners of i and j  with atomicMin
//   check    a candidate holding all its stamps is independent of every EARLIER candidate; the batch is the
//            longest prefix of independent hard-core / bond events
//   exec     the batch is executed in parallel, one warp per event, directly on the shared state, logging
//            every calendar entry it overwrites (Undo) and the earliest event time it creates
//   commit   event q stays valid iff no event created by events 0..q-1 is earlier than t_q (the reference's
//            own rule, main.F90:970-993); the first invalid event and everything after it is rolled back
//
// EXACTNESS.  A committed batch is exactly what the serial loop does: events of a batch have pairwise disjoint
// footprints, so each reads precisely the state it would read in time order (the only writes between its serial
// position and the round start come from earlier batch members, which touch none of its beads); cascaded
// re-predictions read beads two hops away, but a bead whose STATE is written (i', j' of another event) has all
// its neighbours inside that event's footprint, so a cascade on l (inside mine) reading i' would make l a shared
// bead -- excluded by the claims.  Newly created events earlier than a later batch member make the serial order
// differ, hence the commit rule.  Anything that is not a plain type 1/2/3 pair event (H-bond events, ghost,
// interval, output) is executed alone at the head of the calendar by the serial code of dmd_engine.h.
// The committed sequence, every time and every state bit equal the warp engine's and the oracle's.
#pragma once
#include "dmd_engine.h"

namespace dmd {

constexpr int BK_MAXW = 16;   // warps per CTA = events executed per round
constexpr int BK_CAND = 64;   // candidate capacity per round
constexpr int BK_UNDO = 96;   // calendar entries one speculative event may overwrite
constexpr int FP_SIDE = 32;   // neighbour-list entries staged per list when an event's footprint is claimed
constexpr uint32_t CLAIM_FREE = 0xffffffffu;

#if defined(DMD_HOST_TRACE)
// test scaffolding: virtual warps are host threads meeting at a pthread barrier (tests/host_trace/trace_lib.cpp)
void blk_sync();
inline uint32_t blk_atomic_min(uint32_t* p, uint32_t v) {
  uint32_t o = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
  }
  return o;
}
inline int blk_atomic_add(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline void blk_atomic_add64(long long* p, long long v) { __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
#else
__device__ __forceinline__ void blk_sync() { __syncthreads(); }
__device__ __forceinline__ uint32_t blk_atomic_min(uint32_t* p, uint32_t v) { return atomicMin(p, v); }
__device__ __forceinline__ int blk_atomic_add(int* p, int v) { return atomicAdd(p, v); }
__device__ __forceinline__ void blk_atomic_add64(long long* p, long long v) {
  atomicAdd((unsigned long long*)p, (unsigned long long)v);
}
#endif

struct BlkSlot {  // one event of the batch
  double t;       // its time (= tfalse while it executes)
  double newmin;  // earliest event time it created
  int owner, j;
  int win;        // holds all its claims
  int ct, code;   // executed type and ev_code, for the log
  int n_undo;
};

struct BlkShared {
  double wmin_t[BK_MAXW];
  double cand_t[BK_CAND];
  double moved_far[BK_MAXW];
  double window;
  long long coll, target;
  int stop_at_output, pad2;
  long long st_rounds, st_exec, st_rollback, st_conflict, st_cold;  // statistics of the batching
  long long n_pair_pred, n_nbr_visits;
  int wmin_i[BK_MAXW];
  int cand_o[BK_CAND];
  int cand_hot[BK_CAND];  // plain type 1/2/3 pair event (recorded while the calendar is stable)
  int rank[BK_CAND];      // position in the serial processing order
  double tlast;           // time of the last committed event (< 0: unknown, e.g. after an interval event)
  long long nevents[32];  // main.F90:926 tallies of the committed batch events (flushed at the end of the run)
  int n_cand;
  int error, error_info;
  int pad;
  BlkSlot slot[BK_MAXW];
  BeadRec old_i[BK_MAXW], old_j[BK_MAXW];
  CalEnt undo_old[BK_MAXW][BK_UNDO];
  int32_t undo_idx[BK_MAXW][BK_UNDO];
  uint32_t fp[BK_MAXW][4][FP_SIDE];  // staged lists of the slot's event: up(i), dn(i), up(j), dn(j)
  long long cyc[8];                  // clock64 per phase (thread 0): scan, select+rank, claim, check, exec, commit, serial
                                     // head events, of which list rebuilds
};

#if defined(DMD_HOST_TRACE)
inline long long blk_clock() { return 0; }
#else
__device__ __forceinline__ long long blk_clock() { return clock64(); }
#endif

DMD_DEV int warp_min_i(int v) {
#if DMD_W > 1
  v = (int)__reduce_min_sync(0xffffffffu, (unsigned)v);  // callers pass non-negative values
#endif
  return v;
}

DMD_DEV bool blk_is_hot(const Rep& r, int o) {  // plain hard-core / bond event of a bead
  if (o >= r.N) return false;
  const CalEnt e = r.cal[o];
  const int ct = type_of(e.type);
  return e.ptnr >= 0 && ct >= 1 && ct <= 3;
}

// lane-strided walk over the footprint of the event of bead i with partner j; f(b) for every bead in it
template <class F>
DMD_DEV void blk_footprint(const Rep& r, int i, int j, const ListRef& li, const ListRef& lj, F f) {
  for (int side = 0; side < 2; side++) {
    const int a = side ? j : i;
    const ListRef& l = side ? lj : li;
    const BeadRec* ra = &r.rec[a];
    const int total = 1 + l.nu + l.nd + 3;
    for (int p = Warp::lane(); p < total; p += DMD_W) {
      int b;
      if (p == 0) b = a;
      else if (p <= l.nu) b = (int)(l.up[p - 1] & NB_MASK);
      else if (p <= l.nu + l.nd) b = (int)(l.dn[p - 1 - l.nu] & NB_MASK);
      else {
        const int k = p - 1 - l.nu - l.nd;
        b = k == 0 ? ra->er1 : (k == 1 ? ra->er2 : r.er34[2 * a]);
      }
      if (b >= 0) f(b);
    }
  }
}

// the lists of the event's two beads: staged into the slot's shared-memory buffers when they fit (the passes of
// the event then never go to global memory for them), else left where they are.  All eight loads of a lane (four
// list entries, four lengths) are independent and issued together.
DMD_DEV void blk_stage_lists(const Rep& r, int i, int j, uint32_t (*fp)[FP_SIDE], ListRef& li, ListRef& lj) {
  li.nu = r.nup[i]; li.nd = r.ndn[i];
  lj.nu = r.nup[j]; lj.nd = r.ndn[j];
  const size_t bi = (size_t)i * r.cap, bj = (size_t)j * r.cap;
  const bool fit_i = li.nu <= FP_SIDE && li.nd <= FP_SIDE, fit_j = lj.nu <= FP_SIDE && lj.nd <= FP_SIDE;
  for (int p = Warp::lane(); p < FP_SIDE; p += DMD_W) {
    const bool in = p < r.cap;
    const uint32_t a0 = in ? r.up[bi + p] : 0u, a1 = in ? r.dn[bi + p] : 0u;
    const uint32_t a2 = in ? r.up[bj + p] : 0u, a3 = in ? r.dn[bj + p] : 0u;
    fp[0][p] = a0; fp[1][p] = a1; fp[2][p] = a2; fp[3][p] = a3;
  }
  li.up = fit_i ? fp[0] : r.up + bi;
  li.dn = fit_i ? fp[1] : r.dn + bi;
  lj.up = fit_j ? fp[2] : r.up + bj;
  lj.dn = fit_j ? fp[3] : r.dn + bj;
}

// ---- scan + select in one sweep: per-warp arg-min over a block-strided slice of the calendar, and every entry
// with t <= lim becomes a candidate (lim < 0: no selection, the minimum is not known yet)
DMD_DEV void blk_phase_scan(BlkShared& S, const Rep& r, int w, int nw, double lim, bool select) {
  double best = T_PAD;
  int bi = 0x7fffffff;
  for (int k = w * DMD_W + Warp::lane(); k < r.N + 3; k += nw * DMD_W) {
    const CalEnt e = r.cal[k];
    if (e.t < best) {  // ascending k: the first minimum keeps the lowest index
      best = e.t;
      bi = k;
    }
    if (select && e.t <= lim) {
      const int pos = blk_atomic_add(&S.n_cand, 1);
      if (pos < BK_CAND) {
        const int ct = type_of(e.type);
        S.cand_t[pos] = e.t;
        S.cand_o[pos] = k;
        S.cand_hot[pos] = (k < r.N && e.ptnr >= 0 && ct >= 1 && ct <= 3) ? 1 : 0;
        S.rank[pos] = 0;
      }
    }
  }
  warp_argmin(best, bi);
  if (Warp::lane() == 0) {
    S.wmin_t[w] = best;
    S.wmin_i[w] = bi;
  }
}

DMD_DEV double blk_tmin(const BlkShared& S, int nw) {  // redundant per warp
  double t = T_PAD;
  for (int q = Warp::lane(); q < nw; q += DMD_W)
    if (S.wmin_t[q] < t) t = S.wmin_t[q];
  return warp_min(t);
}

// ---- rank by (time, owner index) = the serial processing order (oracle decision D3): all pairs in parallel
DMD_DEV void blk_phase_rank(BlkShared& S, int w, int nw, int nc) {
  for (int q = w * DMD_W + Warp::lane(); q < nc * nc; q += nw * DMD_W) {
    const int k = q / nc, m = q - k * nc;
    const double t = S.cand_t[k], tm = S.cand_t[m];
    if (tm < t || (tm == t && S.cand_o[m] < S.cand_o[k])) blk_atomic_add(&S.rank[k], 1);
  }
}

// the candidate with rank q (warp-uniform result; every rank 0..nc-1 occurs exactly once)
DMD_DEV int blk_cand_of_rank(const BlkShared& S, int nc, int q) {
  int found = 0;
  for (int k = Warp::lane(); k < nc; k += DMD_W)
    if (S.rank[k] == q) found = k + 1;
#if DMD_W > 1
  found = (int)__reduce_max_sync(0xffffffffu, (unsigned)found);
#endif
  return found - 1;
}

// ---- plan (computed redundantly by every warp from the round's immutable candidate arrays -- NOT from the
// calendar, which warp 0 may already be changing on the serial path): length of the leading run of hot events
DMD_DEV int blk_plan(const BlkShared& S, int nc, int maxb) {
  int first_bad = maxb;
  for (int k = Warp::lane(); k < nc; k += DMD_W)
    if (!S.cand_hot[k] && S.rank[k] < first_bad) first_bad = S.rank[k];
  return warp_min_i(first_bad);
}

// ---- claim: stage the lists, stamp the rank on the footprint
DMD_DEV void blk_phase_claim(BlkShared& S, const Rep& r, uint32_t* claim, int w, int nc, ListRef& li, ListRef& lj) {
  const int k = blk_cand_of_rank(S, nc, w);
  const int i = S.cand_o[k];
  const int j = r.cal[i].ptnr;
  blk_stage_lists(r, i, j, S.fp[w], li, lj);
  Warp::sync();
  blk_footprint(r, i, j, li, lj, [&](int b) { blk_atomic_min(&claim[b], (uint32_t)w); });
  if (Warp::lane() == 0) {
    BlkSlot& sl = S.slot[w];
    sl.t = S.cand_t[k];
    sl.owner = i;
    sl.j = j;
    // upper bound of the calendar entries the event can overwrite: i, j, every down entry once as a lowered
    // neighbour and once as a cascade.  An event too large for the undo log only runs at the head.
    const int bound = 2 + 2 * (li.nd + 3) + 2 * (lj.nd + 3);
    sl.win = bound <= BK_UNDO || w == 0;
  }
}

// ---- check: do I hold every stamp?
template <class SH>
DMD_DEV void blk_phase_check(SH& S, const Rep& r, const uint32_t* claim, int w, const ListRef& li, const ListRef& lj) {
  const BlkSlot& sl = S.slot[w];
  int bad = 0;
  blk_footprint(r, sl.owner, sl.j, li, lj, [&](int b) {
    if (claim[b] != (uint32_t)w) bad = 1;
  });
  const bool any_bad = Warp::any(bad != 0);
  if (Warp::lane() == 0 && any_bad) S.slot[w].win = 0;
}

DMD_DEV int blk_n_exec(const BlkShared& S, int batch) {  // redundant per warp: first slot that lost a claim
  int first_lost = batch;
  for (int q = Warp::lane(); q < batch; q += DMD_W)
    if (!S.slot[q].win && q < first_lost) first_lost = q;
  return warp_min_i(first_lost);  // >= 1: rank 0 holds the minimum stamp everywhere
}

// ---- exec: one hot pair event (main.F90:1636 eventdyn + :943 partial_events) with undo logging
template <class SH>
DMD_DEV void blk_exec_event(SH& S, Rep& r, int w, const ListRef& li, const ListRef& lj) {
  BlkSlot& sl = S.slot[w];
  const int i = sl.owner;
  const CalEnt ev = r.cal[i];
  const int j = ev.ptnr;
  int ct = type_of(ev.type);
  r.tfalse = sl.t;
  BeadRec ri = r.rec[i], rj = r.rec[j];
  if (Warp::lane() == 0) {
    S.old_i[w] = ri;
    S.old_j[w] = rj;
  }
  const int code = overlay_code(sc_of(ev.type), i, ri, j, rj);  // ev_code(i,j), main.F90:587
  ct = event_dynamics_hot(r.c, ct, code, ri, rj, r.c.meta[i], ri.bptnr == j, r.tfalse);
  Warp::sync();
  if (Warp::lane() == 0) {
    BeadRec* pi = &r.rec[i];
    BeadRec* pj = &r.rec[j];
    pi->x = ri.x; pi->y = ri.y; pi->z = ri.z; pi->vx = ri.vx; pi->vy = ri.vy; pi->vz = ri.vz;
    pj->x = rj.x; pj->y = rj.y; pj->z = rj.z; pj->vx = rj.vx; pj->vy = rj.vy; pj->vz = rj.vz;
  }
  Undo u;
  u.idx = S.undo_idx[w];
  u.old = S.undo_old[w];
  u.n = 0;
  u.cap = w == 0 ? 0 : BK_UNDO;  // the head event is never rolled back
  u.newmin = T_PAD;
  partial_events_t<true>(r, i, j, false, &u, &li, &lj);
  const double nm = warp_min(u.newmin);
  if (w > 0 && u.n > BK_UNDO) set_error(r, DMD_E_NBR_CAP, u.n);
  if (Warp::lane() == 0) {
    sl.newmin = nm;
    sl.ct = ct;
    sl.code = code;
    sl.n_undo = u.n;
  }
}

template <class SH>
DMD_DEV void blk_rollback(SH& S, Rep& r, int w) {
  if (Warp::lane() == 0) {
    const BlkSlot& sl = S.slot[w];
    for (int q = sl.n_undo - 1; q >= 0; q--) r.cal[S.undo_idx[w][q]] = S.undo_old[w][q];
    r.rec[sl.owner] = S.old_i[w];
    r.rec[sl.j] = S.old_j[w];
  }
}

// reference's validation rule (main.F90:970-993): event q is kept iff every event created by 0..q-1 is later.
// Redundant per warp; lane q holds slot q (n_exec <= BK_MAXW <= 32).
DMD_DEV int blk_n_valid(const BlkShared& S, int n_exec) {
#if DMD_W > 1
  const int q = Warp::lane();
  const double t = q < n_exec ? S.slot[q].t : T_PAD;
  double pm = q < n_exec ? S.slot[q].newmin : T_PAD;  // -> inclusive prefix minimum of newmin
#pragma unroll
  for (int d = 1; d < BK_MAXW; d <<= 1) {
    const double o = Warp::shfl(pm, q >= d ? q - d : q);
    if (q >= d && o < pm) pm = o;
  }
  const double before = Warp::shfl(pm, q > 0 ? q - 1 : 0);  // minimum over slots 0..q-1
  const bool bad = q > 0 && q < n_exec && !(t < before);
  const unsigned m = Warp::ballot(bad);
  return m ? dmd_ffs(m) - 1 : n_exec;
#else
  double rm = S.slot[0].newmin;
  int v = 1;
  while (v < n_exec && S.slot[v].t < rm) {
    if (S.slot[v].newmin < rm) rm = S.slot[v].newmin;
    v++;
  }
  return v;
#endif
}

// warp 0: account for the committed events in serial order (main.F90:639, 926) and log them; lane q takes slot q
DMD_DEV void blk_commit(BlkShared& S, Rep& r, int n_valid) {
  const int log_cap = r.c.sys->log_cap;
  for (int q = Warp::lane(); q < n_valid; q += DMD_W) {
    const BlkSlot& sl = S.slot[q];
    if (sl.ct >= 0 && sl.ct < 32) blk_atomic_add64(&S.nevents[sl.ct], 1);
    if (r.n_log + q < log_cap) {
      EventLogRec e;
      e.t = r.t + sl.t;
      e.i = sl.owner + 1;
      e.j = sl.j + 1;
      e.type = sl.ct;
      e.evcode = sl.code;
      r.log[r.n_log + q] = e;
    }
  }
  if (r.n_log < log_cap) r.n_log = r.n_log + n_valid < log_cap ? r.n_log + n_valid : log_cap;  // like log_event()
  r.coll += n_valid;
  r.tfalse = S.slot[n_valid - 1].t;
  r.old_tfalse = r.tfalse;
}

// keep a warp's private counters across a scalar reload
DMD_DEV void blk_reload_scalars(Rep& r) {
  const int64_t a = r.n_pair_pred, b = r.n_nbr_visits;
  const int e = r.error, ei = r.error_info;
  rep_load_scalars(r);
  r.n_pair_pred = a;
  r.n_nbr_visits = b;
  if (e) {
    r.error = e;
    r.error_info = ei;
  }
}

#if !defined(DMD_HOST_TRACE)
// main.F90:1126-1187 by the whole CTA (the serial code is interval_event_cold in dmd_engine.h).  Every thread
// holds identical scalars on entry and on exit.
__device__ __noinline__ void blk_interval_event(BlkShared& S, Rep& r, int w, int nw, double tev) {
  const SysConst& s = *r.c.sys;
  const int N = r.N;
  const int tid = threadIdx.x, nt = blockDim.x;
  const double tf = tev;
  r.tfalse = tf;
  if (w == 0) r.coll += 1;
  r.t = r.t + tf;
  for (int k = tid; k < N + 3; k += nt) r.cal[k].t = r.cal[k].t - tf;  // :1133-1135
  r.interval_max = r.interval_max - tf;
  double moved_far = 0.0;
  for (int k = tid; k < N; k += nt) {  // :1140-1144 + displ.f:20-33
    BeadRec* p = &r.rec[k];
    double x = p->x + p->vx * tf, y = p->y + p->vy * tf, z = p->z + p->vz * tf;
    p->x = x; p->y = y; p->z = z;
    double a = r.oldr[3 * k] - x, b = r.oldr[3 * k + 1] - y, cc = r.oldr[3 * k + 2] - z;
    double dis = a * a + b * b + cc * cc;
    double moved = dis / s.hdelr;
    if (moved > moved_far) moved_far = moved;
  }
  moved_far = warp_max(moved_far);
  if (Warp::lane() == 0) S.moved_far[w] = moved_far;
  __syncthreads();
  moved_far = S.moved_far[0];
  for (int q = 1; q < nw; q++)
    if (S.moved_far[q] > moved_far) moved_far = S.moved_far[q];
  r.tfalse = 0.0;
  bool update = false;
  if (moved_far >= 0.1) {  // displ.f:37-46
    update = true;
    if (moved_far >= 1.25 * 1.25) {
      r.t_fact = r.t_fact / 1.01;
      r.interval = r.t_fact / dmd_sqrt(r.setemp);
    }
  }
  if (update || r.interval > r.interval_max) {  // :1150-1179
    if (!update) {
      if (tid == 0) r.sc->nforcedupdate += 1;
      r.n_forced = r.n_forced * 1.01;
    }
    r.interval_max = r.interval * r.n_forced;
    if (tid == 0) r.sc->nupdates += 1;
    for (int k = tid; k < N; k += nt) {
      BeadRec* p = &r.rec[k];
      double x = p->x - dmd_round(p->x), y = p->y - dmd_round(p->y), z = p->z - dmd_round(p->z);
      p->x = x; p->y = y; p->z = z;
      r.oldr[3 * k] = x; r.oldr[3 * k + 1] = y; r.oldr[3 * k + 2] = z;
    }
    __syncthreads();
    const long long tr = blk_clock();
    {
      // the cell grid of the rebuild lives in the (idle) undo / footprint buffers when it fits: list heads of the
      // coarse grid, chain links and packed cell coordinates then cost a shared-memory access per hop instead of
      // an L2 round trip (the walk is a chain of dependent loads)
      Rep q = r;
      const int ncc = coarse_dim(s.ncr);
      const size_t nheads = (size_t)ncc * ncc * ncc;
      const size_t avail = sizeof(S.undo_old) + sizeof(S.undo_idx) + sizeof(S.fp);  // consecutive members
      const bool fits = (nheads + 2 * (size_t)N) * 4 <= avail;
      if (fits) {
        int32_t* scratch = reinterpret_cast<int32_t*>(&S.undo_old[0][0]);
        q.cellhead = scratch;
        q.cnext = scratch + nheads;
        q.cpk = reinterpret_cast<uint32_t*>(scratch + nheads + N);
        for (size_t k = tid; k < nheads; k += nt) scratch[k] = -1;
        __syncthreads();
      }
      cell_build(q, tid, nt);
      __syncthreads();
      nbor_build(q, tid, nt);
      __syncthreads();
      if (!fits) cell_clear(q, tid, nt);
      if (q.error) set_error(r, q.error, q.error_info);
    }
    for (int l = tid; l < N; l += nt) redo_lane(r, l);  // events(): every bead from interval_max + ltstep
    __syncthreads();
    if (tid == 0) S.cyc[7] += blk_clock() - tr;
    if (r.error && Warp::lane() == 0) {
      S.error = r.error;
      S.error_info = r.error_info;
    }
  }
  if (tid == 0) r.cal[N + 1].t = r.interval * 0.999;  // :1181
  if (w == 0) log_event(r, N + 1, -1, -2, 0);
  r.old_tfalse = r.tfalse;
  __syncthreads();
}
#endif

// ---- the round loop; every thread of the CTA runs it with its warp's view r (warp 0 holds the master copy of
// coll / log position / rng counter).  claim[] has N+3 entries preset to CLAIM_FREE.
DMD_DEV void blk_run(BlkShared& S, Rep& r, uint32_t* claim, int w, int nw) {
  const bool t0 = w == 0 && Warp::lane() == 0;
  while (true) {
    blk_sync();
    if (S.error || S.coll >= S.target) break;
    const long long remaining = S.target - S.coll;
    long long tc = blk_clock();
    // one sweep finds the minimum and, when the time of the last committed event is known, the candidates
    double window = S.window;
    const bool known = S.tlast >= 0.0;
    blk_phase_scan(S, r, w, nw, S.tlast + window, known);
    blk_sync();
    if (t0) S.cyc[0] += blk_clock() - tc, tc = blk_clock();
    const double tmin = blk_tmin(S, nw);
    if (!(tmin < 1e299)) {
      if (t0) {
        S.error = DMD_E_CAL_EMPTY;
        S.error_info = 0;
      }
      continue;
    }
    int nc = S.n_cand;
    int tries = 0;
    while (nc == 0 || nc > BK_CAND) {  // nothing selected yet, or too many: select again around the minimum
      blk_sync();                      // everybody has read n_cand
      if (nc > BK_CAND) window = window * 0.5;
      tries++;
      if (t0) {
        S.n_cand = 0;
        S.window = window;
      }
      blk_sync();
      blk_phase_scan(S, r, w, nw, tries > 60 ? tmin : tmin + window, true);
      blk_sync();
      nc = S.n_cand;
      if (tries > 60 && nc > BK_CAND) nc = BK_CAND;  // a massive exact tie: any BK_CAND of them contain the head
    }
    blk_phase_rank(S, w, nw, nc);
    blk_sync();
    int maxb = nc < nw ? nc : nw;
    if (remaining < maxb) maxb = (int)remaining;
    const int batch = blk_plan(S, nc, maxb);
    if (batch == 0) {
      // ---- the head of the calendar is an H-bond-related pair event or a pseudo-event: serial code, alone
      const int k0 = blk_cand_of_rank(S, nc, 0);
      const int o = S.cand_o[k0];
#if !defined(DMD_HOST_TRACE)
      if (o == r.N + 1) {
        blk_interval_event(S, r, w, nw, S.cand_t[k0]);
        if (t0) {
          S.coll = r.coll;
          S.n_cand = 0;
          S.tlast = -1.0;
          S.st_cold += 1;
          S.cyc[6] += blk_clock() - tc;
        }
        continue;
      }
#endif
      if (w == 0) {
        const CalEnt ev = r.cal[o];
        process_one(r, o, ev);
        rep_save(r);
        if (Warp::lane() == 0) {
          S.coll = r.coll;
          S.n_cand = 0;
          S.tlast = -1.0;
          S.st_cold += 1;
          if (S.stop_at_output && o == r.N + 2) S.target = r.coll;  // the host takes over after the output event
          if (r.error) {
            S.error = r.error;
            S.error_info = r.error_info;
          }
        }
      }
      blk_sync();
      if (w != 0) blk_reload_scalars(r);
      if (t0) S.cyc[6] += blk_clock() - tc;
      continue;
    }
    ListRef li, lj;
    li.up = li.dn = lj.up = lj.dn = nullptr;
    li.nu = li.nd = lj.nu = lj.nd = 0;
    if (t0) S.cyc[1] += blk_clock() - tc, tc = blk_clock();
    if (w < batch) blk_phase_claim(S, r, claim, w, nc, li, lj);
    blk_sync();
    if (t0) S.cyc[2] += blk_clock() - tc, tc = blk_clock();
    if (w < batch) blk_phase_check(S, r, claim, w, li, lj);
    blk_sync();
    if (t0) S.cyc[3] += blk_clock() - tc, tc = blk_clock();
    const int n_exec = blk_n_exec(S, batch);
    if (w < batch) {
      const BlkSlot& sl = S.slot[w];
      blk_footprint(r, sl.owner, sl.j, li, lj, [&](int b) { claim[b] = CLAIM_FREE; });
    }
    if (w < n_exec) blk_exec_event(S, r, w, li, lj);
    blk_sync();
    if (t0) S.cyc[4] += blk_clock() - tc, tc = blk_clock();
    const int n_valid = blk_n_valid(S, n_exec);
    if (w >= n_valid && w < n_exec) blk_rollback(S, r, w);
    if (w == 0) {
      blk_commit(S, r, n_valid);
      if (Warp::lane() == 0) {
        S.coll = r.coll;
        S.n_cand = 0;
        S.tlast = S.slot[n_valid - 1].t;
        S.st_rounds += 1;
        S.st_exec += n_exec;
        S.st_rollback += n_exec - n_valid;
        S.st_conflict += batch - n_exec;
        // steer the window towards ~1.5 candidates per warp
        if (nc < nw + nw / 2) S.window = window * 1.25;
        else if (nc > 2 * nw) S.window = window * 0.8;
        else S.window = window;
        S.cyc[5] += blk_clock() - tc;
      }
    }
    if (r.error && Warp::lane() == 0) {
      S.error = r.error;
      S.error_info = r.error_info;
    }
  }
}

}  // namespace dmd
