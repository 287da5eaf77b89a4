// dmd_engine.h -- the warp-per-replica DMD engine (device code; also compiled 1-lane by tests/host_trace).
//
// One warp runs the whole serial-semantics event loop of main.F90:484-1258 for its replica:
//   pop_min        two-level min-reduction calendar (replaces add_tbin.f/del_tbin.f + main.F90:496-545)
//   pair_event     resolution main.F90:1487-1634, eventdyn.f, bookkeeping main.F90:1638-1937
//   partial_events partial_events.f:16-201: ONE pass per colliding bead, lanes over its up-list, aux slots
//                  and down-list at once; cascaded full re-predictions go through the same pass
//   ghost/interval/output pseudo-events main.F90:997-1049, 1126-1187, 1191-1246  (cold, out of line)
//   cell_build/nbor_build/predict_all (cell_add.f, nbor.f, events.f) with one lane per bead (cold)
// The hot loop is kept small (it must live in the instruction cache while 16+ warps per SM sit at different
// program counters); everything rare is __noinline__ and takes the replica view by value.
// Every loop is written for DMD_W lanes per replica (dmd_warp.h: 32 or 16 on the device, 1 in the host trace build).
// No include guard: dmd_cuda.cu includes this file once per lane count, each time in its own namespace.
#include "dmd_physics.h"
#include "dmd_topology.h"
#include "dmd_types.h"
#include "dmd_warp.h"

#if defined(DMD_HOST_TRACE)
#define DMD_COLD inline
#else
#define DMD_COLD __device__ __noinline__
#endif

namespace dmd {
DMD_VARIANT_BEGIN

constexpr int WLOG = DMD_W == 32 ? 5 : (DMD_W == 16 ? 4 : (DMD_W == 8 ? 3 : 0));  // log2(DMD_W)
static_assert((1 << WLOG) == DMD_W, "DMD_W must be 1, 8, 16 or 32");
constexpr double T_PAD = 1e300;  // calendar padding entries
#ifndef DMD_CQ_CAP
#define DMD_CQ_CAP 96
#endif
constexpr int CQ_CAP = DMD_CQ_CAP;  // per-warp scratch (shared memory on the device): the cascade queue ...
constexpr int CQ_Q = CQ_CAP - 8;    // ... of CQ_Q entries (<= one per down-list candidate of the two main passes) and,
                                    // behind it, the warp's work counters and stale-group masks: updated in the hot
                                    // pass, they would otherwise sit in (and be spilled from) registers -- the
                                    // counters alone measured +7 % events/s

// optional phase profile of the warp-per-replica loop (-DDMD_PHASE_PROF, tools only): lane 0 of every warp adds the
// SM clocks since its previous mark to a per-warp shared-memory slot; the kernel sums the slots into g_phase_cyc
#if defined(DMD_PHASE_PROF) && !defined(DMD_HOST_TRACE)
__device__ unsigned long long g_phase_cyc[16];
__device__ unsigned long long g_phase_acc[160 * 4 * 28 * 16];  // per-replica accumulators of the 8-lane build
__device__ unsigned long long g_cta_end[512];
__device__ unsigned g_cta_sm[512];
#if DMD_PHASE_PROF == 2  // only the finishing times of the CTAs: the loop itself runs as in the product build
#define DMD_PROF_MARK(r, k) do { } while (0)
#else
#define DMD_PROF_MARK(r, k)                                                    \
  do {                                                                         \
    if ((r).prof && Warp::lane() == 0) {                                       \
      const unsigned long long now_ = (unsigned long long)clock64();          \
      (r).prof[k] += now_ - (r).prof[15];                                      \
      (r).prof[15] = now_;                                                     \
    }                                                                          \
  } while (0)
#endif
#else
#define DMD_PROF_MARK(r, k) do { } while (0)
#endif

struct Rep {
  Ctx c;
#if defined(DMD_PHASE_PROF) && !defined(DMD_HOST_TRACE)
  unsigned long long* prof = nullptr;
#endif
  int N, cap, G;
  BeadRec* rec;
  CalEnt* cal;
  int32_t* er34;
  uint32_t *up, *dn;
  uint16_t *nup, *ndn;
  double* oldr;
  int32_t *cellhead, *cnext, *cellof;
  uint32_t* cpk;  // packed fine cell coordinates of each bead (10 bits per dimension)
  unsigned long long* tmin1;  // per-group minima as ordered images (ord_bits64): lowered with an integer atomicMin
  RepScalars* sc;
  EventLogRec* log;
  OutRec* out;
  int32_t* cq;  // cascade queue storage (shared memory on the device)
  int32_t* svc;  // this replica's request word of the list-rebuild service, or nullptr: rebuild in place
  unsigned long long* svc_ctl;  // service counters ([2] requests taken back, [3] cycles spent waiting) and request queue
  int32_t svc_id;               // the replica's index in the request words (what the request queue carries)
  // scalars cached in registers (identical in every lane)
  double t, tfalse, old_tfalse, setemp, interval, t_fact, interval_max, n_forced, avegtime;
  int64_t coll;
  uint64_t seed, ctr;
  int64_t n_pair_pred, n_nbr_visits;
  int32_t n_log, n_out, error, error_info;
};

DMD_DEV void rep_load_scalars(Rep& r) {
  const RepScalars& q = *r.sc;
  r.t = q.t; r.tfalse = q.tfalse; r.old_tfalse = q.old_tfalse; r.setemp = q.setemp; r.interval = q.interval;
  r.t_fact = q.t_fact; r.interval_max = q.interval_max; r.n_forced = q.n_forced; r.avegtime = q.avegtime;
  r.coll = q.coll; r.seed = q.rng_seed; r.ctr = q.rng_ctr;
  r.n_pair_pred = q.n_pair_pred; r.n_nbr_visits = q.n_nbr_visits;
  r.n_log = q.n_log; r.n_out = q.n_out; r.error = q.error; r.error_info = q.error_info;
}

struct Staged {  // the shared-memory copies of the read-only tables (global pointers in the host trace build)
  const HotTables* tab;
  const HotConst* hot;
  const double* bl;
};
DMD_DEV Staged staged_global(const DevArrays& d) {
  Staged st;
  st.tab = reinterpret_cast<const HotTables*>(d.tables);  // HotTables is a prefix of PairTables
  st.hot = d.hot;
  st.bl = d.bl;
  return st;
}

// pair predictions / list entries visited since the last harvest (segmented_pass adds, rep_save collects)
// (32-bit: harvested at every pseudo-event, i.e. every few hundred events, and shared-memory 32-bit adds are native)
DMD_DEV unsigned* warp_counters(const Rep& r) { return reinterpret_cast<unsigned*>(r.cq + CQ_Q); }
// calendar groups 0..63 / 64..127 whose minimum is stale: lane 0 writes, flush_dirty reads after a warp sync
DMD_DEV unsigned long long* warp_dirty(const Rep& r) { return reinterpret_cast<unsigned long long*>(r.cq + CQ_Q + 2); }
DMD_DEV void clear_dirty(Rep& r) {
  if (r.cq && Warp::lane() == 0) warp_dirty(r)[0] = warp_dirty(r)[1] = 0ull;
}
DMD_DEV void harvest_counters(Rep& r) {
  if (!r.cq) return;
  Warp::sync();
  unsigned* c = warp_counters(r);
  const unsigned a = c[0], b = c[1];
  Warp::sync();
  if (Warp::lane() == 0) c[0] = c[1] = 0u;
  Warp::sync();
  r.n_pair_pred += (int64_t)a;
  r.n_nbr_visits += (int64_t)b;
}

DMD_DEV void rep_bind(Rep& r, const DevArrays& d, const Staged& st, int32_t* cq, int rid) {
  // sizes and strides come from the kernel parameters (constant bank), not from *d.sys: every base below is then
  // a function of rid and constants only, which the compiler can recompute instead of spilling
  r.c.sys = d.sys;
  r.c.tab = st.tab;
  r.c.hot = st.hot;
  r.c.bl = st.bl;
  r.c.meta = d.meta;
  r.c.chain = d.chain;
  r.c.sctab = d.sctab;
  const int N = d.n_beads, cap = d.cap;
  r.N = N;
  r.cap = cap;
  r.G = d.ngroups;
  const size_t rr = (size_t)rid;
  r.rec = d.rec + rr * N;
  r.cal = d.cal + rr * d.cal_stride;
  r.er34 = d.er34 + rr * 2 * N;
  r.up = d.up + rr * N * cap;
  r.dn = d.dn + rr * N * cap;
  r.nup = d.nup + rr * N;
  r.ndn = d.ndn + rr * N;
  r.oldr = d.oldr + rr * 3 * N;
  r.cellhead = d.cellhead + rr * (size_t)d.ncc3;
  r.cpk = d.cpk + rr * N;
  r.cnext = d.cnext + rr * N;
  r.cellof = d.cellof + rr * N;
  r.tmin1 = reinterpret_cast<unsigned long long*>(d.tmin1) + rr * d.ngroups;
  r.sc = d.scal + rr;
  r.log = d.log + rr * (d.log_cap > 0 ? d.log_cap : 1);
  r.out = d.out + rr * d.out_cap;
  r.cq = cq;
  if (cq) {
    if (Warp::lane() == 0) {
      warp_counters(r)[0] = warp_counters(r)[1] = 0u;
      warp_dirty(r)[0] = warp_dirty(r)[1] = 0ull;
    }
    Warp::sync();
  }
  r.svc = nullptr;
  r.svc_ctl = nullptr;
  r.svc_id = 0;
  rep_load_scalars(r);
}

DMD_DEV void rep_save(Rep& r) {
  harvest_counters(r);
  if (Warp::lane() == 0) {
    RepScalars& q = *r.sc;
    q.t = r.t; q.tfalse = r.tfalse; q.old_tfalse = r.tfalse; q.setemp = r.setemp; q.interval = r.interval;
    q.t_fact = r.t_fact; q.interval_max = r.interval_max; q.n_forced = r.n_forced; q.avegtime = r.avegtime;
    q.coll = r.coll; q.rng_ctr = r.ctr; q.n_pair_pred = r.n_pair_pred; q.n_nbr_visits = r.n_nbr_visits;
    q.n_log = r.n_log; q.n_out = r.n_out; q.error = r.error; q.error_info = r.error_info;
  }
  Warp::sync();
}

DMD_DEV void set_error(Rep& r, int code, int info) {
  if (r.error == 0) {
    r.error = code;
    r.error_info = info;
  }
}

// ---------------------------------------------------------------------------------------------------------
// calendar: cal[] in groups of 32 entries with a per-group minimum tmin1[]; the event to process is the
// global arg-min (ties -> lowest bead index).  Replaces the bucket lists of add_tbin.f / del_tbin.f;
// dropping the "tim < interval_max" filter is semantics-neutral (SURVEY.md 8a note C).
// ---------------------------------------------------------------------------------------------------------
DMD_DEV void group_min_update(Rep& r, int g) {
  double v = T_PAD;
  for (int q = Warp::lane(); q < 32; q += DMD_W) {
    double x = r.cal[g * 32 + q].t;
    if (x < v) v = x;
  }
  v = warp_min(v);
  if (Warp::lane() == 0) r.tmin1[g] = ord_bits64(v);
}

DMD_DEV void mark_dirty(Rep& r, int g) {  // g warp-uniform
  if (g < 64) {
    if (Warp::lane() == 0) warp_dirty(r)[0] |= 1ull << g;
  } else if (g < 128) {
    if (Warp::lane() == 0) warp_dirty(r)[1] |= 1ull << (g - 64);
  } else {  // very large systems: refresh immediately
    Warp::sync();
    group_min_update(r, g);
  }
}

DMD_DEV int pop_lowest_bit(uint64_t& m) {
#if defined(DMD_HOST_TRACE)
  int g = __builtin_ctzll(m);
#else
  int g = __ffsll((long long)m) - 1;
#endif
  m &= m - 1;
  return g;
}

DMD_DEV void flush_dirty(Rep& r) {
  Warp::sync();
  uint64_t dirty0 = warp_dirty(r)[0], dirty1 = warp_dirty(r)[1];
  Warp::sync();
  clear_dirty(r);
#if DMD_W > 1
  while (dirty0) {  // two groups per round so that their loads overlap
    const int g0 = pop_lowest_bit(dirty0);
    const int g1 = dirty0 ? pop_lowest_bit(dirty0) : -1;
    double x0 = T_PAD, x1 = T_PAD;
#pragma unroll
    for (int q = Warp::lane(); q < 32; q += DMD_W) {
      const double y0 = r.cal[g0 * 32 + q].t;
      const double y1 = g1 >= 0 ? r.cal[g1 * 32 + q].t : T_PAD;
      x0 = y0 < x0 ? y0 : x0;
      x1 = y1 < x1 ? y1 : x1;
    }
    x0 = warp_min(x0);
    if (g1 >= 0) x1 = warp_min(x1);
    if (Warp::lane() == 0) {
      r.tmin1[g0] = ord_bits64(x0);
      if (g1 >= 0) r.tmin1[g1] = ord_bits64(x1);
    }
  }
#else
  while (dirty0) group_min_update(r, pop_lowest_bit(dirty0));
#endif
  while (dirty1) group_min_update(r, 64 + pop_lowest_bit(dirty1));
  Warp::sync();
}

// every lane may have changed the entry of a different bead l (l < 0: none): each such lane sets the bit of its
// group in the warp's shared-memory masks (a 32-bit shared-memory atomic; only the lanes that changed something
// take part), groups >= 128 (very large systems) are refreshed at once
DMD_DEV void mark_dirty_lanes(Rep& r, int l) {
  const int g = l >> 5;  // negative when l < 0
#if DMD_W > 1
  if (g >= 0 && g < 128) atomicOr(reinterpret_cast<unsigned*>(warp_dirty(r)) + (g >> 5), 1u << (g & 31));
#else
  if (g >= 0 && g < 128) warp_dirty(r)[g >> 6] |= 1ull << (g & 63);
#endif
  if (r.G > 128) {  // uniform branch
    unsigned m = Warp::ballot(g >= 128);
    while (m) {
      int src = dmd_ffs(m) - 1;
      m &= m - 1;
      mark_dirty(r, Warp::shfl(g, src));
    }
  }
}

DMD_DEV void rebuild_all_groups(Rep& r) {
  Warp::sync();
  for (int g = 0; g < r.G; g++) group_min_update(r, g);
  clear_dirty(r);
  Warp::sync();
}

// returns the owner index of the earliest entry (or -1) and the entry itself
DMD_DEV int pop_min(Rep& r, CalEnt& ev) {
  unsigned long long best = ord_bits64(T_PAD);
  int bg = 0x7fffffff;
  for (int g = Warp::lane(); g < r.G; g += DMD_W) {
    const unsigned long long v = r.tmin1[g];
    if (v < best) {  // ascending g: the first minimum keeps the lowest group
      best = v;
      bg = g;
    }
  }
  warp_argmin_ord(best, bg);
  if (bg == 0x7fffffff || !(ord_value64(best) < 1e299)) return -1;
  double v = T_PAD;
  int key = 0x7fffffff, pt = -1, ty = -1;
  for (int q = Warp::lane(); q < 32; q += DMD_W) {
    CalEnt e = r.cal[bg * 32 + q];
    if (e.t < v) {
      v = e.t;
      key = q;
      pt = e.ptnr;
      ty = e.type;
    }
  }
  double wv = v;
  int wkey = key;
  warp_argmin(wv, wkey);
#if DMD_W > 1
  const int src = wkey & (DMD_W - 1);  // the lane that scanned the winning entry holds it as its own best
  pt = Warp::shfl(pt, src);
  ty = Warp::shfl(ty, src);
#endif
  ev.t = wv;
  ev.ptnr = pt;
  ev.type = ty;
  return bg * 32 + wkey;
}

DMD_DEV int pack_type(int type, int sc) { return (type & 0xff) | (sc << 8); }
DMD_DEV int type_of(int packed) { return (int)(int8_t)(packed & 0xff); }
DMD_DEV int sc_of(int packed) { return (packed >> 8) & 0xff; }

// ---------------------------------------------------------------------------------------------------------
// partial_events.f:16-201 -- re-prediction after an event on (i, j).
//
// The Fortran does full(i), full(j), down(i), down(j), with a full re-prediction of a down-neighbour l whenever
// nptnr(l) was i or j (a "cascade").  Here each colliding bead gets ONE pass whose lanes cover all its items, and
// the cascades of both beads are collected and done at the end, several beads per pass.  The items of a bead a,
// in the Fortran's evaluation order:
//   [0, nu)            up-list partners b > a           } events.f:26-57 / eventredo_up.f : owner a ("full")
//   [nu, nu+3)         aux slots extra_repuls(a,1:3) > a }
//   [nF, nF+nd)        down-list beads l < a            } eventredo_down.f : owner l, or cascade when
//   [nF+nd, nF+nd+3)   aux slots extra_repuls(a,1:3) < a }   nptnr(l) == a (partial_events.f:73-96)
// Pair times depend on bead records only, never on the calendar.  The calendar updates are applied for i, then
// (after a warp sync, re-reading the entries) for j; cascades run last with the lanes split into segments.
//
// Equivalence with the Fortran order (exact ties in time between different pairs excepted): a cascade rewrites
// cal[l] with a pure function of the records, so it gives the same entry whenever it runs, and any earlier
// lowering of that entry is overwritten; compare-and-lower operations on an entry whose partner is neither i nor
// j commute; an entry whose partner was i (j) is either queued by i's (j's) operation, or -- if j's (i's)
// operation came first and replaced it -- was beaten by a pair time that is then also the full minimum.
// ---------------------------------------------------------------------------------------------------------
// Block engine (dmd_block.h) only: an event executed speculatively records every calendar entry it overwrites
// (index + old entry) so that it can be rolled back, and the earliest event time it creates.
struct Undo {
  int32_t* idx;
  CalEnt* old;
  int n, cap;
  double newmin;
};

// where a pass finds the two lists of its bead: the replica's own arrays, or (block engine) the copy staged in
// shared memory when the event's footprint was claimed
struct ListRef {
  const uint32_t* up;
  const uint32_t* dn;
  int nu, nd;
};

// the two halves of a bead record: integer part (partners, identity, overlay codes) / positions and velocities
DMD_DEV void rec_load_tail(BeadRec& dst, const BeadRec* src) {
#if DMD_W > 1
  const int4 t = *reinterpret_cast<const int4*>(reinterpret_cast<const char*>(src) + 48);
  dst.bptnr = t.x; dst.er1 = t.y; dst.er2 = t.z;
  dst.ident = (uint8_t)(t.w & 0xff); dst.ov1 = (uint8_t)((t.w >> 8) & 0xff); dst.ov2 = (uint8_t)((t.w >> 16) & 0xff);
  dst.pad = 0;
#else
  dst.bptnr = src->bptnr; dst.er1 = src->er1; dst.er2 = src->er2;
  dst.ident = src->ident; dst.ov1 = src->ov1; dst.ov2 = src->ov2; dst.pad = 0;
#endif
}
DMD_DEV void rec_load_head(BeadRec& dst, const BeadRec* src) {
#if DMD_W > 1
  asm volatile("" ::: "memory");  // a load the compiler must not hoist out of the caller's loop
  const double2* p = reinterpret_cast<const double2*>(src);
  const double2 p0 = p[0], p1 = p[1], p2 = p[2];
  dst.x = p0.x; dst.y = p0.y; dst.z = p1.x; dst.vx = p1.y; dst.vy = p2.x; dst.vz = p2.y;
#else
  dst.x = src->x; dst.y = src->y; dst.z = src->z; dst.vx = src->vx; dst.vy = src->vy; dst.vz = src->vz;
#endif
}

// ONE pass = ONE copy of the prediction code in the hot loop (the loop must stay resident in the instruction
// cache while every warp of the SM sits at a different program counter).  A lane works for ONE bead `a` during the
// whole pass and takes the items q = q0, q0 + stride, ... < T of that bead:
//   [0, nu)            up-list partners b > a: "full" items, they feed the running minimum of a (events.f:26-57)
//   [nu, nu+nd)        down-list beads l < a (main pass only): they lower cal[l] -- eventredo_down.f:70-77 -- or
//                      queue l for a cascade when nptnr(l) == a (partial_events.f:73-96)
//   [nu+nd, nu+nd+3)   aux slots extra_repuls(a,1:3), present only when one of them is set: b > a is a full item
//                      (events.f:77), b < a a down item (partial_events.f:100,166)
// Lane mappings:
//   main pass, two beads  the items of BOTH colliding beads in one trip when they fit the lanes (lanes [0, T_i) bead
//                         i, [T_i, T_i + T_j) bead j).  The down items of i are applied first, then -- after a warp
//                         sync, re-reading the entries -- those of j: the order of the Fortran (down(i), then
//                         down(j)).  Returns false without touching j when the items do not fit, or when j holds i
//                         in an aux slot (its down item on cal[i] must see the entry full(i) writes at the end of
//                         the pass): the caller then gives j a pass of its own.
//   main pass, one bead   all lanes (ghost event; bead j after a `false` above; every main pass of the 1-lane build)
//   cascade pass          1, 2 or 4 queued beads at once, full items only (events.f:26-57 for one bead each)
template <bool BLK>
DMD_DEV bool prediction_pass(Rep& r, const bool main_pass, const int i, const int j, const int skip1, const ListRef* li,
                             const ListRef* lj, const int idx, const int rem, int& cqn, Undo* u) {
  const int lane = Warp::lane();
  const int cap = r.cap;
  int a, q0, ssh, nu, nd, er3, skip, T;
  bool act = true, second = false, two = false;
  unsigned segmask = WARP_ALL;  // the lanes (bit k = lane k) that work for the same bead as this lane
  uint32_t e0 = 0, ma;
  BeadRec ra;  // of the lane's own bead only the integer part (partners, identity, overlay) is held across the pass;
               // positions and velocities are fetched again (L1) where the geometry is formed
  const ListRef* lra = nullptr;  // where later trips find the lists of the lane's bead (nullptr: the replica's arrays)
  if (main_pass) {
    two = DMD_W > 1 && j >= 0;
    const int jj = two ? j : i;
    const uint32_t* const upi = li ? li->up : r.up + (size_t)i * cap;
    const uint32_t* const dni = li ? li->dn : r.dn + (size_t)i * cap;
    const uint32_t* const upj = two ? (lj ? lj->up : r.up + (size_t)jj * cap) : upi;
    const uint32_t* const dnj = two ? (lj ? lj->dn : r.dn + (size_t)jj * cap) : dni;
    // ---- level-1 loads, all independent: before the list lengths are known every lane fetches "its" entry of the
    // four lists; the lane that ends up with item q of a list takes it from lane q by a shuffle
#if DMD_W > 1
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    if (lane < cap) {
      s0 = upi[lane];
      s1 = dni[lane];
      if (two) {
        s2 = upj[lane];
        s3 = dnj[lane];
      }
    }
#endif
    BeadRec ti, tj;
    (rec_load_tail)(ti, &r.rec[i]);
    (rec_load_tail)(tj, &r.rec[jj]);
    const int nu_i = li ? li->nu : (int)r.nup[i], nd_i = li ? li->nd : (int)r.ndn[i];
    const int nu_j = lj ? lj->nu : (int)r.nup[jj], nd_j = lj ? lj->nd : (int)r.ndn[jj];
    const int e3i = r.er34[2 * i], e3j = r.er34[2 * jj];
    const uint32_t mi = r.c.meta[i], mj = r.c.meta[jj];
    int Ti = nu_i + nd_i + ((ti.er1 >= 0 || ti.er2 >= 0 || e3i >= 0) ? 3 : 0);
    int Tj = nu_j + nd_j + ((tj.er1 >= 0 || tj.er2 >= 0 || e3j >= 0) ? 3 : 0);
    Ti = Ti > 0 ? Ti : 1;  // a bead without items still needs a lane that writes its calendar entry
    Tj = Tj > 0 ? Tj : 1;
    if (two && (Ti + Tj > DMD_W || tj.er1 == i || tj.er2 == i || e3j == i)) two = false;  // j gets its own pass
    {  // work counters (roofline input): slots and list entries of the bead(s) of this pass
      unsigned* c = warp_counters(r);
      const unsigned slots = (unsigned)(nu_i + nd_i + 6 + (two ? nu_j + nd_j + 6 : 0));
      const unsigned visits = (unsigned)(nu_i + nd_i + (two ? nu_j + nd_j : 0));
      if (lane == 0) {
#if DMD_W > 1
        atomicAdd(&c[0], slots);
        atomicAdd(&c[1], visits);
#else
        c[0] += slots;
        c[1] += visits;
#endif
      }
    }
    q0 = lane;
    ssh = WLOG;
    if (two) {
      second = lane >= Ti;
      q0 = second ? lane - Ti : lane;
      const unsigned m1 = (1u << Ti) - 1u;  // Ti < DMD_W here
      segmask = second ? WARP_ALL & ~m1 : m1;
    }
    a = second ? j : i;
    nu = second ? nu_j : nu_i;
    nd = second ? nd_j : nd_i;
    T = second ? Tj : Ti;
    er3 = second ? e3j : e3i;
    ma = second ? mj : mi;
    skip = second ? i : skip1;
    ra.bptnr = second ? tj.bptnr : ti.bptnr;
    ra.er1 = second ? tj.er1 : ti.er1;
    ra.er2 = second ? tj.er2 : ti.er2;
    ra.ident = second ? tj.ident : ti.ident;
    ra.ov1 = second ? tj.ov1 : ti.ov1;
    ra.ov2 = second ? tj.ov2 : ti.ov2;
    ra.pad = 0;
    lra = second ? lj : li;
#if DMD_W > 1
    {  // the entry of the lane's first item
      const bool isup = q0 < nu;
      const int src = (isup ? q0 : q0 - nu) & (DMD_W - 1);
      const uint32_t t0 = (uint32_t)Warp::shfl((int)s0, src), t1 = (uint32_t)Warp::shfl((int)s1, src);
      const uint32_t t2 = (uint32_t)Warp::shfl((int)s2, src), t3 = (uint32_t)Warp::shfl((int)s3, src);
      e0 = second ? (isup ? t2 : t3) : (isup ? t0 : t1);
    }
#endif
  } else {
    // cascades, several beads per pass: segment count by the queue length only (no size look-up: that would put a
    // dependent load in front of the pass); a list longer than its segment takes another trip
#if DMD_W > 1
    const int sh = rem >= 3 ? 2 : (rem == 2 ? 1 : 0);
    const int SEG = DMD_W >> sh;
    const int g = lane >> (WLOG - sh);
    segmask = (sh == 0 ? WARP_ALL : ((1u << SEG) - 1u)) << (g * SEG);
    q0 = lane & (SEG - 1);
    ssh = WLOG - sh;
#else
    const int g = 0;
    q0 = 0;
    ssh = 0;
#endif
    act = g < rem;
    a = r.cq[idx + (act ? g : 0)];
#if DMD_W > 1
    if (q0 < cap) e0 = r.up[(size_t)a * cap + q0];
#endif
    (rec_load_tail)(ra, &r.rec[a]);
    ma = r.c.meta[a];
    nu = (int)r.nup[a];
    nd = 0;
    er3 = r.er34[2 * a];
    T = act ? nu + ((ra.er1 >= 0 || ra.er2 >= 0 || er3 >= 0) ? 3 : 0) : 0;
    if (act && T == 0) T = 1;
    skip = -1;
    if (q0 == 0 && act) {
      unsigned* c = warp_counters(r);
#if DMD_W > 1
      atomicAdd(&c[0], (unsigned)(nu + 3));
      atomicAdd(&c[1], (unsigned)nu);
#else
      c[0] += (unsigned)(nu + 3);
      c[1] += (unsigned)nu;
#endif
    }
  }
  int ntrip = T > q0 ? (T - q0 + (1 << ssh) - 1) >> ssh : 0;
  ntrip = warp_max_u(ntrip);
  double best = r.interval_max + LTSTEP - r.tfalse;
  int bpos = 0x7fffffff, bj = -1, btype = -1;
#pragma unroll 1
  for (int trip = 0, q = q0; trip < ntrip; trip++, q += 1 << ssh) {
    int b = -1, sc = 1;  // the other bead of the pair and the pair's static class
    bool full = false;
    if (q < nu + nd) {
      const bool isup = q < nu;
      full = isup;
      uint32_t e = e0;
      if (DMD_W == 1 || trip > 0) {  // later trips (long lists): a dependent load
        const uint32_t* const lst = lra ? (isup ? lra->up : lra->dn) : (isup ? r.up : r.dn) + (size_t)a * cap;
        e = lst[isup ? q : q - nu];
      }
      b = (int)(e & NB_MASK);
      sc = (int)(e >> NB_SHIFT);
      if (!isup && b == skip) b = -1;  // partial_events.f:136
    } else if (q < T) {
      const int k = q - nu - nd;
      b = k == 0 ? ra.er1 : (k == 1 ? ra.er2 : er3);
      full = b > a;                                 // events.f:77
      if (b < 0 || (!full && !main_pass)) b = -1;  // partial_events.f:100,166
    }
    if (q >= T) b = -1;  // an idle cascade segment (T == 0) has no items
    const bool down = b >= 0 && !full;
    const bool late = two && second;  // j's down items are decided after i's, on re-read entries
    double tij = T_NONE;
    int type = -1;
    CalEnt eb;
    eb.t = 0.0; eb.ptnr = -1; eb.type = -1;
    if (b >= 0) {
      // level-2 loads, all depending on b only (+ the position / velocity part of a's own record)
      const BeadRec rb = r.rec[b];
      (rec_load_head)(ra, &r.rec[a]);
      uint32_t mlo = ma;
      if (!full) {
        mlo = r.c.meta[b];
        if (!late) eb = r.cal[b];
      }
      const int code = overlay_code(sc, a, ra, b, rb);
      {  // one prediction site for both orientations (owner = lower index: a when full, b otherwise)
        const Geom gm = pair_geom(ra, rb, r.tfalse);
        const double rijsq = gm.rx * gm.rx + gm.ry * gm.ry + gm.rz * gm.rz;
        const double vijsq = gm.vx * gm.vx + gm.vy * gm.vy + gm.vz * gm.vz;
        const int idlo = full ? ra.ident : rb.ident, idhi = full ? rb.ident : ra.ident;
        const bool bonded = full ? ra.bptnr == b : rb.bptnr == a;
        pair_time_core(r.c, code, gm.bij, rijsq, vijsq, idlo, idhi, mlo, bonded, tij, type);
      }
      if (full) {
        if (tij < best) {  // strict: first in evaluation order wins (events.f:53)
          best = tij;
          bpos = q;
          bj = b;
          btype = pack_type(type, sc);
        }
      } else {
        tij = tij + r.tfalse;
      }
    }
    if (main_pass) {
#pragma unroll 1
      for (int ph = 0; ph < (two ? 2 : 1); ph++) {
        const bool mine = down && late == (ph == 1);
        if (ph == 1) {
          Warp::sync();  // j's operations re-read the entries i's operations may have lowered
          if (mine) eb = r.cal[b];
        }
        const bool need_full = mine && eb.ptnr == a;  // l's next event was with a: full re-prediction of l (cascade)
        const bool lower = mine && !need_full && tij < eb.t;  // eventredo_down.f:70-77
        if (lower) {
          CalEnt ne;
          ne.t = tij;
          ne.ptnr = a;
          ne.type = pack_type(type, sc);
          r.cal[b] = ne;
          if (BLK && tij < u->newmin) u->newmin = tij;
        }
        if (BLK) {  // undo log of the lowered entries (eb = the old entry of this lane's bead)
          const unsigned mc = Warp::ballot(lower);
          if (mc) {
            const int pos = u->n + dmd_popc(mc & ((1u << lane) - 1u));
            if (lower && pos < u->cap) {
              u->idx[pos] = b;
              u->old[pos] = eb;
            }
            u->n += dmd_popc(mc);
          }
        } else if (lower) {
          // an entry that was only LOWERED: the group minimum follows with an integer atomicMin on its ordered image
          // (exact -- no rescan of the group's 32 entries); entries that may rise (the writer below) mark the group
#if DMD_W > 1
          atomicMin(&r.tmin1[b >> 5], ord_bits64(tij));
#else
          if (ord_bits64(tij) < r.tmin1[b >> 5]) r.tmin1[b >> 5] = ord_bits64(tij);
#endif
        }
        const unsigned m = Warp::ballot(need_full);
        if (m) {
          const int pos = cqn + dmd_popc(m & ((1u << lane) - 1u));
          if (need_full && pos < CQ_Q) r.cq[pos] = b;
          cqn += dmd_popc(m);
        }
      }
    }
  }
  // ---- the minimum of each bead's full items -> cal[a]; the lane that holds it writes the entry (a bead without
  // any event: its first lane)
  double wbest = best;
  int wpos = bpos;
  seg_argmin(wbest, wpos, segmask);
  const bool mine = act && bpos == wpos && bpos != 0x7fffffff;
  const unsigned any_mine = Warp::ballot(mine) & segmask;
  const bool writer = act && (any_mine ? mine : q0 == 0);
  CalEnt ne;
  ne.t = wbest + r.tfalse;
  ne.ptnr = any_mine ? bj : -1;
  ne.type = any_mine ? btype : -1;
  CalEnt old;
  old.t = 0.0; old.ptnr = -1; old.type = -1;
  if (writer) {
    if (BLK) old = r.cal[a];
    r.cal[a] = ne;
  }
  if (BLK) {
    const unsigned mw = Warp::ballot(writer);
    const int pos = u->n + dmd_popc(mw & ((1u << lane) - 1u));
    if (writer && pos < u->cap) {
      u->idx[pos] = a;
      u->old[pos] = old;
    }
    u->n += dmd_popc(mw);
    if (writer && ne.t < u->newmin) u->newmin = ne.t;
  } else {
    mark_dirty_lanes(r, writer ? a : -1);
  }
  return two;
}

DMD_DEV void repuls_del_b(Rep& r, int n, int cb);

template <bool BLK>
DMD_DEV void partial_events_t(Rep& r, int i, int j, bool xpulse_del, Undo* u, const ListRef* li, const ListRef* lj) {
  Warp::sync();
  int cqn = 0, idx = 0;
  int stage = 0;  // 0 main pass (both beads, or bead i), 1 bead j alone, 2 prepare the cascade queue, 3 cascades
#pragma unroll 1
  while (true) {
    int pi = i, pj = -1, skip = -1, rem = 0;
    const ListRef *l1 = li, *l2 = nullptr;
    if (stage == 0) {
      pj = j;  // j < 0: ghost event, one bead only (main.F90:1049)
      l2 = lj;
    } else if (stage == 1) {
      pi = j;
      skip = i;
      l1 = lj;
    } else {
      if (stage == 2) {  // all down items are done: prepare the cascade queue
        stage = 3;
        if (cqn > CQ_Q) {
          set_error(r, DMD_E_NBR_CAP, cqn);
          cqn = 0;
        }
        if (cqn > 1) {  // a bead queued by both i and j is re-predicted once
          int keep_n = 0;
          for (int base = 0; base < cqn; base += DMD_W) {
            const int k = base + Warp::lane();
            bool keep = false;
            int v = -1;
            if (k < cqn) {
              v = r.cq[k];
              keep = true;
              for (int m = 0; m < k; m++)
                if (r.cq[m] == v) keep = false;
            }
            Warp::sync();
            const unsigned mk = Warp::ballot(keep);
            if (keep) r.cq[keep_n + dmd_popc(mk & ((1u << Warp::lane()) - 1u))] = v;
            keep_n += dmd_popc(mk);
            Warp::sync();
          }
          cqn = keep_n;
        }
      }
      DMD_PROF_MARK(r, 4);
      if (idx >= cqn) break;
      rem = cqn - idx;
    }
    const bool both = prediction_pass<BLK>(r, stage < 2, pi, pj, skip, l1, l2, idx, rem, cqn, u);
    Warp::sync();  // later passes re-read the entries this one may have lowered
    DMD_PROF_MARK(r, stage < 2 ? 3 : 5);
    if (stage == 0) stage = (j >= 0 && !both) ? 1 : 2;
    else if (stage == 1) stage = 2;
    else idx += DMD_W > 1 ? (rem >= 3 ? (rem < 4 ? rem : 4) : rem) : 1;
  }
  if (xpulse_del) {
    if (Warp::lane() == 0) {
      if (r.rec[i].ident < r.rec[j].ident) repuls_del_b(r, i, j);
      else repuls_del_b(r, j, i);
    }
    Warp::sync();
  }
}

DMD_DEV void partial_events(Rep& r, int i, int j, bool xpulse_del) { partial_events_t<false>(r, i, j, xpulse_del, nullptr, nullptr, nullptr); }


// one lane does the whole list of bead l (bulk path: events.f:23-107 with one lane per bead)
DMD_DEV void redo_lane(Rep& r, int l) {
  const BeadRec rl = r.rec[l];
  const uint32_t ml = r.c.meta[l];
  const int nu = r.nup[l];
  const int er3 = r.er34[2 * l];
  double best = r.interval_max + LTSTEP - r.tfalse;
  int bj = -1, btype = -1;
  for (int p = 0; p < nu + 3; p++) {
    int j, sc = 1;
    if (p < nu) {
      uint32_t e = r.up[(size_t)l * r.cap + p];
      j = (int)(e & NB_MASK);
      sc = (int)(e >> NB_SHIFT);
    } else {
      int k = p - nu;
      j = k == 0 ? rl.er1 : (k == 1 ? rl.er2 : er3);
      if (j <= l) continue;
    }
    const BeadRec rj = r.rec[j];
    const int code = overlay_code(sc, l, rl, j, rj);
    double tij = T_NONE;
    int type = -1;
    pair_time(r.c, code, rl, rj, ml, rl.bptnr == j, r.tfalse, tij, type);
    if (tij < best) {
      best = tij;
      bj = j;
      btype = pack_type(type, sc);
    }
  }
  CalEnt ne;
  ne.t = best + r.tfalse;
  ne.ptnr = bj;
  ne.type = btype;
  r.cal[l] = ne;
}

#if !defined(DMD_HOST_TRACE)
// events.f:23-107 for the list-rebuild service: the items of 32 consecutive beads are spread over the 32 lanes of a
// hardware warp (one pair prediction per lane and trip) instead of one bead per lane -- a bead has 0 ... ~15 items, so
// the per-bead loop of redo_lane() leaves more than half of the lanes idle.  Same items, same arithmetic, and the same
// winner: the earliest time, ties to the item that comes first in the bead's own order (events.f:53 is a strict <).
DMD_DEV void predict_all_flat(Rep& r, int warp, int nwarps) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, cap = r.cap;
  const double best0 = r.interval_max + LTSTEP - r.tfalse;
  for (int l0 = warp * 32; l0 < r.N; l0 += nwarps * 32) {
    const int l = l0 + lane;
    const bool valid = l < r.N;
    const int ll = valid ? l : r.N - 1;
    const int e1 = r.rec[ll].er1, e2 = r.rec[ll].er2, e3 = r.er34[2 * ll];
    const int nu = valid ? (int)r.nup[ll] : 0;
    const bool a0 = valid && e1 > l, a1 = valid && e2 > l, a2 = valid && e3 > l;  // events.f:77 (skips "none" = -1 too)
    const int n = nu + (int)a0 + (int)a1 + (int)a2;
    int inc = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(FULL, inc, d);
      if (lane >= d) inc += v;
    }
    const int off = inc - n, T = __shfl_sync(FULL, inc, 31);
    double best = best0;
    int bj = -1, bty = -1;
    for (int b = 0; b < T; b += 32) {
      const int x = b + lane;
      const bool have = x < T;
      int o = 0;  // the lane whose bead owns item x: the largest o with off(o) <= x
#pragma unroll
      for (int st = 16; st >= 1; st >>= 1) {
        const int v = __shfl_sync(FULL, off, o + st);
        if (v <= x) o += st;
      }
      const int p = x - __shfl_sync(FULL, off, o), nuo = __shfl_sync(FULL, nu, o), lo = l0 + o;
      const int oe1 = __shfl_sync(FULL, e1, o), oe2 = __shfl_sync(FULL, e2, o), oe3 = __shfl_sync(FULL, e3, o);
      int j = -1, sc = 1;
      if (have) {
        if (p < nuo) {
          const uint32_t e = r.up[(size_t)lo * cap + p];
          j = (int)(e & NB_MASK);
          sc = (int)(e >> NB_SHIFT);
        } else {  // the (p - nu)-th auxiliary partner above the bead, in the order er1, er2, er3
          int c = p - nuo;
          if (oe1 > lo) { if (c == 0) j = oe1; c--; }
          if (oe2 > lo) { if (c == 0) j = oe2; c--; }
          if (oe3 > lo) { if (c == 0) j = oe3; c--; }
        }
      }
      double tij = T_NONE;
      int ty = -1;
      if (j >= 0) {
        const BeadRec ro = r.rec[lo], rj = r.rec[j];
        const int code = overlay_code(sc, lo, ro, j, rj);
        int type = -1;
        pair_time(r.c, code, ro, rj, r.c.meta[lo], ro.bptnr == j, r.tfalse, tij, type);
        ty = pack_type(type, sc);
      }
      // segmented arg-min scan over the lanes (items of one bead are neighbours); the left item wins a tie
      unsigned hi, lw;
      ord_split(tij, hi, lw);
      int src = lane;
      const int key = have ? o : 32 + lane;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned h2 = __shfl_up_sync(FULL, hi, d), w2 = __shfl_up_sync(FULL, lw, d);
        const int s2 = __shfl_up_sync(FULL, src, d), k2 = __shfl_up_sync(FULL, key, d);
        if (lane >= d && k2 == key && (h2 < hi || (h2 == hi && w2 <= lw))) {
          hi = h2;
          lw = w2;
          src = s2;
        }
      }
      // every lane, as the OWNER of its bead, reads the result at the last lane of its segment in this trip
      const int seg_end = off + n < b + 32 ? off + n : b + 32;
      const bool seg = n > 0 && seg_end > b && off < b + 32 && seg_end > off;
      const int last = seg ? seg_end - 1 - b : 0;
      const unsigned sh = __shfl_sync(FULL, hi, last), sw = __shfl_sync(FULL, lw, last);
      const int ss = __shfl_sync(FULL, src, last);
      const int cj = __shfl_sync(FULL, j, ss), cty = __shfl_sync(FULL, ty, ss);
      const double ts = ord_join(sh, sw);
      if (seg && ts < best) {  // strict: an earlier trip holds earlier items
        best = ts;
        bj = cj;
        bty = cty;
      }
    }
    if (valid) {
      CalEnt ne;
      ne.t = best + r.tfalse;
      ne.ptnr = bj;
      ne.type = bty;
      r.cal[l] = ne;
    }
  }
}
#endif

// events.f:23-123 for the whole replica (every bead is re-derived from interval_max + ltstep)
DMD_DEV void predict_all(Rep& r) {
  Warp::sync();
  for (int l = Warp::lane(); l < r.N; l += DMD_W) redo_lane(r, l);
  rebuild_all_groups(r);
}

// ---------------------------------------------------------------------------------------------------------
// H-bond auxiliary-shoulder bookkeeping (lane 0 only, directly on global memory)
// ---------------------------------------------------------------------------------------------------------
struct AuxIdx {
  int ncim1, ncai, ncaj, nnjp1;
};
// repuls_add.f:14-28 index arithmetic; n = the N bead, cb = the C bead (0-based)
DMD_DEV AuxIdx aux_indices(const Rep& r, int n, int cb) {
  const SysConst& s = *r.c.sys;
  int Ln = s.chnln[meta_sp(r.c.meta[n])], Lc = s.chnln[meta_sp(r.c.meta[cb])];
  AuxIdx a;
  a.ncim1 = n + Ln - 1;
  a.ncai = n - Ln;
  a.ncaj = cb - 2 * Lc;
  a.nnjp1 = cb - Lc + 1;
  return a;
}

// write matrix entry ev_code(a,b) = code (and ev_code(b,a) = mirror) into whichever bead lists the other
DMD_DEV void set_pair_code(Rep& r, int a, int b, int code) {
  BeadRec* pa = &r.rec[a];
  BeadRec* pb = &r.rec[b];
  if (pa->er1 == b) pa->ov1 = (uint8_t)code;
  if (pa->er2 == b) pa->ov2 = (uint8_t)code;
  if (pb->er1 == a) pb->ov1 = (uint8_t)ov_mirror(code);
  if (pb->er2 == a) pb->ov2 = (uint8_t)ov_mirror(code);
}

// repuls_add.f:14-47
DMD_DEV void repuls_add(Rep& r, int n, int cb) {
  AuxIdx a = aux_indices(r, n, cb);
  r.rec[n].er1 = a.ncaj;
  r.rec[n].er2 = a.nnjp1;
  r.rec[cb].er1 = a.ncai;
  r.rec[cb].er2 = a.ncim1;
  r.rec[n].ov1 = 1; r.rec[n].ov2 = 1; r.rec[cb].ov1 = 1; r.rec[cb].ov2 = 1;
  set_pair_code(r, n, a.ncaj, 40);
  set_pair_code(r, n, a.nnjp1, 40);
  set_pair_code(r, cb, a.ncai, 40);
  set_pair_code(r, cb, a.ncim1, 40);
  r.er34[2 * a.ncaj] = n;
  r.er34[2 * a.nnjp1] = n;
  r.er34[2 * a.ncai] = cb;
  r.er34[2 * a.ncim1] = cb;
  r.er34[2 * n + 1] = cb;
  r.er34[2 * cb + 1] = n;
}

// repuls_del_a.f:14-37
DMD_DEV void repuls_del_a(Rep& r, int n, int cb) {
  AuxIdx a = aux_indices(r, n, cb);
  set_pair_code(r, n, a.ncaj, 1);
  set_pair_code(r, n, a.nnjp1, 1);
  set_pair_code(r, cb, a.ncai, 1);
  set_pair_code(r, cb, a.ncim1, 1);
}

// repuls_del_b.f:14-39
DMD_DEV void repuls_del_b(Rep& r, int n, int cb) {
  AuxIdx a = aux_indices(r, n, cb);
  r.rec[n].er1 = -1; r.rec[n].er2 = -1; r.rec[n].ov1 = 1; r.rec[n].ov2 = 1;
  r.rec[cb].er1 = -1; r.rec[cb].er2 = -1; r.rec[cb].ov1 = 1; r.rec[cb].ov2 = 1;
  r.er34[2 * a.ncaj] = -1;
  r.er34[2 * a.nnjp1] = -1;
  r.er34[2 * a.ncai] = -1;
  r.er34[2 * a.ncim1] = -1;
  r.er34[2 * n + 1] = -1;
  r.er34[2 * cb + 1] = -1;
}

DMD_DEV void log_event(Rep& r, int i, int j, int type, int code) {
  if (r.n_log < r.c.sys->log_cap) {
    if (Warp::lane() == 0) {
      EventLogRec e;
      e.t = r.t + r.tfalse;
      e.i = i + 1;
      e.j = j + 1;
      e.type = type;
      e.evcode = code;
      r.log[r.n_log] = e;
    }
    r.n_log++;
  }
}

// the four auxiliary distances of repuls_check.f:32-80 / repuls_check_3.f:34-102; lanes 0..3 take one each.
// excl = auxiliary bead to leave out (repuls_check_3) or -1.  Returns the number of distances that exceed
// their shoulder diameter (the excluded one is not counted).
DMD_DEV int aux_clear_count(Rep& r, int n, int cb, int excl) {
  AuxIdx a = aux_indices(r, n, cb);
  int cnt = 0;
  for (int q = Warp::lane(); q < 4; q += DMD_W) {
    int p = q < 2 ? n : cb;
    int x = q == 0 ? a.ncaj : (q == 1 ? a.nnjp1 : (q == 2 ? a.ncai : a.ncim1));
    if (x != excl) {
      double d = pair_dist(r.rec[p], r.rec[x], r.tfalse);
      if (d > r.c.sys->shder[q]) cnt++;
    }
  }
  return warp_sum(cnt);
}

// ---- cold part of a pair event: H-bond related types (anything but core / bond events), < 1 % of events.
// Resolution main.F90:1487-1634, eventdyn.f (types 4-13) / bumped.f, bookkeeping main.F90:1638-1937.
// Takes the replica view by value (the hot loop keeps its copy in registers); rng counter goes through r.sc.
struct ColdRes {
  int ct;
  int xpulse;
  uint64_t ctr;
};
DMD_COLD ColdRes pair_event_cold(Rep r, int i, int j, int ct, int code) {
  const SysConst& s = *r.c.sys;
  BeadRec ri = r.rec[i], rj = r.rec[j];
  const uint32_t mi = r.c.meta[i], mj = r.c.meta[j];
  const bool bonded = ri.bptnr == j;
  const bool tok = !is_terminal_bead(s, mi) && !is_terminal_bead(s, mj);
  const int er4i = r.er34[2 * i + 1], er4j = r.er34[2 * j + 1];
  if (ct == 7) {
    if (er4i < 0 && er4j < 0) {
      if (tok) {
        int n = ri.ident < rj.ident ? i : j, cb = ri.ident < rj.ident ? j : i;
        ct = aux_clear_count(r, n, cb, -1) == 4 ? 4 : 14;  // repuls_check.f:77-80
      } else {
        double ran_non = rng_uniform(r.seed, r.ctr);
        ct = ran_non <= 0.2 ? 4 : 9;
      }
    } else {
      ct = 9;
    }
  } else if (ct == 10 || ct == 12) {
    const int x = code < 45 ? i : j;      // the H-bond bead of the auxiliary pair
    const int other = code < 45 ? j : i;  // the auxiliary bead
    const int hb = code < 45 ? er4i : er4j;
    if (hb < 0) {
      ct = 15;  // cannot happen while the overlay is consistent; treated as a no-event nudge
    } else {
      const BeadRec rx = code < 45 ? ri : rj;
      const BeadRec rh = r.rec[hb];
      if (ct == 10) {
        int n = rx.ident < rh.ident ? x : hb, cb = rx.ident < rh.ident ? hb : x;
        ct = aux_clear_count(r, n, cb, other) == 3 ? 5 : 15;  // repuls_check_3.f:98-102
      } else if (rx.bptnr == hb) {  // check_sigma.f:12-29
        Geom g = pair_geom(rx, rh, r.tfalse);
        double rijsq = g.rx * g.rx + g.ry * g.ry + g.rz * g.rz;
        double diff = rijsq - r.c.tab->sigma_sq[tix(rx.ident, rh.ident)];
        ct = diff < 0.0 ? 13 : 6;
      } else {
        ct = 15;
      }
    }
  }
  if (ct < 14) ct = event_dynamics(r.c, ct, code, ri, rj, mi, bonded, r.tfalse);  // main.F90:1636
  else bump_off(r.c, code, ri, rj, r.tfalse);                                       // :1829,1881,1884
  Warp::sync();
  if (Warp::lane() == 0) {
    BeadRec* pi = &r.rec[i];
    BeadRec* pj = &r.rec[j];
    pi->x = ri.x; pi->y = ri.y; pi->z = ri.z; pi->vx = ri.vx; pi->vy = ri.vy; pi->vz = ri.vz;
    pj->x = rj.x; pj->y = rj.y; pj->z = rj.z; pj->vx = rj.vx; pj->vy = rj.vy; pj->vz = rj.vz;
    const int n = ri.ident < rj.ident ? i : j, cb = ri.ident < rj.ident ? j : i;
    if (ct == 20) {
      if (ri.ident + rj.ident == 5) {
        pi->bptnr = j; pj->bptnr = i;
        pi->ident = (uint8_t)(ri.ident + 4); pj->ident = (uint8_t)(rj.ident + 4);
        if (tok) repuls_add(r, n, cb);
      }
    } else if (ct == 21) {
      if (ri.ident <= 8) {
        if (tok) repuls_del_a(r, n, cb);
        if (bonded) {
          pi->bptnr = -1; pj->bptnr = -1;
          pi->ident = (uint8_t)(ri.ident - 4); pj->ident = (uint8_t)(rj.ident - 4);
        }
      }
    } else if (ct == 24 || ct == 25) {
      const int x = code < 45 ? i : j;
      const int hb = code < 45 ? er4i : er4j;
      if (hb >= 0) {
        BeadRec* px = &r.rec[x];
        BeadRec* ph = &r.rec[hb];
        if (ct == 24) {
          if (px->ident + ph->ident == 5) {
            px->bptnr = hb; ph->bptnr = x;
            px->ident = (uint8_t)(px->ident + 4); ph->ident = (uint8_t)(ph->ident + 4);
          }
        } else if (px->ident >= 5) {
          px->bptnr = -1; ph->bptnr = -1;
          px->ident = (uint8_t)(px->ident - 4); ph->ident = (uint8_t)(ph->ident - 4);
        }
      }
    } else if (ct == 14) {
      if (tok) repuls_add(r, n, cb);
    } else if (ct == 16) {
      if (tok) repuls_del_a(r, n, cb);
    }
  }
  ColdRes res;
  res.xpulse = ((ct == 21 && ri.ident <= 8 && tok) || (ct == 16 && tok)) ? 1 : 0;
  res.ct = ct;
  res.ctr = r.ctr;
  Warp::sync();
  return res;
}

// worker block main.F90:1429-1959 with current state, then master main.F90:926; the re-prediction (:943) is done
// by the caller.  Returns xpulse_del.
DMD_DEV bool pair_event(Rep& r, int i, const CalEnt& ev) {
  const int j = ev.ptnr;
  int ct = type_of(ev.type);
  bool xpulse_del = false;
  int code;
  if (ct >= 1 && ct <= 3) {  // hot: hard-core and bond events (> 99 % of all events)
    BeadRec ri = r.rec[i], rj = r.rec[j];
    code = overlay_code(sc_of(ev.type), i, ri, j, rj);  // ev_code(i,j), main.F90:587
    ct = event_dynamics_hot(r.c, ct, code, ri, rj, r.c.meta[i], ri.bptnr == j, r.tfalse);
    Warp::sync();
    if (Warp::lane() == 0) {
      BeadRec* pi = &r.rec[i];
      BeadRec* pj = &r.rec[j];
      pi->x = ri.x; pi->y = ri.y; pi->z = ri.z; pi->vx = ri.vx; pi->vy = ri.vy; pi->vz = ri.vz;
      pj->x = rj.x; pj->y = rj.y; pj->z = rj.z; pj->vx = rj.vx; pj->vy = rj.vy; pj->vz = rj.vz;
    }
  } else {
    code = overlay_code(sc_of(ev.type), i, r.rec[i], j, r.rec[j]);
    ColdRes cr = pair_event_cold(r, i, j, ct, code);
    ct = cr.ct;
    xpulse_del = cr.xpulse != 0;
    r.ctr = cr.ctr;
  }
  if (Warp::lane() == 0 && ct >= 0 && ct < 32) {  // main.F90:926 (a reduction without return value: nothing to wait for)
#if DMD_W > 1
    atomicAdd(reinterpret_cast<unsigned long long*>(&r.sc->nevents[ct]), 1ull);
#else
    r.sc->nevents[ct] += 1;
#endif
  }
  log_event(r, i, j, ct, code);
  return xpulse_del;
}

// ---------------------------------------------------------------------------------------------------------
// cell grid + neighbour lists (cell_add.f:12-28, nbor.f:33-137), one lane per bead
// ---------------------------------------------------------------------------------------------------------
DMD_DEV void cell_coords(const SysConst& s, const BeadRec& b, int& cx, int& cy, int& cz) {
  cx = (int)((b.x + s.half) / s.width);  // cell_add.f:22 -- a true fp64 division, truncation toward zero
  cy = (int)((b.y + s.half) / s.width);
  cz = (int)((b.z + s.half) / s.width);
}

DMD_DEV int exch(int32_t* p, int v) {
#if defined(DMD_HOST_TRACE)
  int o = *p;
  *p = v;
  return o;
#else
  return atomicExch(p, v);
#endif
}

// The reference bins beads into cells of width rl/2 and searches the 5 x 5 x 5 block around a bead's cell
// (n_wrap = 2, cell_link.f:16-94).  At the shipped concentrations that grid is almost empty (0.02 beads per
// cell), so 125 list heads are probed to find ~15 candidates.  Here the linked lists hang off a COARSE grid of
// 2 x 2 x 2 fine cells (at most 4 distinct coarse cells per dimension cover the fine range c-2..c+2, 27 in the
// typical case), and a candidate is kept iff its FINE cell -- computed with the reference's own fp64 division,
// cell_add.f:22 -- lies in the reference's 5 x 5 x 5 block.  The neighbour SETS are therefore the reference's.
DMD_DEV uint32_t cpk_pack(int cx, int cy, int cz) { return (uint32_t)cx | ((uint32_t)cy << 10) | ((uint32_t)cz << 20); }
DMD_DEV int coarse_dim(int ncr) { return (ncr + 1) >> 1; }

// the distinct coarse indices of the fine cells c-2 .. c+2 (periodic) -> out[0..n)
DMD_DEV int coarse_span(int c, int ncr, int* out) {
  int n = 0;
#pragma unroll
  for (int d = -2; d <= 2; d++) {
    int f = c + d;
    f = f < 0 ? f + ncr : (f >= ncr ? f - ncr : f);
    const int cc = f >> 1;
    // consecutive fine cells: a repeat is either the previous value or (after wrapping round a small ring) the first
    if (n == 0 || (out[n - 1] != cc && out[0] != cc)) out[n++] = cc;
  }
  return n;
}

DMD_DEV bool in_fine_stencil(uint32_t pk, uint32_t pj, int ncr) {
  int dx = (int)(pk & 1023u) - (int)(pj & 1023u), dy = (int)((pk >> 10) & 1023u) - (int)((pj >> 10) & 1023u),
      dz = (int)(pk >> 20) - (int)(pj >> 20);
  dx = dx < 0 ? -dx : dx; dy = dy < 0 ? -dy : dy; dz = dz < 0 ? -dz : dz;
  dx = dx > ncr - dx ? ncr - dx : dx; dy = dy > ncr - dy ? ncr - dy : dy; dz = dz > ncr - dz ? ncr - dz : dz;
  return dx <= 2 && dy <= 2 && dz <= 2;
}

// f(j) for every bead j != k whose cell lies in the 5 x 5 x 5 block around the cell of bead k; beads with index in
// [skip_lo, skip_hi) are passed over (nbor_build handles the beads of k's own chain without the cell grid)
template <class F>
DMD_DEV void stencil_visit(const Rep& r, int k, F f, int skip_lo = 0, int skip_hi = 0) {
  const int ncr = r.c.sys->ncr, ncc = coarse_dim(ncr);
  const uint32_t pk = r.cpk[k];
  int xs[4], ys[4], zs[4];
  const int nx = coarse_span((int)(pk & 1023u), ncr, xs), ny = coarse_span((int)((pk >> 10) & 1023u), ncr, ys),
            nz = coarse_span((int)(pk >> 20), ncr, zs);
  // ONE visit site, no unrolling: the lanes of a warp walk different cells and chains, and they can only
  // re-converge on the (expensive) body of f if there is a single copy of it
  const int nrow = ny * nz;
#pragma unroll 1
  for (int row = 0; row < nrow; row++) {
    const int iz = row / ny, iy = row - iz * ny;
    const int rowbase = (ys[iy] + zs[iz] * ncc) * ncc;
    int h0 = r.cellhead[rowbase + xs[0]], h1 = -1, h2 = -1, h3 = -1;  // the row's heads: loads in flight together
    if (nx > 1) h1 = r.cellhead[rowbase + xs[1]];
    if (nx > 2) h2 = r.cellhead[rowbase + xs[2]];
    if (nx > 3) h3 = r.cellhead[rowbase + xs[3]];
    int ix = 0, j = h0;
#pragma unroll 1
    while (true) {
      if (j < 0) {  // next chain of the row
        ix++;
        if (ix >= nx) break;
        j = ix == 1 ? h1 : (ix == 2 ? h2 : h3);
        continue;
      }
      if (j != k && (j < skip_lo || j >= skip_hi) && in_fine_stencil(pk, r.cpk[j], ncr)) f(j);
      j = r.cnext[j];
    }
  }
}

// (t0, ts): first bead and stride of the calling thread -- (lane, DMD_W) when one warp owns the replica,
// (thread index, CTA size) when a whole CTA does (dmd_block.h), (global thread, "infinite") in the bulk kernels
DMD_DEV void cell_build(Rep& r, int t0 = Warp::lane(), int ts = DMD_W) {
  const SysConst& s = *r.c.sys;
  const int ncr = s.ncr, nc = s.num_cell, nw = s.n_wrap, ncc = coarse_dim(ncr);
  Warp::sync();
  for (int k = t0; k < r.N; k += ts) {
    int cx, cy, cz;
    (cell_coords)(s, r.rec[k], cx, cy, cz);
    r.cellof[k] = 1 + (cx + nw) + (cy + nw) * nc + (cz + nw) * nc * nc;  // cell_add.f:25
    if (cx < 0 || cy < 0 || cz < 0 || cx >= ncr || cy >= ncr || cz >= ncr) {
      // the reference would file the bead in a ghost cell that is never looked up
      r.cnext[k] = -2;
      r.cpk[k] = 0;
      continue;
    }
    r.cpk[k] = cpk_pack(cx, cy, cz);
    const int cidx = (cx >> 1) + ((cy >> 1) + (cz >> 1) * ncc) * ncc;
    r.cnext[k] = exch(&r.cellhead[cidx], k);
  }
  Warp::sync();
}

DMD_DEV void cell_clear(Rep& r, int t0 = Warp::lane(), int ts = DMD_W) {
  const int ncc = coarse_dim(r.c.sys->ncr);
  Warp::sync();
  for (int k = t0; k < r.N; k += ts) {
    if (r.cnext[k] == -2) continue;
    const uint32_t pk = r.cpk[k];
    r.cellhead[(int)((pk & 1023u) >> 1) + ((int)(((pk >> 10) & 1023u) >> 1) + (int)((pk >> 20) >> 1) * ncc) * ncc] = -1;
  }
  Warp::sync();
}

// nbor.f:33-137: class rule nbor.f:60 (bonded-class pairs are neighbours whenever found in the stencil) /
// distance rule nbor.f:97-105; up list (partners > k) and down list (partners < k) of every bead
DMD_DEV void nbor_build(Rep& r, int t0 = Warp::lane(), int ts = DMD_W) {
  const SysConst& s = *r.c.sys;
  const int cap = r.cap;
  int overflow = 0;
  for (int k = t0; k < r.N; k += ts) {
    const BeadRec rk = r.rec[k];
    const uint32_t mk = r.c.meta[k];
    const int ck = r.c.chain[k];
    // same-chain pairs (most candidates): class from the per-species table instead of the closed form
    const int sct0 = r.c.sctab ? s.sct_off[meta_sp(mk)] : -1;
    const uint8_t* const sct_row = sct0 >= 0 ? r.c.sctab + sct0 + (size_t)meta_local(mk) * s.numbeads[meta_sp(mk)] : nullptr;
    int nu = 0, nd = 0;
    auto append_with_class = [&](int j, int sc) {
      bool in;
      if (code_is_bonded_class(sc)) {
        in = true;  // nbor.f:60
      } else {
        const BeadRec rj = r.rec[j];
        const int code = overlay_code(sc, k, rk, j, rj);
        double rx = rk.x - rj.x, ry = rk.y - rj.y, rz = rk.z - rj.z;  // nbor.f:97-103
        rx = rx - dmd_round(rx);
        ry = ry - dmd_round(ry);
        rz = rz - dmd_round(rz);
        double rijsq = rx * rx + ry * ry + rz * rz;
        in = rijsq <= s.rlsq[code];  // nbor.f:105
      }
      if (in) {
        uint32_t e = ((uint32_t)sc << NB_SHIFT) | (uint32_t)j;
        if (j > k) {
          if (nu < cap) r.up[(size_t)k * cap + nu] = e;
          nu++;
        } else {
          if (nd < cap) r.dn[(size_t)k * cap + nd] = e;
          nd++;
        }
      }
    };
    auto test_and_append = [&](int j) {
      const uint32_t mj = r.c.meta[j];
      const int cj = r.c.chain[j];
      const int sc = (sct_row && cj == ck) ? (int)sct_row[meta_local(mj)] : static_code(s, mk, ck, k, mj, cj, j);
      append_with_class(j, sc);
    };
    if (r.cnext[k] != -2) {
      // (0) short chains: the beads of k's own chain -- most of its neighbours, a contiguous index range with known
      // classes -- are tested directly (same criterion: inside the 5 x 5 x 5 block, then class / distance rule);
      // the lanes of a warp hold beads of the same chain, so these loads are broadcasts and the loop is convergent
      int own_lo = 0, own_hi = 0;
      if (sct_row && s.numbeads[meta_sp(mk)] <= 64) {
        const int nb = s.numbeads[meta_sp(mk)], ncr = s.ncr;
        own_lo = k - meta_local(mk);
        own_hi = own_lo + nb;
        const uint32_t pk = r.cpk[k];
        // entries of the down list go to the front of the row the candidates are parked in: parked ones start behind
        for (int lj = 0; lj < nb; lj++) {
          const int j = own_lo + lj;
          if (j == k || r.cnext[j] == -2 || !in_fine_stencil(pk, r.cpk[j], ncr)) continue;
          append_with_class(j, (int)sct_row[lj]);
        }
      }
      // two steps, so that the lanes of a warp run the expensive test in lockstep: (1) walk the cells and park
      // the candidates in the bead's down-list row behind the entries it already holds, (2) test them one after
      // the other.  Writing entry nd of the row is safe: nd never exceeds the number of slots already consumed.
      uint32_t* park = r.dn + (size_t)k * cap;
      const int p0 = nd;  // down entries of the own chain stay in front
      int nc = 0;
      stencil_visit(r, k, [&](int j) {
        if (p0 + nc < cap) park[p0 + nc] = (uint32_t)j;
        nc++;
      }, own_lo, own_hi);
      if (p0 + nc <= cap) {
#pragma unroll 1
        for (int t = 0; t < nc; t++) test_and_append((int)park[p0 + t]);
      } else {  // more candidates than a row holds (very dense region): test them on the fly
        stencil_visit(r, k, test_and_append, own_lo, own_hi);
      }
    }
    if (nu > cap || nd > cap) {
      overflow = nu > nd ? nu : nd;
      nu = nu > cap ? cap : nu;
      nd = nd > cap ? cap : nd;
    }
    r.nup[k] = (uint16_t)nu;
    r.ndn[k] = (uint16_t)nd;
  }
  unsigned m = Warp::ballot(overflow != 0);
  if (m) set_error(r, DMD_E_NBR_CAP, Warp::shfl(overflow, dmd_ffs(m) - 1));
  Warp::sync();
}

DMD_DEV void nbor(Rep& r) {  // nbor.f:33-137
  cell_build(r);
  nbor_build(r);
  cell_clear(r);
}

#if !defined(DMD_HOST_TRACE)
// ---------------------------------------------------------------------------------------------------------
// Chain-wise list rebuild (service CTAs, small systems: SysConst.chainwise).  The cell walk above makes the lanes
// of a warp chase different linked lists (measured: ~530 warp instructions per bead for ~15 candidates).  Here a
// bead's candidates are the beads of its own chain plus the beads of the chains whose bounding sphere comes within
// the largest list cut-off of its chain's sphere -- loops over contiguous index ranges that the lanes of a warp
// (32 consecutive beads, i.e. one or two chains) run in step.  The neighbour SETS are the reference's:
//   same chain     found in the 5 x 5 x 5 fine-cell block (cell_add.f:22 coordinates), then class rule nbor.f:60 /
//                  distance rule nbor.f:97-105 -- as in nbor_build()
//   other chains   classes 1 / 15 / 16 only (never bonded): distance rule; r <= rl <= 2 cell widths puts the pair
//                  inside the block, the block test is kept as a cheap integer pre-filter.  The cut-off of a class-1
//                  pair does not depend on its 40 / 50 overlay (SysConst.chainwise is set only if rlsq agrees)
// and a pair closer than rl_max has sphere centres closer than R_a + R_b + rl_max for ANY periodic image.
// Entry order (own chain ascending, then chains ascending) is deterministic.
// ---------------------------------------------------------------------------------------------------------
struct ChainBound {
  double cx, cy, cz, rad;
};
constexpr uint32_t CPK_OUT = 0xffffffffu;  // bead outside the cell grid: the reference never looks it up
DMD_DEV int chain_first(const SysConst& s, int c) { return c < s.nch[0] ? c * s.numbeads[0] : s.nop1 + (c - s.nch[0]) * s.numbeads[1]; }
DMD_DEV int chain_len(const SysConst& s, int c) { return c < s.nch[0] ? s.numbeads[0] : s.numbeads[1]; }

// step 1 (thread per bead): reference cell id + packed fine cell coordinates
DMD_DEV void chainwise_cells(Rep& r, uint32_t* cpk, int tid, int nt) {
  const SysConst& s = *r.c.sys;
  const int ncr = s.ncr, nc = s.num_cell, nw = s.n_wrap;
  for (int k = tid; k < r.N; k += nt) {
    int cx, cy, cz;
    (cell_coords)(s, r.rec[k], cx, cy, cz);
    r.cellof[k] = 1 + (cx + nw) + (cy + nw) * nc + (cz + nw) * nc * nc;  // cell_add.f:25
    const bool out = cx < 0 || cy < 0 || cz < 0 || cx >= ncr || cy >= ncr || cz >= ncr;
    cpk[k] = out ? CPK_OUT : cpk_pack(cx, cy, cz);
  }
}
// step 2 (hardware warp per chain): bounding sphere around the mean bead position (minimum image to the first bead)
DMD_DEV void chainwise_bounds(const Rep& r, ChainBound* cb, int nch, int warp, int nwarps) {
  const SysConst& s = *r.c.sys;
  const int hl = threadIdx.x & 31;
  for (int c = warp; c < nch; c += nwarps) {
    const int f = (chain_first)(s, c), nb = (chain_len)(s, c);
    const BeadRec* p0 = &r.rec[f];
    const double x0 = p0->x, y0 = p0->y, z0 = p0->z;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int j = hl; j < nb; j += 32) {
      const BeadRec* p = &r.rec[f + j];
      double dx = p->x - x0, dy = p->y - y0, dz = p->z - z0;
      sx += dx - dmd_round(dx); sy += dy - dmd_round(dy); sz += dz - dmd_round(dz);
    }
    for (int m = 16; m >= 1; m >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, m);
      sy += __shfl_xor_sync(0xffffffffu, sy, m);
      sz += __shfl_xor_sync(0xffffffffu, sz, m);
    }
    const double cx = x0 + sx / nb, cy = y0 + sy / nb, cz = z0 + sz / nb;
    double rad = 0.0;
    for (int j = hl; j < nb; j += 32) {
      const BeadRec* p = &r.rec[f + j];
      double dx = p->x - cx, dy = p->y - cy, dz = p->z - cz;
      dx -= dmd_round(dx); dy -= dmd_round(dy); dz -= dmd_round(dz);
      const double d = dmd_sqrt(dx * dx + dy * dy + dz * dz);
      rad = d > rad ? d : rad;
    }
    for (int m = 16; m >= 1; m >>= 1) {
      const double o = __shfl_xor_sync(0xffffffffu, rad, m);
      rad = o > rad ? o : rad;
    }
    if (hl == 0) {
      ChainBound b;
      b.cx = cx; b.cy = cy; b.cz = cz; b.rad = rad;
      cb[c] = b;
    }
  }
}
// step 3 (thread per ordered chain pair): near[a] bit b <=> the spheres come within the largest list cut-off
DMD_DEV void chainwise_near(const Rep& r, const ChainBound* cb, unsigned* near, int nch, int tid, int nt) {
  const double reach = r.c.sys->rl_max * (1.0 + 1e-9) + 1e-12;
  for (int p = tid; p < nch * nch; p += nt) {
    const int a = p / nch, b = p - a * nch;
    if (a == b) continue;
    const ChainBound A = cb[a], B = cb[b];
    double dx = A.cx - B.cx, dy = A.cy - B.cy, dz = A.cz - B.cz;
    dx -= dmd_round(dx); dy -= dmd_round(dy); dz -= dmd_round(dz);
    const double lim = A.rad + B.rad + reach;
    if (dx * dx + dy * dy + dz * dz <= lim * lim) atomicOr(&near[2 * a + (b >> 5)], 1u << (b & 31));
  }
}
// step 4 (thread per bead; whole hardware warps): the two lists of every bead
DMD_DEV void chainwise_lists(Rep& r, const uint32_t* cpk, const unsigned* near, int tid, int nt) {
  const SysConst& s = *r.c.sys;
  const int cap = r.cap, ncr = s.ncr;
  int overflow = 0;
  for (int base = 0; base < r.N; base += nt) {  // the same trip count for every lane of a hardware warp
    const int k = base + tid;
    const bool valid = k < r.N;
    const int kk = valid ? k : 0;
    const BeadRec rk = r.rec[kk];
    const uint32_t mk = r.c.meta[kk];
    const int ck = r.c.chain[kk];
    const uint32_t pk = cpk[kk];
    const bool live = valid && pk != CPK_OUT;
    int nu = 0, nd = 0;
    auto append = [&](int j, int sc) {
      const uint32_t e = ((uint32_t)sc << NB_SHIFT) | (uint32_t)j;
      if (j > k) {
        if (nu < cap) r.up[(size_t)k * cap + nu] = e;
        nu++;
      } else {
        if (nd < cap) r.dn[(size_t)k * cap + nd] = e;
        nd++;
      }
    };
    // Two steps per chain, so that the lanes stay together and the gathers overlap: (1) the integer block test
    // (cell words from shared memory, no global loads) marks the candidates in a bit mask; (2) the mask is drained four
    // candidates at a time -- four position gathers in flight, then the class / distance rule (nbor.f:60, :97-105 with
    // the cut-off of the static class) and the appends in ascending order.
    auto drain = [&](unsigned long long cand, const int jbase, auto&& classify) {
      while (cand) {
        int jq[4], sq[4];
        double d2[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          jq[q] = cand ? jbase + (__ffsll((long long)cand) - 1) : -1;
          cand &= cand - 1ull;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int jj = jq[q] >= 0 ? jq[q] : kk;  // an always valid address: the loads are unconditional
          sq[q] = classify(jj);
          const BeadRec* pj = &r.rec[jj];
          double rx = rk.x - pj->x, ry = rk.y - pj->y, rz = rk.z - pj->z;
          rx = rx - dmd_round(rx);
          ry = ry - dmd_round(ry);
          rz = rz - dmd_round(rz);
          d2[q] = rx * rx + ry * ry + rz * rz;
        }
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (jq[q] >= 0 && (code_is_bonded_class(sq[q]) || d2[q] <= s.rlsq[sq[q]])) append(jq[q], sq[q]);
      }
    };
    // (a) own chain
    {
      const int sp = meta_sp(mk), nb = s.numbeads[sp];
      const int own_lo = kk - meta_local(mk);
      const uint8_t* const sct_row = r.c.sctab + s.sct_off[sp] + (size_t)meta_local(mk) * nb;
      const int nbmax = s.numbeads[0] > s.numbeads[1] ? s.numbeads[0] : s.numbeads[1];
      unsigned long long cand = 0ull;
#pragma unroll 4
      for (int lj = 0; lj < nbmax; lj++) {
        const int j = own_lo + (lj < nb ? lj : 0);
        const uint32_t pj = cpk[j];
        const bool ok = live && lj < nb && j != k && pj != CPK_OUT && in_fine_stencil(pk, pj, ncr);
        cand |= (unsigned long long)ok << lj;
      }
      drain(cand, own_lo, [&](int j) { return (int)sct_row[j - own_lo]; });
    }
    // (b) the chains near the own one: the union over the hardware warp keeps the loops in step
    unsigned m0 = live ? near[2 * ck] : 0u, m1 = live ? near[2 * ck + 1] : 0u;
    unsigned u0 = __reduce_or_sync(0xffffffffu, m0), u1 = __reduce_or_sync(0xffffffffu, m1);
    for (int half = 0; half < 2; half++) {
      unsigned u = half ? u1 : u0;
      const unsigned mine_mask = half ? m1 : m0;
      while (u) {
        const int bit = __ffs((int)u) - 1;
        u &= u - 1;
        const int c = half * 32 + bit;
        const bool mine = (mine_mask >> bit) & 1u;
        const int f = (chain_first)(s, c), nb = (chain_len)(s, c);
        unsigned long long cand = 0ull;
#pragma unroll 4
        for (int lj = 0; lj < nb; lj++) {
          const uint32_t pj = cpk[f + lj];
          const bool ok = mine && pj != CPK_OUT && in_fine_stencil(pk, pj, ncr);
          cand |= (unsigned long long)ok << lj;
        }
        drain(cand, f, [&](int j) { return static_code(s, mk, ck, k, r.c.meta[j], c, j); });  // other chain: 1, 15 or 16
      }
    }
    if (valid) {
      if (nu > cap || nd > cap) {
        overflow = nu > nd ? nu : nd;
        nu = nu > cap ? cap : nu;
        nd = nd > cap ? cap : nd;
      }
      r.nup[k] = (uint16_t)nu;
      r.ndn[k] = (uint16_t)nd;
    }
  }
  if (overflow) {  // reported through the stored scalars (the service CTA has no replica view of its own)
    if (atomicCAS(&r.sc->error, 0, DMD_E_NBR_CAP) == 0) r.sc->error_info = overflow;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Sorted-grid list rebuild (service CTAs; any density).  The beads are counting-sorted by COARSE cell (2 x 2 x 2 fine
// cells) into a 16-bit index array in shared memory, with 16-bit cell end offsets beside it: a bead's candidates are
// then short contiguous runs of that array -- no linked lists through global memory, no per-candidate chain of
// dependent loads -- filtered by the exact fine-cell test and classified from the same-chain table or the closed
// inter-chain rule.  Same neighbour sets as nbor_build() / nbor.f:33-137; entries ordered by (coarse cell in scan
// order, bead index): deterministic.  Needs N < 65536 and (ncc^3 + 1 + N) 16-bit words of scratch.
// ---------------------------------------------------------------------------------------------------------
struct SortedGrid {
  uint16_t* end;     // ncc3 + 1 entries: end[c] = one past the last slot of coarse cell c (start = end[c - 1], 0 for c = 0)
  uint16_t* sorted;  // N entries: bead indices grouped by coarse cell, ascending inside a cell
  unsigned* tot;     // one word per thread of the group (prefix sum over the cells)
};
DMD_DEV unsigned sg_add16(uint16_t* a, int c, unsigned v) {  // 16-bit atomic add in shared memory; returns the old value
  const unsigned old = atomicAdd(reinterpret_cast<unsigned*>(a) + (c >> 1), v << (16 * (c & 1)));
  return (old >> (16 * (c & 1))) & 0xffffu;
}
// steps 1-4, all threads of the group (gsync = the group's barrier)
template <class Sync>
DMD_DEV void sorted_grid_build(Rep& r, SortedGrid g, int tid, int nt, Sync gsync) {
  const SysConst& s = *r.c.sys;
  const int ncr = s.ncr, nc = s.num_cell, nw = s.n_wrap, ncc = coarse_dim(ncr), ncc3 = ncc * ncc * ncc;
  for (int c = tid; c < (ncc3 + 2) / 2; c += nt) reinterpret_cast<unsigned*>(g.end)[c] = 0u;
  gsync();
  for (int k = tid; k < r.N; k += nt) {  // fine coordinates (cell_add.f:22), reference cell id, coarse cell population
    int cx, cy, cz;
    (cell_coords)(s, r.rec[k], cx, cy, cz);
    r.cellof[k] = 1 + (cx + nw) + (cy + nw) * nc + (cz + nw) * nc * nc;  // cell_add.f:25
    if (cx < 0 || cy < 0 || cz < 0 || cx >= ncr || cy >= ncr || cz >= ncr) {
      r.cpk[k] = CPK_OUT;  // the reference files the bead in a ghost cell that is never looked up
      continue;
    }
    r.cpk[k] = cpk_pack(cx, cy, cz);
    sg_add16(g.end, (cx >> 1) + ((cy >> 1) + (cz >> 1) * ncc) * ncc, 1u);
  }
  gsync();
  // exclusive prefix sum in place: a contiguous slice per thread, the slice totals scanned by thread 0, then every
  // slice adds its base
  const int per = (ncc3 + nt - 1) / nt;
  const int c0 = tid * per, c1 = c0 + per < ncc3 ? c0 + per : ncc3;
  unsigned sum = 0;
  for (int c = c0; c < c1; c++) sum += g.end[c];
  unsigned* const tot = g.tot;
  tot[tid] = sum;
  gsync();
  if (tid == 0) {
    unsigned run = 0;
    for (int t = 0; t < nt; t++) {
      const unsigned v = tot[t];
      tot[t] = run;
      run += v;
    }
  }
  gsync();
  unsigned run = tot[tid];
  gsync();
  for (int c = c0; c < c1; c++) {
    const unsigned v = g.end[c];
    g.end[c] = (uint16_t)run;  // start of the cell; the scatter below advances it to the cell's end
    run += v;
  }
  gsync();
  for (int k = tid; k < r.N; k += nt) {
    const uint32_t pk = r.cpk[k];
    if (pk == CPK_OUT) continue;
    const int c = (int)((pk & 1023u) >> 1) + ((int)(((pk >> 10) & 1023u) >> 1) + (int)((pk >> 20) >> 1) * ncc) * ncc;
    g.sorted[sg_add16(g.end, c, 1u)] = (uint16_t)k;
  }
  gsync();
  for (int c = tid; c < ncc3; c += nt) {  // ascending bead index inside every cell: the scatter order is arbitrary
    const int a = c ? g.end[c - 1] : 0, b = g.end[c];
    for (int i = a + 1; i < b; i++) {
      const uint16_t v = g.sorted[i];
      int j = i - 1;
      while (j >= a && g.sorted[j] > v) {
        g.sorted[j + 1] = g.sorted[j];
        j--;
      }
      g.sorted[j + 1] = v;
    }
  }
  gsync();
}
// step 5 (one HARDWARE warp per bead, lanes over its candidates): both lists of every bead.  The coarse cells of the
// bead's stencil are taken by the lanes (one cell each), their populations prefix-summed over the warp, and the
// candidates -- the concatenation of the cells' runs of sorted[] -- tested 32 at a time: cell word, fine-cell filter,
// class, distance, all lanes in flight together; the survivors are appended in candidate order (ballot + popc).
DMD_DEV void sorted_grid_lists(Rep& r, SortedGrid g, int warp, int nwarps) {
  const SysConst& s = *r.c.sys;
  const int cap = r.cap, ncr = s.ncr, ncc = coarse_dim(ncr);
  const int lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  int overflow = 0;
  for (int k = warp; k < r.N; k += nwarps) {  // warp-uniform
    const uint32_t pk = r.cpk[k];
    int nu = 0, nd = 0;
    if (pk != CPK_OUT) {
      const BeadRec* pkr = &r.rec[k];
      const double xk = pkr->x, yk = pkr->y, zk = pkr->z;
      const uint32_t mk = r.c.meta[k];
      const int ck = r.c.chain[k];
      const int sp = meta_sp(mk), nb = s.numbeads[sp];
      const int own_lo = k - meta_local(mk);
      const uint8_t* const sct_row = r.c.sctab + s.sct_off[sp] + (size_t)meta_local(mk) * nb;
      int xs[4], ys[4], zs[4];
      const int nx = coarse_span((int)(pk & 1023u), ncr, xs), ny = coarse_span((int)((pk >> 10) & 1023u), ncr, ys),
                nz = coarse_span((int)(pk >> 20), ncr, zs);
      const int ncell = nx * ny * nz;  // <= 64
      for (int q0 = 0; q0 < ncell; q0 += 32) {
        // ---- one coarse cell per lane: its run [a, a + cnt) of sorted[]
        const int q = q0 + lane;
        int a = 0, cnt = 0;
        if (q < ncell) {
          const int iz = q / (nx * ny), rem = q - iz * nx * ny, iy = rem / nx, ix = rem - iy * nx;
          const int c = xs[ix] + (ys[iy] + zs[iz] * ncc) * ncc;
          a = c ? g.end[c - 1] : 0;
          cnt = (int)g.end[c] - a;
        }
        int incl = cnt;  // inclusive prefix sum of the populations over the lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int o = __shfl_up_sync(FULL, incl, d);
          if (lane >= d) incl += o;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        const int excl = incl - cnt;
        for (int t0 = 0; t0 < total; t0 += 32) {
          const int idx = t0 + lane;
          // the lane whose run holds candidate idx: the first lane with incl > idx (binary search over the warp)
          int lo = 0;
#pragma unroll
          for (int step = 16; step >= 1; step >>= 1) {
            const int probe = __shfl_sync(FULL, incl, lo + step - 1);
            if (probe <= idx) lo += step;
          }
          const int src = lo & 31;
          const int a_src = __shfl_sync(FULL, a, src), excl_src = __shfl_sync(FULL, excl, src);
          bool in = false;
          int j = -1, sc = 1;
          if (idx < total) {
            j = g.sorted[a_src + (idx - excl_src)];
            if (j != k && in_fine_stencil(pk, r.cpk[j], ncr)) {
              if (j >= own_lo && j < own_lo + nb) {
                sc = (int)sct_row[j - own_lo];
                in = code_is_bonded_class(sc);  // nbor.f:60
              } else {
                sc = static_code(s, mk, ck, k, r.c.meta[j], ck + 1, j);  // another chain: 1, 15 or 16
              }
              if (!in) {  // nbor.f:97-105 (the cut-off of class 1 also holds for its 40 / 50 overlay: SysConst.chainwise)
                const BeadRec* pj = &r.rec[j];
                double rx = xk - pj->x, ry = yk - pj->y, rz = zk - pj->z;
                rx = rx - dmd_round(rx);
                ry = ry - dmd_round(ry);
                rz = rz - dmd_round(rz);
                in = rx * rx + ry * ry + rz * rz <= s.rlsq[sc];
              }
            }
          }
          const unsigned mu = __ballot_sync(FULL, in && j > k), md = __ballot_sync(FULL, in && j < k);
          const unsigned below = (1u << lane) - 1u;
          if (in) {
            const uint32_t e = ((uint32_t)sc << NB_SHIFT) | (uint32_t)j;
            if (j > k) {
              const int pos = nu + __popc(mu & below);
              if (pos < cap) r.up[(size_t)k * cap + pos] = e;
            } else {
              const int pos = nd + __popc(md & below);
              if (pos < cap) r.dn[(size_t)k * cap + pos] = e;
            }
          }
          nu += __popc(mu);
          nd += __popc(md);
        }
      }
    }
    if (nu > cap || nd > cap) {
      overflow = nu > nd ? nu : nd;
      nu = nu > cap ? cap : nu;
      nd = nd > cap ? cap : nd;
    }
    if (lane == 0) {
      r.nup[k] = (uint16_t)nu;
      r.ndn[k] = (uint16_t)nd;
    }
  }
  if (overflow && lane == 0) {
    if (atomicCAS(&r.sc->error, 0, DMD_E_NBR_CAP) == 0) r.sc->error_info = overflow;
  }
}

// ---- list-rebuild service (device only).  A warp whose replica needs nbor() + events() publishes the request
// in its svc word, queues its index in the ticket ring behind the service counters (dmd_types.h: SVC_Q_*) and sleeps;
// a free group of a service CTA on another SM claims the oldest ticket and the word (1 -> 2), rebuilds with all its
// threads and clears the word.  release/acquire at gpu scope on the word orders the replica's arrays between the two SMs (the
// acquire also drops the stale L1 lines).  A request nobody claims within SVC_PATIENCE cycles is taken back
// (1 -> 3) and served in place, so the loop never depends on a service CTA being resident.
#ifndef DMD_SVC_PATIENCE
#define DMD_SVC_PATIENCE 80000000  // ~40 ms: a rebuild done in place drags the rebuild code through the event-loop SM's
                                   // instruction cache and stalls the warp's other replica (measured: 2 ms -> 40 ms +15 %)
#endif
constexpr long long SVC_PATIENCE = DMD_SVC_PATIENCE;
constexpr long long SVC_TIMEOUT = 4000000000ll;  // ~2 s in state 2: report an error instead of hanging
DMD_DEV int svc_ld_relaxed(const int32_t* p) {  // polling: no L1 invalidation (the other warps of the SM keep their lines)
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
DMD_DEV void svc_fence_acquire() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
DMD_DEV void svc_st_release(int32_t* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
DMD_DEV void svc_st_release64(unsigned long long* p, unsigned long long v) { asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
DMD_DEV unsigned long long svc_ld_acquire64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
DMD_DEV unsigned long long svc_ld_relaxed64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
DMD_DEV int svc_cas_acq_rel(int32_t* p, int cmp, int val) {
  int old;
  asm volatile("atom.acq_rel.gpu.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
  return old;
}
// returns true when a service CTA has rebuilt lists and calendar, false when the warp has to do it in place
DMD_COLD bool svc_request(Rep r) {
  __threadfence();  // every lane's writes (wrapped positions, oldr, scalars) before the request becomes visible
  Warp::sync();
  int res = 0;  // 0 served, 1 in place, 2 time-out
  if (Warp::lane() == 0) {
    svc_st_release(r.svc, 1);
    {  // queue the request: ticket, then the tagged slot (release: whoever reads the slot also sees the word above)
      unsigned long long* const q = r.svc_ctl;
      const unsigned long long tk = atomicAdd(&q[SVC_Q_TAIL], 1ull);
      svc_st_release64(&q[SVC_Q_RING + tk % q[SVC_Q_CAP]], ((tk + 1ull) << 24) | (unsigned long long)(unsigned)r.svc_id);
    }
    const long long t0 = clock64();
    while (true) {
      const int v = svc_ld_relaxed(r.svc);
      if (v == 0) {
        svc_fence_acquire();  // once: orders the service CTA's writes before this warp's reads, drops stale L1 lines
        break;
      }
      const long long dt = clock64() - t0;
      if (v == 1 && dt > SVC_PATIENCE && svc_cas_acq_rel(r.svc, 1, 3) == 1) {
        res = 1;
        break;
      }
      if (dt > SVC_TIMEOUT) {
        res = 2;
        break;
      }
      __nanosleep(1500);
    }
    if (r.svc_ctl) {
      if (res == 1) atomicAdd(&r.svc_ctl[2], 1ull);
      atomicAdd(&r.svc_ctl[3], (unsigned long long)(clock64() - t0));
    }
  }
  Warp::sync();  // a memory-ordering point for all lanes: they read the arrays the service CTA wrote after lane 0's acquire
  res = Warp::shfl(res, 0);
  if (res == 2) {  // time-out while being served: the service CTA may still be writing -- report it, touch nothing else
    if (Warp::lane() == 0) {
      r.sc->error = DMD_E_SERVICE;
      r.sc->error_info = 0;
    }
    Warp::sync();
  }
  return res != 1;
}
#endif

// ---- cold pseudo-events: they work on a by-value copy of the view and hand the scalars back through r.sc
// main.F90:997-1048; returns the bead that got its new velocity -- the caller re-predicts it (:1049) with the hot
// loop's own copy of partial_events, so that a ghost event does not drag a second copy through the instruction cache
DMD_COLD int ghost_event_cold(Rep r, double prev_tfalse) {
  const int N = r.N;
  int i;
  do {
    i = (int)(rng_uniform(r.seed, r.ctr) * N);
  } while (i == N);
  BeadRec b = r.rec[i];
  const double bmi = r.c.hot->bmass[b.ident];
  b.x = b.x + b.vx * r.tfalse;
  b.y = b.y + b.vy * r.tfalse;
  b.z = b.z + b.vz * r.tfalse;
  double v1, v2, rr, fact;
  do {
    v1 = 2.0 * rng_uniform(r.seed, r.ctr) - 1.0;
    v2 = 2.0 * rng_uniform(r.seed, r.ctr) - 1.0;
    rr = v1 * v1 + v2 * v2;
  } while (rr == 0.0 || rr >= 1.0);
  fact = dmd_sqrt(-2.0 * r.setemp * bmi * dmd_log(rr) / rr);
  b.vx = v1 * fact / bmi;
  b.vy = v2 * fact / bmi;
  do {
    v1 = 2.0 * rng_uniform(r.seed, r.ctr) - 1.0;
    v2 = 2.0 * rng_uniform(r.seed, r.ctr) - 1.0;
    rr = v1 * v1 + v2 * v2;
  } while (rr == 0.0 || rr >= 1.0);
  fact = dmd_sqrt(-2.0 * r.setemp * bmi * dmd_log(rr) / rr);
  b.vz = v1 * fact / bmi;
  b.x = b.x - b.vx * r.tfalse;
  b.y = b.y - b.vy * r.tfalse;
  b.z = b.z - b.vz * r.tfalse;
  double tgho = 0.0;
  while (tgho < 1e-18 || tgho == 1.0) tgho = rng_uniform(r.seed, r.ctr);
  const double tnext = -1.0 * dmd_log(tgho) * r.avegtime + r.tfalse;
  Warp::sync();
  if (Warp::lane() == 0) {
    BeadRec* p = &r.rec[i];
    p->x = b.x; p->y = b.y; p->z = b.z; p->vx = b.vx; p->vy = b.vy; p->vz = b.vz;
    r.cal[N].t = tnext;
    r.sc->numghosts += 1;
  }
  if (r.tfalse < prev_tfalse) r.tfalse = prev_tfalse;  // main.F90:1047 (old_tfalse = time of the previous event)
  log_event(r, N, i, -2, 0);
  rep_save(r);
  return i;
}

// main.F90:1126-1187
DMD_COLD void interval_event_cold(Rep r) {
  const SysConst& s = *r.c.sys;
  const int N = r.N;
  const double tf = r.tfalse;
  r.t = r.t + tf;
  Warp::sync();
  for (int k = Warp::lane(); k < N + 3; k += DMD_W) r.cal[k].t = r.cal[k].t - tf;  // :1133-1135
  r.interval_max = r.interval_max - tf;
  double moved_far = 0.0;
  for (int k = Warp::lane(); k < N; k += DMD_W) {  // :1140-1144 + displ.f:20-33
    BeadRec* p = &r.rec[k];
    double x = p->x + p->vx * tf, y = p->y + p->vy * tf, z = p->z + p->vz * tf;
    p->x = x; p->y = y; p->z = z;
    double a = r.oldr[3 * k] - x, b = r.oldr[3 * k + 1] - y, cc = r.oldr[3 * k + 2] - z;
    double dis = a * a + b * b + cc * cc;
    double moved = dis / s.hdelr;
    if (moved > moved_far) moved_far = moved;
  }
  moved_far = warp_max(moved_far);
  r.tfalse = 0.0;
  bool update = false;
  if (moved_far >= 0.1) {  // displ.f:37-46
    update = true;
    if (moved_far >= 1.25 * 1.25) {
      r.t_fact = r.t_fact / 1.01;
      r.interval = r.t_fact / dmd_sqrt(r.setemp);
    }
  }
  Warp::sync();
  if (update || r.interval > r.interval_max) {  // :1150-1179
    if (!update) {
      if (Warp::lane() == 0) r.sc->nforcedupdate += 1;
      r.n_forced = r.n_forced * 1.01;
    }
    r.interval_max = r.interval * r.n_forced;
    if (Warp::lane() == 0) r.sc->nupdates += 1;
    for (int k = Warp::lane(); k < N; k += DMD_W) {
      BeadRec* p = &r.rec[k];
      double x = p->x - dmd_round(p->x), y = p->y - dmd_round(p->y), z = p->z - dmd_round(p->z);
      p->x = x; p->y = y; p->z = z;
      r.oldr[3 * k] = x; r.oldr[3 * k + 1] = y; r.oldr[3 * k + 2] = z;
    }
    Warp::sync();
    bool in_place = true;
#if !defined(DMD_HOST_TRACE)
    if (r.svc) {  // hand nbor() + events() to a service CTA (another SM); the view's scalars travel through r.sc
      rep_save(r);
      DMD_PROF_MARK(r, 6);
      in_place = !svc_request(r);
      DMD_PROF_MARK(r, 11);  // waiting for the list-rebuild service
      r.error = r.sc->error;  // a service CTA reports list overflow etc. through the stored scalars
      r.error_info = r.sc->error_info;
    }
#endif
    if (in_place) {
      nbor(r);
      predict_all(r);  // events(); every bead's (tim, nptnr, coltype) is re-derived from interval_max+ltstep
#if !defined(DMD_HOST_TRACE)
      if (r.svc && Warp::lane() == 0) *r.svc = 0;  // taken back (state 3): nobody else touches the word
#endif
    }
  }
  if (Warp::lane() == 0) r.cal[N + 1].t = r.interval * 0.999;  // :1181
  Warp::sync();
  rebuild_all_groups(r);
  log_event(r, N + 1, -1, -2, 0);
  rep_save(r);
}

// energy.f:25-101 on the neighbour lists instead of the O(N^2) ev_code scan (rl(16) >= every well diameter
// and list validity is policed by displ, so every pair inside its well is on an up-list)
DMD_DEV void energy_of(Rep& r, OutRec& o) {
  const SysConst& s = *r.c.sys;
  int hb_ii = 0, hb_ij = 0, hb_alpha = 0;
  double ehh_ii = 0.0, ehh_ij = 0.0, sumvel = 0.0;
  Warp::sync();
  for (int k = Warp::lane(); k < r.N; k += DMD_W) {
    const BeadRec rk = r.rec[k];
    const uint32_t mk = r.c.meta[k];
    if (rk.bptnr > k) {
      if (r.c.chain[k] == r.c.chain[rk.bptnr]) hb_ii++; else hb_ij++;
    }
    // energy.f:60-72: N_r (r >= 5) bonded to C_{r-4} of the same chain
    if (meta_cls(mk) == 1 && meta_res(mk) >= 5 && rk.bptnr == k + s.chnln[meta_sp(mk)] - 4) hb_alpha++;
    const int nu = r.nup[k];
    for (int p = 0; p < nu; p++) {
      uint32_t e = r.up[(size_t)k * r.cap + p];
      if ((int)(e >> NB_SHIFT) != 16) continue;
      int j = (int)(e & NB_MASK);
      const BeadRec rj = r.rec[j];
      Geom g = pair_geom(rk, rj, r.tfalse);
      double rijsq = g.rx * g.rx + g.ry * g.ry + g.rz * g.rz;
      if (rijsq <= r.c.tab->welldia_sq[tix(rk.ident, rj.ident)]) {
        double ep = r.c.sys->ep_sqrt[tix(rk.ident, rj.ident)];
        if (r.c.chain[k] == r.c.chain[j]) ehh_ii = ehh_ii + ep; else ehh_ij = ehh_ij + ep;
      }
    }
    sumvel = sumvel + s.bmass[rk.ident] * (rk.vx * rk.vx + rk.vy * rk.vy + rk.vz * rk.vz);
  }
#if DMD_W > 1
  for (int m = DMD_W / 2; m >= 1; m >>= 1) {  // fixed-order tree: deterministic
    ehh_ii += Warp::shfl_xor(ehh_ii, m);
    ehh_ij += Warp::shfl_xor(ehh_ij, m);
    sumvel += Warp::shfl_xor(sumvel, m);
  }
#endif
  hb_ii = warp_sum(hb_ii);
  hb_ij = warp_sum(hb_ij);
  hb_alpha = warp_sum(hb_alpha);
  double sumeps = -((hb_ii + hb_ij) * s.eps_hb + (ehh_ii + ehh_ij));
  o.coll = r.coll;
  o.t = r.t + r.tfalse;
  o.ered = 0.5 * sumvel + sumeps;
  o.tred = sumvel / 3.0 / (double)r.N;
  o.sumvel = sumvel;
  o.ehh_ii = ehh_ii;
  o.ehh_ij = ehh_ij;
  o.hb_alpha = hb_alpha;
  o.hb_ii = hb_ii;
  o.hb_ij = hb_ij;
  o.pad = 0;
}

// main.F90:1191-1246
DMD_COLD void output_event_cold(Rep r) {
  const int N = r.N;
  OutRec o;
  energy_of(r, o);
  if (r.n_out < r.c.sys->out_cap) {
    if (Warp::lane() == 0) r.out[r.n_out] = o;
    r.n_out++;
  }
  if (Warp::lane() == 0) r.cal[N + 2].t = 3.3 / (dmd_sqrt(r.setemp)) + 5 + r.tfalse;
  mark_dirty(r, (N + 2) >> 5);
  log_event(r, N + 2, -1, -2, 0);
  flush_dirty(r);
  rep_save(r);
}

// one iteration of main.F90:484-1258 with serial semantics
// process calendar entry o (already popped): main.F90:639-1246
DMD_DEV void process_one(Rep& r, int o, const CalEnt& ev) {
  const double prev_tfalse = r.tfalse;  // old_tfalse of main.F90: between two events it equals tfalse, so it needs no register
  r.tfalse = ev.t;
  r.coll += 1;
  int pi = o, pj = ev.ptnr;  // the bead(s) whose events have to be re-predicted (partial_events.f)
  bool xpulse_del = false, redo = true;
  if (o < r.N) {
    xpulse_del = pair_event(r, o, ev);
    DMD_PROF_MARK(r, 2);
  } else {
    rep_save(r);  // hand the scalars to the out-of-line handler through r.sc ...
    if (o == r.N) {
      pi = ghost_event_cold(r, prev_tfalse);
      pj = -1;
    } else {
      redo = false;
      if (o == r.N + 1) interval_event_cold(r);
      else output_event_cold(r);
    }
    Warp::sync();
    rep_load_scalars(r);  // ... and take them back
    clear_dirty(r);
    Warp::sync();
    if (redo) mark_dirty(r, r.N >> 5);  // the next ghost time
    DMD_PROF_MARK(r, 6);
  }
  if (redo) partial_events(r, pi, pj, xpulse_del);  // main.F90:943, :1049 -- the only call site in the loop
}

// returns the owner index of the processed calendar entry, or -1 on error
DMD_DEV int step(Rep& r) {
  DMD_PROF_MARK(r, 7);
  flush_dirty(r);
  DMD_PROF_MARK(r, 0);
  CalEnt ev;
  const int o = pop_min(r, ev);
  DMD_PROF_MARK(r, 1);
  if (o < 0) {
    set_error(r, DMD_E_CAL_EMPTY, 0);
    return -1;
  }
  process_one(r, o, ev);
  return r.error == 0 ? o : -1;
}

// stop_at_output: return right after the output pseudo-event (main.F90:1191-1246) has been processed, so that the
// host can write the .energy line and the .config / .bptnr / .lastvel records exactly where the reference does
DMD_DEV void run_events(Rep& r, int64_t n_events, bool stop_at_output) {
  const int64_t coll_end = r.coll + n_events;  // coll counts every processed calendar entry (main.F90:639)
  while (r.coll < coll_end) {
    const int o = step(r);
    if (o < 0 || (stop_at_output && o == r.N + 2)) break;
  }
  flush_dirty(r);
}

// Replica-exchange temperature change on resident state (new functionality): advance to true
// positions, rescale velocities by sqrt(T_new/T_old), reset the time constants of main.F90:143-156 for the new
// temperature and re-derive the calendar.  The neighbour lists are KEPT: positions do not change, so the lists stay
// valid -- the positions they were built at (oldr) stay the reference of the displacement test (displ.f), and the
// positions are not wrapped here for the same reason (the interval event wraps when it rebuilds).  Rebuilding the
// lists of every swapped replica at the same moment also made their next rebuild requests arrive in one burst.
// H-bond state (bptnr, identity, extra_repuls, overlay) is kept.
DMD_DEV void retemp(Rep& r, double tstar_new) {
  const int N = r.N;
  const double tf = r.tfalse;
  const double setemp_new = tstar_new * 12.0;
  const double scale = dmd_sqrt(setemp_new / r.setemp);
  r.t = r.t + tf;
  Warp::sync();
  for (int k = Warp::lane(); k < N; k += DMD_W) {
    BeadRec* p = &r.rec[k];
    const double x = p->x + p->vx * tf, y = p->y + p->vy * tf, z = p->z + p->vz * tf;
    p->x = x; p->y = y; p->z = z;
    p->vx = p->vx * scale; p->vy = p->vy * scale; p->vz = p->vz * scale;
  }
  r.tfalse = 0.0;
  r.old_tfalse = 0.0;
  r.setemp = setemp_new;
  r.t_fact = 0.00005;
  r.n_forced = 150.0;
  r.interval = r.t_fact / dmd_sqrt(r.setemp);
  r.interval_max = r.n_forced * r.interval;
  r.avegtime = 0.00005 / dmd_sqrt(r.setemp);
  double tg = 1000000000.0;
  if (r.c.sys->canon) {
    double tgho = 0.0;
    while (tgho < 1e-18 || tgho == 1.0) tgho = rng_uniform(r.seed, r.ctr);
    tg = -1.0 * dmd_log(tgho) * r.avegtime;
  }
  Warp::sync();
  if (Warp::lane() == 0) {
    r.cal[N].t = tg;
    r.cal[N + 1].t = r.interval;
    r.cal[N + 2].t = 3.3 / (dmd_sqrt(r.setemp)) + 5;
  }
  Warp::sync();
  predict_all(r);
}

// ---------------------------------------------------------------------------------------------------------
// beta-sheet observables of one replica from its resident state, the definitions of the reference's
// post-processing program results/r/fibril_list_assign.f (one peptide species, as there):
//   hb_contact(a,b)  inter-chain backbone H-bonds of peptides a, b: bptnr over the N and C beads of each chain except
//                    the first N and the last C (:51-60)
//   sheet partners   hb_contact(a,b) >= chnln/2 + 1 (:89); a sheet = a connected component of that relation (:228)
// out[0] inter-chain H-bonds, [1] sheet-partner pairs, [2] sheets (>= 2 peptides), [3] largest sheet, [4] peptides in
// sheets, [5] intra-chain H-bonds, [6..7] 0.  hbm: nc x nc bytes of scratch, lab: nc ints (shared memory on the device).
// ---------------------------------------------------------------------------------------------------------
DMD_DEV void sheet_observables(const Rep& r, int32_t* out, uint8_t* hbm, int32_t* lab) {
  const SysConst& s = *r.c.sys;
  const int L = s.chnln[0], nbd = s.numbeads[0];
  const int nc = r.N / nbd;
  Warp::sync();
  for (int k = Warp::lane(); k < nc * nc; k += DMD_W) hbm[k] = 0;
  for (int k = Warp::lane(); k < nc; k += DMD_W) lab[k] = k;
  Warp::sync();
  int hb_inter = 0, hb_intra = 0;
  for (int aa = Warp::lane(); aa < r.N; aa += DMD_W) {
    const int bb = r.rec[aa].bptnr;
    if (bb <= aa) continue;  // every bond once (bptnr is symmetric)
    const int ca = aa / nbd, cb = bb / nbd;
    if (ca == cb) {
      hb_intra++;
      continue;
    }
    hb_inter++;
    const int la = aa - ca * nbd, lb = bb - cb * nbd;  // 0-based index in the chain; inner = 1-based L+2 .. 3L-1
    if (la >= L + 1 && la <= 3 * L - 2 && lb >= L + 1 && lb <= 3 * L - 2) {
#if DMD_W > 1
      // byte counters in shared memory: one 32-bit atomic add on the containing word
      const int idx = ca * nc + cb;
      atomicAdd(reinterpret_cast<unsigned*>(hbm) + (idx >> 2), 1u << (8 * (idx & 3)));
#else
      hbm[ca * nc + cb]++;
#endif
    }
  }
  hb_inter = warp_sum(hb_inter);
  hb_intra = warp_sum(hb_intra);
  Warp::sync();
  const int need = L / 2 + 1;
  int dimers = 0;
  for (int k = Warp::lane(); k < nc * nc; k += DMD_W) {
    const int a = k / nc, b = k - a * nc;
    if (a < b && hbm[a * nc + b] + hbm[b * nc + a] >= need) dimers++;
  }
  dimers = warp_sum(dimers);
  // connected components by label propagation (labels only decrease; at most nc sweeps)
  for (int sweep = 0; sweep < nc; sweep++) {
    bool changed = false;
    for (int a = Warp::lane(); a < nc; a += DMD_W) {
      int m = lab[a];
      for (int b = 0; b < nc; b++)
        if (b != a && hbm[a * nc + b] + hbm[b * nc + a] >= need && lab[b] < m) m = lab[b];
      if (m < lab[a]) {
        lab[a] = m;
        changed = true;
      }
    }
    Warp::sync();
    if (!Warp::any(changed)) break;
  }
  int sheets = 0, largest = 0, in_sheets = 0;
  for (int a = Warp::lane(); a < nc; a += DMD_W) {
    if (lab[a] != a) continue;  // a is the root of its component
    int size = 0;
    for (int b = 0; b < nc; b++) size += lab[b] == a ? 1 : 0;
    if (size >= 2) {
      sheets++;
      in_sheets += size;
      largest = size > largest ? size : largest;
    }
  }
  sheets = warp_sum(sheets);
  in_sheets = warp_sum(in_sheets);
  largest = warp_max_u(largest);
  if (Warp::lane() == 0) {
    out[0] = hb_inter; out[1] = dimers; out[2] = sheets; out[3] = largest; out[4] = in_sheets; out[5] = hb_intra;
    out[6] = out[7] = 0;
  }
  Warp::sync();
}

// main.F90:1288-1295 (then the state must be re-initialised by the host before running on)
DMD_DEV void sync_positions(Rep& r) {
  for (int k = Warp::lane(); k < r.N; k += DMD_W) {
    BeadRec* p = &r.rec[k];
    double x = p->x + p->vx * r.tfalse, y = p->y + p->vy * r.tfalse, z = p->z + p->vz * r.tfalse;
    p->x = x - dmd_round(x);
    p->y = y - dmd_round(y);
    p->z = z - dmd_round(z);
  }
  Warp::sync();
}

// ---------------------------------------------------------------------------------------------------------
// run start on the device (dmdb_set_state / dmdb_set_state_all): wrap (inputinfo.f:89-91, main.F90:206-208),
// time constants (main.F90:143-156), per-bead reset (:205-234), restart fix-up from bptnr (:249-321) and the
// pseudo-event times (:408-423; the ghost time is drawn by the start kernel).  Writes every per-replica array
// the event loop reads except the neighbour lists, which nbor() builds next.
//   sv      N x 6 (x,y,z,vx,vy,vz per bead, box units) as uploaded;  bptnr1  N 1-based partners or nullptr
//   nc      ascending indices of the N and C beads (the only H-bond capable ones), n_nc of them
// The fix-up is order dependent (identity changes made on the way affect later tests, and a later repuls_add
// overwrites an earlier one), so it is done in two steps: all lanes collect the pairs the literal k < k_j double
// loop can act on -- N-C pairs inside their well with ev_code 15 and non-terminal beads, and bonded pairs -- into
// scratch (the not yet built up-list array), then they are ranked lexicographically and lane 0 replays them.
// ---------------------------------------------------------------------------------------------------------
// per-bead part of the run start: wrap, record, er34, oldr; returns false on a bad bptnr entry
DMD_DEV bool init_bead(Rep& r, int k, const double* sv, const int32_t* bptnr1) {
  BeadRec b;
  double x = sv[6 * (size_t)k], y = sv[6 * (size_t)k + 1], z = sv[6 * (size_t)k + 2];
  x = x - dmd_round(x); y = y - dmd_round(y); z = z - dmd_round(z);  // inputinfo.f:89-91
  x = x - dmd_round(x); y = y - dmd_round(y); z = z - dmd_round(z);  // main.F90:206-208
  b.x = x; b.y = y; b.z = z;
  b.vx = sv[6 * (size_t)k + 3]; b.vy = sv[6 * (size_t)k + 4]; b.vz = sv[6 * (size_t)k + 5];
  int bp = bptnr1 ? bptnr1[k] - 1 : -1;
  const bool ok = bp >= -1 && bp < r.N;
  b.bptnr = ok ? bp : -1;
  b.er1 = b.er2 = -1;
  b.ident = (uint8_t)meta_id0(r.c.meta[k]);
  b.ov1 = b.ov2 = 1;
  b.pad = 0;
  r.rec[k] = b;
  r.er34[2 * k] = -1;
  r.er34[2 * k + 1] = -1;
  r.oldr[3 * k] = x; r.oldr[3 * k + 1] = y; r.oldr[3 * k + 2] = z;
  return ok;
}

// does the literal k < k_j loop of main.F90:249-321 act on the pair (k, kj)?  Either the geometric test of
// :262-283 (N-C pair with identities 1 + 4 inside its well, ev_code 15, both beads non-terminal) or a bonded pair
DMD_DEV bool fixup_pair_hit(const Rep& r, int k, const BeadRec& a, uint32_t mk, int ck, int kj) {
  if (kj == a.bptnr) return true;
  const BeadRec b = r.rec[kj];
  if (a.ident + b.ident != 5) return false;
  const SysConst& s = *r.c.sys;
  double rx = a.x - b.x, ry = a.y - b.y, rz = a.z - b.z;
  rx = rx - dmd_round(rx); ry = ry - dmd_round(ry); rz = rz - dmd_round(rz);
  const double rijsq = rx * rx + ry * ry + rz * rz;
  const double diff = rijsq - r.c.tab->welldia_sq[tix(a.ident, b.ident)];
  const uint32_t mj = r.c.meta[kj];
  return diff < 0.0 && static_code(s, mk, ck, k, mj, r.c.chain[kj], kj) == 15 && !is_terminal_bead(s, mk) &&
         !is_terminal_bead(s, mj);
}

// the fix-up is order dependent (identity changes made on the way affect later tests, and a later repuls_add
// overwrites an earlier one): rank the collected pairs lexicographically, then lane 0 replays the loop body
DMD_DEV void fixup_replay(Rep& r, int32_t* hits, int nh, int hit_cap) {
  Warp::sync();
  int32_t* sorted = hits + 2 * (size_t)hit_cap;
  for (int h = Warp::lane(); h < nh; h += DMD_W) {
    const int k = hits[2 * h], kj = hits[2 * h + 1];
    int rank = 0;
    for (int m = 0; m < nh; m++) {
      const int k2 = hits[2 * m], kj2 = hits[2 * m + 1];
      if (k2 < k || (k2 == k && kj2 < kj)) rank++;
    }
    sorted[2 * rank] = k;
    sorted[2 * rank + 1] = kj;
  }
  Warp::sync();
  if (Warp::lane() == 0) {
    const SysConst& s = *r.c.sys;
    for (int h = 0; h < nh; h++) {
      const int k = sorted[2 * h], kj = sorted[2 * h + 1];
      BeadRec* a = &r.rec[k];
      BeadRec* b = &r.rec[kj];
      if (a->ident + b->ident == 5) {
        double rx = a->x - b->x, ry = a->y - b->y, rz = a->z - b->z;
        rx = rx - dmd_round(rx); ry = ry - dmd_round(ry); rz = rz - dmd_round(rz);
        const double rijsq = rx * rx + ry * ry + rz * rz;
        const double diff = rijsq - r.c.tab->welldia_sq[tix(a->ident, b->ident)];
        const uint32_t mk = r.c.meta[k], mj = r.c.meta[kj];
        if (diff < 0.0 && static_code(s, mk, r.c.chain[k], k, mj, r.c.chain[kj], kj) == 15 && !is_terminal_bead(s, mk) &&
            !is_terminal_bead(s, mj)) {
          if (a->ident == 1) repuls_add(r, k, kj);
          else repuls_add(r, kj, k);
        }
      }
      if (kj == a->bptnr) {
        if (a->ident == 1) { a->ident = 5; b->ident = 8; }
        else { a->ident = 8; b->ident = 5; }
      }
    }
  }
  Warp::sync();
}

// time constants and tallies (main.F90:127, 143-156) in the view's registers + the stored tallies
DMD_DEV void init_scalars(Rep& r, double tstar, uint64_t seed) {
  r.t = 0.0; r.tfalse = 0.0; r.old_tfalse = 0.0;
  r.setemp = tstar * 12.0;
  r.t_fact = 0.00005;
  r.n_forced = 150.0;
  r.interval = r.t_fact / dmd_sqrt(r.setemp);
  r.interval_max = r.n_forced * r.interval;
  r.avegtime = 0.00005 / dmd_sqrt(r.setemp);
  r.coll = 0;
  r.seed = seed;
  r.ctr = 0;
  r.n_pair_pred = r.n_nbr_visits = 0;
  r.n_log = r.n_out = 0;
  if (Warp::lane() == 0) {
    RepScalars& q = *r.sc;
    for (int k = 0; k < 32; k++) q.nevents[k] = 0;
    q.numghosts = q.nupdates = q.nforcedupdate = 0;
    q.rng_seed = seed;
  }
}

// calendar entry k at run start: main.F90:212-234, 408-423; the first ghost time is drawn here (:408-416)
DMD_DEV CalEnt init_cal_entry(const Rep& r, int k, double tghost) {
  CalEnt e;
  e.t = T_PAD; e.ptnr = -1; e.type = -1;
  if (k < r.N) e.t = r.interval_max + 1e-10;
  else if (k < r.N + 3) {
    e.t = k == r.N ? tghost : (k == r.N + 1 ? r.interval : 3.3 / (dmd_sqrt(r.setemp)) + 5);
    e.ptnr = -2;
    e.type = -2;
  }
  return e;
}
DMD_DEV double first_ghost_time(Rep& r) {  // main.F90:408-416 (advances the replica's RNG counter)
  if (!r.c.sys->canon) return 1000000000.0;
  double tgho = 0.0;
  while (tgho < 1e-18 || tgho == 1.0) tgho = rng_uniform(r.seed, r.ctr);
  return -1.0 * dmd_log(tgho) * r.avegtime * .0000001;
}

// the whole run start by ONE warp (host trace build; the CUDA backend spreads the same steps over the grid)
DMD_DEV void init_replica(Rep& r, const double* sv, const int32_t* bptnr1, double tstar, uint64_t seed,
                          const int32_t* nc, int n_nc, int cal_stride) {
  const int N = r.N;
  r.error = 0;
  r.error_info = 0;
  Warp::sync();
  int bad = -1;
  for (int k = Warp::lane(); k < N; k += DMD_W)
    if (!init_bead(r, k, sv, bptnr1)) bad = k;
  {
    const unsigned m = Warp::ballot(bad >= 0);
    if (m) {
      r.error = DMD_E_BAD_INPUT;  // reported by the host after the launch
      r.error_info = Warp::shfl(bad, dmd_ffs(m) - 1);
    }
  }
  Warp::sync();
  // ---- main.F90:249-321, step 1: collect the pairs (k < kj) the loop acts on into scratch (the not yet built
  // up-list array: first half unsorted, second half sorted)
  int32_t* hits = reinterpret_cast<int32_t*>(r.up);
  const int hit_cap = (int)(((size_t)N * r.cap) / 4);
  int nh = 0;
  for (int base = 0; base < n_nc; base += DMD_W) {
    const int ik = base + Warp::lane();
    const int k = ik < n_nc ? nc[ik] : -1;
    BeadRec a;
    uint32_t mk = 0;
    int ck = 0;
    if (k >= 0) {
      a = r.rec[k];
      mk = r.c.meta[k];
      ck = r.c.chain[k];
    }
    int ij = ik + 1;
    while (true) {
      bool found = false;  // next pair of this lane that qualifies
      int kj = -1;
      while (k >= 0 && ij < n_nc) {
        kj = nc[ij];
        ij++;
        if (fixup_pair_hit(r, k, a, mk, ck, kj)) {
          found = true;
          break;
        }
      }
      const unsigned m = Warp::ballot(found);
      if (!m) break;
      const int pos = nh + dmd_popc(m & ((1u << Warp::lane()) - 1u));
      if (found && pos < hit_cap) {
        hits[2 * pos] = k;
        hits[2 * pos + 1] = kj;
      }
      nh += dmd_popc(m);
    }
  }
  if (nh > hit_cap) {
    set_error(r, DMD_E_NBR_CAP, nh);
    nh = 0;
  }
  fixup_replay(r, hits, nh, hit_cap);
  init_scalars(r, tstar, seed);
  const double tghost = first_ghost_time(r);
  for (int k = Warp::lane(); k < cal_stride; k += DMD_W) r.cal[k] = init_cal_entry(r, k, tghost);
  Warp::sync();
}

DMD_VARIANT_END
}  // namespace dmd
