// dmd_lockstep.h -- several replicas per hardware warp, executed in LOCKSTEP (device only; included by dmd_cuda.cu
// right after the 8-lane (or 16-lane) build of dmd_engine.h, inside the same namespace).
//
// Why.  The warp-per-replica loop is bound by the latency of its dependent gathers and by instruction supply, not by
// lanes or issue slots: a pass over the ~11 items of a bead leaves most of 32 lanes idle, the
// register file limits an SM to 28 hardware warps, and 28 instruction streams at 28 different places thrash the 32 KB
// instruction cache.  Giving a replica LK_W = 8 lanes puts LK_SUB = 4 replicas into one warp (16 lanes: 2): four times
// the events in flight per register, and ONE instruction stream (one fetch, one issue slot) serves all of them.
//
// How.  The LK_W-lane engine code of dmd_engine.h names the lanes of its own segment in every collective, so the
// segments of a warp may run it independently -- but a collective on part of a warp splits the warp, and the parts then
// take turns instead of sharing instructions (measured: no gain).  The functions below are the hot path -- calendar
// pop, hard-core / bond event, partial_events -- rewritten so that all replicas of the warp execute the SAME
// instructions at the same time: every collective is a full-warp one (shuffles of width LK_W, full ballots read by
// segments, butterfly minima inside the aligned 8-lane segment / REDUX with the other lanes neutralised in the 16-lane
// build) and every loop bound is a full-warp vote.  The rule that keeps this deadlock-free: between two full-warp
// collectives the replicas diverge only into code that has none (or into the serial LK_W-lane engine, whose
// collectives name one segment).  A replica that has anything but a hard-core / bond event at the head of its calendar
// (H-bond events, ghost, interval incl. list rebuild, output; ~1.6 % of the events) processes it through the serial
// engine code in a divergent branch while the others wait; all meet again at the next vote.
// Results are those of the serial code: same items, same order of compare-and-lower operations, same tie rules.
// ("both replicas" / "half" in the comments below date from the 16-lane version: read "all replicas of the warp" /
// "the replica's lane segment".)
#if (DMD_W != 16 && DMD_W != 8) || defined(DMD_HOST_TRACE)
#error "dmd_lockstep.h is the several-replicas-per-warp hot path: include it in the 16- or 8-lane device build only"
#endif

namespace dmd {
DMD_VARIANT_BEGIN

// LK_W lanes per replica, LK_SUB replicas ("subs") per hardware warp
constexpr int LK_W = DMD_W, LK_SUB = 32 / DMD_W;
constexpr unsigned LK_MASK = (1u << LK_W) - 1u;
struct Lk {
  static __device__ __forceinline__ int lane() { return threadIdx.x & (LK_W - 1); }
  static __device__ __forceinline__ int half() { return (threadIdx.x >> WLOG) & (LK_SUB - 1); }  // which replica of the warp
  static __device__ __forceinline__ void sync() { __syncwarp(); }
  // bit k = lane k of the caller's replica
  static __device__ __forceinline__ unsigned ballot(bool p) {
    return (__ballot_sync(0xffffffffu, p) >> (threadIdx.x & 31u & ~(unsigned)(LK_W - 1))) & LK_MASK;
  }
  static __device__ __forceinline__ int shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src, LK_W); }
  static __device__ __forceinline__ bool any2(bool p) { return __any_sync(0xffffffffu, p); }  // over ALL replicas of the warp
  static __device__ __forceinline__ bool all2(bool p) { return __all_sync(0xffffffffu, p); }
  static __device__ __forceinline__ unsigned max2(unsigned v) { return __reduce_max_sync(0xffffffffu, v); }
  // minimum over the lanes of the caller's replica: one REDUX per replica of the warp, the others' lanes neutralised
  static __device__ __forceinline__ unsigned hmin(unsigned v) {
#if DMD_W == 8  // three butterfly steps inside the aligned 8-lane segment: 6 instructions instead of 4 x (SEL, REDUX, MOV, SEL)
    v = min(v, __shfl_xor_sync(0xffffffffu, v, 4));
    v = min(v, __shfl_xor_sync(0xffffffffu, v, 2));
    return min(v, __shfl_xor_sync(0xffffffffu, v, 1));
#else
    const int h = half();
    unsigned res = 0;
#pragma unroll
    for (int s = 0; s < LK_SUB; s++) {
      const unsigned m = __reduce_min_sync(0xffffffffu, h == s ? v : 0xffffffffu);
      if (h == s) res = m;
    }
    return res;
#endif
  }
  // minimum over the lanes of the caller's replica that are in the caller's group (16-lane build: two groups)
  static __device__ __forceinline__ unsigned gmin(unsigned v, int grp) {
#if DMD_W == 16
    const int s = half() * 2 + grp;
    const unsigned m0 = __reduce_min_sync(0xffffffffu, s == 0 ? v : 0xffffffffu);
    const unsigned m1 = __reduce_min_sync(0xffffffffu, s == 1 ? v : 0xffffffffu);
    const unsigned m2 = __reduce_min_sync(0xffffffffu, s == 2 ? v : 0xffffffffu);
    const unsigned m3 = __reduce_min_sync(0xffffffffu, s == 3 ? v : 0xffffffffu);
    return s < 2 ? (s == 0 ? m0 : m1) : (s == 2 ? m2 : m3);
#else
    return hmin(v);  // one bead per replica and pass: no groups
#endif
  }
};

// (ordered image, key) arg-min over the caller's half: smallest value, ties -> smallest key (key >= 0)
DMD_DEV void lk_hargmin_ord(unsigned& hi, unsigned& lo, int& key) {
  const unsigned mhi = Lk::hmin(hi);
  const bool c1 = hi == mhi;
  const unsigned mlo = Lk::hmin(c1 ? lo : 0xffffffffu);
  const bool c2 = c1 && lo == mlo;
  const unsigned mkey = Lk::hmin(c2 ? (unsigned)key : 0xffffffffu);
  hi = mhi;
  lo = mlo;
  key = (int)mkey;
}
DMD_DEV double lk_hmin_d(double v) {
  unsigned hi, lo;
  ord_split(v, hi, lo);
  const unsigned mhi = Lk::hmin(hi);
  const unsigned mlo = Lk::hmin(hi == mhi ? lo : 0xffffffffu);
  return ord_join(mhi, mlo);
}
// (value, key) arg-min over the lanes of the caller's group
DMD_DEV void lk_gargmin(double& v, int& key, int grp) {
  unsigned hi, lo;
  ord_split(v, hi, lo);
  const unsigned mhi = Lk::gmin(hi, grp);
  const bool c1 = hi == mhi;
  const unsigned mlo = Lk::gmin(c1 ? lo : 0xffffffffu, grp);
  const bool c2 = c1 && lo == mlo;
  const unsigned mkey = Lk::gmin(c2 ? (unsigned)key : 0xffffffffu, grp);
  v = ord_join(mhi, mlo);
  key = (int)mkey;
}

// flush_dirty() for both replicas: the loop runs while either has a stale group
DMD_DEV void lk_flush_dirty(Rep& r) {
  const int lane = Lk::lane();
  Lk::sync();
  uint64_t dirty0 = warp_dirty(r)[0], dirty1 = warp_dirty(r)[1];
  Lk::sync();
  if (lane == 0) warp_dirty(r)[0] = warp_dirty(r)[1] = 0ull;
  while (Lk::any2(dirty0 != 0)) {  // two groups per round so that their loads overlap
    const bool have = dirty0 != 0;
    const int g0 = have ? pop_lowest_bit(dirty0) : 0;
    const int g1 = dirty0 ? pop_lowest_bit(dirty0) : -1;
    double x0 = T_PAD, x1 = T_PAD;
#pragma unroll
    for (int q = lane; q < 32; q += LK_W) {
      const double y0 = have ? r.cal[g0 * 32 + q].t : T_PAD;
      const double y1 = g1 >= 0 ? r.cal[g1 * 32 + q].t : T_PAD;
      x0 = y0 < x0 ? y0 : x0;
      x1 = y1 < x1 ? y1 : x1;
    }
    x0 = lk_hmin_d(x0);
    x1 = lk_hmin_d(x1);
    if (lane == 0) {
      if (have) r.tmin1[g0] = ord_bits64(x0);
      if (g1 >= 0) r.tmin1[g1] = ord_bits64(x1);
    }
  }
  while (dirty1) group_min_update(r, 64 + pop_lowest_bit(dirty1));  // systems beyond 2048 beads: serial code
  Lk::sync();
}

// pop_min() for both replicas; no early exit between the collectives (an empty calendar returns -1 at the end)
DMD_DEV int lk_pop_min(Rep& r, CalEnt& ev) {
  const int lane = Lk::lane();
  unsigned long long best = ord_bits64(T_PAD);
  int bg = 0x7fffffff;
  for (int g = lane; g < r.G; g += LK_W) {
    const unsigned long long v = r.tmin1[g];
    if (v < best) {  // ascending g: the first minimum keeps the lowest group
      best = v;
      bg = g;
    }
  }
  unsigned bhi = (unsigned)(best >> 32), blo = (unsigned)best;
  lk_hargmin_ord(bhi, blo, bg);
  best = ((unsigned long long)bhi << 32) | blo;
  const bool empty = bg == 0x7fffffff || !(ord_value64(best) < 1e299);
  const int sg = empty ? 0 : bg;
  double v = T_PAD;
  int key = 0x7fffffff, pt = -1, ty = -1;
#pragma unroll
  for (int q = lane; q < 32; q += LK_W) {
    const CalEnt e = r.cal[sg * 32 + q];
    if (e.t < v) {
      v = e.t;
      key = q;
      pt = e.ptnr;
      ty = e.type;
    }
  }
  unsigned vhi, vlo;
  ord_split(v, vhi, vlo);
  int wkey = key;
  lk_hargmin_ord(vhi, vlo, wkey);
  const int src = wkey & (LK_W - 1);  // the lane that scanned the winning entry holds it as its own best
  pt = Lk::shfl(pt, src);
  ty = Lk::shfl(ty, src);
  ev.t = ord_join(vhi, vlo);
  ev.ptnr = pt;
  ev.type = ty;
  return empty || wkey == 0x7fffffff ? -1 : sg * 32 + wkey;
}

// prediction_pass() for both replicas: ONE bead per replica and pass.  `mode` is the same for both replicas
// (0: main pass of bead i, 1: main pass of bead j -- its down list skips i, partial_events.f:136 --, 2: cascade pass
// over one or two queued beads); `act` says whether this replica takes part.  With 16 lanes the items of both
// colliding beads rarely fit one trip, so there is no two-bead mapping here: down(i) is applied by pass 0, down(j) by
// pass 1 on re-read entries -- the Fortran's order.
DMD_DEV void lk_pass(Rep& r, const int mode, const bool act, const int i, const int j, const int idx, const int rem,
                     int& cqn) {
  const int lane = Lk::lane();
  const int cap = r.cap;
  int a, q0, ssh, nu, nd, er3, skip, T, grp = 0;
  bool lact = act;
  unsigned segmask = LK_MASK;  // the lanes of this replica (bit k = lane k) that work for the same bead as this lane
  uint32_t e0 = 0, ma;
  BeadRec ra;
  if (mode != 2) {  // ---- main pass (same items as prediction_pass: see there)
    a = mode == 0 ? i : j;
    a = act ? a : 0;
    uint32_t s0 = 0, s1 = 0;  // before the list lengths are known every lane fetches "its" entry of both lists
    if (lane < cap) {
      s0 = r.up[(size_t)a * cap + lane];
      s1 = r.dn[(size_t)a * cap + lane];
    }
    (rec_load_tail)(ra, &r.rec[a]);
    nu = (int)r.nup[a];
    nd = (int)r.ndn[a];
    er3 = r.er34[2 * a];
    ma = r.c.meta[a];
    T = nu + nd + ((ra.er1 >= 0 || ra.er2 >= 0 || er3 >= 0) ? 3 : 0);
    T = act ? (T > 0 ? T : 1) : 0;  // a bead without items still needs a lane that writes its calendar entry
    if (lane == 0 && act) {        // work counters (roofline input)
      unsigned* c = warp_counters(r);
      atomicAdd(&c[0], (unsigned)(nu + nd + 6));
      atomicAdd(&c[1], (unsigned)(nu + nd));
    }
    q0 = lane;
    ssh = WLOG;
    skip = mode == 1 ? i : -1;
    {  // the entry of the lane's first item
      const bool isup = q0 < nu;
      const int src = (isup ? q0 : q0 - nu) & (LK_W - 1);
      const uint32_t t0 = (uint32_t)Lk::shfl((int)s0, src), t1 = (uint32_t)Lk::shfl((int)s1, src);
      e0 = isup ? t0 : t1;
    }
  } else {  // ---- cascade pass: one bead on all lanes of the replica, or (16-lane build) two on 8 lanes each
    const int sh = (LK_W == 16 && rem >= 2) ? 1 : 0;  // rem is this replica's own queue length; an idle replica has rem <= 0
    const int SEG = LK_W >> sh;
    const int g = lane >> (WLOG - sh);
    segmask = (sh == 0 ? LK_MASK : ((1u << SEG) - 1u)) << (g * SEG);
    grp = g;
    q0 = lane & (SEG - 1);
    ssh = WLOG - sh;
    lact = act && g < rem;
    a = lact ? r.cq[idx + g] : 0;
    if (q0 < cap) e0 = r.up[(size_t)a * cap + q0];
    (rec_load_tail)(ra, &r.rec[a]);
    ma = r.c.meta[a];
    nu = (int)r.nup[a];
    nd = 0;
    er3 = r.er34[2 * a];
    T = lact ? nu + ((ra.er1 >= 0 || ra.er2 >= 0 || er3 >= 0) ? 3 : 0) : 0;
    if (lact && T == 0) T = 1;
    skip = -1;
    if (q0 == 0 && lact) {
      unsigned* c = warp_counters(r);
      atomicAdd(&c[0], (unsigned)(nu + 3));
      atomicAdd(&c[1], (unsigned)nu);
    }
  }
  const bool main_pass = mode != 2;
  const int ntrip = (int)Lk::max2((unsigned)(T > q0 ? (T - q0 + (1 << ssh) - 1) >> ssh : 0));  // over both replicas
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
      if (trip > 0) e = ((isup ? r.up : r.dn) + (size_t)a * cap)[isup ? q : q - nu];  // long lists: a dependent load
      b = (int)(e & NB_MASK);
      sc = (int)(e >> NB_SHIFT);
      if (!isup && b == skip) b = -1;  // partial_events.f:136
    } else if (q < T) {
      const int k = q - nu - nd;
      b = k == 0 ? ra.er1 : (k == 1 ? ra.er2 : er3);
      full = b > a;                                 // events.f:77
      if (b < 0 || (!full && !main_pass)) b = -1;  // partial_events.f:100,166
    }
    if (q >= T) b = -1;
    const bool down = b >= 0 && !full;
    double tij = T_NONE;
    int type = -1;
    CalEnt eb;
    eb.t = 0.0; eb.ptnr = -1; eb.type = -1;
    if (b >= 0) {
      const BeadRec rb = r.rec[b];
      (rec_load_head)(ra, &r.rec[a]);
      uint32_t mlo = ma;
      if (!full) {
        mlo = r.c.meta[b];
        eb = r.cal[b];
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
      const bool need_full = down && eb.ptnr == a;          // l's next event was with a: cascade
      const bool lower = down && !need_full && tij < eb.t;  // eventredo_down.f:70-77
      if (lower) {
        CalEnt ne;
        ne.t = tij;
        ne.ptnr = a;
        ne.type = pack_type(type, sc);
        r.cal[b] = ne;
        atomicMin(&r.tmin1[b >> 5], ord_bits64(tij));  // a LOWERED entry: the group minimum follows exactly
      }
      const unsigned m = Lk::ballot(need_full);
      if (m) {
        const int pos = cqn + dmd_popc(m & ((1u << lane) - 1u));
        if (need_full && pos < CQ_Q) r.cq[pos] = b;
        cqn += dmd_popc(m);
      }
    }
  }
  // ---- the minimum of each bead's full items -> cal[a]; the lane that holds it writes the entry
  double wbest = best;
  int wpos = bpos;
  lk_gargmin(wbest, wpos, grp);
  const bool mine = lact && bpos == wpos && bpos != 0x7fffffff;
  const unsigned any_mine = Lk::ballot(mine) & segmask;
  const bool writer = lact && (any_mine ? mine : q0 == 0);
  if (writer) {
    CalEnt ne;
    ne.t = wbest + r.tfalse;
    ne.ptnr = any_mine ? bj : -1;
    ne.type = any_mine ? btype : -1;
    r.cal[a] = ne;
    const int g = a >> 5;  // mark_dirty_lanes(): the group is rescanned before the next pop (G <= 128 here)
    atomicOr(reinterpret_cast<unsigned*>(warp_dirty(r)) + (g >> 5), 1u << (g & 31));
  }
}

// partial_events() for both replicas; (i, j) are the caller's own beads (j < 0: one bead, main.F90:1049),
// act == false: this replica has nothing to re-predict (interval / output pseudo-event)
DMD_DEV void lk_partial_events(Rep& r, const int i, const int j, const bool xpulse_del, const bool act) {
  Lk::sync();
  int cqn = 0, idx = 0;
  int mode = 0;  // the same for both replicas: 0 bead i, 1 bead j, 2 cascades
#pragma unroll 1
  while (true) {
    const bool pact = mode == 0 ? act : (mode == 1 ? act && j >= 0 : idx < cqn);
    const int rem = cqn - idx;
    lk_pass(r, mode, pact, i, j, idx, rem, cqn);
    Lk::sync();  // later passes re-read the entries this one may have lowered
    DMD_PROF_MARK(r, mode < 2 ? 3 : 5);
    if (mode == 2) {
      if (pact) idx += (LK_W == 16 && rem >= 2) ? 2 : 1;
    } else {
      mode = (mode == 0 && Lk::any2(act && j >= 0)) ? 1 : 2;
      if (mode == 2) {  // all down items are done: prepare the cascade queue (each replica for itself, 16-lane code)
        if (cqn > CQ_Q) {
          set_error(r, DMD_E_NBR_CAP, cqn);
          cqn = 0;
        }
        if (cqn > 1) {  // a bead queued by both i and j is re-predicted once
          int keep_n = 0;
          for (int base = 0; base < cqn; base += LK_W) {
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
        DMD_PROF_MARK(r, 4);
      }
    }
    if (mode == 2 && !Lk::any2(idx < cqn)) break;
  }
  if (xpulse_del && Lk::lane() == 0) {  // partial_events.f:195-201
    if (r.rec[i].ident < r.rec[j].ident) repuls_del_b(r, i, j);
    else repuls_del_b(r, j, i);
  }
  Lk::sync();
}

// process_one() for both replicas: calendar entry o (already popped) of each.  Hard-core / bond events (> 98 %) take
// the branch without collectives; anything else is handled by the serial 16-lane code of dmd_engine.h inside the
// divergent branch (its collectives name one half of the warp); the re-prediction runs for both replicas together.
DMD_DEV void lk_process(Rep& r, const int o, const CalEnt& ev) {
  const double prev_tfalse = r.tfalse;
  r.tfalse = ev.t;
  r.coll += 1;
  int pi = o, pj = ev.ptnr;
  int ct = type_of(ev.type), code = 0;
  bool xpulse_del = false, redo = true;
  const bool hot = o < r.N && pj >= 0 && ct >= 1 && ct <= 3;
  BeadRec ri, rj;
  if (hot) {
    ri = r.rec[pi];
    rj = r.rec[pj];
    code = overlay_code(sc_of(ev.type), pi, ri, pj, rj);  // ev_code(i,j), main.F90:587
    ct = event_dynamics_hot(r.c, ct, code, ri, rj, r.c.meta[pi], ri.bptnr == pj, r.tfalse);
  } else if (o < r.N) {  // H-bond related pair event: resolution, eventdyn, bookkeeping (cold, out of line)
    code = overlay_code(sc_of(ev.type), pi, r.rec[pi], pj, r.rec[pj]);
    DMD_PROF_MARK(r, 2);
    const ColdRes cr = pair_event_cold(r, pi, pj, ct, code);
    ct = cr.ct;
    xpulse_del = cr.xpulse != 0;
    r.ctr = cr.ctr;
    DMD_PROF_MARK(r, 8);
  } else {  // pseudo-events
    DMD_PROF_MARK(r, 2);
    rep_save(r);
    if (o == r.N) {
      pi = ghost_event_cold(r, prev_tfalse);
      pj = -1;
    } else {
      redo = false;
      if (o == r.N + 1) interval_event_cold(r);
      else output_event_cold(r);
    }
    Warp::sync();
    rep_load_scalars(r);
    clear_dirty(r);
    Warp::sync();
    if (redo) mark_dirty(r, r.N >> 5);  // the next ghost time
    DMD_PROF_MARK(r, o == r.N ? 9 : 6);
  }
  DMD_PROF_MARK(r, 2);
  Lk::sync();  // every lane has read the two records
  DMD_PROF_MARK(r, 10);  // (waiting for the warp's other replica)
  if (o < r.N) {
    if (hot && Lk::lane() == 0) {
      BeadRec* qi = &r.rec[pi];
      BeadRec* qj = &r.rec[pj];
      qi->x = ri.x; qi->y = ri.y; qi->z = ri.z; qi->vx = ri.vx; qi->vy = ri.vy; qi->vz = ri.vz;
      qj->x = rj.x; qj->y = rj.y; qj->z = rj.z; qj->vx = rj.vx; qj->vy = rj.vy; qj->vz = rj.vz;
    }
    if (Lk::lane() == 0 && ct >= 0 && ct < 32)
      atomicAdd(reinterpret_cast<unsigned long long*>(&r.sc->nevents[ct]), 1ull);  // main.F90:926
    log_event(r, pi, pj, ct, code);
    DMD_PROF_MARK(r, 2);
  }
  lk_partial_events(r, pi, pj, xpulse_del, redo);  // main.F90:943, :1049
}

// run_events() for the replicas of a hardware warp
DMD_DEV void lk_run_events(Rep& r, int64_t n_events, bool stop_at_output) {
  const int64_t coll_end = r.coll + n_events;
  bool live = r.error == 0;
  const bool lockstep = r.G <= 128;  // the same for both replicas (one system); larger systems: serial code only
#pragma unroll 1
  while (lockstep && Lk::all2(live && r.coll < coll_end)) {  // both replicas have work
    DMD_PROF_MARK(r, 7);
    lk_flush_dirty(r);
    DMD_PROF_MARK(r, 0);
    CalEnt ev;
    const int o = lk_pop_min(r, ev);
    DMD_PROF_MARK(r, 1);
    if (Lk::any2(o < 0)) {  // an empty calendar: report it and leave the lockstep loop (the other replica goes on in
      if (o < 0) {          // the serial loop below; its popped entry is still in place)
        set_error(r, DMD_E_CAL_EMPTY, 0);
        live = false;
      }
      break;
    } else {
      lk_process(r, o, ev);
      if (stop_at_output && o == r.N + 2) live = false;
    }
    if (r.error) live = false;
  }
  // the other replica of the warp has stopped, or there is none: the serial loop for what is left
  if (live && r.coll < coll_end) run_events(r, coll_end - r.coll, stop_at_output);
  else flush_dirty(r);
}

DMD_VARIANT_END
}  // namespace dmd
