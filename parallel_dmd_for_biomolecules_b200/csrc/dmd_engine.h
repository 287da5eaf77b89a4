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
// Every loop is written for DMD_W lanes (32 on the device, 1 in the host trace build).
#pragma once
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

constexpr double T_PAD = 1e300;  // calendar padding entries
constexpr int CQ_CAP = 160;      // cascade queue (<= one entry per down-list candidate of a pass)

struct Rep {
  Ctx c;
  int N, cap, G;
  BeadRec* rec;
  CalEnt* cal;
  int32_t* er34;
  uint32_t *up, *dn;
  uint16_t *nup, *ndn;
  double* oldr;
  int32_t *cellhead, *cnext, *cellof;
  double* tmin1;
  RepScalars* sc;
  EventLogRec* log;
  OutRec* out;
  int32_t* cq;  // cascade queue storage (shared memory on the device)
  // scalars cached in registers (identical in every lane)
  double t, tfalse, old_tfalse, setemp, interval, t_fact, interval_max, n_forced, avegtime;
  int64_t coll;
  uint64_t seed, ctr;
  int64_t n_pair_pred, n_nbr_visits;
  int32_t n_log, n_out, error, error_info;
  uint64_t dirty0, dirty1;  // calendar groups 0..127 whose minimum is stale
};

DMD_DEV void rep_load_scalars(Rep& r) {
  const RepScalars& q = *r.sc;
  r.t = q.t; r.tfalse = q.tfalse; r.old_tfalse = q.old_tfalse; r.setemp = q.setemp; r.interval = q.interval;
  r.t_fact = q.t_fact; r.interval_max = q.interval_max; r.n_forced = q.n_forced; r.avegtime = q.avegtime;
  r.coll = q.coll; r.seed = q.rng_seed; r.ctr = q.rng_ctr;
  r.n_pair_pred = q.n_pair_pred; r.n_nbr_visits = q.n_nbr_visits;
  r.n_log = q.n_log; r.n_out = q.n_out; r.error = q.error; r.error_info = q.error_info;
}

DMD_DEV void rep_bind(Rep& r, const DevArrays& d, const PairTables* tab, int32_t* cq, int rid) {
  const SysConst* s = d.sys;
  r.c.sys = s;
  r.c.tab = tab;
  r.c.meta = d.meta;
  r.c.chain = d.chain;
  const int N = s->N;
  r.N = N;
  r.cap = s->cap;
  r.G = s->ngroups;
  const size_t rr = (size_t)rid;
  r.rec = d.rec + rr * N;
  r.cal = d.cal + rr * d.cal_stride;
  r.er34 = d.er34 + rr * 2 * N;
  r.up = d.up + rr * N * s->cap;
  r.dn = d.dn + rr * N * s->cap;
  r.nup = d.nup + rr * N;
  r.ndn = d.ndn + rr * N;
  r.oldr = d.oldr + rr * 3 * N;
  const size_t nc3 = (size_t)s->ncr * s->ncr * s->ncr;
  r.cellhead = d.cellhead + rr * nc3;
  r.cnext = d.cnext + rr * N;
  r.cellof = d.cellof + rr * N;
  r.tmin1 = d.tmin1 + rr * s->ngroups;
  r.sc = d.scal + rr;
  r.log = d.log + rr * (s->log_cap > 0 ? s->log_cap : 1);
  r.out = d.out + rr * s->out_cap;
  r.cq = cq;
  r.dirty0 = r.dirty1 = 0;
  rep_load_scalars(r);
}

DMD_DEV void rep_save(Rep& r) {
  if (Warp::lane() == 0) {
    RepScalars& q = *r.sc;
    q.t = r.t; q.tfalse = r.tfalse; q.old_tfalse = r.old_tfalse; q.setemp = r.setemp; q.interval = r.interval;
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
  if (Warp::lane() == 0) r.tmin1[g] = v;
}

DMD_DEV void mark_dirty(Rep& r, int g) {  // g warp-uniform
  if (g < 64) r.dirty0 |= 1ull << g;
  else if (g < 128) r.dirty1 |= 1ull << (g - 64);
  else {  // very large systems: refresh immediately
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
#if DMD_W > 1
  while (r.dirty0) {  // two groups per round so that their loads overlap
    const int g0 = pop_lowest_bit(r.dirty0);
    const int g1 = r.dirty0 ? pop_lowest_bit(r.dirty0) : -1;
    double x0 = r.cal[g0 * 32 + Warp::lane()].t;
    double x1 = g1 >= 0 ? r.cal[g1 * 32 + Warp::lane()].t : 0.0;
    x0 = warp_min(x0);
    if (g1 >= 0) x1 = warp_min(x1);
    if (Warp::lane() == 0) {
      r.tmin1[g0] = x0;
      if (g1 >= 0) r.tmin1[g1] = x1;
    }
  }
#else
  while (r.dirty0) group_min_update(r, pop_lowest_bit(r.dirty0));
#endif
  while (r.dirty1) group_min_update(r, 64 + pop_lowest_bit(r.dirty1));
  Warp::sync();
}

// every lane may have changed the entry of a different bead l (l < 0: none): record the groups uniformly
DMD_DEV void mark_dirty_lanes(Rep& r, int l) {
  const int g = l >> 5;  // negative when l < 0
  unsigned a0 = (g >= 0 && g < 32) ? 1u << g : 0u;
  unsigned a1 = (g >= 32 && g < 64) ? 1u << (g - 32) : 0u;
  a0 = warp_or(a0);
  a1 = warp_or(a1);
  r.dirty0 |= (uint64_t)a0 | ((uint64_t)a1 << 32);
  if (r.G > 64) {  // larger systems (uniform branch)
    unsigned b0 = (g >= 64 && g < 96) ? 1u << (g - 64) : 0u;
    unsigned b1 = (g >= 96 && g < 128) ? 1u << (g - 96) : 0u;
    b0 = warp_or(b0);
    b1 = warp_or(b1);
    r.dirty1 |= (uint64_t)b0 | ((uint64_t)b1 << 32);
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
  r.dirty0 = r.dirty1 = 0;
  Warp::sync();
}

// returns the owner index of the earliest entry (or -1) and the entry itself
DMD_DEV int pop_min(Rep& r, CalEnt& ev) {
  double best = T_PAD;
  int bg = 0x7fffffff;
  for (int g = Warp::lane(); g < r.G; g += DMD_W) {
    double v = r.tmin1[g];
    if (v < best) {  // ascending g: the first minimum keeps the lowest group
      best = v;
      bg = g;
    }
  }
  warp_argmin(best, bg);
  if (bg == 0x7fffffff || !(best < 1e299)) return -1;
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
  const int src = wkey & 31;  // the owning lane (one entry per lane when DMD_W == 32)
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
// prediction pass for bead a (hot): lanes cover, in this order,
//   [0, nu)            up-list partners j > a          } events.f:26-57 / eventredo_up.f : owner a
//   [nu, nu+3)         aux slots extra_repuls(a,1:3)>a }
//   [nF, nF+nd)        down-list beads l < a           } eventredo_down.f : owner l, or cascade when
//   [nF+nd, nF+nd+3)   aux slots extra_repuls(a,1:3)<a }   nptnr(l) == a (partial_events.f:73-96)
// with_down = false restricts the pass to the first two ranges (a cascaded full re-prediction).
// Lanes of the last two ranges that need a cascade push their bead on the queue r.cq.
// ---------------------------------------------------------------------------------------------------------
DMD_DEV void predict_pass(Rep& r, int a, bool with_down, int skip, int& cqn) {
  // level-1 loads, all independent: the record of a, its list lengths, and (speculatively, before the lengths
  // are known) the first 32 entries of both lists -- one entry per lane
  const size_t lbase = (size_t)a * r.cap;
#if DMD_W > 1
  uint32_t eu0 = 0, ed0 = 0;
  if (Warp::lane() < r.cap) {
    eu0 = r.up[lbase + Warp::lane()];
    if (with_down) ed0 = r.dn[lbase + Warp::lane()];
  }
#endif
  const BeadRec ra = r.rec[a];
  const uint32_t ma = r.c.meta[a];
  const int nu = r.nup[a];
  const int nd = with_down ? (int)r.ndn[a] : 0;
  const int er3 = r.er34[2 * a];
  const int nF = nu + 3, total = with_down ? nF + nd + 3 : nF;
  double best = r.interval_max + LTSTEP - r.tfalse;
  int bpos = 0x7fffffff, bj = -1, btype = -1;
  r.n_pair_pred += total;
  r.n_nbr_visits += nu + nd;
  for (int base = 0; base < total; base += DMD_W) {
    const int p = base + Warp::lane();
    int b = -1, sc = 1;  // the other bead of the pair and the pair's static class
    const bool full = p < nF;
    const int q = p - nF;  // index into the down list
#if DMD_W > 1
    const uint32_t edq = Warp::shfl((int)ed0, q & 31);
#endif
    if (p < nu) {
#if DMD_W > 1
      uint32_t e = base == 0 ? eu0 : r.up[lbase + p];
#else
      uint32_t e = r.up[lbase + p];
#endif
      b = (int)(e & NB_MASK);
      sc = (int)(e >> NB_SHIFT);
    } else if (p < nF) {
      int k = p - nu;
      b = k == 0 ? ra.er1 : (k == 1 ? ra.er2 : er3);
      if (b <= a) b = -1;  // events.f:77
    } else if (q < nd) {
#if DMD_W > 1
      uint32_t e = q < 32 ? edq : r.dn[lbase + q];
#else
      uint32_t e = r.dn[lbase + q];
#endif
      b = (int)(e & NB_MASK);
      sc = (int)(e >> NB_SHIFT);
      if (b == skip) b = -1;  // partial_events.f:136
    } else if (p < total) {
      int k = q - nd;
      b = k == 0 ? ra.er1 : (k == 1 ? ra.er2 : er3);
      if (!(b >= 0 && b < a)) b = -1;  // partial_events.f:100,166
    }
    bool need_full = false;
    int changed = -1;
    if (b >= 0) {
      // level-2 loads, all depending on b only
      const BeadRec rb = r.rec[b];
      CalEnt eb;
      eb.t = 0.0; eb.ptnr = -1; eb.type = -1;
      uint32_t mlo = ma;
      if (!full) {
        eb = r.cal[b];
        mlo = r.c.meta[b];
      }
      if (!full && eb.ptnr == a) {
        need_full = true;  // l's next event was with a: full re-prediction of l (cascade)
      } else {
        const int code = overlay_code(sc, a, ra, b, rb);
        double tij = T_NONE;
        int type = -1;
        {  // one prediction site for both orientations (owner = lower index: a when full, b otherwise)
          const Geom g = pair_geom(ra, rb, r.tfalse);
          const double rijsq = g.rx * g.rx + g.ry * g.ry + g.rz * g.rz;
          const double vijsq = g.vx * g.vx + g.vy * g.vy + g.vz * g.vz;
          const int idlo = full ? ra.ident : rb.ident, idhi = full ? rb.ident : ra.ident;
          const bool bonded = full ? ra.bptnr == b : rb.bptnr == a;
          pair_time_core(r.c, code, g.bij, rijsq, vijsq, idlo, idhi, mlo, bonded, tij, type);
        }
        if (full) {
          if (tij < best) {  // strict: first in evaluation order wins (events.f:53)
            best = tij;
            bpos = p;
            bj = b;
            btype = pack_type(type, sc);
          }
        } else {  // eventredo_down.f:70-77
          tij = tij + r.tfalse;
          if (tij < eb.t) {
            CalEnt ne;
            ne.t = tij;
            ne.ptnr = a;
            ne.type = pack_type(type, sc);
            r.cal[b] = ne;
            changed = b;
          }
        }
      }
    }
    if (with_down) {
      mark_dirty_lanes(r, changed);
      unsigned m = Warp::ballot(need_full);
      if (m) {
        int pos = cqn + dmd_popc(m & ((1u << Warp::lane()) - 1u));
        if (need_full && pos < CQ_CAP) r.cq[pos] = b;
        cqn += dmd_popc(m);
      }
    }
  }
  double wbest = best;
  int wpos = bpos;
  warp_argmin(wbest, wpos);
  unsigned owner = Warp::ballot(bpos == wpos && bpos != 0x7fffffff);
  if (owner) {
    int src = dmd_ffs(owner) - 1;
    bj = Warp::shfl(bj, src);
    btype = Warp::shfl(btype, src);
  } else {
    bj = -1;
    btype = -1;
  }
  if (Warp::lane() == 0) {
    CalEnt ne;
    ne.t = wbest + r.tfalse;
    ne.ptnr = bj;
    ne.type = btype;
    r.cal[a] = ne;
  }
  mark_dirty(r, a >> 5);
}

// partial_events.f:16-201.  Order used: full(i), down(i) [+cascades], full(j), down(j) [+cascades]; this is
// equivalent to the Fortran's full(i), full(j), down(i), down(j) because full(j) only rewrites entry j, which
// down(i) never reads (its beads are < i < j), see DESIGN.md "pass order".
DMD_DEV void repuls_del_b(Rep& r, int n, int cb);

DMD_DEV void partial_events(Rep& r, int i, int j, bool xpulse_del) {
  Warp::sync();
  int cqn = 0, stage = 0;
  while (true) {
    int a, skip = -1;
    bool with_down;
    if (cqn > 0) {
      if (cqn > CQ_CAP) {
        set_error(r, DMD_E_NBR_CAP, cqn);
        break;
      }
      Warp::sync();
      a = r.cq[--cqn];
      with_down = false;
    } else if (stage < 2) {
      a = stage == 0 ? i : j;
      skip = stage == 0 ? -1 : i;
      stage++;
      if (a < 0) continue;
      with_down = true;
    } else {
      break;
    }
    predict_pass(r, a, with_down, skip, cqn);
    Warp::sync();
  }
  if (xpulse_del) {
    if (Warp::lane() == 0) {
      if (r.rec[i].ident < r.rec[j].ident) repuls_del_b(r, i, j);
      else repuls_del_b(r, j, i);
    }
    Warp::sync();
  }
}

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

// worker block main.F90:1429-1959 with current state, then master main.F90:926,943
DMD_DEV void pair_event(Rep& r, int i, const CalEnt& ev) {
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
  if (Warp::lane() == 0 && ct >= 0 && ct < 32) r.sc->nevents[ct] += 1;  // main.F90:926
  log_event(r, i, j, ct, code);
  partial_events(r, i, j, xpulse_del);  // main.F90:943
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

DMD_DEV void cell_build(Rep& r) {
  const SysConst& s = *r.c.sys;
  const int ncr = s.ncr, nc = s.num_cell, nw = s.n_wrap;
  Warp::sync();
  for (int k = Warp::lane(); k < r.N; k += DMD_W) {
    int cx, cy, cz;
    cell_coords(s, r.rec[k], cx, cy, cz);
    r.cellof[k] = 1 + (cx + nw) + (cy + nw) * nc + (cz + nw) * nc * nc;  // cell_add.f:25
    if (cx < 0 || cy < 0 || cz < 0 || cx >= ncr || cy >= ncr || cz >= ncr) {
      // the reference would file the bead in a ghost cell that is never looked up
      r.cnext[k] = -2;
      continue;
    }
    int cidx = cx + (cy + cz * ncr) * ncr;
    r.cnext[k] = exch(&r.cellhead[cidx], k);
  }
  Warp::sync();
}

DMD_DEV void cell_clear(Rep& r) {
  const SysConst& s = *r.c.sys;
  const int ncr = s.ncr;
  Warp::sync();
  for (int k = Warp::lane(); k < r.N; k += DMD_W) {
    if (r.cnext[k] == -2) continue;
    int cx, cy, cz;
    cell_coords(s, r.rec[k], cx, cy, cz);
    r.cellhead[cx + (cy + cz * ncr) * ncr] = -1;
  }
  Warp::sync();
}

DMD_DEV void nbor_build(Rep& r) {
  const SysConst& s = *r.c.sys;
  const int ncr = s.ncr, cap = r.cap;
  int overflow = 0;
  for (int k = Warp::lane(); k < r.N; k += DMD_W) {
    const BeadRec rk = r.rec[k];
    const uint32_t mk = r.c.meta[k];
    const int ck = r.c.chain[k];
    int nu = 0, nd = 0;
    if (r.cnext[k] != -2) {
      int cx, cy, cz;
      cell_coords(s, rk, cx, cy, cz);
      for (int dz = -2; dz <= 2; dz++) {
        int z = cz + dz;
        z = z < 0 ? z + ncr : (z >= ncr ? z - ncr : z);
        for (int dy = -2; dy <= 2; dy++) {
          int y = cy + dy;
          y = y < 0 ? y + ncr : (y >= ncr ? y - ncr : y);
          const int rowbase = (y + z * ncr) * ncr;
          for (int dx = -2; dx <= 2; dx++) {
            int x = cx + dx;
            x = x < 0 ? x + ncr : (x >= ncr ? x - ncr : x);
            for (int j = r.cellhead[rowbase + x]; j >= 0; j = r.cnext[j]) {
              if (j == k) continue;
              const int sc = static_code(s, mk, ck, k, r.c.meta[j], r.c.chain[j], j);
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
            }
          }
        }
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

// ---- cold pseudo-events: they work on a by-value copy of the view and hand the scalars back through r.sc
// main.F90:997-1049
DMD_COLD void ghost_event_cold(Rep r) {
  const int N = r.N;
  int i;
  do {
    i = (int)(rng_uniform(r.seed, r.ctr) * N);
  } while (i == N);
  BeadRec b = r.rec[i];
  const double bmi = r.c.sys->bmass[b.ident];
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
  mark_dirty(r, N >> 5);
  if (r.tfalse < r.old_tfalse) r.tfalse = r.old_tfalse;  // main.F90:1047
  log_event(r, N, i, -2, 0);
  partial_events(r, i, -1, false);
  flush_dirty(r);
  rep_save(r);
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
    nbor(r);
    predict_all(r);  // events(); every bead's (tim, nptnr, coltype) is re-derived from interval_max+ltstep
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
        double ep = r.c.tab->ep_sqrt[tix(rk.ident, rj.ident)];
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
DMD_DEV bool step(Rep& r) {
  flush_dirty(r);
  CalEnt ev;
  const int o = pop_min(r, ev);
  if (o < 0) {
    set_error(r, DMD_E_CAL_EMPTY, 0);
    return false;
  }
  r.tfalse = ev.t;
  r.coll += 1;
  if (o < r.N) {
    pair_event(r, o, ev);
  } else {
    rep_save(r);  // hand the scalars to the out-of-line handler through r.sc ...
    if (o == r.N) ghost_event_cold(r);
    else if (o == r.N + 1) interval_event_cold(r);
    else output_event_cold(r);
    Warp::sync();
    rep_load_scalars(r);  // ... and take them back
    r.dirty0 = r.dirty1 = 0;
  }
  r.old_tfalse = r.tfalse;
  return r.error == 0;
}

DMD_DEV void run_events(Rep& r, int64_t n_events) {
  for (int64_t n = 0; n < n_events; n++)
    if (!step(r)) break;
  flush_dirty(r);
}

// Replica-exchange temperature change on resident state (new functionality): advance to true
// positions, rescale velocities by sqrt(T_new/T_old), reset the time constants of main.F90:143-156 for the new
// temperature and rebuild lists + calendar.  H-bond state (bptnr, identity, extra_repuls, overlay) is kept.
DMD_DEV void retemp(Rep& r, double tstar_new) {
  const int N = r.N;
  const double tf = r.tfalse;
  const double setemp_new = tstar_new * 12.0;
  const double scale = dmd_sqrt(setemp_new / r.setemp);
  r.t = r.t + tf;
  Warp::sync();
  for (int k = Warp::lane(); k < N; k += DMD_W) {
    BeadRec* p = &r.rec[k];
    double x = p->x + p->vx * tf, y = p->y + p->vy * tf, z = p->z + p->vz * tf;
    x = x - dmd_round(x); y = y - dmd_round(y); z = z - dmd_round(z);
    p->x = x; p->y = y; p->z = z;
    p->vx = p->vx * scale; p->vy = p->vy * scale; p->vz = p->vz * scale;
    r.oldr[3 * k] = x; r.oldr[3 * k + 1] = y; r.oldr[3 * k + 2] = z;
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
  nbor(r);
  predict_all(r);
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

}  // namespace dmd
