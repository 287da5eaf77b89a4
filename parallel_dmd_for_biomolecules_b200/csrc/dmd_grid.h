// dmd_grid.h -- the whole-GPU engine for ONE large system: the batched conservative commit of dmd_block.h with
// the rounds spread over every SM (device code only).
//
// A 10^6-bead box has ~35 000 peptides: the serial order still has to be respected, but hundreds of the earliest
// events have pairwise disjoint footprints.  148 CTAs x 16 warps run the same round as the CTA-per-replica engine
// -- scan + window selection, ranking in serial order (time, bead index), footprint claims with atomicMin, the
// longest prefix of independent hard-core / bond events, parallel execution with undo logs, the reference's
// validation rule (main.F90:970-993), rollback of what it rejects -- with the state in HBM / L2 (64 B records:
// 64 MB for 10^6 beads, L2-resident on a B200), the round's bookkeeping in a global workspace and a grid-wide
// barrier between the phases.  One warp executes one event, so a round commits up to 2368 events.
// Anything that is not a plain type 1/2/3 pair event ends the kernel: the host lets the warp-per-replica engine
// process that one calendar entry (ghost, interval incl. list rebuild, output, H-bond events) and relaunches.
// Exactness: identical to dmd_block.h (same claims, same prefix rule, same validation).
#pragma once
#include "dmd_block.h"

#if !defined(DMD_HOST_TRACE)
namespace dmd {

constexpr int GK = 4096;  // candidate capacity per round
constexpr int GW = 4096;  // slot capacity (>= warps of the grid)

struct GridShared {
  double window, tlast;
  long long coll, target;
  long long st_rounds, st_exec, st_rollback, st_conflict;
  long long nevents[32];
  long long cyc[8];              // SM clocks per phase (thread 0): scan, rank, claim, check, exec, commit
  unsigned long long tmin_bits;  // order-preserving image of the calendar minimum (atomicMin)
  unsigned barrier;              // grid barrier arrival counter
  int n_cand, first_cold, first_lost;
  int status;                    // why the kernel returned: 0 target reached, 1 cold event at the head, 2 error
  int error, error_info;
  int n_log;
  int head_owner;                // owner of the calendar head when status == 1
  double cand_t[GK];
  double slot_t[GW], slot_newmin[GW];  // compact copies of slot[].t / slot[].newmin for the validation sweep
  int cand_o[GK];
  int by_rank[GK];
  BlkSlot slot[GW];
  BeadRec old_i[GW], old_j[GW];
  CalEnt undo_old[GW][BK_UNDO];
  int32_t undo_idx[GW][BK_UNDO];
};

__device__ __forceinline__ unsigned long long ord_bits(double v) {
  unsigned hi, lo;
  ord_split(v, hi, lo);
  return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ double ord_value(unsigned long long b) { return ord_join((unsigned)(b >> 32), (unsigned)b); }

// all CTAs of the (co-resident, cooperatively launched) grid meet here; `epoch` counts this CTA's arrivals
__device__ __forceinline__ void grid_barrier(GridShared& S, unsigned& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    __threadfence();
    atomicAdd(&S.barrier, 1u);
    while (*((volatile unsigned*)&S.barrier) < epoch) {
    }
    __threadfence();
  }
  __syncthreads();
}

// the round loop; gw / ngw = this warp's index in the grid / warps in the grid.  Warp 0 of the grid holds the
// master copy of coll / log position.
__device__ void grid_run(GridShared& S, Rep& r, uint32_t* claim, int gw, int ngw) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gnt = gridDim.x * blockDim.x;
  const bool t0 = gtid == 0;
  unsigned epoch = 0;
  constexpr int GRID_STAGE = 256;
  __shared__ int s_nvalid, s_ncand, s_base;
  __shared__ double s_ct[GRID_STAGE];
  __shared__ int s_co[GRID_STAGE];
  while (true) {
    grid_barrier(S, epoch);
    if (S.error || S.coll >= S.target) {
      if (t0) S.status = S.error ? 2 : 0;
      break;
    }
    const long long remaining = S.target - S.coll;
    long long tc = clock64();
    // ---- scan + select (one sweep when the time of the last committed event is known)
    double window = S.window;
    bool known = S.tlast >= 0.0;
    int nc = 0;
    int tries = 0;
    double tmin = 0.0;
    while (true) {
      const double lim = known ? S.tlast + window : (tries ? tmin + window : -1.0);
      double best = T_PAD;
      if (threadIdx.x == 0) s_ncand = 0;
      __syncthreads();
      for (int k = gtid; k < r.N + 3; k += gnt) {
        const CalEnt e = r.cal[k];
        if (e.t < best) best = e.t;
        if (e.t <= lim) {  // candidates are staged per CTA: one global reservation per CTA instead of one per candidate
          const int ct = type_of(e.type);
          const int code = (k < r.N && e.ptnr >= 0 && ct >= 1 && ct <= 3) ? k : -1 - k;  // negative: not a hot event
          const int pos = atomicAdd(&s_ncand, 1);
          if (pos < GRID_STAGE) {
            s_ct[pos] = e.t;
            s_co[pos] = code;
          } else {  // (a very dense window: straight to the global list)
            const int gp = atomicAdd(&S.n_cand, 1);
            if (gp < GK) {
              S.cand_t[gp] = e.t;
              S.cand_o[gp] = code;
            }
          }
        }
      }
      __syncthreads();
      {
        const int ns = s_ncand < GRID_STAGE ? s_ncand : GRID_STAGE;
        if (threadIdx.x == 0) s_base = ns ? atomicAdd(&S.n_cand, ns) : 0;
        __syncthreads();
        for (int q = threadIdx.x; q < ns; q += blockDim.x)
          if (s_base + q < GK) {
            S.cand_t[s_base + q] = s_ct[q];
            S.cand_o[s_base + q] = s_co[q];
          }
      }
      best = warp_min(best);
      if (Warp::lane() == 0) atomicMin(&S.tmin_bits, ord_bits(best));
      grid_barrier(S, epoch);
      tmin = ord_value(S.tmin_bits);
      nc = S.n_cand;
      if (nc > 0 && nc <= GK) break;
      if (!(tmin < 1e299) || tries > 80) break;
      grid_barrier(S, epoch);  // everybody has read n_cand
      if (nc > GK) window = window * 0.5;
      known = false;
      tries++;
      if (t0) {
        S.n_cand = 0;
        S.window = window;
      }
      grid_barrier(S, epoch);
    }
    if (!(tmin < 1e299) || nc <= 0 || nc > GK) {
      if (t0) {
        S.error = DMD_E_CAL_EMPTY;
        S.error_info = nc;
      }
      continue;
    }
    if (t0) S.cyc[0] += clock64() - tc, tc = clock64();
    // ---- rank in serial order (time, bead index): one warp per candidate
    for (int k = gw; k < nc; k += ngw) {
      const double t = S.cand_t[k];
      const int ok = S.cand_o[k], o = ok >= 0 ? ok : -1 - ok;
      int rank = 0;
      for (int m = Warp::lane(); m < nc; m += 32) {
        const double tm = S.cand_t[m];
        const int om = S.cand_o[m] >= 0 ? S.cand_o[m] : -1 - S.cand_o[m];
        if (tm < t || (tm == t && om < o)) rank++;
      }
      rank = warp_sum(rank);
      if (Warp::lane() == 0) {
        S.by_rank[rank] = k;
        if (ok < 0) atomicMin(&S.first_cold, rank);
      }
    }
    grid_barrier(S, epoch);
    int batch = nc < ngw ? nc : ngw;
    if (remaining < batch) batch = (int)remaining;
    if (S.first_cold < batch) batch = S.first_cold;
    if (batch == 0) {  // the head of the calendar is not a plain pair event: hand it to the host
      if (t0) {
        const int ok = S.cand_o[S.by_rank[0]];
        S.status = 1;
        S.head_owner = ok >= 0 ? ok : -1 - ok;
      }
      break;
    }
    if (t0) S.cyc[1] += clock64() - tc, tc = clock64();
    // ---- claim
    ListRef li, lj;
    li.up = li.dn = lj.up = lj.dn = nullptr;
    li.nu = li.nd = lj.nu = lj.nd = 0;
    if (gw < batch) {
      const int k = S.by_rank[gw];
      const int i = S.cand_o[k];
      const int j = r.cal[i].ptnr;
      li.nu = r.nup[i]; li.nd = r.ndn[i]; li.up = r.up + (size_t)i * r.cap; li.dn = r.dn + (size_t)i * r.cap;
      lj.nu = r.nup[j]; lj.nd = r.ndn[j]; lj.up = r.up + (size_t)j * r.cap; lj.dn = r.dn + (size_t)j * r.cap;
      blk_footprint(r, i, j, li, lj, [&](int b) { atomicMin(&claim[b], (uint32_t)gw); });
      if (Warp::lane() == 0) {
        BlkSlot& sl = S.slot[gw];
        sl.t = S.cand_t[k];
        sl.owner = i;
        sl.j = j;
        const int bound = 2 + 2 * (li.nd + 3) + 2 * (lj.nd + 3);
        sl.win = bound <= BK_UNDO || gw == 0;
      }
    }
    grid_barrier(S, epoch);
    if (t0) S.cyc[2] += clock64() - tc, tc = clock64();
    // ---- check
    if (gw < batch) {
      blk_phase_check(S, r, claim, gw, li, lj);
      if (Warp::lane() == 0 && !S.slot[gw].win) atomicMin(&S.first_lost, gw);
    }
    grid_barrier(S, epoch);
    if (t0) S.cyc[3] += clock64() - tc, tc = clock64();
    // ---- exec
    const int n_exec = S.first_lost < batch ? S.first_lost : batch;  // >= 1
    if (gw < batch) {
      const BlkSlot& sl = S.slot[gw];
      blk_footprint(r, sl.owner, sl.j, li, lj, [&](int b) { claim[b] = CLAIM_FREE; });
    }
    if (gw < n_exec) {
      blk_exec_event(S, r, gw, li, lj);
      if (Warp::lane() == 0) {
        S.slot_t[gw] = S.slot[gw].t;
        S.slot_newmin[gw] = S.slot[gw].newmin;
      }
    }
    grid_barrier(S, epoch);
    if (t0) S.cyc[4] += clock64() - tc, tc = clock64();
    // ---- validate (main.F90:970-993; warp 0 of every CTA redundantly: running prefix minimum of newmin)
    if (threadIdx.x < 32) {
      int nv = n_exec;
      double carry = T_PAD;  // minimum of newmin over all earlier slots
      for (int base = 0; base < n_exec && nv == n_exec; base += 32 * 8) {
        // eight chunks of 32 slots per trip: their loads are independent and issued together
        double tt[8], mm[8];
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const int q = base + c * 32 + Warp::lane();
          tt[c] = q < n_exec ? S.slot_t[q] : T_PAD;
          mm[c] = q < n_exec ? S.slot_newmin[q] : T_PAD;
        }
#pragma unroll
        for (int c = 0; c < 8; c++) {
          if (nv != n_exec || base + c * 32 >= n_exec) break;
          const int q = base + c * 32 + Warp::lane();
          double pm = mm[c];
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const double o = Warp::shfl(pm, Warp::lane() >= d ? Warp::lane() - d : Warp::lane());
            if (Warp::lane() >= d && o < pm) pm = o;
          }
          double before = Warp::shfl(pm, Warp::lane() > 0 ? Warp::lane() - 1 : 0);
          if (Warp::lane() == 0) before = T_PAD;
          if (carry < before) before = carry;
          const bool bad = q > 0 && q < n_exec && !(tt[c] < before);
          const unsigned m = Warp::ballot(bad);
          if (m) nv = base + c * 32 + dmd_ffs(m) - 1;
          const double last = Warp::shfl(pm, 31);
          if (last < carry) carry = last;
        }
      }
      if (Warp::lane() == 0) s_nvalid = nv;
    }
    __syncthreads();
    const int n_valid = s_nvalid;
    if (gw >= n_valid && gw < n_exec) blk_rollback(S, r, gw);
    if (gw == 0) {  // commit: tallies, log, counters (lanes over the slots)
      const int log_cap = r.c.sys->log_cap;
      int c1 = 0, c2 = 0, c3 = 0;  // batch events are types 1, 2, 3 only
      for (int q = Warp::lane(); q < n_valid; q += 32) {
        const BlkSlot& sl = S.slot[q];
        c1 += sl.ct == 1;
        c2 += sl.ct == 2;
        c3 += sl.ct == 3;
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
      c1 = warp_sum(c1); c2 = warp_sum(c2); c3 = warp_sum(c3);
      if (Warp::lane() == 0) {
        S.nevents[1] += c1;
        S.nevents[2] += c2;
        S.nevents[3] += c3;
      }
      if (r.n_log < log_cap) r.n_log = r.n_log + n_valid < log_cap ? r.n_log + n_valid : log_cap;
      r.coll += n_valid;
      r.tfalse = S.slot[n_valid - 1].t;
      r.old_tfalse = r.tfalse;
      if (Warp::lane() == 0) {
        S.coll = r.coll;
        S.n_cand = 0;
        S.first_cold = 0x7fffffff;
        S.first_lost = 0x7fffffff;
        S.tmin_bits = ~0ull;
        S.tlast = r.tfalse;
        S.st_rounds += 1;
        S.st_exec += n_exec;
        S.st_rollback += n_exec - n_valid;
        S.st_conflict += batch - n_exec;
        // steer the window towards ~2 candidates per expected batch member (the batch is cut at the first conflict)
        const int want = n_exec * 2 + 64;
        if (nc < want) S.window = window * 1.25;
        else if (nc > 2 * want) S.window = window * 0.8;
        else S.window = window;
        S.cyc[5] += clock64() - tc;
      }
    }
    if (r.error && Warp::lane() == 0) {
      S.error = r.error;
      S.error_info = r.error_info;
    }
  }
}

}  // namespace dmd
#endif
