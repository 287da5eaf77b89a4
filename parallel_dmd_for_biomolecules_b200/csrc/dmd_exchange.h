// dmd_exchange.h -- the replica-exchange decision (host + device; new functionality, SURVEY.md 8e: the reference
// runs its temp_0xx files one after the other by hand, qfile/script.sh:11-18).
//
// M = world x R replicas are gathered rank-major: entry g = rank * R + r holds (E_pot, T*) of local replica r of
// that rank.  Ladders of L replicas are cut from the SLOT order s = r * world + rank, so that the members of a ladder
// -- and in particular temperature neighbours -- sit on different GPUs whenever world > 1.  Inside a ladder the
// replicas are ordered by their current temperature (stable), and the neighbours (k, k+1), k of the step's parity,
// swap TEMPERATURES with probability min(1, exp((beta_a - beta_b)(E_a - E_b))), beta = 1 / (12 T*) (the engine's
// energy unit: k_B T = setemp = 12 T*, main.F90:127).  The uniform for pair k of the ladder starting at slot s0 is
// draw number s0 * 131 + k + 1 of the counter RNG (dmd_physics.h) seeded with seed + 7919 * step: a pure function of
// the gathered data, so every rank reaches the same decision without talking.
#pragma once
#include <math.h>

#include "dmd_physics.h"

namespace dmd {

constexpr int XCH_MAX_LADDER = 32;

struct XchCounts {
  int32_t attempted, accepted, changed_local, ladders;
};

DMD_HD int xch_slot_to_gathered(int s, int world, int R) { return (s % world) * R + s / world; }

// one ladder: reads et[2 g] = E_pot, et[2 g + 1] = T*; writes tnew[g] for its members
DMD_DEV void xch_decide_ladder(const double* et, double* tnew, int ladder, int L, int world, int R, long long step,
                               uint64_t seed, int& attempted, int& accepted) {
  int g[XCH_MAX_LADDER], order[XCH_MAX_LADDER];
  double T[XCH_MAX_LADDER], E[XCH_MAX_LADDER];
  const int s0 = ladder * L;
  for (int m = 0; m < L; m++) {
    g[m] = xch_slot_to_gathered(s0 + m, world, R);
    E[m] = et[2 * g[m]];
    T[m] = et[2 * g[m] + 1];
    int p = m;  // stable insertion by temperature
    while (p > 0 && T[order[p - 1]] > T[m]) {
      order[p] = order[p - 1];
      p--;
    }
    order[p] = m;
  }
  for (int k = (int)(step & 1); k + 1 < L; k += 2) {
    const int a = order[k], b = order[k + 1];
    attempted++;
    const double ta = T[a], tb = T[b];
    if (ta == tb) continue;
    const double delta = (1.0 / (12.0 * ta) - 1.0 / (12.0 * tb)) * (E[a] - E[b]);
    bool swap = delta >= 0.0;
    if (!swap) {
      uint64_t ctr = (uint64_t)s0 * 131u + (uint64_t)k;  // rng_uniform pre-increments: draw number s0 * 131 + k + 1
      swap = rng_uniform(seed + 7919u * (uint64_t)step, ctr) < exp(delta);
    }
    if (swap) {
      T[a] = tb;
      T[b] = ta;
      accepted++;
    }
  }
  for (int m = 0; m < L; m++) tnew[g[m]] = T[m];
}

}  // namespace dmd
