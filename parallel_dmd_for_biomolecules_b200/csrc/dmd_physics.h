// dmd_physics.h -- scalar fp64 physics of the PRIME20 DMD hot path, written for the device (one lane = one
// pair).  Every function names the reference routine it replaces (paths relative to
// /root/reference/parallel-dmd-PRIME20/code).  Arithmetic contract (DESIGN.md "fp64 discipline"): IEEE fp64,
// operations in the order the Fortran writes them, NO FMA contraction (nvcc -fmad=false), dnint == round().
#pragma once
#include "dmd_math.h"
#include "dmd_types.h"

namespace dmd {

constexpr double LTSTEP = 1e-10;   // def.h:1
constexpr double SMDIST = 5e-12;   // def.h:2
constexpr double T_NONE = 1000000000.0;  // events.f:28

struct Ctx {                 // read-only context of one warp
  const SysConst* sys;       // global memory (cold fields only)
  const HotTables* tab;      // shared-memory copy of sigma_sq / welldia_sq (the other two tables: sys, global memory)
  const HotConst* hot;       // shared-memory copy of the squeeze factors, bond windows of codes 4-9, masses
  const double* bl;          // per-residue side-chain bond windows (shared memory when nres <= HOT_MAX_RES)
  const uint32_t* meta;
  const int32_t* chain;
  const uint8_t* sctab;      // same-chain static classes (list rebuild), or nullptr
};

DMD_DEV int tix(int idi, int idj) { return (idi - 1) * 28 + (idj - 1); }

// minimum-image pair geometry at the current false time (core.f:14-23 and every other predictor)
struct Geom {
  double vx, vy, vz, rx, ry, rz, bij;
};
DMD_DEV Geom pair_geom(const BeadRec& a, const BeadRec& b, double tfalse) {
  Geom g;
  g.vx = a.vx - b.vx;
  g.vy = a.vy - b.vy;
  g.vz = a.vz - b.vz;
  g.rx = a.x - b.x + g.vx * tfalse;
  g.ry = a.y - b.y + g.vy * tfalse;
  g.rz = a.z - b.z + g.vz * tfalse;
  g.rx = g.rx - dmd_round(g.rx);
  g.ry = g.ry - dmd_round(g.ry);
  g.rz = g.rz - dmd_round(g.rz);
  g.bij = g.rx * g.vx + g.ry * g.vy + g.rz * g.vz;
  return g;
}

// x / m for a bead mass m, bit-identical to the IEEE division the Fortran (and the oracle) performs, in five
// multiply-adds instead of the ~25-instruction division sequence -- the collision update of eventdyn.f:366-381
// divides by a mass twelve times per event.  minv = RN(1/m) (HotConst.binv).  q0 = RN(x minv) is within 2 ulp of
// x/m; one residual step makes it faithful, the second gives the correctly rounded quotient (Markstein's theorem:
// y correctly rounded, q faithful, r = x - q m exact by FMA  =>  RN(q + r y) = RN(x/m)).  Checked against `/` on
// 1.9e9 random operands for every mass of parameters/mass.data, and by every event-sequence test.
DMD_DEV double div_mass(double x, double m, double minv) {
  const double q0 = x * minv;
  const double r0 = dmd_fma(-q0, m, x);
  const double q1 = dmd_fma(r0, minv, q0);
  const double r1 = dmd_fma(-q1, m, x);
  return dmd_fma(r1, minv, q1);
}

// bond-length window of bond.f:30-41 / :80-91 for event classes 4-12; `mi` = topology word of the
// lower-index (backbone) bead.
DMD_DEV void bond_limits(const Ctx& c, int code, uint32_t mi, double& blmin, double& blmax) {
  if (code >= 10) {
    const int r = (meta_sp(mi) ? c.hot->chnln0 : 0) + meta_res(mi) - 1;
    blmin = c.bl[r * 6 + 2 * (code - 10)];
    blmax = c.bl[r * 6 + 2 * (code - 10) + 1];
  } else {
    blmin = c.hot->ev_param2[code];
    blmax = c.hot->ev_param3[code];
  }
}

// hard-core diameter^2 of core.f:27-31 / eventdyn.f:70-74 / checkover.f:38-42
DMD_DEV double core_sigsq(const Ctx& c, int code, int idi, int idj) {
  double f = c.hot->ev_param1[code];
  double sigsq = c.tab->sigma_sq[tix(idi, idj)] * (f * f);
  if (code >= 22 && code <= 26) {
    int k = idi > idj ? idi : idj;
    double q = c.hot->sqz610[(code - 22) * 29 + k];
    sigsq = sigsq * (q * q);
  }
  return sigsq;
}

// One pair-time prediction = the dispatch of events.f:30-48 + core.f / bond.f / sqwel.f / nc_sqwel.f /
// sqshlder.f.  `a` is the lower-index bead (event owner), `bonded` = (bptnr(a) == b).  Leaves tij/type
// untouched when the pair has no event (the Fortran leaves tij at its 1e9 preset).
// The predictors only use bij, |r|^2 and |v|^2, which are bitwise identical for (a,b) and (b,a) (negation is
// exact), so the caller may form the geometry in either order; idi/idj/meta_a/bonded refer to the lower-index
// bead first, as in the Fortran call core(i,j,...).
//
// All five Fortran predictors are the same computation with different radii and type labels:
//   d1 = b^2 - v^2 (r^2 - R1^2)   inner radius R1: hard core (core.f:34), bond minimum (bond.f:46)
//   d2 = b^2 - v^2 (r^2 - R2^2)   outer radius R2: bond maximum (bond.f:53), well (sqwel.f:41), shoulder
//   t  = (-b -/+ sqrt(d)) / v^2,  or the cancellation-free -(r^2-R2^2)/(sqrt(d2)+b) for a receding bond (bond.f:63)
// so the lanes of a warp (which hold pairs of different classes) select radii / root / type with predicates and
// execute ONE sqrt and ONE division instead of diverging through five routines.  Every selected expression is
// exactly the one the Fortran evaluates for that case.
DMD_DEV void pair_time_core(const Ctx& c, int code, double bij, double rijsq, double vijsq, int idi, int idj,
                            uint32_t meta_a, bool bonded, double& tij, int& type) {
  double R1sq, R2sq = 0.0;
  int t_inner = 1, t_leave = 0, t_enter = 0;
  int cls;                 // 0 core only, 1 bond, 2 well-like
  int inside_force = 0;    // +1 treat as inside the well, -1 treat as outside
  if (code <= 3 || (code >= 17 && code <= 26)) {  // core.f
    cls = 0;
    R1sq = core_sigsq(c, code, idi, idj);
  } else if (code <= 12) {  // bond.f
    cls = 1;
    double blmin, blmax;
    bond_limits(c, code, meta_a, blmin, blmax);
    R1sq = blmin * blmin;
    R2sq = blmax * blmax;
    t_inner = 2;
  } else {
    cls = 2;
    const double sig = c.tab->sigma_sq[tix(idi, idj)];
    R1sq = sig;
    if (code == 16) {  // sqwel.f
      R2sq = c.tab->welldia_sq[tix(idi, idj)];
      t_leave = 8;
      t_enter = 4;
    } else if (code == 15) {  // nc_sqwel.f
      R2sq = c.tab->welldia_sq[tix(idi, idj)];
      if (idi + idj == 5) {  // free N and free C
        t_leave = 16;
        t_enter = 7;
      } else if (bonded) {  // bound to each other: core at the 1.05*(2.24 A) factor, exit 8, no inside test
        const double f = c.hot->ev_param1[15];
        R1sq = sig * f * f;
        t_leave = 8;
        inside_force = 1;
      } else {  // not eligible: hard wall at the well diameter
        t_enter = 9;
        inside_force = -1;
      }
    } else {  // sqshlder.f
      R2sq = c.sys->shlddia_sq[tix(idi, idj)];
      t_leave = 10;
      t_enter = 12;
    }
  }
  const bool approaching = bij < 0.0;
  const double diff2 = rijsq - R2sq;
  const bool inside = inside_force > 0 || (inside_force == 0 && diff2 < 0.0);
  const double bb = bij * bij;
  const double d1 = bb - vijsq * (rijsq - R1sq);
  const double d2 = bb - vijsq * diff2;
  bool valid, alt = false;
  double d, sgn;
  int ty;
  if (approaching && d1 > 0.0 && (cls != 2 || inside)) {  // inner root: core hit / bond minimum
    valid = true;
    d = d1;
    sgn = -1.0;
    ty = t_inner;
  } else if (cls == 0) {
    valid = false;
    d = 1.0;
    sgn = 1.0;
    ty = -1;
  } else if (cls == 1) {  // bond maximum
    valid = d2 > 0.0;
    d = d2;
    sgn = 1.0;
    ty = 3;
    alt = !approaching;
  } else if (approaching && !inside) {  // capture / enter (or wall) at R2 from outside
    valid = d2 > 0.0;
    d = d2;
    sgn = -1.0;
    ty = t_enter;
  } else {  // leave the well / shoulder at R2 from inside
    valid = inside;
    d = d2;
    sgn = 1.0;
    ty = t_leave;
  }
  if (valid) {
    const double sq = dmd_sqrt(d);
    const double num = alt ? -diff2 : (-bij + sgn * sq);
    const double den = alt ? (sq + bij) : vijsq;
    tij = num / den;
    type = ty;
  }
}

DMD_DEV void pair_time(const Ctx& c, int code, const BeadRec& a, const BeadRec& b, uint32_t meta_a, bool bonded,
                       double tfalse, double& tij, int& type) {
  const Geom g = pair_geom(a, b, tfalse);
  const double rijsq = g.rx * g.rx + g.ry * g.ry + g.rz * g.rz;
  const double vijsq = g.vx * g.vx + g.vy * g.vy + g.vz * g.vz;
  pair_time_core(c, code, g.bij, rijsq, vijsq, a.ident, b.ident, meta_a, bonded, tij, type);
}

// distance between two beads at the current false time (repuls_check.f:32-42)
DMD_DEV double pair_dist(const BeadRec& a, const BeadRec& b, double tfalse) {
  Geom g = pair_geom(a, b, tfalse);
  double d = g.rx * g.rx + g.ry * g.ry + g.rz * g.rz;
  return dmd_sqrt(d);
}

// eventdyn.f:18-381.  `a` owner (lower index), `b` partner; ct = resolved coltype (< 14).  Returns the
// executed type (20..27 or the unchanged 1/2/3).  Updates positions (bump + rewind) and velocities.
DMD_DEV int event_dynamics(const Ctx& c, int ct, int code, BeadRec& a, BeadRec& b, uint32_t meta_a, bool bonded,
                           double tfalse) {
  const Geom g = pair_geom(a, b, tfalse);
  const double rxij = g.rx, ryij = g.ry, rzij = g.rz, bij = g.bij;
  const int idi = a.ident, idj = b.ident;
  const double bmi = c.hot->bmass[idi], bmj = c.hot->bmass[idj];
  const double rmass = 2 * bmi * bmj / (bmi + bmj);
  double ratio = 0.0, bumpdist = 0.0, sgn = 0.0;  // sgn +1: a += bump*r, b -= ; -1: the opposite
  if (ct == 2 || ct == 3) {
    double blmin, blmax;
    bond_limits(c, code, meta_a, blmin, blmax);
    ratio = ct == 2 ? rmass * bij / (blmin * blmin) : rmass * bij / (blmax * blmax);
  } else if (ct == 1) {
    double sigsq;
    if (code == 15) {
      double f = c.hot->ev_param1[15];
      sigsq = bonded ? c.tab->sigma_sq[tix(idi, idj)] * (f * f) : c.tab->sigma_sq[tix(idi, idj)];
    } else {
      sigsq = core_sigsq(c, code, idi, idj);
    }
    ratio = rmass * bij / sigsq;
  } else if (ct == 4 || ct == 8 || ct == 9) {
    double wellsq = c.tab->welldia_sq[tix(idi, idj)];
    double epsave = c.sys->ep_sqrt[tix(idi, idj)];
    double del_pe = 4.0 * wellsq * epsave / rmass;
    bumpdist = SMDIST * dmd_sqrt(wellsq);
    if (ct == 4) {
      if (bij * bij + del_pe > 0.0) {
        ratio = rmass * (dmd_sqrt((4.0 * wellsq * epsave / rmass) + bij * bij) + bij) / (2.0 * wellsq);
        ct = 20;
        sgn = -1.0;
      } else {
        ratio = rmass * bij / wellsq;
        ct = 22;
        sgn = 1.0;
      }
    } else if (ct == 8) {
      if (bij * bij > del_pe) {
        ratio = rmass * (-dmd_sqrt(-del_pe + bij * bij) + bij) / (2.0 * wellsq);
        ct = 21;
        sgn = 1.0;
      } else {
        ratio = rmass * bij / wellsq;
        ct = 22;
        sgn = -1.0;
      }
    } else {
      ratio = rmass * bij / wellsq;
      ct = 23;
      sgn = 1.0;
    }
  } else if (ct == 5 || ct == 6 || ct == 13) {
    double wellsq = c.sys->shlddia_sq[tix(idi, idj)];
    double epsave = -c.sys->eps1;
    double del_pe = 4.0 * wellsq * epsave / rmass;
    bumpdist = SMDIST * dmd_sqrt(wellsq);
    if (ct == 5) {
      ratio = rmass * (-dmd_sqrt(-del_pe + bij * bij) + bij) / (2.0 * wellsq);
      ct = 24;
      sgn = 1.0;
    } else if (ct == 6) {
      if (bij * bij > -del_pe) {
        ratio = rmass * (dmd_sqrt(del_pe + bij * bij) + bij) / (2.0 * wellsq);
        ct = 25;
        sgn = -1.0;
      } else {
        ratio = rmass * bij / wellsq;
        ct = 26;
        sgn = 1.0;
      }
    } else {
      ratio = rmass * bij / wellsq;
      ct = 27;
      sgn = 1.0;
    }
  }
  if (sgn != 0.0) {  // the 5e-12*d "bump" off the discontinuity, eventdyn.f:85-91 etc.
    a.x = a.x + sgn * (bumpdist * rxij);
    a.y = a.y + sgn * (bumpdist * ryij);
    a.z = a.z + sgn * (bumpdist * rzij);
    b.x = b.x - sgn * (bumpdist * rxij);
    b.y = b.y - sgn * (bumpdist * ryij);
    b.z = b.z - sgn * (bumpdist * rzij);
  }
  const double delvx = ratio * rxij, delvy = ratio * ryij, delvz = ratio * rzij;  // eventdyn.f:366-381
  const double ii = c.hot->binv[idi], ij = c.hot->binv[idj];
  a.vx = a.vx - div_mass(delvx, bmi, ii);
  b.vx = b.vx + div_mass(delvx, bmj, ij);
  a.vy = a.vy - div_mass(delvy, bmi, ii);
  b.vy = b.vy + div_mass(delvy, bmj, ij);
  a.vz = a.vz - div_mass(delvz, bmi, ii);
  b.vz = b.vz + div_mass(delvz, bmj, ij);
  a.x = a.x + div_mass(delvx * tfalse, bmi, ii);
  a.y = a.y + div_mass(delvy * tfalse, bmi, ii);
  a.z = a.z + div_mass(delvz * tfalse, bmi, ii);
  b.x = b.x - div_mass(delvx * tfalse, bmj, ij);
  b.y = b.y - div_mass(delvy * tfalse, bmj, ij);
  b.z = b.z - div_mass(delvz * tfalse, bmj, ij);
  return ct;
}

// the > 99 % case of eventdyn.f: hard-core (1) and bond (2, 3) events -- no bump, no type change.
// Same arithmetic as event_dynamics(); kept separate so the hot loop stays small.
DMD_DEV int event_dynamics_hot(const Ctx& c, int ct, int code, BeadRec& a, BeadRec& b, uint32_t meta_a, bool bonded,
                               double tfalse) {
  const Geom g = pair_geom(a, b, tfalse);
  const double rxij = g.rx, ryij = g.ry, rzij = g.rz, bij = g.bij;
  const int idi = a.ident, idj = b.ident;
  const double bmi = c.hot->bmass[idi], bmj = c.hot->bmass[idj];
  const double rmass = 2 * bmi * bmj / (bmi + bmj);
  double ratio;
  if (ct == 1) {
    double sigsq;
    if (code == 15) {
      double f = c.hot->ev_param1[15];
      sigsq = bonded ? c.tab->sigma_sq[tix(idi, idj)] * (f * f) : c.tab->sigma_sq[tix(idi, idj)];
    } else {
      sigsq = core_sigsq(c, code, idi, idj);
    }
    ratio = rmass * bij / sigsq;
  } else {
    double blmin, blmax;
    bond_limits(c, code, meta_a, blmin, blmax);
    ratio = ct == 2 ? rmass * bij / (blmin * blmin) : rmass * bij / (blmax * blmax);
  }
  const double delvx = ratio * rxij, delvy = ratio * ryij, delvz = ratio * rzij;  // eventdyn.f:366-381
  const double ii = c.hot->binv[idi], ij = c.hot->binv[idj];
  a.vx = a.vx - div_mass(delvx, bmi, ii);
  b.vx = b.vx + div_mass(delvx, bmj, ij);
  a.vy = a.vy - div_mass(delvy, bmi, ii);
  b.vy = b.vy + div_mass(delvy, bmj, ij);
  a.vz = a.vz - div_mass(delvz, bmi, ii);
  b.vz = b.vz + div_mass(delvz, bmj, ij);
  a.x = a.x + div_mass(delvx * tfalse, bmi, ii);
  a.y = a.y + div_mass(delvy * tfalse, bmi, ii);
  a.z = a.z + div_mass(delvz * tfalse, bmi, ii);
  b.x = b.x - div_mass(delvx * tfalse, bmj, ij);
  b.y = b.y - div_mass(delvy * tfalse, bmj, ij);
  b.z = b.z - div_mass(delvz * tfalse, bmj, ij);
  return ct;
}

// bumped.f:12-43
DMD_DEV void bump_off(const Ctx& c, int code, BeadRec& a, BeadRec& b, double tfalse) {
  const Geom g = pair_geom(a, b, tfalse);
  double dsq = code >= 40 ? c.sys->shlddia_sq[tix(a.ident, b.ident)] : c.tab->welldia_sq[tix(a.ident, b.ident)];
  double bumpdist = SMDIST * dmd_sqrt(dsq);
  double sgn = g.bij < 0.0 ? -1.0 : 1.0;
  a.x = a.x + sgn * (bumpdist * g.rx);
  a.y = a.y + sgn * (bumpdist * g.ry);
  a.z = a.z + sgn * (bumpdist * g.rz);
  b.x = b.x - sgn * (bumpdist * g.rx);
  b.y = b.y - sgn * (bumpdist * g.ry);
  b.z = b.z - sgn * (bumpdist * g.rz);
}

// counter RNG replacing Intel IFPORT drandm (main.F90:173,412,998,1009-1010,1041,1510): splitmix64 of
// (seed + n*golden) -> 53-bit uniform in [0,1).  Shared bit-for-bit with the oracle.
DMD_DEV double rng_uniform(uint64_t seed, uint64_t& ctr) {
  ctr += 1;
  uint64_t z = seed + ctr * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// natural log for the ghost thermostat (main.F90:1016,1044): the fdlibm e_log.c algorithm with plain
// (unfused) fp64 operations so host oracle and device agree bit for bit.  x must be positive and finite.
DMD_DEV double dmd_log(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
               Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
               Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
  int hx = dmd_hi(x);
  unsigned lx = dmd_lo(x);
  int k = 0;
  if (hx < 0x00100000) {
    x *= 1.80143985094819840000e+16;
    k -= 54;
    hx = dmd_hi(x);
    lx = dmd_lo(x);
  }
  k += (hx >> 20) - 1023;
  hx &= 0x000fffff;
  int i = (hx + 0x95f64) & 0x100000;
  x = dmd_hi_lo(hx | (i ^ 0x3ff00000), lx);
  k += (i >> 20);
  double f = x - 1.0, dk;
  if ((0x000fffff & (2 + hx)) < 3) {
    if (f == 0.0) {
      if (k == 0) return 0.0;
      dk = (double)k;
      return dk * ln2_hi + dk * ln2_lo;
    }
    double R = f * f * (0.5 - 0.33333333333333333 * f);
    if (k == 0) return f - R;
    dk = (double)k;
    return dk * ln2_hi - ((R - dk * ln2_lo) - f);
  }
  double s = f / (2.0 + f);
  dk = (double)k;
  double z = s * s;
  i = hx - 0x6147a;
  double w = z * z;
  int j = 0x6b851 - hx;
  double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  i |= j;
  double R = t2 + t1;
  if (i > 0) {
    double hfsq = 0.5 * f * f;
    if (k == 0) return f - (hfsq - s * (hfsq + R));
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
  }
  if (k == 0) return f - s * (f - R);
  return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

}  // namespace dmd
