// dmd_topology.h -- the static event class of a bead pair as a FUNCTION of topology.
//
// The reference stores ev_code as an N x N int8 matrix (header.f:14; 1.8 MB at N=1344, 1 TB at 1e6 beads)
// filled by make_code.f:73-566, where later loops overwrite earlier ones.  The closed form below is the net
// effect of that assignment order and is checked pair-by-pair against the oracle's literal
// matrix in tests/test_evcode.py.  The dynamic 40/50 overlay of repuls_add.f:30-37 lives in BeadRec.ov1/ov2.
#pragma once
#include "dmd_types.h"

namespace dmd {

DMD_HD bool code_is_bonded_class(int code) {  // nbor.f:60: neighbours unconditionally when found in the stencil
  return (code >= 4 && code <= 12) || (code >= 17 && code < 27);
}

// ma/mb topology words, ca/cb global chain ids, ia/ib bead indices (only their order is used)
DMD_HD int static_code(const SysConst& s, uint32_t ma, int ca, int ia, uint32_t mb, int cb, int ib) {
  int cls_a = meta_cls(ma), cls_b = meta_cls(mb);
  if (ca != cb) {
    if (cls_a == 3 && cls_b == 3) return (meta_hp(ma) && meta_hp(mb)) ? 16 : 1;  // make_code.f:83-110
    if (!s.no_hbs && ((cls_a == 1 && cls_b == 2) || (cls_a == 2 && cls_b == 1))) {  // make_code.f:117-125
      // proline N-H exclusion, make_code.f:146-168: effective only when the N bead precedes the C bead
      int n_is_a = cls_a == 1;
      uint32_t mn = n_is_a ? ma : mb;
      int in = n_is_a ? ia : ib, ic = n_is_a ? ib : ia;
      if (meta_proex(mn) && in < ic) return 1;
      return 15;
    }
    return 1;
  }
  // same chain: order the pair by class so that (cls_a <= cls_b)
  int ra = meta_res(ma), rb = meta_res(mb);
  if (cls_a > cls_b) {
    int t = cls_a; cls_a = cls_b; cls_b = t;
    t = ra; ra = rb; rb = t;
    uint32_t tm = ma; ma = mb; mb = tm;
  }
  int d = rb - ra;  // residue of the higher class minus residue of the lower class
  switch (cls_a * 4 + cls_b) {
    case 0:  // Ca-Ca, make_code.f:247-249
      return (d == 1 || d == -1) ? 9 : 1;
    case 1:  // Ca_r - N_s : 4 (s=r), 7 (s=r+1), 18 (s=r-1)   make_code.f:211-213, 229-231, 277-279
      return d == 0 ? 4 : (d == 1 ? 7 : (d == -1 ? 18 : 1));
    case 2:  // Ca_r - C_s : 5 (s=r), 8 (s=r-1), 17 (s=r+1)   make_code.f:217-219, 235-237, 271-273
      return d == 0 ? 5 : (d == -1 ? 8 : (d == 1 ? 17 : 1));
    case 3:  // Ca_r - R_s : 10 (s=r), 24 (s=r+1), 25 (s=r-1)  make_code.f:253-255, 313-321
      return d == 0 ? 10 : (d == 1 ? 24 : (d == -1 ? 25 : 1));
    case 5:  // N-N, make_code.f:289-291
      return (d == 1 || d == -1) ? 20 : 1;
    case 6:  // N_r - C_s : 8 (s=r), 6 (s=r-1), 19 (s=r-2), 15 (|s-r|>=4)  make_code.f:194-206, 223-225, 241-243, 283-285
      if (d == 0) return 8;
      if (d == -1) return 6;
      if (d == -2) return 19;
      if ((d >= 4 || d <= -4) && !s.no_hbs) return 15;
      return 1;
    case 7:  // N_r - R_s : 11 (s=r), 23 (s=r-1)   make_code.f:259-261, 307-309
      return d == 0 ? 11 : (d == -1 ? 23 : 1);
    case 10:  // C-C, make_code.f:295-297
      return (d == 1 || d == -1) ? 21 : 1;
    case 11:  // C_r - R_s : 12 (s=r), 22 (s=r+1), 26 (s=r+2)   make_code.f:265-267, 301-303, 325-327
      return d == 0 ? 12 : (d == 1 ? 22 : (d == 2 ? 26 : 1));
    case 15:  // R-R same chain: residue separation >= 4 and both hydrophobic, make_code.f:178-190
      return ((d >= 4 || d <= -4) && meta_hp(ma) && meta_hp(mb)) ? 16 : 1;
  }
  return 1;
}

DMD_HD int ov_mirror(int code) { return code == 40 ? 50 : (code == 50 ? 40 : code); }

// ev_code(a,b) (ROW a) including the 40/50 overlay; sc = static class of the pair
DMD_HD int overlay_code(int sc, int ia, const BeadRec& a, int ib, const BeadRec& b) {
  if (sc != 1) return sc;
  if (a.er1 == ib) return a.ov1;
  if (a.er2 == ib) return a.ov2;
  if (b.er1 == ia) return ov_mirror(b.ov1);
  if (b.er2 == ia) return ov_mirror(b.ov2);
  return 1;
}

// terminal N (residue 1) / terminal C (residue L) test of main.F90:1491,1518,1545
DMD_HD bool is_terminal_bead(const SysConst& s, uint32_t m) {
  int cls = meta_cls(m), r = meta_res(m);
  return (cls == 1 && r == 1) || (cls == 2 && r == s.chnln[meta_sp(m)]);
}

}  // namespace dmd
