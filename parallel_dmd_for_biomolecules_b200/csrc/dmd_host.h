// dmd_host.h -- host-side model set-up of libdmdb200: turns the reference's raw parameter tables and chain
// topology into the constant blocks the device engine reads.  Replaces (runs once, not on the hot path):
//   inputinfo.f:105-411 (tables, topology), scale_down.f:27-79, make_code.f:18-66 (ev_param),
//   nbor_setup.f:13-118 (cut-offs), main.F90:389-396 (cell grid).  The run start itself (main.F90:143-156, 205-321,
//   408-423) is done on the device by init_replica() in dmd_engine.h.
#pragma once
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dmdb200.h"
#include "dmd_topology.h"
#include "dmd_types.h"

namespace dmd {

struct HostModel {
  SysConst sys;
  PairTables tab;
  HotConst hot;
  std::vector<double> bl;  // nres x 6
  std::vector<uint32_t> meta;
  std::vector<int32_t> chain;
  std::vector<uint8_t> id0;
  std::vector<int32_t> nc_beads;  // indices of the N and C beads (ascending): the only H-bond capable ones
  std::vector<uint8_t> sctab;     // static_code() over one chain of each species (chains of <= 256 beads)
  dmdb_params params;
};

inline double hsq(double x) { return x * x; }

inline void build_model(const dmdb_params& p, const dmdb_topology& topo, const dmdb_tables& tab, HostModel& m) {
  if (topo.n_species < 1 || topo.n_species > 2) throw std::runtime_error("n_species must be 1 or 2");
  if (!(p.boxl > 0) || !(p.tstar > 0)) throw std::runtime_error("boxl and tstar must be positive");
  m.params = p;
  SysConst& s = m.sys;
  std::memset(&s, 0, sizeof(s));
  s.n_species = topo.n_species;
  int N = 0;
  for (int k = 0; k < topo.n_species; k++) {
    s.nch[k] = topo.n_chains[k];
    s.chnln[k] = topo.chnln[k];
    s.numbeads[k] = topo.numbeads[k];
    if (s.chnln[k] < 1 || s.numbeads[k] < 3 * s.chnln[k] || s.numbeads[k] > 4 * s.chnln[k] || s.nch[k] < 0)
      throw std::runtime_error("inconsistent chain sizes");
    if (s.chnln[k] > 1000) throw std::runtime_error("chain too long for the topology word (max 1000 residues)");
    N += s.nch[k] * s.numbeads[k];
  }
  if (s.chnln[0] + s.chnln[1] > MAX_RES) throw std::runtime_error("too many residues per chain (MAX_RES)");
  if (N >= (1 << NB_SHIFT)) throw std::runtime_error("too many beads for the neighbour-entry packing");
  s.N = N;
  s.nop1 = s.nch[0] * s.numbeads[0];
  s.n_wrap = p.n_wrap ? p.n_wrap : 2;
  if (s.n_wrap != 2) throw std::runtime_error("only n_wrap=2 (the shipped build, qfile/script.sh:7) is supported");
  s.canon = p.canon != 0;
  s.no_hbs = p.no_hbs != 0;
  s.cap = p.nbr_capacity > 0 ? p.nbr_capacity : 64;
  if (s.cap > 65535) throw std::runtime_error("nbr_capacity too large");
  s.ngroups = (N + 3 + 31) / 32;
  s.log_cap = p.log_capacity > 0 ? p.log_capacity : 0;
  s.out_cap = 64;

  // ---- inputinfo.f:162-207, 373-378 then scale_down.f:27-30
  double sigma[29] = {0}, welldia[29] = {0}, epsilon[29] = {0};
  for (int k = 1; k <= 4; k++) {
    sigma[k] = tab.protein[k - 1];
    welldia[k] = tab.protein[4 + k - 1];
    epsilon[k] = tab.protein[8 + k - 1];
    sigma[k + 4] = sigma[k];
    welldia[k + 4] = welldia[k];
    epsilon[k + 4] = epsilon[k];
  }
  auto T = [](const double* a, int i, int j) { return a[(i - 9) * 20 + (j - 9)]; };
  for (int i = 9; i <= 28; i++) {
    sigma[i] = 1.00 * T(tab.bds, i, i);
    welldia[i] = 1.5 * sigma[i];
  }
  sigma[9] = 1.00 * T(tab.bds, 9, 9) - 1.2;
  // squeeze ratios use the UNSCALED diameters (inputinfo.f:382-389; file columns sz8,sz6,sz7,sz9,sz10)
  for (int i = 1; i <= 20; i++) {
    const double* r = &tab.sqz6to10[(i - 1) * 5];
    double sz8 = r[0], sz6 = r[1], sz7 = r[2], sz9 = r[3], sz10 = r[4];
    s.sqz610[0 * 29 + i + 8] = sz6 * 2 / (sigma[i + 8] + sigma[4]);
    s.sqz610[1 * 29 + i + 8] = sz7 * 2 / (sigma[i + 8] + sigma[1]);
    s.sqz610[2 * 29 + i + 8] = sz8 * 2 / (sigma[i + 8] + sigma[2]);
    s.sqz610[3 * 29 + i + 8] = sz9 * 2 / (sigma[i + 8] + sigma[2]);
    s.sqz610[4 * 29 + i + 8] = sz10 * 2 / (sigma[i + 8] + sigma[4]);
  }
  for (int id = 1; id <= 28; id++) s.bmass[id] = tab.mass[id - 1];  // inputinfo.f:395-404
  s.bmass[3] = s.bmass[20];
  for (int i = 1; i <= 4; i++) s.bmass[i + 4] = s.bmass[i];
  const double boxl_orig = p.boxl;
  s.boxl_orig = boxl_orig;
  for (int k = 1; k <= 28; k++) {
    sigma[k] = sigma[k] / boxl_orig;
    welldia[k] = welldia[k] / boxl_orig;
  }
  s.shder[0] = 5.00 / boxl_orig;
  s.shder[1] = 4.74 / boxl_orig;
  s.shder[2] = 4.86 / boxl_orig;
  s.shder[3] = 4.83 / boxl_orig;
  // ---- scale_down.f:37-67
  double sigma_2b[29][29];
  PairTables& pt = m.tab;
  for (int k = 1; k <= 28; k++)
    for (int kk = 1; kk <= 28; kk++) {
      const int x = (k - 1) * 28 + (kk - 1);
      pt.sigma_sq[x] = 0.25 * hsq(sigma[k] + sigma[kk]);
      sigma_2b[k][kk] = 0.50 * (sigma[k] + sigma[kk]);
      pt.welldia_sq[x] = 0.25 * hsq(welldia[k] + welldia[kk]);
      pt.ep_sqrt[x] = std::sqrt(epsilon[k] * epsilon[kk]);
      pt.shlddia_sq[x] = hsq(s.shder[3]);
    }
  for (int k = 9; k <= 28; k++)
    for (int kk = 9; kk <= 28; kk++) {
      const int x = (k - 1) * 28 + (kk - 1);
      pt.ep_sqrt[x] = epsilon[1] * (-T(tab.ep, k, kk));  // ep = -file value, inputinfo.f:287
      pt.sigma_sq[x] = hsq(T(tab.bds, k, kk)) / hsq(boxl_orig);
      sigma_2b[k][kk] = T(tab.bds, k, kk) / boxl_orig;
      pt.welldia_sq[x] = hsq(T(tab.wel, k, kk)) / hsq(boxl_orig);
    }
  auto SH = [&](int a, int b, double v) { pt.shlddia_sq[(a - 1) * 28 + (b - 1)] = v; };
  SH(1, 2, hsq(s.shder[0])); SH(2, 1, hsq(s.shder[0])); SH(5, 2, hsq(s.shder[0])); SH(2, 5, hsq(s.shder[0]));
  SH(1, 1, hsq(s.shder[1])); SH(5, 5, hsq(s.shder[1])); SH(1, 5, hsq(s.shder[1])); SH(5, 1, hsq(s.shder[1]));
  SH(2, 4, hsq(s.shder[2])); SH(4, 2, hsq(s.shder[2])); SH(2, 8, hsq(s.shder[2])); SH(8, 2, hsq(s.shder[2]));
  s.eps1 = epsilon[1];
  s.eps_hb = pt.ep_sqrt[(5 - 1) * 28 + (8 - 1)];
  std::memcpy(s.shlddia_sq, pt.shlddia_sq, sizeof(s.shlddia_sq));
  std::memcpy(s.ep_sqrt, pt.ep_sqrt, sizeof(s.ep_sqrt));
  // ---- make_code.f:18-66 (ev_param)
  const double del = 0.02375;
  for (int l = 0; l <= 50; l++) s.ev_param1[l] = s.ev_param2[l] = s.ev_param3[l] = 1;
  s.ev_param1[15] = 1.05 * ((2.24 / boxl_orig) / ((sigma[1] + sigma[4]) / 2.0));
  s.ev_param1[17] = 1.1436; s.ev_param1[18] = 0.88; s.ev_param1[19] = 0.87829; s.ev_param1[20] = 0.8;
  s.ev_param1[21] = 0.7713; s.ev_param1[27] = 1.0;
  const double dnc = 1.46, dcc = 1.51, dcn = 1.33, dcaca = 3.8, dtie = 2.41, dtie2 = 2.45;
  s.ev_param2[4] = dnc * (1.0 - del) / boxl_orig; s.ev_param2[5] = dcc * (1.0 - del) / boxl_orig;
  s.ev_param2[6] = dcn * (1.0 - del) / boxl_orig; s.ev_param2[7] = dtie * (1.0 - del) / boxl_orig;
  s.ev_param2[8] = dtie2 * (1.0 - del) / boxl_orig; s.ev_param2[9] = dcaca * (1.0 - del) / boxl_orig;
  s.ev_param3[4] = dnc * (1.0 + del) / boxl_orig; s.ev_param3[5] = dcc * (1.0 + del) / boxl_orig;
  s.ev_param3[6] = dcn * (1.0 + del) / boxl_orig; s.ev_param3[7] = (1.0 + del) * dtie / boxl_orig;
  s.ev_param3[8] = (1.0 + del) * dtie2 / boxl_orig; s.ev_param3[9] = (1.0 + del) * dcaca / boxl_orig;

  // ---- per-bead topology words (inputinfo.f:105-132, 209-261) and side-chain bond windows (:291-339, bond.f:82-91)
  m.meta.assign(N, 0);
  m.chain.assign(N, 0);
  m.id0.assign(N, 0);
  int bead = 0, chain = 0;
  for (int sp = 0; sp < topo.n_species; sp++) {
    const int L = s.chnln[sp], nbd = s.numbeads[sp];
    std::vector<int> sc_of_res(L + 1, 0), res_of_sc;  // ordinal of the residue's side chain (1-based) / inverse
    int nsc = 0;
    for (int r = 1; r <= L; r++)
      if (topo.firstside[sp][r - 1] != 0) {
        sc_of_res[r] = ++nsc;
        res_of_sc.push_back(r);
      }
    if (3 * L + nsc != nbd) throw std::runtime_error("numbeads does not match firstside flags");
    for (int r = 1; r <= L; r++) {  // bond windows; Gly (no bead) gets row 1 of rcarnrco like inputinfo.f:313-320
      int iii = 1;
      if (sc_of_res[r]) iii = topo.identity[sp][3 * L + sc_of_res[r] - 1] - 8;
      if (iii < 1 || iii > 20) throw std::runtime_error("side-chain identity out of range 9..28");
      const double* row = &tab.rcarnrco[(iii - 1) * 6];
      const int x = (sp ? s.chnln[0] : 0) + r - 1;
      for (int kind = 0; kind < 3; kind++) {
        double len = row[kind] / boxl_orig;  // scale_down.f:69-79
        double dl = row[3 + kind] < del ? del : row[3 + kind];
        s.blmin_sc[kind][x] = (1.0 - dl) * len;
        s.blmax_sc[kind][x] = (1.0 + dl) * len;
      }
    }
    for (int c = 0; c < s.nch[sp]; c++, chain++)
      for (int l = 0; l < nbd; l++, bead++) {
        int cls, res;
        if (l < 3 * L) {
          cls = l / L;
          res = l % L + 1;
        } else {
          cls = 3;
          res = res_of_sc[l - 3 * L];
        }
        const int id = topo.identity[sp][l];
        const int expect = cls == 0 ? 2 : (cls == 1 ? 1 : (cls == 2 ? 4 : id));
        if (id != expect || (cls == 3 && (id < 10 || id > 28)))
          throw std::runtime_error("identity array must be Ca x L (2), N x L (1), C x L (4), side chains 10..28");
        // proline quirk make_code.f:146-168: N bead number x (ordinal!) of a chain whose x-th side chain is Pro
        int proex = 0;
        if (cls == 1 && res <= nsc && topo.identity[sp][3 * L + res - 1] == 17) proex = 1;
        m.meta[bead] = meta_pack(cls, res, topo.hp[sp][l] == 1, id, sp, proex, l);
        m.chain[bead] = chain;
        m.id0[bead] = (uint8_t)id;
        if (cls == 1 || cls == 2) m.nc_beads.push_back(bead);
      }
  }
  // ---- same-chain class table for the list rebuild: static_code() of every bead pair of the first chain of a species
  s.sct_off[0] = s.sct_off[1] = -1;
  m.sctab.clear();
  for (int sp = 0, first = 0; sp < topo.n_species; first += s.nch[sp] * s.numbeads[sp], sp++) {
    const int nbd = s.numbeads[sp];
    if (s.nch[sp] < 1 || nbd > 256) continue;
    s.sct_off[sp] = (int)m.sctab.size();
    for (int a = 0; a < nbd; a++)
      for (int b = 0; b < nbd; b++)
        m.sctab.push_back(a == b ? 0 : (uint8_t)static_code(s, m.meta[first + a], m.chain[first + a], first + a, m.meta[first + b],
                                                           m.chain[first + b], first + b));
  }
  // ---- nbor_setup.f:13-118 (first chain of each species against itself and each other)
  double sig_max[51] = {0}, sig_max_all = 0.0;
  auto scan = [&](int i0, int i1, int j0, int j1, bool tri) {
    for (int i = i0; i <= i1; i++)
      for (int j = tri ? i + 1 : j0; j <= j1; j++) {
        const int idi = m.id0[i], idj = m.id0[j];
        const int code = static_code(s, m.meta[i], m.chain[i], i, m.meta[j], m.chain[j], j);
        double sig_ij;
        if (code <= 3) sig_ij = sigma_2b[idi][idj];
        else if (code == 15) sig_ij = 0.5 * (welldia[idi] + welldia[idj]);
        else if (code == 16) sig_ij = T(tab.wel, idi, idj) / boxl_orig;
        else sig_ij = 0.0;
        if (sig_ij > sig_max[code]) {
          sig_max[code] = sig_ij;
          if (sig_max[code] > sig_max_all) sig_max_all = sig_max[code];
        }
      }
  };
  if (s.nch[0] > 0) scan(0, s.numbeads[0] - 2, 0, s.numbeads[0] - 1, true);
  if (topo.n_species == 2 && s.nch[1] > 0) {
    scan(s.nop1, s.nop1 + s.numbeads[1] - 2, 0, s.nop1 + s.numbeads[1] - 1, true);
    if (s.nch[0] > 0) scan(0, s.numbeads[0] - 1, s.nop1, s.nop1 + s.numbeads[1] - 1, false);
  }
  double sigij = 0.0, sig = 0.0;
  for (int i = 1; i <= 4; i++)
    for (int j = 1; j <= 4; j++) {
      if (i != 3 || j != 3) sig = sigma[i] + sigma[j];
      if (sig > sigij) sigij = sig;
    }
  for (int i = 40; i <= 50; i++) {
    sig_max[i] = 0.5 * sigij * 1.0;
    if (sig_max[i] > sig_max_all) sig_max_all = sig_max[i];
  }
  const double rl_const = 1.2;
  for (int i = 1; i <= 50; i++) {
    s.rlsq[i] = 0.0;
    if (sig_max[i] != 0.0) s.rlsq[i] = hsq((rl_const - 1) * sig_max_all + sig_max[i]);
  }
  {  // the chain-wise list rebuild of the service CTAs (dmd_engine.h, nbor_chainwise): its pair test uses the static
     // class only, which is exact when the overlay codes 40..50 share the cut-off of code 1
    double rm = 0.0;
    bool same = true;
    for (int i = 1; i <= 50; i++) {
      if (s.rlsq[i] > rm) rm = s.rlsq[i];
      if (i >= 40 && s.rlsq[i] != s.rlsq[1]) same = false;
    }
    s.rl_max = std::sqrt(rm);
    const int nchains = s.nch[0] + (topo.n_species == 2 ? s.nch[1] : 0);
    const bool short_chains = s.numbeads[0] <= 64 && (topo.n_species < 2 || s.numbeads[1] <= 64);
    s.chainwise = (same && nchains <= 64 && short_chains && s.sct_off[0] >= 0 && (topo.n_species < 2 || s.nch[1] == 0 || s.sct_off[1] >= 0)) ? 1 : 0;
    s.pad_cw = 0;
  }
  const double rl = rl_const * sig_max_all;
  s.hdelr = hsq(0.4 * (rl - sig_max_all));
  s.sig_max_all = sig_max_all;
  // ---- main.F90:389-396
  const double boxl = 1.0;
  s.num_cell = (int)(boxl / (sig_max_all * rl_const) * s.n_wrap) + 2 * s.n_wrap;
  s.ncr = s.num_cell - 2 * s.n_wrap;
  if (s.ncr < 5) throw std::runtime_error("box too small: fewer than 5 real cells per dimension");
  if (s.ncr > 1023) throw std::runtime_error("box too large: more than 1023 cells per dimension");
  s.width = boxl / (double)(s.num_cell - 2 * s.n_wrap);
  s.half = boxl / 2.0;
  // ---- the hot copies (HotConst + per-residue bond windows)
  std::memset(&m.hot, 0, sizeof(m.hot));
  std::memcpy(m.hot.ev_param1, s.ev_param1, sizeof(s.ev_param1));
  std::memcpy(m.hot.ev_param2, s.ev_param2, sizeof(s.ev_param2));
  std::memcpy(m.hot.ev_param3, s.ev_param3, sizeof(s.ev_param3));
  std::memcpy(m.hot.sqz610, s.sqz610, sizeof(s.sqz610));
  std::memcpy(m.hot.bmass, s.bmass, sizeof(s.bmass));
  for (int id = 0; id < 29; id++) m.hot.binv[id] = s.bmass[id] != 0.0 ? 1.0 / s.bmass[id] : 0.0;
  m.hot.chnln0 = s.chnln[0];
  m.hot.nres = s.chnln[0] + (topo.n_species == 2 ? s.chnln[1] : 0);
  m.bl.assign((size_t)m.hot.nres * 6, 0.0);
  for (int x = 0; x < m.hot.nres; x++)
    for (int kind = 0; kind < 3; kind++) {
      m.bl[(size_t)x * 6 + 2 * kind] = s.blmin_sc[kind][x];
      m.bl[(size_t)x * 6 + 2 * kind + 1] = s.blmax_sc[kind][x];
    }
}

}  // namespace dmd
