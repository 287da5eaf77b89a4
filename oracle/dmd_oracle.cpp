// dmd_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code).  See dmd_oracle.hpp for the
// parity status and the documented deviations D1-D4.  Each routine cites the reference file:line it follows
// (paths relative to /root/reference/parallel-dmd-PRIME20/code unless noted).
#include "dmd_oracle.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace dmdo {

// ---- constants of def.h:1-57
static const double ltstep = 1e-10, smdist = 5e-12;
static const double dnc = 1.46, dcc = 1.51, dcn = 1.33, dcaca = 3.8, dtie = 2.41, dtie2 = 2.45;
static const double sqz1 = 1.1436, sqz2 = 0.88, sqz3 = 0.87829, sqz4 = 0.8, sqz5 = 0.7713, sqz11 = 1.0;
static const int n_b_hydro = 3, n_b_hbond = 3;
static const double del = 0.02375, rl_const = 1.2;
static const int xrepuls1 = 40, xrepuls2 = 50;

static inline double dnint(double x) { return std::round(x); }
static inline double sq(double x) { return x * x; }

// D4: fdlibm __ieee754_log (Sun Microsystems, e_log.c), for positive finite normal x. No FMA.
double fdlibm_log(double x) {
  static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                      Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                      Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                      Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                      Lg7 = 1.479819860511658591e-01;
  uint64_t bits;
  std::memcpy(&bits, &x, 8);
  int32_t hx = (int32_t)(bits >> 32);
  uint32_t lx = (uint32_t)bits;
  int32_t k = 0;
  if (hx < 0x00100000) {  // subnormal: scale up
    x *= 1.80143985094819840000e+16;
    k -= 54;
    std::memcpy(&bits, &x, 8);
    hx = (int32_t)(bits >> 32);
    lx = (uint32_t)bits;
  }
  k += (hx >> 20) - 1023;
  hx &= 0x000fffff;
  int32_t i = (hx + 0x95f64) & 0x100000;
  bits = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | lx;
  std::memcpy(&x, &bits, 8);
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
  int32_t j = 0x6b851 - hx;
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

// D2: counter RNG replacing Intel IFPORT drandm (main.F90:173,412,998,...).
double Oracle::rng_uniform() {
  rng_ctr += 1;
  uint64_t z = seed + rng_ctr * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

Oracle::Oracle(const dmdb_params& p, const dmdb_topology& topo, const dmdb_tables& tab) {
  if (topo.n_species < 1 || topo.n_species > 2) throw std::runtime_error("n_species must be 1 or 2");
  chnln1 = topo.chnln[0];
  numbeads1 = topo.numbeads[0];
  nch1 = topo.n_chains[0];
  nop1 = nch1 * numbeads1;
  if (topo.n_species == 2) {
    chnln2 = topo.chnln[1];
    numbeads2 = topo.numbeads[1];
    nch2 = topo.n_chains[1];
  } else {
    chnln2 = 0;
    numbeads2 = 0;
    nch2 = 0;
  }
  nop2 = nch2 * numbeads2;
  noptotal = nop1 + nop2;
  n_wrap = p.n_wrap ? p.n_wrap : 2;
  n_nab_cell = (n_wrap == 1) ? 13 : 62;  // def.h:31-37
  canon = p.canon != 0;
  no_hbs = p.no_hbs != 0;
  seed = p.seed;
  boxl = p.boxl;
  setemp = p.tstar * 12.0;  // main.F90:127
  if (p.nbr_capacity > 0) maxnbs = p.nbr_capacity;
  log_capacity = (size_t)p.log_capacity;
  const int N = noptotal;
  identity.assign(N + 1, 0);
  chnnum.assign(N + 1, 0);
  bptnr.assign(N + 1, 0);
  coltype.assign(N + 4, -1);
  nptnr.assign(N + 4, -1);
  extra_repuls.assign((size_t)(N + 1) * 5, 0);
  sv.assign((size_t)(N + 1) * 6, 0.0);
  bm.assign(N + 1, 0.0);
  tim.assign(N + 4, 0.0);
  old_rx.assign(N + 1, 0.0);
  old_ry.assign(N + 1, 0.0);
  old_rz.assign(N + 1, 0.0);
  ev_code.assign((size_t)(N + 1) * (N + 1), 0);
  npt.assign(N + 2, 0);
  npt_dn.assign(N + 2, 0);
  na_npt.assign(N + 1, 0);
  nnabdn.assign(N + 1, 0);
  nb.assign((size_t)maxnbs * N + 2, 0);
  dnnab.assign((size_t)maxnbs * N + 2, 0);
  tlinks.assign(N + 4, 0);
  tlinks2.assign(N + 4, 0);
  bin.assign(numbin + 2, 0);
  clinks.assign(N + 1, 0);
  map.assign(n_nab_cell + 1, 0);
  inputinfo(topo, tab);
  make_code();
}

// inputinfo.f:105-411 (tables and topology; the restart read :76-101 is set_state) then scale_down (:665).
void Oracle::inputinfo(const dmdb_topology& topo, const dmdb_tables& tab) {
  const int N = noptotal;
  fside1.assign(chnln1 + 1, 0);
  fside2.assign(chnln2 + 1, 0);
  for (int k = 1; k <= chnln1; k++) fside1[k] = topo.firstside[0][k - 1];
  {  // inputinfo.f:108-116
    int l = 1;
    for (int k = 1; k <= chnln1; k++)
      if (fside1[k] != 0) { fside1[k] = 3 * chnln1 + l; l++; }
  }
  if (nop2 > 0) {  // inputinfo.f:120-132
    for (int k = 1; k <= chnln2; k++) fside2[k] = topo.firstside[1][k - 1];
    int l = 1;
    for (int k = 1; k <= chnln2; k++)
      if (fside2[k] != 0) { fside2[k] = nop1 + 3 * chnln2 + l; l++; }
  }
  // inputinfo.f:162-207
  for (int k = 0; k < 29; k++) sigma[k] = welldia[k] = epsilon[k] = bmass[k] = 0.0;
  for (int k = 1; k <= 4; k++) {
    sigma[k] = tab.protein[k - 1];
    welldia[k] = tab.protein[4 + k - 1];
    epsilon[k] = tab.protein[8 + k - 1];
  }
  for (int k = 1; k <= 4; k++) {
    sigma[k + 4] = sigma[k];
    epsilon[k + 4] = epsilon[k];
    welldia[k + 4] = welldia[k];
  }
  // inputinfo.f:209-228
  aa.assign(numbeads1 + numbeads2 + 1, 0);
  for (int k = 1; k <= numbeads1; k++) aa[k] = topo.identity[0][k - 1];
  for (int k = 1; k <= numbeads2; k++) aa[numbeads1 + k] = topo.identity[1][k - 1];
  for (int l = 1; l <= nop1; l += numbeads1)
    for (int k = 1; k <= numbeads1; k++) identity[l + k - 1] = aa[k];
  for (int l = nop1 + 1; l <= nop1 + nop2; l += numbeads2)
    for (int k = 1; k <= numbeads2; k++) identity[l + k - 1] = aa[numbeads1 + k];
  // inputinfo.f:236-261
  hp1.assign(numbeads1 + 1, 0);
  hp2.assign(numbeads2 + 1, 0);
  hp.assign(numbeads1 + numbeads2 + 1, 0);
  for (int i = 1; i <= numbeads1; i++) hp[i] = hp1[i] = topo.hp[0][i - 1];
  for (int i = 1; i <= numbeads2; i++) hp[numbeads1 + i] = hp2[i] = topo.hp[1][i - 1];
  for (int k = 1; k <= nop1; k++) chnnum[k] = (k - 1) / numbeads1 + 1;
  for (int k = 1; k <= nop2; k++) chnnum[nop1 + k] = nch1 + (k - 1) / numbeads2 + 1;
  // inputinfo.f:282-288 ep = -file, :361-369 bds / wel
  std::memset(ep, 0, sizeof(ep));
  std::memset(bds, 0, sizeof(bds));
  std::memset(wel, 0, sizeof(wel));
  for (int i = 9; i <= 28; i++)
    for (int j = 9; j <= 28; j++) {
      ep[i][j] = -tab.ep[(i - 9) * 20 + (j - 9)];
      bds[i][j] = tab.bds[(i - 9) * 20 + (j - 9)];
      wel[i][j] = tab.wel[(i - 9) * 20 + (j - 9)];
    }
  // inputinfo.f:291-339
  double drca[21], drnh[21], drco[21], del_rca[21], del_rnh[21], del_rco[21];
  for (int i = 1; i <= 20; i++) {
    const double* r = &tab.rcarnrco[(i - 1) * 6];
    drca[i] = r[0]; drnh[i] = r[1]; drco[i] = r[2];
    del_rca[i] = r[3]; del_rnh[i] = r[4]; del_rco[i] = r[5];
    if (del_rca[i] < del) del_rca[i] = del;
    if (del_rnh[i] < del) del_rnh[i] = del;
    if (del_rco[i] < del) del_rco[i] = del;
  }
  const int nres = chnln1 + chnln2;
  bdln.assign(nres + 1, 0); bl_rn.assign(nres + 1, 0); bl_rc.assign(nres + 1, 0);
  del_bdln.assign(nres + 1, 0); del_blrn.assign(nres + 1, 0); del_blrc.assign(nres + 1, 0);
  for (int i = 1; i <= chnln1; i++) {
    int iii = (fside1[i] != 0) ? aa[fside1[i]] - 8 : 1;
    bdln[i] = drca[iii]; bl_rn[i] = drnh[iii]; bl_rc[i] = drco[iii];
    del_bdln[i] = del_rca[iii]; del_blrn[i] = del_rnh[iii]; del_blrc[i] = del_rco[iii];
  }
  for (int i = 1; i <= chnln2; i++) {
    int iii = (fside2[i] != 0) ? identity[fside2[i]] - 8 : 1;
    bdln[chnln1 + i] = drca[iii]; bl_rn[chnln1 + i] = drnh[iii]; bl_rc[chnln1 + i] = drco[iii];
    del_bdln[chnln1 + i] = del_rca[iii]; del_blrn[chnln1 + i] = del_rnh[iii]; del_blrc[chnln1 + i] = del_rco[iii];
  }
  // inputinfo.f:373-378
  for (int i = 9; i <= 28; i++) {
    sigma[i] = 1.00 * bds[i][i];
    welldia[i] = 1.5 * sigma[i];
  }
  sigma[9] = 1.00 * bds[9][9] - 1.2;
  // inputinfo.f:382-389 (file columns are sz8,sz6,sz7,sz9,sz10)
  std::memset(sqz610, 0, sizeof(sqz610));
  for (int i = 1; i <= 20; i++) {
    const double* r = &tab.sqz6to10[(i - 1) * 5];
    double sz8 = r[0], sz6 = r[1], sz7 = r[2], sz9 = r[3], sz10 = r[4];
    sqz610[1][i + 8] = sz6 * 2 / (sigma[i + 8] + sigma[4]);
    sqz610[2][i + 8] = sz7 * 2 / (sigma[i + 8] + sigma[1]);
    sqz610[3][i + 8] = sz8 * 2 / (sigma[i + 8] + sigma[2]);
    sqz610[4][i + 8] = sz9 * 2 / (sigma[i + 8] + sigma[2]);
    sqz610[5][i + 8] = sz10 * 2 / (sigma[i + 8] + sigma[4]);
  }
  // inputinfo.f:395-411
  for (int id = 1; id <= 28; id++) bmass[id] = tab.mass[id - 1];
  bmass[3] = bmass[20];
  for (int i = 1; i <= 4; i++) bmass[i + 4] = bmass[i];
  for (int i = 1; i <= N; i++) bm[i] = bmass[identity[i]];
  scale_down();
}

// scale_down.f:27-79 (positions are already in box units on the restart path, -Drunr)
void Oracle::scale_down() {
  boxl_orig = boxl;
  for (int k = 1; k <= 28; k++) {
    sigma[k] = sigma[k] / boxl_orig;
    welldia[k] = welldia[k] / boxl_orig;
  }
  shder_dist1 = 5.00 / boxl_orig;
  shder_dist2 = 4.74 / boxl_orig;
  shder_dist3 = 4.86 / boxl_orig;
  shder_dist4 = 4.83 / boxl_orig;
  for (int k = 1; k <= 28; k++)
    for (int kk = 1; kk <= 28; kk++) {
      sigma_sq[k][kk] = 0.25 * sq(sigma[k] + sigma[kk]);
      sigma_2b[k][kk] = 0.50 * (sigma[k] + sigma[kk]);
      welldia_sq[k][kk] = 0.25 * sq(welldia[k] + welldia[kk]);
      ep_sqrt[k][kk] = std::sqrt(epsilon[k] * epsilon[kk]);
      shlddia_sq[k][kk] = sq(shder_dist4);
    }
  for (int k = 9; k <= 28; k++)
    for (int kk = 9; kk <= 28; kk++) {
      ep_sqrt[k][kk] = epsilon[1] * ep[k][kk];
      sigma_sq[k][kk] = sq(bds[k][kk]) / sq(boxl_orig);
      sigma_2b[k][kk] = bds[k][kk] / boxl_orig;
      welldia_sq[k][kk] = sq(wel[k][kk]) / sq(boxl_orig);
    }
  shlddia_sq[1][2] = sq(shder_dist1); shlddia_sq[2][1] = sq(shder_dist1);
  shlddia_sq[5][2] = sq(shder_dist1); shlddia_sq[2][5] = sq(shder_dist1);
  shlddia_sq[1][1] = sq(shder_dist2); shlddia_sq[5][5] = sq(shder_dist2);
  shlddia_sq[1][5] = sq(shder_dist2); shlddia_sq[5][1] = sq(shder_dist2);
  shlddia_sq[2][4] = sq(shder_dist3); shlddia_sq[4][2] = sq(shder_dist3);
  shlddia_sq[2][8] = sq(shder_dist3); shlddia_sq[8][2] = sq(shder_dist3);
  for (int k = 1; k <= chnln1 + chnln2; k++) {
    bdln[k] = bdln[k] / boxl_orig;
    bl_rn[k] = bl_rn[k] / boxl_orig;
    bl_rc[k] = bl_rc[k] / boxl_orig;
  }
  boxl = 1.0;
  half = boxl / 2.0;
}

void Oracle::setec(int k, int l, int v) {
  if (k < 1 || l < 1 || k > noptotal || l > noptotal) return;  // the Fortran would write out of bounds
  ev_code[(size_t)k * (noptotal + 1) + l] = (int8_t)v;
}

int Oracle::local_index(int i) const {  // main.F90:1468-1474
  if (i <= nop1) return i - ((chnnum[i] - 1) * numbeads1);
  return i - nop1 - ((chnnum[i] - nch1 - 1) * numbeads2) + numbeads1;
}

// make_code.f:18-566, literal (assignment ORDER matters: later loops overwrite earlier ones).
void Oracle::make_code() {
  const int N = noptotal;
  for (int k = 1; k <= 3; k++)
    for (int l = 1; l <= 50; l++) ev_param[k][l] = 1;
  ev_param[1][15] = 1.05 * ((2.24 / boxl_orig) / ((sigma[1] + sigma[4]) / 2.0));
  ev_param[1][17] = sqz1; ev_param[1][18] = sqz2; ev_param[1][19] = sqz3;
  ev_param[1][20] = sqz4; ev_param[1][21] = sqz5; ev_param[1][27] = sqz11;
  ev_param[2][4] = dnc * (1.0 - del) / boxl_orig;
  ev_param[2][5] = dcc * (1.0 - del) / boxl_orig;
  ev_param[2][6] = dcn * (1.0 - del) / boxl_orig;
  ev_param[2][7] = dtie * (1.0 - del) / boxl_orig;
  ev_param[2][8] = dtie2 * (1.0 - del) / boxl_orig;
  ev_param[2][9] = dcaca * (1.0 - del) / boxl_orig;
  ev_param[2][10] = (1.0 - del); ev_param[2][11] = (1.0 - del); ev_param[2][12] = (1.0 - del);
  ev_param[3][4] = dnc * (1.0 + del) / boxl_orig;
  ev_param[3][5] = dcc * (1.0 + del) / boxl_orig;
  ev_param[3][6] = dcn * (1.0 + del) / boxl_orig;
  ev_param[3][7] = (1.0 + del) * dtie / boxl_orig;
  ev_param[3][8] = (1.0 + del) * dtie2 / boxl_orig;
  ev_param[3][9] = (1.0 + del) * dcaca / boxl_orig;
  ev_param[3][10] = (1.0 + del); ev_param[3][11] = (1.0 + del); ev_param[3][12] = (1.0 + del);

  for (int k = 1; k <= N; k++)
    for (int l = 1; l <= N; l++) setec(k, l, 1);
  // :83-110 inter-chain hydrophobic side chains
  for (int m = 1; m <= N - 1; m++) {
    int mm = local_index(m);
    for (int l = m + 1; l <= N; l++) {
      int ll = local_index(l);
      if (identity[m] > 8 && identity[l] > 8)
        if (hp[mm] == 1 && hp[ll] == 1)
          if (chnnum[l] != chnnum[m]) setec(m, l, 16);
    }
  }
  if (!no_hbs) {
    // :117-125
    for (int m = 1; m <= N - 1; m++)
      for (int l = m + 1; l <= N; l++)
        if ((identity[m] == 1 && identity[l] == 4) || (identity[m] == 4 && identity[l] == 1)) setec(m, l, 15);
    // :128-143 reset same-chain pairs
    for (int k = 1; k <= nch1; k++) {
      int kk = (k - 1) * numbeads1;
      for (int m = kk + 1; m <= kk + numbeads1 - 1; m++)
        for (int l = m + 1; l <= kk + numbeads1; l++) setec(m, l, 1);
    }
    for (int k = nch1 + 1; k <= nch1 + nch2; k++) {
      int kk = nop1 + (k - nch1 - 1) * numbeads2;
      for (int m = kk + 1; m <= kk + numbeads2 - 1; m++)
        for (int l = m + 1; l <= kk + numbeads2; l++) setec(m, l, 1);
    }
    // :146-168 proline N-H does not hydrogen bond
    for (int m = 1; m <= N - 1; m++)
      for (int l = m + 1; l <= N; l++) {
        if (m <= nop1 && l <= nop1) {
          if (identity[m] == 17 && identity[l] == 4) setec(m - chnln1 * 2, l, 1);
          else if (identity[m] == 4 && identity[l] == 17) setec(l - chnln1 * 2, m, 1);
        } else if (m <= nop1 && l > nop1) {
          if (identity[m] == 17 && identity[l] == 4) setec(m - chnln1 * 2, l, 1);
          else if (identity[m] == 4 && identity[l] == 17) setec(l - chnln2 * 2, m, 1);
        } else if (m > nop1 && l > nop1) {
          if (identity[m] == 17 && identity[l] == 4) setec(m - chnln2 * 2, l, 1);
          else if (identity[m] == 4 && identity[l] == 17) setec(l - chnln2 * 2, m, 1);
        }
      }
  }
  // per-chain blocks: species 1 :174-328, species 2 :330-487 (chaptype 1)
  for (int sp = 0; sp < 2; sp++) {
    const int nch = sp == 0 ? nch1 : nch2;
    const int L = sp == 0 ? chnln1 : chnln2;
    const int nbd = sp == 0 ? numbeads1 : numbeads2;
    const int off = sp == 0 ? 0 : nop1;  // fside2 holds nop1 + local index (inputinfo.f:127)
    const std::vector<int>& fs = sp == 0 ? fside1 : fside2;
    const std::vector<int>& hps = sp == 0 ? hp1 : hp2;
    for (int ll = 1; ll <= nch; ll++) {
      int lll = (sp == 0 ? 0 : nop1) + (ll - 1) * nbd;
      int ncount = 2 + n_b_hydro;
      for (int k = 1; k <= L; k++) {
        if (fs[k] != 0)
          for (int l = ncount; l <= L; l++)
            if (fs[l] != 0)
              if (hps[fs[k] - off] == 1 && hps[fs[l] - off] == 1) setec(lll + fs[k] - off, lll + fs[l] - off, 16);
        ncount++;
      }
      if (!no_hbs) {
        for (int k = lll + L + 1; k <= lll + 2 * L; k++)
          for (int l = lll + 2 * L + 1; l <= lll + 3 * L; l++) setec(k, l, 15);
        for (int k = lll + L + 1; k <= lll + 2 * L; k++)
          for (int l = k + L - n_b_hbond; l <= k + L + n_b_hbond; l++) setec(k, l, 1);
      }
      for (int k = lll + 1; k <= lll + L; k++) setec(k, L + k, 4);
      for (int k = lll + 1; k <= lll + L; k++) setec(k, 2 * L + k, 5);
      for (int k = lll + L + 2; k <= lll + 2 * L; k++) setec(k, k + L - 1, 6);
      for (int k = lll + 1; k <= lll + L - 1; k++) setec(k, L + k + 1, 7);
      for (int k = lll + 2; k <= lll + L; k++) setec(k, 2 * L + k - 1, 8);
      for (int k = lll + L + 1; k <= lll + 2 * L; k++) setec(k, k + L, 8);
      for (int k = lll + 1; k <= lll + L - 1; k++) setec(k, k + 1, 9);
      for (int k = lll + 1; k <= lll + L; k++)
        if (fs[k - lll] != 0) setec(k, lll + fs[k - lll] - off, 10);
      for (int k = lll + L + 1; k <= lll + 2 * L; k++)
        if (fs[k - lll - L] != 0) setec(k, lll + fs[k - lll - L] - off, 11);
      for (int k = lll + 2 * L + 1; k <= lll + 3 * L; k++)
        if (fs[k - lll - 2 * L] != 0) setec(k, lll + fs[k - lll - 2 * L] - off, 12);
      for (int k = lll + 1; k <= lll + L - 1; k++) setec(k, k + 2 * L + 1, 17);
      for (int k = lll + 2; k <= lll + L; k++) setec(k, k + L - 1, 18);
      for (int k = lll + L + 3; k <= lll + 2 * L; k++) setec(k, k + L - 2, 19);
      for (int k = lll + L + 1; k <= lll + 2 * L - 1; k++) setec(k, k + 1, 20);
      for (int k = lll + 2 * L + 1; k <= lll + 3 * L - 1; k++) setec(k, k + 1, 21);
      for (int k = lll + 2 * L + 1; k <= lll + 3 * L - 1; k++)
        if (fs[k - lll - 2 * L + 1] != 0) setec(k, lll + fs[k - lll - 2 * L + 1] - off, 22);
      for (int k = lll + L + 2; k <= lll + 2 * L; k++)
        if (fs[k - lll - L - 1] != 0) setec(k, lll + fs[k - lll - L - 1] - off, 23);
      for (int k = lll + 1; k <= lll + L - 1; k++)
        if (fs[k - lll + 1] != 0) setec(k, lll + fs[k - lll + 1] - off, 24);
      for (int k = lll + 2; k <= lll + L; k++)
        if (fs[k - lll - 1] != 0) setec(k, lll + fs[k - lll - 1] - off, 25);
      for (int k = lll + 2 * L + 1; k <= lll + 3 * L - 2; k++)
        if (fs[k - lll - 2 * L + 2] != 0) setec(k, lll + fs[k - lll - 2 * L + 2] - off, 26);
    }
  }
  // :561-566 copy upper triangle into lower
  for (int i = 1; i <= N - 1; i++)
    for (int j = i + 1; j <= N; j++) setec(j, i, ev(i, j));
}

// nbor_setup.f:13-118
void Oracle::nbor_setup() {
  double sig_max[51];
  sig_max_all = 0.0;
  for (int i = 0; i <= 50; i++) sig_max[i] = 0.0;
  auto scan = [&](int i0, int i1, int j0, int j1, bool tri) {
    for (int i = i0; i <= i1; i++)
      for (int j = tri ? i + 1 : j0; j <= j1; j++) {
        double welli = welldia[identity[i]], wellj = welldia[identity[j]], sig_ij;
        int evcode = ev(j, i);
        if (evcode <= 26) {
          if (evcode <= 3) sig_ij = sigma_2b[identity[i]][identity[j]];
          else if (evcode == 15) sig_ij = 0.5 * (welli + wellj);
          else if (evcode == 16) sig_ij = wel[identity[i]][identity[j]] / boxl_orig;
          else sig_ij = 0.0;
          if (sig_ij > sig_max[evcode]) {
            sig_max[evcode] = sig_ij;
            if (sig_max[evcode] > sig_max_all) sig_max_all = sig_max[evcode];
          }
        }
      }
  };
  scan(1, numbeads1 - 1, 0, numbeads1, true);                    // :20-41
  if (nop2 > 0) {
    scan(nop1 + 1, nop1 + numbeads2 - 1, 0, nop1 + numbeads2, true);  // :44-65
    scan(1, numbeads1, nop1 + 1, nop1 + numbeads2, false);            // :68-89
  }
  double sigij = 0.0, sig = 0.0;
  for (int i = 1; i <= 4; i++)
    for (int j = 1; j <= 4; j++) {
      if (i != 3 || j != 3) sig = sigma[i] + sigma[j];
      if (sig > sigij) sigij = sig;
    }
  for (int i = 40; i <= 50; i++) {
    double sig_ij = 0.5 * sigij * 1.0;
    sig_max[i] = sig_ij;
    if (sig_max[i] > sig_max_all) sig_max_all = sig_max[i];
  }
  for (int i = 1; i <= 50; i++) {
    rlsq[i] = 0.0;
    if (sig_max[i] != 0.0) rlsq[i] = sq((rl_const - 1) * sig_max_all + sig_max[i]);
  }
  double rl = rl_const * sig_max_all;
  hdelr = sq(0.4 * (rl - sig_max_all));
}

// cell_link.f:16-94
void Oracle::cell_link() {
  const int nc = num_cell;
  static const int o[6] = {0, 0, 1, -1, 2, -2};
  static const int oy[4] = {0, 0, 1, 2};
  auto i_cell = [&](int ix, int iy, int iz) {
    return 1 + (ix - 1 + nc) % nc + ((iy - 1 + nc) % nc) * nc + ((iz - 1 + nc) % nc) * nc * nc;
  };
  int d_cell = 1;
  for (int iz = 1; iz <= 3; iz++) {
    int dz = o[iz];
    for (int iy = 1; iy <= 2; iy++) {
      int dy = o[iy];
      for (int ix = 1; ix <= 3; ix++) {
        int dx = o[ix];
        if (dy == 0) {
          if (dz < 1) { if (dx < 1) continue; }
          else if (dx < 0) continue;
        }
        map[d_cell] = dx + (dy + dz * nc) * nc;
        d_cell++;
      }
    }
  }
  if (n_wrap == 2) {
    for (int iz = 1; iz <= 5; iz++) {
      int dz = o[iz];
      for (int iy = 1; iy <= 3; iy++) {
        int dy = oy[iy];
        int min_ix = (iz < 4 && iy != 3) ? 4 : 1;
        for (int ix = min_ix; ix <= 5; ix++) {
          int dx = o[ix];
          if (dy == 0) {
            if (dz < 1) { if (dx < 1) continue; }
            else if (dx < 0) continue;
          }
          map[d_cell] = dx + (dy + dz * nc) * nc;
          d_cell++;
        }
      }
    }
  }
  if (d_cell - 1 != n_nab_cell) throw std::runtime_error("cell_link: stencil size mismatch");
  auto wrap1 = [&](int ix) {
    if (ix <= n_wrap) return ix + nc - 2 * n_wrap;
    if (nc - ix < n_wrap) return ix - nc + 2 * n_wrap;
    return ix;
  };
  for (int ix = 1; ix <= nc; ix++)
    for (int iy = 1; iy <= nc; iy++)
      for (int iz = 1; iz <= nc; iz++) wrap_map[i_cell(ix, iy, iz)] = i_cell(wrap1(ix), wrap1(iy), wrap1(iz));
}

int Oracle::cell_of(int k) const {  // cell_add.f:22-25
  int x = (int)((S(1, k) + half) / width) + n_wrap;
  int y = (int)((S(2, k) + half) / width) + n_wrap;
  int z = (int)((S(3, k) + half) / width) + n_wrap;
  return 1 + x + y * num_cell + z * num_cell * num_cell;
}

// cell_add.f:12-28
void Oracle::cell_add() {
  for (int k = 1; k <= noptotal; k++) clinks[k] = 0;
  std::fill(cell.begin(), cell.end(), 0);
  for (int k = 1; k <= noptotal; k++) {
    int cell_k = cell_of(k);
    clinks[k] = cell[cell_k];
    cell[cell_k] = k;
  }
}

// nbor.f:33-137
void Oracle::nbor() {
  const int N = noptotal, nc2 = num_cell * num_cell;
  for (int l = 1; l <= N; l++) {
    na_npt[l] = npt[l];
    nnabdn[l] = npt_dn[l];
  }
  cell_add();
  std::vector<int> n_cell(N + 1);
  auto push = [&](int lo, int hi) {  // lo < hi : hi goes on lo's up-list, lo on hi's down-list
    int l = na_npt[lo];
    if (l < lo * maxnbs) { nb[l] = hi; na_npt[lo] = l + 1; }
    else throw std::runtime_error("oracle: up-list capacity exceeded");
    l = nnabdn[hi];
    if (l < hi * maxnbs) { dnnab[l] = lo; nnabdn[hi] = l + 1; }
    else throw std::runtime_error("oracle: down-list capacity exceeded");
  };
  for (int c = n_wrap * nc2; c <= nc2 * (num_cell - n_wrap); c++) {
    int n_bead = 0, bead = cell[c];
    if (bead == 0) continue;
    while (bead != 0) { n_cell[++n_bead] = bead; bead = clinks[bead]; }
    int f_bead = n_bead;
    for (int n = 1; n <= n_nab_cell; n++) {
      int ncl = wrap_map[c + map[n]];
      bead = cell[ncl];
      while (bead != 0) { n_cell[++n_bead] = bead; bead = clinks[bead]; }
    }
    for (int i_n = 1; i_n <= f_bead; i_n++) {
      int i = n_cell[i_n];
      for (int j_n = i_n + 1; j_n <= n_bead; j_n++) {
        int j = n_cell[j_n];
        int evcode = ev(j, i);
        if ((evcode >= 4 && evcode <= 12) || (evcode >= 17 && evcode < 27)) {
          if (j > i) push(i, j); else push(j, i);
        } else {
          double rxij = S(1, i) - S(1, j), ryij = S(2, i) - S(2, j), rzij = S(3, i) - S(3, j);
          rxij = rxij - dnint(rxij); ryij = ryij - dnint(ryij); rzij = rzij - dnint(rzij);
          double rijsq = rxij * rxij + ryij * ryij + rzij * rzij;
          if (rijsq <= rlsq[evcode]) { if (j > i) push(i, j); else push(j, i); }
        }
      }
    }
  }
  for (int k = 1; k <= N; k++) {
    na_npt[k] = na_npt[k] - npt[k];
    nnabdn[k] = nnabdn[k] - npt_dn[k];
  }
}

// displ.f:20-46
bool Oracle::displ() {
  double moved_far = 0.0;
  for (int i = 1; i <= noptotal; i++) {
    double a = old_rx[i] - S(1, i), b = old_ry[i] - S(2, i), c = old_rz[i] - S(3, i);
    double dis = a * a + b * b + c * c;
    double moved = dis / hdelr;
    if (moved > moved_far) moved_far = moved;
  }
  if (moved_far >= 0.1) {
    if (moved_far >= 1.25 * 1.25) {
      t_fact = t_fact / 1.01;
      interval = t_fact / std::sqrt(setemp);
    }
    return true;
  }
  return false;
}

// ---- pair geometry shared by the predictors (core.f:14-23 etc.)
struct PairGeom { double vx, vy, vz, rx, ry, rz, bij; };
static inline PairGeom geom(const double* svi, const double* svj, double tfalse) {
  PairGeom g;
  g.vx = svi[3] - svj[3]; g.vy = svi[4] - svj[4]; g.vz = svi[5] - svj[5];
  g.rx = svi[0] - svj[0] + g.vx * tfalse;
  g.ry = svi[1] - svj[1] + g.vy * tfalse;
  g.rz = svi[2] - svj[2] + g.vz * tfalse;
  g.rx = g.rx - dnint(g.rx); g.ry = g.ry - dnint(g.ry); g.rz = g.rz - dnint(g.rz);
  g.bij = g.rx * g.vx + g.ry * g.vy + g.rz * g.vz;
  return g;
}

// core.f:14-40
void Oracle::core(int i, int j, int evcode, double& tij, int& type) const {
  PairGeom g = geom(&sv[(size_t)i * 6], &sv[(size_t)j * 6], tfalse);
  if (g.bij < 0.0) {
    double sigsq = sigma_sq[identity[i]][identity[j]] * sq(ev_param[1][evcode]);
    if (evcode >= 22 && evcode <= 26) {
      int k = identity[i] > identity[j] ? identity[i] : identity[j];
      sigsq = sigsq * sq(sqz610[evcode - 21][k]);
    }
    double rijsq = g.rx * g.rx + g.ry * g.ry + g.rz * g.rz;
    double vijsq = g.vx * g.vx + g.vy * g.vy + g.vz * g.vz;
    double discr = g.bij * g.bij - vijsq * (rijsq - sigsq);
    if (discr > 0.0) {
      tij = (-g.bij - std::sqrt(discr)) / vijsq;
      type = 1;
    }
  }
}

// bond.f:27-126 -- the species-2 branch differs only in the index into bdln/bl_rn/bl_rc (:78-91 vs :128-141)
static void bond_limits(const Oracle& o, int i, int evcode, double& blmin, double& blmax) {
  blmin = o.ev_param[2][evcode];
  blmax = o.ev_param[3][evcode];
  if (evcode < 10 || evcode > 12) return;
  int r;
  if (i <= o.nop1) {
    int ii = i - ((o.chnnum[i] - 1) * o.numbeads1);
    r = evcode == 10 ? ii : (evcode == 11 ? ii - o.chnln1 : ii - 2 * o.chnln1);
  } else {
    int ii = i - o.nop1 - ((o.chnnum[i] - o.nch1 - 1) * o.numbeads2) + o.numbeads1;
    int base = ii - 3 * o.chnln1 + (o.chnln1 * 4 - o.numbeads1);
    r = evcode == 10 ? base : (evcode == 11 ? base - o.chnln2 : base - 2 * o.chnln2);
  }
  const double *len, *dl;
  if (evcode == 10) { len = o.bdln.data(); dl = o.del_bdln.data(); }
  else if (evcode == 11) { len = o.bl_rn.data(); dl = o.del_blrn.data(); }
  else { len = o.bl_rc.data(); dl = o.del_blrc.data(); }
  blmin = (1.0 - dl[r]) * len[r];
  blmax = (1.0 + dl[r]) * len[r];
}

void Oracle::bond(int i, int j, int evcode, double& tij, int& type) const {
  double blmin, blmax;
  bond_limits(*this, i, evcode, blmin, blmax);
  PairGeom g = geom(&sv[(size_t)i * 6], &sv[(size_t)j * 6], tfalse);
  double rijsq = g.rx * g.rx + g.ry * g.ry + g.rz * g.rz;
  double vijsq = g.vx * g.vx + g.vy * g.vy + g.vz * g.vz;
  if (g.bij < 0.0) {
    double discr1 = g.bij * g.bij - vijsq * (rijsq - blmin * blmin);
    if (discr1 > 0.0) {
      tij = (-g.bij - std::sqrt(discr1)) / vijsq;
      type = 2;
    } else {
      double discr2 = g.bij * g.bij - vijsq * (rijsq - blmax * blmax);
      if (discr2 > 0.0) {
        tij = (-g.bij + std::sqrt(discr2)) / vijsq;
        type = 3;
      }
    }
  } else {
    double discr2 = g.bij * g.bij - vijsq * (rijsq - blmax * blmax);
    if (discr2 > 0.0) {
      tij = -(rijsq - blmax * blmax) / (std::sqrt(discr2) + g.bij);
      type = 3;
    }
  }
}

// sqwel.f:15-64
void Oracle::sqwel(int i, int j, int, double& tij, int& type) const {
  PairGeom g = geom(&sv[(size_t)i * 6], &sv[(size_t)j * 6], tfalse);
  double rijsq = g.rx * g.rx + g.ry * g.ry + g.rz * g.rz;
  double vijsq = g.vx * g.vx + g.vy * g.vy + g.vz * g.vz;
  if (g.bij < 0.0) {
    double diff = rijsq - welldia_sq[identity[i]][identity[j]];
    if (diff < 0.0) {
      double corediscr = g.bij * g.bij - vijsq * (rijsq - sigma_sq[identity[i]][identity[j]]);
      if (corediscr > 0.0) {
        tij = (-g.bij - std::sqrt(corediscr)) / vijsq;
        type = 1;
      } else {
        double welldiscr = g.bij * g.bij - vijsq * diff;
        tij = (-g.bij + std::sqrt(welldiscr)) / vijsq;
        type = 8;
      }
    } else {
      double welldiscr = g.bij * g.bij - vijsq * diff;
      if (welldiscr > 0.0) {
        tij = (-g.bij - std::sqrt(welldiscr)) / vijsq;
        type = 4;
      }
    }
  } else {
    double diff = rijsq - welldia_sq[identity[i]][identity[j]];
    if (diff < 0.0) {
      double welldiscr = g.bij * g.bij - vijsq * diff;
      tij = (-g.bij + std::sqrt(welldiscr)) / vijsq;
      type = 8;
    }
  }
}

// nc_sqwel.f:19-122
void Oracle::nc_sqwel(int i, int j, int, double& tij, int& type) const {
  PairGeom g = geom(&sv[(size_t)i * 6], &sv[(size_t)j * 6], tfalse);
  double rijsq = g.rx * g.rx + g.ry * g.ry + g.rz * g.rz;
  double vijsq = g.vx * g.vx + g.vy * g.vy + g.vz * g.vz;
  double diff = rijsq - welldia_sq[identity[i]][identity[j]];
  if (identity[i] + identity[j] == 5) {
    if (g.bij < 0.0) {
      if (diff < 0.0) {
        double corediscr = g.bij * g.bij - vijsq * (rijsq - sigma_sq[identity[i]][identity[j]]);
        if (corediscr > 0.0) {
          tij = (-g.bij - std::sqrt(corediscr)) / vijsq;
          type = 1;
        } else {
          double welldiscr = g.bij * g.bij - vijsq * diff;
          tij = (-g.bij + std::sqrt(welldiscr)) / vijsq;
          type = 16;
        }
      } else {
        double welldiscr = g.bij * g.bij - vijsq * diff;
        if (welldiscr > 0.0) {
          tij = (-g.bij - std::sqrt(welldiscr)) / vijsq;
          type = 7;
        }
      }
    } else {
      if (diff < 0.0) {
        double welldiscr = g.bij * g.bij - vijsq * diff;
        tij = (-g.bij + std::sqrt(welldiscr)) / vijsq;
        type = 16;
      }
    }
  } else if (bptnr[i] == j) {
    if (g.bij < 0.0) {
      double fac_sigsq = sigma_sq[identity[i]][identity[j]] * ev_param[1][15] * ev_param[1][15];
      double fac_cored = g.bij * g.bij - vijsq * (rijsq - fac_sigsq);
      if (fac_cored > 0.0) {
        tij = (-g.bij - std::sqrt(fac_cored)) / vijsq;
        type = 1;
      } else {
        double welldiscr = g.bij * g.bij - vijsq * diff;
        tij = (-g.bij + std::sqrt(welldiscr)) / vijsq;
        type = 8;
      }
    } else {
      double welldiscr = g.bij * g.bij - vijsq * diff;
      tij = (-g.bij + std::sqrt(welldiscr)) / vijsq;
      type = 8;
    }
  } else {
    if (g.bij < 0.0) {
      double welldiscr = g.bij * g.bij - vijsq * diff;
      if (welldiscr > 0.0) {
        tij = (-g.bij - std::sqrt(welldiscr)) / vijsq;
        type = 9;
      }
    }
  }
}

// sqshlder.f:15-63
void Oracle::sqshlder(int i, int j, int, double& tij, int& type) const {
  PairGeom g = geom(&sv[(size_t)i * 6], &sv[(size_t)j * 6], tfalse);
  double rijsq = g.rx * g.rx + g.ry * g.ry + g.rz * g.rz;
  double vijsq = g.vx * g.vx + g.vy * g.vy + g.vz * g.vz;
  if (g.bij < 0.0) {
    double diff = rijsq - shlddia_sq[identity[i]][identity[j]];
    if (diff < 0.0) {
      double corediscr = g.bij * g.bij - vijsq * (rijsq - sigma_sq[identity[i]][identity[j]]);
      if (corediscr > 0.0) {
        tij = (-g.bij - std::sqrt(corediscr)) / vijsq;
        type = 1;
      } else {
        double shlddiscr = g.bij * g.bij - vijsq * diff;
        tij = (-g.bij + std::sqrt(shlddiscr)) / vijsq;
        type = 10;
      }
    } else {
      double shlddiscr = g.bij * g.bij - vijsq * diff;
      if (shlddiscr > 0.0) {
        tij = (-g.bij - std::sqrt(shlddiscr)) / vijsq;
        type = 12;
      }
    }
  } else {
    double diff = rijsq - shlddia_sq[identity[i]][identity[j]];
    if (diff < 0.0) {
      double shlddiscr = g.bij * g.bij - vijsq * diff;
      tij = (-g.bij + std::sqrt(shlddiscr)) / vijsq;
      type = 10;
    }
  }
}

// the dispatch of events.f:28-48 / eventredo_up.f:25-46 / eventredo_down.f:25-58
static inline void dispatch(const Oracle& o, int i, int j, int evcode, double& tij, int& type) {
  if (evcode <= 3) o.core(i, j, evcode, tij, type);
  else if (evcode >= 4 && evcode <= 12) o.bond(i, j, evcode, tij, type);
  else if (evcode == 15) o.nc_sqwel(i, j, evcode, tij, type);
  else if (evcode == 16) o.sqwel(i, j, evcode, tij, type);
  else if (evcode >= 17 && evcode <= 26) o.core(i, j, evcode, tij, type);
  else if (evcode >= 40) o.sqshlder(i, j, evcode, tij, type);
  else throw std::runtime_error("error in ev_code matrix (events)");
}

// add_tbin.f:12-32 with D3 (bucket index clamped to [nbin, numbin])
void Oracle::add_tbin(int i) {
  int j = (int)((tim[i] + tbin_off) / sortsize) + 1;
  if (j > numbin) j = numbin;  // -Ddebugging clamp, add_tbin.f:20-23
  if (j < nbin) j = nbin;      // D3
  tlinks[i] = bin[j];
  tlinks2[i] = j + noptotal + 3;
  if (bin[j] != 0) tlinks2[bin[j]] = i;
  bin[j] = i;
}

// del_tbin.f:12-20
void Oracle::del_tbin(int i) {
  if (tlinks2[i] > noptotal + 3) bin[tlinks2[i] - noptotal - 3] = tlinks[i];
  else tlinks[tlinks2[i]] = tlinks[i];
  if (tlinks[i] != 0) tlinks2[tlinks[i]] = tlinks2[i];
  tlinks2[i] = 0;
}

// events.f:23-123
void Oracle::events() {
  const int N = noptotal;
  for (int i = 1; i <= N; i++) {
    int kstart = (i - 1) * maxnbs + 1, kend = kstart + na_npt[i] - 1;
    for (int k = kstart; k <= kend; k++) {
      int j = nb[k], type = 0;
      double tij = 1000000000.0;
      dispatch(*this, i, j, ev(j, i), tij, type);
      if (tij < tim[i]) { tim[i] = tij; nptnr[i] = j; coltype[i] = type; }
    }
    for (int k = 1; k <= 3; k++) {
      int j = ER(i, k);
      if (j > i) {
        double tij = 1000000000.0;
        int type = 0, evcode = ev(j, i);
        if (evcode == 1) core(i, j, evcode, tij, type);
        else sqshlder(i, j, evcode, tij, type);
        if (tij < tim[i]) { nptnr[i] = j; coltype[i] = type; tim[i] = tij; }
      }
    }
  }
  for (int i = 1; i <= numbin + 1; i++) bin[i] = 0;
  for (int i = 1; i <= N + 3; i++) {
    tlinks[i] = 0;
    tlinks2[i] = 0;
    if (tim[i] < interval_max) add_tbin(i);
  }
}

// eventredo_up.f:25-56
void Oracle::eventredo_up(int i, int j) {
  double tij = 1000000000.0;
  int type = 0;
  dispatch(*this, i, j, ev(j, i), tij, type);
  n_pair_predictions++;
  if (tij < tim[i]) { tim[i] = tij; nptnr[i] = j; coltype[i] = type; }
}

// eventredo_down.f:25-78
void Oracle::eventredo_down(int i, int j) {
  double tij = 1000000000.0;
  int type = 0;
  dispatch(*this, i, j, ev(i, j), tij, type);
  n_pair_predictions++;
  tij = tij + tfalse;
  if (tij < tim[i]) {
    if (tlinks2[i] != 0) del_tbin(i);
    tim[i] = tij; nptnr[i] = j; coltype[i] = type;
    if (tim[i] < interval_max) add_tbin(i);
  }
}

// the block repeated at partial_events.f:16-35, :41-60, :78-95, ...
void Oracle::redo_full(int l) {
  if (tlinks2[l] != 0) del_tbin(l);
  tim[l] = interval_max + ltstep - tfalse;
  coltype[l] = -1;
  nptnr[l] = -1;
  int kstart = (l - 1) * maxnbs + 1, kend = kstart + na_npt[l] - 1;
  n_nbr_visits += na_npt[l];
  for (int ll = kstart; ll <= kend; ll++) eventredo_up(l, nb[ll]);
  for (int ll = 1; ll <= 3; ll++)
    if (ER(l, ll) > l) eventredo_up(l, ER(l, ll));
  tim[l] = tim[l] + tfalse;
  if (tim[l] < interval_max) add_tbin(l);
}

// partial_events.f:16-201
void Oracle::partial_events(int i, int j, bool xpulse_del) {
  redo_full(i);
  if (j != 0) redo_full(j);
  auto down = [&](int a, int skip) {
    int kstart = (a - 1) * maxnbs + 1, kend = kstart + nnabdn[a] - 1;
    n_nbr_visits += nnabdn[a];
    for (int kk = kstart; kk <= kend; kk++) {
      int l = dnnab[kk];
      if (l == skip) continue;
      if (nptnr[l] != a) eventredo_down(l, a);
      else redo_full(l);
    }
    for (int kk = 1; kk <= 3; kk++) {
      int l = ER(a, kk);
      if (l < a && l != 0) {
        if (nptnr[l] != a) eventredo_down(l, a);
        else redo_full(l);
      }
    }
  };
  down(i, 0);
  if (j != 0) down(j, i);  // partial_events.f:136 skips l==i in the list loop only
  if (xpulse_del) {
    if (identity[i] < identity[j]) repuls_del_b(i, j);
    else repuls_del_b(j, i);
  }
}

// eventdyn.f:18-381
void Oracle::eventdyn(int i, int j, int evcode) {
  PairGeom g = geom(&sv[(size_t)i * 6], &sv[(size_t)j * 6], tfalse);
  const double rxij = g.rx, ryij = g.ry, rzij = g.rz, bij = g.bij;
  double rmass = 2 * bm[i] * bm[j] / (bm[i] + bm[j]);
  double ratio = 0.0, blmin, blmax;
  auto bump = [&](double bumpdist, double sgn) {  // sgn=+1: i += , j -= (move apart); -1: together
    S(1, i) = S(1, i) + sgn * (bumpdist * rxij); S(2, i) = S(2, i) + sgn * (bumpdist * ryij);
    S(3, i) = S(3, i) + sgn * (bumpdist * rzij);
    S(1, j) = S(1, j) - sgn * (bumpdist * rxij); S(2, j) = S(2, j) - sgn * (bumpdist * ryij);
    S(3, j) = S(3, j) - sgn * (bumpdist * rzij);
  };
  const int ct = coltype[i];
  const int idi = identity[i], idj = identity[j];
  if (ct == 2) {
    bond_limits(*this, i, evcode, blmin, blmax);
    ratio = rmass * bij / (blmin * blmin);
  } else if (ct == 3) {
    bond_limits(*this, i, evcode, blmin, blmax);
    ratio = rmass * bij / (blmax * blmax);
  } else if (ct == 1) {
    double sigsq;
    if (evcode == 15) {
      if (bptnr[i] == j) sigsq = sigma_sq[idi][idj] * sq(ev_param[1][evcode]);
      else sigsq = sigma_sq[idi][idj];
    } else {
      sigsq = sigma_sq[idi][idj] * sq(ev_param[1][evcode]);
      if (evcode >= 22 && evcode <= 26) {
        int k = idi > idj ? idi : idj;
        sigsq = sigsq * sq(sqz610[evcode - 21][k]);
      }
    }
    ratio = rmass * bij / sigsq;
  } else if (ct == 4) {
    double wellsq = welldia_sq[idi][idj], epsave = ep_sqrt[idi][idj];
    double del_pe = 4.0 * wellsq * epsave / rmass;
    if (bij * bij + del_pe > 0.0) {
      ratio = rmass * (std::sqrt((4.0 * wellsq * epsave / rmass) + bij * bij) + bij) / (2.0 * wellsq);
      coltype[i] = 20;
      bump(smdist * std::sqrt(wellsq), -1.0);
    } else {
      ratio = rmass * bij / wellsq;
      coltype[i] = 22;
      bump(smdist * std::sqrt(wellsq), +1.0);
    }
  } else if (ct == 8) {
    double wellsq = welldia_sq[idi][idj], epsave = ep_sqrt[idi][idj];
    double del_pe = 4.0 * wellsq * epsave / rmass;
    double bumpdist = smdist * std::sqrt(wellsq);
    if (bij * bij > del_pe) {
      ratio = rmass * (-std::sqrt(-del_pe + bij * bij) + bij) / (2.0 * wellsq);
      coltype[i] = 21;
      bump(bumpdist, +1.0);
    } else {
      ratio = rmass * bij / wellsq;
      coltype[i] = 22;
      bump(bumpdist, -1.0);
    }
  } else if (ct == 9) {
    double wellsq = welldia_sq[idi][idj];
    ratio = rmass * bij / wellsq;
    coltype[i] = 23;
    bump(smdist * std::sqrt(wellsq), +1.0);
  } else if (ct == 5) {
    double wellsq = shlddia_sq[idi][idj], epsave = -epsilon[1];
    double del_pe = 4.0 * wellsq * epsave / rmass;
    double bumpdist = smdist * std::sqrt(wellsq);
    ratio = rmass * (-std::sqrt(-del_pe + bij * bij) + bij) / (2.0 * wellsq);
    coltype[i] = 24;
    bump(bumpdist, +1.0);
  } else if (ct == 6) {
    double wellsq = shlddia_sq[idi][idj], epsave = -epsilon[1];
    double del_pe = 4.0 * wellsq * epsave / rmass;
    double bumpdist = smdist * std::sqrt(wellsq);
    if (bij * bij > -del_pe) {
      ratio = rmass * (std::sqrt(del_pe + bij * bij) + bij) / (2.0 * wellsq);
      coltype[i] = 25;
      bump(bumpdist, -1.0);
    } else {
      ratio = rmass * bij / wellsq;
      coltype[i] = 26;
      bump(bumpdist, +1.0);
    }
  } else if (ct == 13) {
    double wellsq = shlddia_sq[idi][idj];
    ratio = rmass * bij / wellsq;
    coltype[i] = 27;
    bump(smdist * std::sqrt(wellsq), +1.0);
  }
  double delvx = ratio * rxij, delvy = ratio * ryij, delvz = ratio * rzij;
  S(4, i) = S(4, i) - delvx / bm[i]; S(4, j) = S(4, j) + delvx / bm[j];
  S(5, i) = S(5, i) - delvy / bm[i]; S(5, j) = S(5, j) + delvy / bm[j];
  S(6, i) = S(6, i) - delvz / bm[i]; S(6, j) = S(6, j) + delvz / bm[j];
  S(1, i) = S(1, i) + delvx * tfalse / bm[i];
  S(2, i) = S(2, i) + delvy * tfalse / bm[i];
  S(3, i) = S(3, i) + delvz * tfalse / bm[i];
  S(1, j) = S(1, j) - delvx * tfalse / bm[j];
  S(2, j) = S(2, j) - delvy * tfalse / bm[j];
  S(3, j) = S(3, j) - delvz * tfalse / bm[j];
}

// bumped.f:12-43
void Oracle::bumpoff(int i, int j, int evcode) {
  PairGeom g = geom(&sv[(size_t)i * 6], &sv[(size_t)j * 6], tfalse);
  double bumpdist;
  if (evcode >= 40) bumpdist = smdist * std::sqrt(shlddia_sq[identity[i]][identity[j]]);
  else bumpdist = smdist * std::sqrt(welldia_sq[identity[i]][identity[j]]);
  double sgn = g.bij < 0.0 ? -1.0 : +1.0;
  S(1, i) = S(1, i) + sgn * (bumpdist * g.rx); S(2, i) = S(2, i) + sgn * (bumpdist * g.ry);
  S(3, i) = S(3, i) + sgn * (bumpdist * g.rz);
  S(1, j) = S(1, j) - sgn * (bumpdist * g.rx); S(2, j) = S(2, j) - sgn * (bumpdist * g.ry);
  S(3, j) = S(3, j) - sgn * (bumpdist * g.rz);
}

// the index arithmetic shared by repuls_add.f:14-28, repuls_check.f:17-31, repuls_del_a/b
#define AUX_INDICES(i, j)                                           \
  if ((i) <= nop1) { ncim1 = (i) + chnln1 - 1; ncai = (i)-chnln1; } \
  else { ncim1 = (i) + chnln2 - 1; ncai = (i)-chnln2; }             \
  if ((j) <= nop1) { ncaj = (j)-2 * chnln1; nnjp1 = (j)-chnln1 + 1; } \
  else { ncaj = (j)-2 * chnln2; nnjp1 = (j)-chnln2 + 1; }

// repuls_add.f:14-47   (i = the N bead, j = the C bead)
void Oracle::repuls_add(int i, int j) {
  AUX_INDICES(i, j)
  setec(ncaj, i, xrepuls2); setec(i, ncaj, xrepuls1);
  setec(i, nnjp1, xrepuls1); setec(nnjp1, i, xrepuls2);
  setec(ncai, j, xrepuls2); setec(j, ncai, xrepuls1);
  setec(j, ncim1, xrepuls1); setec(ncim1, j, xrepuls2);
  ER(i, 1) = ncaj; ER(i, 2) = nnjp1;
  ER(ncaj, 3) = i; ER(nnjp1, 3) = i;
  ER(j, 1) = ncai; ER(j, 2) = ncim1;
  ER(ncai, 3) = j; ER(ncim1, 3) = j;
  ER(i, 4) = j; ER(j, 4) = i;
}

// repuls_del_a.f:14-37
void Oracle::repuls_del_a(int i, int j) {
  AUX_INDICES(i, j)
  setec(ncaj, i, 1); setec(i, ncaj, 1); setec(i, nnjp1, 1); setec(nnjp1, i, 1);
  setec(ncai, j, 1); setec(j, ncai, 1); setec(ncim1, j, 1); setec(j, ncim1, 1);
}

// repuls_del_b.f:14-39
void Oracle::repuls_del_b(int i, int j) {
  AUX_INDICES(i, j)
  ER(i, 1) = 0; ER(i, 2) = 0; ER(ncaj, 3) = 0; ER(nnjp1, 3) = 0;
  ER(j, 1) = 0; ER(j, 2) = 0; ER(ncai, 3) = 0; ER(ncim1, 3) = 0;
  ER(i, 4) = 0; ER(j, 4) = 0;
}

static inline double pdist(const double* a, const double* b, double tfalse) {  // repuls_check.f:32-42
  double vx = a[3] - b[3], vy = a[4] - b[4], vz = a[5] - b[5];
  double rx = a[0] - b[0] + vx * tfalse, ry = a[1] - b[1] + vy * tfalse, rz = a[2] - b[2] + vz * tfalse;
  rx = rx - dnint(rx); ry = ry - dnint(ry); rz = rz - dnint(rz);
  double d = rx * rx + ry * ry + rz * rz;
  return std::sqrt(d);
}

// repuls_check.f:17-81
double Oracle::repuls_check(int i, int j) const {
  AUX_INDICES(i, j)
  const double* s = sv.data();
  double d1 = pdist(s + (size_t)i * 6, s + (size_t)ncaj * 6, tfalse);
  double d2 = pdist(s + (size_t)i * 6, s + (size_t)nnjp1 * 6, tfalse);
  double d3 = pdist(s + (size_t)j * 6, s + (size_t)ncai * 6, tfalse);
  double d4 = pdist(s + (size_t)j * 6, s + (size_t)ncim1 * 6, tfalse);
  if (d1 > shder_dist1 && d2 > shder_dist2 && d3 > shder_dist3 && d4 > shder_dist4) return 10.0;
  return 11.0;
}

// repuls_check_3.f:16-103
double Oracle::repuls_check_3(int i, int j, int k) const {
  AUX_INDICES(i, j)
  const double* s = sv.data();
  int m = 0;
  double d1 = pdist(s + (size_t)i * 6, s + (size_t)ncaj * 6, tfalse);
  if (ncaj != k && d1 > shder_dist1) m++;
  double d2 = pdist(s + (size_t)i * 6, s + (size_t)nnjp1 * 6, tfalse);
  if (nnjp1 != k && d2 > shder_dist2) m++;
  double d3 = pdist(s + (size_t)j * 6, s + (size_t)ncai * 6, tfalse);
  if (ncai != k && d3 > shder_dist3) m++;
  double d4 = pdist(s + (size_t)j * 6, s + (size_t)ncim1 * 6, tfalse);
  if (ncim1 != k && d4 > shder_dist4) m++;
  return m == 3 ? 10.0 : 11.0;
}

// check_sigma.f:12-29
double Oracle::check_sigma(int i, int j) const {
  PairGeom g = geom(&sv[(size_t)i * 6], &sv[(size_t)j * 6], tfalse);
  double rijsq = g.rx * g.rx + g.ry * g.ry + g.rz * g.rz;
  double diff = rijsq - sigma_sq[identity[i]][identity[j]];
  return diff < 0.0 ? 10.0 : 11.0;
}

// energy.f:25-101
EnergyRec Oracle::energy() const {
  EnergyRec e{};
  const int N = noptotal;
  for (int i = 1; i <= N - 1; i++)
    for (int j = i + 1; j <= N; j++) {
      if (bptnr[i] == j) {
        if (chnnum[i] == chnnum[j]) e.hb_ii++; else e.hb_ij++;
      } else if (ev(j, i) == 16) {
        PairGeom g = geom(&sv[(size_t)i * 6], &sv[(size_t)j * 6], tfalse);
        double rijsq = g.rx * g.rx + g.ry * g.ry + g.rz * g.rz;
        double wellsq = welldia_sq[identity[i]][identity[j]];
        double ep_depth = ep_sqrt[identity[i]][identity[j]];
        if (rijsq <= wellsq) {
          if (chnnum[i] == chnnum[j]) e.ehh_ii = e.ehh_ii + ep_depth;
          else e.ehh_ij = e.ehh_ij + ep_depth;
        }
      }
    }
  for (int k = 1; k <= nch1; k++)
    for (int i = (k - 1) * numbeads1 + chnln1 + 5; i <= (k - 1) * numbeads1 + 2 * chnln1; i++)
      if (bptnr[i] == i + chnln1 - 4) e.hb_alpha++;
  for (int k = 1; k <= nch2; k++)
    for (int i = nop1 + (k - 1) * numbeads2 + chnln2 + 5; i <= nop1 + (k - 1) * numbeads2 + 2 * chnln2; i++)
      if (bptnr[i] == i + chnln2 - 4) e.hb_alpha++;
  double eps_hb = ep_sqrt[5][8];
  int sum_hb = e.hb_ii + e.hb_ij;
  double sum_ehh = e.ehh_ii + e.ehh_ij;
  double sumeps = -(sum_hb * eps_hb + sum_ehh);
  double sumvel = 0.0;
  for (int i = 1; i <= N; i++)
    sumvel = sumvel + bm[i] * (S(4, i) * S(4, i) + S(5, i) * S(5, i) + S(6, i) * S(6, i));
  e.sumvel = sumvel;
  e.ered = 0.5 * sumvel + sumeps;
  e.tred = sumvel / 3.0 / (double)N;
  e.coll = coll;
  e.t = t + tfalse;
  return e;
}

// check_nc_int.f:21-360.  The four species branches of the Fortran (:31-50, :118-124, :205-211, :292-298) differ only in
// the "no end bead" test, which is terminal_ok() (same expressions as main.F90:1490-1491 etc.).  Loop over the up-lists
// (nb / na_npt, :24-29), pairs with ev_code(j,i) == 15 only.
Oracle::NcAudit Oracle::check_nc_int() const {
  NcAudit a;
  for (int i = 1; i <= noptotal; i++) {
    const int kstart = (i - 1) * maxnbs + 1, kend = kstart + na_npt[i] - 1;
    for (int k = kstart; k <= kend; k++) {
      const int j = nb[k];
      if (ev(j, i) != 15) continue;
      a.pairs15++;
      const int ii = local_index(i), jj = local_index(j);
      const bool inner = terminal_ok(i, ii, j, jj);
      if (bptnr[i] == j) {  // counted as a hydrogen bond: the four auxiliary distances must be legal (:43-57)
        double rating = 10.0;
        if (inner) rating = identity[i] < identity[j] ? repuls_check(i, j) : repuls_check(j, i);
        if (rating > 10.0) a.boundbad++;
        if (inner) {
          if (ER(i, 4) == j) a.n_ss++;
          else a.no_ss++;  // print*, 'no ss for bond'
        }
      } else {  // not bonded: inside the well with good geometry should not happen (:66-105)
        PairGeom g = geom(&sv[(size_t)i * 6], &sv[(size_t)j * 6], tfalse);
        const double rijsq = g.rx * g.rx + g.ry * g.ry + g.rz * g.rz;
        const double diff = rijsq - welldia_sq[identity[i]][identity[j]];
        if (diff < 0.0) {
          double rating = 11.0;  // an end bead is involved: "we'll let this slide"
          if (inner) rating = identity[i] < identity[j] ? repuls_check(i, j) : repuls_check(j, i);
          if (rating <= 10.0) a.unboundbad++;
          if (inner) {
            if (ER(i, 4) == j) a.n_ss++;
            else a.no_ss++;  // print*, 'no ss for non-bond'
          }
        }
      }
      if (ER(i, 4) == j) a.m_ss++;  // :107-110
    }
  }
  return a;
}

void Oracle::adopt_state(const double* sv6xN, double tfalse_, const int* bptnr_, const int* identity_, const int* er,
                         const int* off, const int* lst) {
  const int N = noptotal;
  tfalse = tfalse_;
  for (int k = 1; k <= N; k++) {
    for (int c = 0; c < 6; c++) sv[(size_t)k * 6 + c] = sv6xN[(size_t)(k - 1) * 6 + c];
    bptnr[k] = bptnr_[k - 1];
    identity[k] = identity_[k - 1];
    for (int s4 = 1; s4 <= 4; s4++) ER(k, s4) = er[(size_t)(s4 - 1) * N + (k - 1)];
    const int cnt = off[k] - off[k - 1];
    if (cnt > maxnbs) throw std::runtime_error("adopt_state: list longer than maxnbs");
    na_npt[k] = cnt;
    for (int q = 0; q < cnt; q++) nb[(size_t)(k - 1) * maxnbs + 1 + q] = lst[off[k - 1] + q];
  }
}

// checkover.f:21-131 (returns true when an overlap / bond violation exists)
bool Oracle::checkover(std::string* why) const {
  bool over = false;
  char buf[256];
  const int N = noptotal;
  for (int i = 1; i <= N - 1; i++)
    for (int j = i + 1; j <= N; j++) {
      int evcode = ev(j, i);
      if (evcode == 1 || evcode >= 15) {
        double sigsq = sigma_sq[identity[i]][identity[j]] * sq(ev_param[1][evcode]);
        if (evcode >= 22 && evcode <= 26) {
          int k = identity[i] > identity[j] ? identity[i] : identity[j];
          sigsq = sigsq * sq(sqz610[evcode - 21][k]);
        }
        PairGeom g = geom(&sv[(size_t)i * 6], &sv[(size_t)j * 6], tfalse);
        double rijsq = g.rx * g.rx + g.ry * g.ry + g.rz * g.rz;
        rijsq = rijsq * 1.0000000001;
        if (rijsq <= sigsq) {
          over = true;
          if (why && why->size() < 2000) {
            snprintf(buf, sizeof buf, "overlap %d %d code %d rij=%.6f sig=%.6f; ", i, j, evcode,
                     std::sqrt(rijsq) * boxl_orig, std::sqrt(sigsq) * boxl_orig);
            *why += buf;
          }
        }
      } else if (evcode >= 4 && evcode <= 12) {
        double blmin, blmax;
        bond_limits(*this, i, evcode, blmin, blmax);
        blmin = blmin * blmin;
        blmax = blmax * blmax;
        PairGeom g = geom(&sv[(size_t)i * 6], &sv[(size_t)j * 6], tfalse);
        double rijsq = g.rx * g.rx + g.ry * g.ry + g.rz * g.rz;
        double rijsq_min = rijsq * 1.0000000001, rijsq_max = rijsq * 0.9999999999;
        if (rijsq_max > blmax || rijsq_min < blmin) {
          over = true;
          if (why && why->size() < 2000) {
            snprintf(buf, sizeof buf, "bond %d %d code %d rij=%.6f [%.6f,%.6f]; ", i, j, evcode,
                     std::sqrt(rijsq) * boxl_orig, std::sqrt(blmin) * boxl_orig, std::sqrt(blmax) * boxl_orig);
            *why += buf;
          }
        }
      }
    }
  return over;
}

bool Oracle::terminal_ok(int i, int ii, int j, int jj) const {  // main.F90:1490-1491,1517-1518,1544-1545
  if (i <= nop1 && j <= nop1)
    return ii != chnln1 + 1 && ii != 3 * chnln1 && jj != chnln1 + 1 && jj != 3 * chnln1;
  if (i <= nop1 && j > nop1)
    return ii != chnln1 + 1 && ii != 3 * chnln1 && jj != numbeads1 + chnln2 + 1 && jj != numbeads1 + 3 * chnln2;
  if (i > nop1 && j > nop1)
    return ii != numbeads1 + chnln2 + 1 && ii != numbeads1 + 3 * chnln2 && jj != numbeads1 + chnln2 + 1 &&
           jj != numbeads1 + 3 * chnln2;
  // i in species 2, j in species 1: the Fortran has no branch (never reached from the event loop, i<j)
  return ii != numbeads1 + chnln2 + 1 && ii != numbeads1 + 3 * chnln2 && jj != chnln1 + 1 && jj != 3 * chnln1;
}

// restart path: inputinfo.f:76-101 (positions wrapped :89-91), main.F90:143-149, 197-234, 241-321, 389-424
void Oracle::set_state(const double* sv6xN, const int* bp) {
  const int N = noptotal;
  for (int k = 1; k <= N; k++) {
    for (int c = 0; c < 6; c++) sv[(size_t)k * 6 + c] = sv6xN[(size_t)(k - 1) * 6 + c];
    for (int c = 1; c <= 3; c++) S(c, k) = S(c, k) - dnint(S(c, k));  // inputinfo.f:89-91
  }
  // a fresh program start: the random-number stream starts over (main.F90:171-174 seeds drandm with the fixed iflag at
  // every `./dmd` run; D2: the counter of the replica's stream goes back to zero), identity and ev_code overlay are reset
  rng_ctr = 0;
  for (int l = 1; l <= nop1; l += numbeads1)
    for (int k = 1; k <= numbeads1; k++) identity[l + k - 1] = aa[k];
  for (int l = nop1 + 1; l <= nop1 + nop2; l += numbeads2)
    for (int k = 1; k <= numbeads2; k++) identity[l + k - 1] = aa[numbeads1 + k];
  make_code();
  t_fact = 0.00005;  // main.F90:144-149
  n_forced = 150.0;
  interval = t_fact / std::sqrt(setemp);
  interval_max = n_forced * interval;
  sortsize = interval_max / (double)numbin;
  avegtime = 0.00005 / std::sqrt(setemp);  // main.F90:156
  t = 0.0; tfalse = 0.0; old_tfalse = 0.0; nbin = 1; tbin_off = 0.0;  // main.F90:197-202
  coll = 0;
  for (int k = 1; k <= N; k++) {  // main.F90:205-223
    for (int c = 1; c <= 3; c++) S(c, k) = S(c, k) - dnint(S(c, k));
    old_rx[k] = S(1, k); old_ry[k] = S(2, k); old_rz[k] = S(3, k);
    tim[k] = interval_max + ltstep;
    coltype[k] = -1; nptnr[k] = -1;
    npt[k] = (k - 1) * maxnbs + 1;
    npt_dn[k] = (k - 1) * maxnbs + 1;
    bptnr[k] = 0;
    for (int kk = 1; kk <= 4; kk++) ER(k, kk) = 0;
  }
  npt[N + 1] = N * maxnbs + 1;
  npt_dn[N + 1] = N * maxnbs + 1;
  for (int k = N + 1; k <= N + 3; k++) { nptnr[k] = -2; coltype[k] = -2; }
  if (bp) for (int k = 1; k <= N; k++) bptnr[k] = bp[k - 1];  // main.F90:241-246
  // main.F90:249-321
  for (int k = 1; k <= N - 1; k++)
    for (int k_j = k + 1; k_j <= N; k_j++) {
      if (identity[k] + identity[k_j] == 5 && ev(k, k_j) == 15) {
        double rxij = S(1, k) - S(1, k_j), ryij = S(2, k) - S(2, k_j), rzij = S(3, k) - S(3, k_j);
        rxij = rxij - dnint(rxij); ryij = ryij - dnint(ryij); rzij = rzij - dnint(rzij);
        double rijsq = rxij * rxij + ryij * ryij + rzij * rzij;
        double diff = rijsq - welldia_sq[identity[k]][identity[k_j]];
        if (diff < 0.0) {
          int kk = local_index(k), kk_j = local_index(k_j);
          if (terminal_ok(k, kk, k_j, kk_j)) {
            if (identity[k] == 1) repuls_add(k, k_j); else repuls_add(k_j, k);
          }
        }
      }
      if (k_j == bptnr[k]) {
        if (identity[k] == 1) { identity[k] = 5; identity[k_j] = 8; }
        else { identity[k] = 8; identity[k_j] = 5; }
      }
    }
  numghosts = 0; nupdates = 0; nforcedupdate = 0;
  std::memset(nevents, 0, sizeof(nevents));
  // main.F90:389-398
  nbor_setup();
  num_cell = (int)(boxl / (sig_max_all * rl_const) * n_wrap) + 2 * n_wrap;
  cell.assign((size_t)num_cell * num_cell * num_cell + 2, 0);
  wrap_map.assign((size_t)num_cell * num_cell * num_cell + 1, 0);
  width = boxl / (double)(num_cell - 2 * n_wrap);
  half = boxl / 2.0;
  cell_link();
  nbor();
  // main.F90:408-423
  if (canon) {
    double tgho = 0.0;
    while (tgho < 1e-18 || tgho == 1.0) tgho = rng_uniform();
    tim[N + 1] = -1.0 * fdlibm_log(tgho) * avegtime * .0000001;
  } else {
    tim[N + 1] = 1000000000.0;
  }
  tim[N + 2] = interval;
  tim[N + 3] = 3.3 / (std::sqrt(setemp)) + 5;
  events();
  log.clear();
  energy_log.clear();
}

void Oracle::set_temperature(double tstar) {
  // a new "./dmd < temp_0xx" run on the current configuration: positions are advanced to true positions,
  // bptnr is kept (restart files), everything else is rebuilt by the start-up path.
  sync_positions();
  setemp = tstar * 12.0;
  const int N = noptotal;
  std::vector<double> s((size_t)N * 6);
  std::vector<int> bp(N);
  for (int k = 1; k <= N; k++) {
    for (int c = 0; c < 6; c++) s[(size_t)(k - 1) * 6 + c] = sv[(size_t)k * 6 + c];
    bp[k - 1] = bptnr[k];
  }
  set_state(s.data(), bp.data());
}

// Replica-exchange temperature change on resident state.  NOT in the reference: defined here and
// in the engine identically -- true positions (not wrapped), velocities scaled by sqrt(T_new/T_old), time constants of
// main.F90:143-156 reset, calendar re-derived; the neighbour lists and the positions they were built at (old_r*) are
// kept (positions do not change, so the lists stay valid), H-bond state kept.
void Oracle::retemp(double tstar_new) {
  const int N = noptotal;
  const double tf = tfalse;
  const double setemp_new = tstar_new * 12.0;
  const double scale = std::sqrt(setemp_new / setemp);
  t = t + tf;
  for (int k = 1; k <= N; k++) {
    const double x = S(1, k) + S(4, k) * tf, y = S(2, k) + S(5, k) * tf, z = S(3, k) + S(6, k) * tf;
    S(1, k) = x; S(2, k) = y; S(3, k) = z;
    S(4, k) = S(4, k) * scale; S(5, k) = S(5, k) * scale; S(6, k) = S(6, k) * scale;
  }
  tfalse = 0.0; old_tfalse = 0.0;
  setemp = setemp_new;
  t_fact = 0.00005; n_forced = 150.0;
  interval = t_fact / std::sqrt(setemp);
  interval_max = n_forced * interval;
  sortsize = interval_max / (double)numbin;
  avegtime = 0.00005 / std::sqrt(setemp);
  tbin_off = 0.0; nbin = 1;
  double tg = 1000000000.0;
  if (canon) {
    double tgho = 0.0;
    while (tgho < 1e-18 || tgho == 1.0) tgho = rng_uniform();
    tg = -1.0 * fdlibm_log(tgho) * avegtime;
  }
  tim[N + 1] = tg;
  tim[N + 2] = interval;
  tim[N + 3] = 3.3 / (std::sqrt(setemp)) + 5;
  for (int k = 1; k <= N; k++) { tim[k] = interval_max + ltstep; coltype[k] = -1; nptnr[k] = -1; }
  events();
}

// main.F90:1288-1297
void Oracle::sync_positions() {
  for (int k = 1; k <= noptotal; k++) {
    S(1, k) = S(1, k) + S(4, k) * tfalse;
    S(2, k) = S(2, k) + S(5, k) * tfalse;
    S(3, k) = S(3, k) + S(6, k) * tfalse;
    S(1, k) = S(1, k) - dnint(S(1, k));
    S(2, k) = S(2, k) - dnint(S(2, k));
    S(3, k) = S(3, k) - dnint(S(3, k));
  }
  // NOTE: the reference does this only at program end (tfalse is then irrelevant).  For a mid-run sync the
  // calendar must stay consistent, so the caller is expected to re-initialise (set_state) afterwards.
}

// worker block main.F90:1429-1959 (truth) with current state (D1), then master :926,943
void Oracle::pair_event(int i) {
  const int j = nptnr[i];
  const int evcode = ev(i, j);
  bool xpulse_del = false;
  int hb_partner = 0;
  const int ii = local_index(i), jj = local_index(j);
  if (coltype[i] == 7) {  // :1487-1574
    if (ER(i, 4) == 0 && ER(j, 4) == 0) {
      if (terminal_ok(i, ii, j, jj)) {
        double rating = identity[i] < identity[j] ? repuls_check(i, j) : repuls_check(j, i);
        coltype[i] = rating <= 10.0 ? 4 : 14;
      } else {
        double ran_non = rng_uniform();
        coltype[i] = ran_non <= 0.2 ? 4 : 9;
      }
    } else {
      coltype[i] = 9;
    }
  } else if (coltype[i] == 10) {  // :1575-1606
    double rating;
    if (evcode < 45) {
      hb_partner = ER(i, 4);
      rating = identity[i] < identity[hb_partner] ? repuls_check_3(i, hb_partner, j) : repuls_check_3(hb_partner, i, j);
    } else {
      hb_partner = ER(j, 4);
      rating = identity[j] < identity[hb_partner] ? repuls_check_3(j, hb_partner, i) : repuls_check_3(hb_partner, j, i);
    }
    coltype[i] = rating <= 10.0 ? 5 : 15;
  } else if (coltype[i] == 12) {  // :1607-1634
    if (evcode < 45) {
      hb_partner = ER(i, 4);
      if (bptnr[i] == hb_partner) coltype[i] = check_sigma(i, hb_partner) > 10.0 ? 6 : 13;
      else coltype[i] = 15;
    } else {
      hb_partner = ER(j, 4);
      if (bptnr[j] == hb_partner) coltype[i] = check_sigma(j, hb_partner) > 10.0 ? 6 : 13;
      else coltype[i] = 15;
    }
  }
  if (coltype[i] < 14) eventdyn(i, j, evcode);  // :1636
  const int ct = coltype[i];
  if (ct == 20) {  // :1638-1697
    if (identity[i] + identity[j] == 5) {
      bptnr[i] = j; bptnr[j] = i;
      identity[i] = identity[i] + 4; identity[j] = identity[j] + 4;
      if (terminal_ok(i, ii, j, jj)) {
        if (identity[i] < identity[j]) repuls_add(i, j); else repuls_add(j, i);
      }
    }
  } else if (ct == 21) {  // :1698-1763
    if (identity[i] <= 8) {
      if (terminal_ok(i, ii, j, jj)) {
        if (identity[i] < identity[j]) repuls_del_a(i, j); else repuls_del_a(j, i);
        xpulse_del = true;
      }
      if (bptnr[i] == j) {
        bptnr[i] = 0; bptnr[j] = 0;
        identity[i] = identity[i] - 4; identity[j] = identity[j] - 4;
      }
    }
  } else if (ct == 24) {  // :1766-1795
    int x = evcode < 45 ? i : j;
    hb_partner = ER(x, 4);
    if (identity[x] + identity[hb_partner] == 5) {
      bptnr[x] = hb_partner; bptnr[hb_partner] = x;
      identity[x] = identity[x] + 4; identity[hb_partner] = identity[hb_partner] + 4;
    }
  } else if (ct == 25) {  // :1797-1825
    int x = evcode < 45 ? i : j;
    hb_partner = ER(x, 4);
    if (identity[x] >= 5) {
      bptnr[x] = 0; bptnr[hb_partner] = 0;
      identity[x] = identity[x] - 4; identity[hb_partner] = identity[hb_partner] - 4;
    }
  } else if (ct == 14) {  // :1827-1877
    bumpoff(i, j, evcode);
    if (terminal_ok(i, ii, j, jj)) {
      if (identity[i] < identity[j]) repuls_add(i, j); else repuls_add(j, i);
    }
  } else if (ct == 15) {  // :1879-1881
    bumpoff(i, j, evcode);
  } else if (ct == 16) {  // :1882-1935
    bumpoff(i, j, evcode);
    if (terminal_ok(i, ii, j, jj)) {
      if (identity[i] < identity[j]) repuls_del_a(i, j); else repuls_del_a(j, i);
      xpulse_del = true;
    }
  }
  if (ct >= 0 && ct < 32) nevents[ct]++;  // main.F90:926
  if (log.size() < log_capacity) log.push_back({t + tfalse, i, j, ct, evcode});
  partial_events(i, j, xpulse_del);  // main.F90:943
}

// main.F90:997-1049
void Oracle::ghost_event() {
  const int N = noptotal;
  int i;
  do { i = (int)(rng_uniform() * N) + 1; } while (i == N + 1);
  numghosts++;
  S(1, i) = S(1, i) + S(4, i) * tfalse;
  S(2, i) = S(2, i) + S(5, i) * tfalse;
  S(3, i) = S(3, i) + S(6, i) * tfalse;
  double v1, v2, r, fact;
  do { v1 = 2.0 * rng_uniform() - 1.0; v2 = 2.0 * rng_uniform() - 1.0; r = v1 * v1 + v2 * v2; } while (r == 0.0 || r >= 1.0);
  fact = std::sqrt(-2.0 * setemp * bm[i] * fdlibm_log(r) / r);
  S(4, i) = v1 * fact / bm[i];
  S(5, i) = v2 * fact / bm[i];
  do { v1 = 2.0 * rng_uniform() - 1.0; v2 = 2.0 * rng_uniform() - 1.0; r = v1 * v1 + v2 * v2; } while (r == 0.0 || r >= 1.0);
  fact = std::sqrt(-2.0 * setemp * bm[i] * fdlibm_log(r) / r);
  S(6, i) = v1 * fact / bm[i];
  S(1, i) = S(1, i) - S(4, i) * tfalse;
  S(2, i) = S(2, i) - S(5, i) * tfalse;
  S(3, i) = S(3, i) - S(6, i) * tfalse;
  if (tlinks2[N + 1] != 0) del_tbin(N + 1);
  double tgho = 0.0;
  while (tgho < 1e-18 || tgho == 1.0) tgho = rng_uniform();
  tim[N + 1] = -1.0 * fdlibm_log(tgho) * avegtime + tfalse;
  if (tim[N + 1] < interval_max) add_tbin(N + 1);
  if (tfalse < old_tfalse) tfalse = old_tfalse;
  if (log.size() < log_capacity) log.push_back({t + tfalse, N + 1, i, -2, 0});
  partial_events(i, 0, false);
}

// main.F90:1126-1187
void Oracle::interval_event() {
  const int N = noptotal;
  t = t + tfalse;
  for (int k = 1; k <= N + 3; k++) tim[k] = tim[k] - tfalse;
  interval_max = interval_max - tfalse;
  tbin_off = tbin_off + tfalse;
  for (int k = 1; k <= N; k++) {
    S(1, k) = S(1, k) + S(4, k) * tfalse;
    S(2, k) = S(2, k) + S(5, k) * tfalse;
    S(3, k) = S(3, k) + S(6, k) * tfalse;
  }
  tfalse = 0.0;
  bool update = displ();
  if (update || interval > interval_max) {
    if (!update) {
      nforcedupdate++;
      n_forced = n_forced * 1.01;
    }
    interval_max = interval * n_forced;
    sortsize = interval_max / (double)numbin;
    tbin_off = 0.0;
    nupdates++;
    for (int k = 1; k <= N; k++) {
      S(1, k) = S(1, k) - dnint(S(1, k));
      S(2, k) = S(2, k) - dnint(S(2, k));
      S(3, k) = S(3, k) - dnint(S(3, k));
      old_rx[k] = S(1, k); old_ry[k] = S(2, k); old_rz[k] = S(3, k);
    }
    nbor();
    for (int k = 1; k <= N; k++) { tim[k] = interval_max + ltstep; coltype[k] = -1; nptnr[k] = -1; }
    nbin = 1;  // (the reference resets nbin after events(); D3's clamp needs it before)
    events();
  }
  if (tlinks2[N + 2] != 0) del_tbin(N + 2);
  tim[N + 2] = interval * 0.999;
  add_tbin(N + 2);
  if (log.size() < log_capacity) log.push_back({t + tfalse, N + 2, 0, -2, 0});
}

// main.F90:1191-1246
void Oracle::output_event() {
  const int N = noptotal;
  energy_log.push_back(energy());
  if (tlinks2[N + 3] != 0) del_tbin(N + 3);
  tim[N + 3] = 3.3 / (std::sqrt(setemp)) + 5 + tfalse;
  if (tim[N + 3] < interval_max) add_tbin(N + 3);
  if (log.size() < log_capacity) log.push_back({t + tfalse, N + 3, 0, -2, 0});
}

// one iteration of main.F90:484-1258 with serial semantics (App. E): pop the earliest calendar entry
void Oracle::step() {
  const int N = noptotal;
  while (bin[nbin] == 0) {
    nbin++;
    if (nbin > numbin + 1) throw std::runtime_error("oracle: calendar empty");
  }
  int o = bin[nbin];
  for (int k = tlinks[o]; k != 0; k = tlinks[k])
    if (tim[k] < tim[o] || (tim[k] == tim[o] && k < o)) o = k;  // D3: earliest, ties -> lowest index
  tfalse = tim[o];
  coll++;
  if (o <= N) pair_event(o);
  else if (o == N + 1) ghost_event();
  else if (o == N + 2) interval_event();
  else output_event();
  old_tfalse = tfalse;
}

void Oracle::run(long long n_events) {
  for (long long n = 0; n < n_events; n++) step();
}

}  // namespace dmdo
