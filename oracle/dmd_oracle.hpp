// dmd_oracle.hpp -- CPU ORACLE (test infrastructure, NOT product code).
//
// A routine-by-routine C++17 restatement of the hot path of
// HallandSantiso-NCSU/Parallel-DMD-for-biomolecules (Fortran, /root/reference/parallel-dmd-PRIME20/code),
// with single-address-space ("serial") semantics: every read sees current state.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// PARITY STATUS: "parity unpinned" for event times / partners / sequences -- the reference ships no golden
// vectors, known-answer tests or fixtures for this path, and it cannot be compiled here (no Fortran compiler,
// Intel-only drandm/dtime, MPI).  What IS pinned (tests/test_oracle_golden.py): the I/O
// formats and system-A snapshot (genconfig/results/run0000.*), masses, sum m v^2 = 12096 (genconfig/checks),
// the derived constants of SURVEY.md App. C, the static ev_code histogram, and the reference's own
// invariants (checkover.f, NVE nint(E) conservation main.F90:928-942) after every committed event.
//
// Deliberate, documented deviations from the Fortran (DESIGN.md "Oracle decisions"):
//   D1 worker-side stale state is dropped: geometry checks read current state (main.F90:1435-1464).
//   D2 Intel drandm is replaced by the counter RNG splitmix64(seed + n*golden) -> 53-bit uniform.
//   D3 the calendar pops the global minimum: add_tbin clamps the bucket index to >= nbin so an event
//      predicted into an already-passed bucket is processed next instead of being delayed to the next
//      rebuild (main.F90:496-498); ties in time are broken by the lowest bead index.
//   D4 log() for the ghost thermostat is the fdlibm algorithm evaluated without FMA, so that the CUDA
//      engine can reproduce it bit for bit.
// All arithmetic is fp64, evaluated left to right as written in the Fortran, no FMA contraction
// (-ffp-contract=off), dnint == round-half-away (std::round), int() == truncation.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../include/dmdb200.h"

namespace dmdo {

struct EventRec {
  double t;
  int i, j, type, evcode;
};

struct EnergyRec {
  long long coll;
  double t, ered, tred, sumvel, ehh_ii, ehh_ij;
  int hb_alpha, hb_ii, hb_ij;
};

class Oracle {
 public:
  Oracle(const dmdb_params& p, const dmdb_topology& topo, const dmdb_tables& tab);

  // ---- restart path: inputinfo.f:76-101 + main.F90:205-321, 347-424
  void set_state(const double* sv6xN, const int* bptnr_or_null);
  void set_temperature(double tstar);
  void retemp(double tstar_new);  // replica-exchange temperature change on resident state (new functionality)

  // ---- reference operator API
  void nbor_setup();                                   // nbor_setup.f:13-118
  void cell_link();                                    // cell_link.f:16-94
  void cell_add();                                     // cell_add.f:12-28
  void nbor();                                         // nbor.f:33-137
  bool displ();                                        // displ.f:20-46 (returns update)
  void events();                                       // events.f:23-123
  void partial_events(int i, int j, bool xpulse_del);  // partial_events.f:16-201
  void eventredo_up(int i, int j);                     // eventredo_up.f:25-56
  void eventredo_down(int i, int j);                   // eventredo_down.f:25-78
  void core(int i, int j, int evcode, double& tij, int& type) const;      // core.f:14-40
  void bond(int i, int j, int evcode, double& tij, int& type) const;      // bond.f:27-126
  void sqwel(int i, int j, int evcode, double& tij, int& type) const;     // sqwel.f:15-64
  void nc_sqwel(int i, int j, int evcode, double& tij, int& type) const;  // nc_sqwel.f:19-122
  void sqshlder(int i, int j, int evcode, double& tij, int& type) const;  // sqshlder.f:15-63
  void eventdyn(int i, int j, int evcode);             // eventdyn.f:18-381
  void bumpoff(int i, int j, int evcode);              // bumped.f:12-43
  void add_tbin(int i);                                // add_tbin.f:12-32 (+D3)
  void del_tbin(int i);                                // del_tbin.f:12-20
  void repuls_add(int i, int j);                       // repuls_add.f:14-47
  void repuls_del_a(int i, int j);                     // repuls_del_a.f:14-37
  void repuls_del_b(int i, int j);                     // repuls_del_b.f:14-39
  double repuls_check(int i, int j) const;             // repuls_check.f:17-81 (returns rating)
  double repuls_check_3(int i, int j, int k) const;    // repuls_check_3.f:16-103
  double check_sigma(int i, int j) const;              // check_sigma.f:12-29
  EnergyRec energy() const;                            // energy.f:25-101
  bool checkover(std::string* why = nullptr) const;    // checkover.f:21-131 (returns over)
  // check_nc_int.f:21-360 (called at main.F90:425 and, under -Ddebugging, at every output event :1199): the audit of
  // the H-bond <-> auxiliary-shoulder state machine.  Adds to boundbad / unboundbad like the Fortran, returns
  // m_ss / n_ss (the reference exits when they differ) and the number of its "no ss for ..." complaints.
  struct NcAudit {
    int boundbad = 0, unboundbad = 0, m_ss = 0, n_ss = 0, no_ss = 0, pairs15 = 0;
  };
  NcAudit check_nc_int() const;
  // test hook: overwrite the live state with one read back from another engine (no restart reconstruction), so that
  // the same audit runs on device state.  Arrays 0-based, bead indices 1-based (0 = none), lists as dmdb_get_nbors.
  void adopt_state(const double* sv6xN, double tfalse_, const int* bptnr_, const int* identity_, const int* extra_repuls_Nx4,
                   const int* nb_offsets, const int* nb_list);

  // ---- main loop, serial semantics (main.F90:484-1258)
  void step();                      // one calendar event
  void run(long long n_events);
  void sync_positions();            // main.F90:1288-1295

  // ---- accessors (1-based arrays, slot 0 unused)
  int N() const { return noptotal; }
  int ev(int i, int j) const { return ev_code[(size_t)i * (noptotal + 1) + j]; }
  int cell_of(int k) const;         // cell id of cell_add.f:25 for the current sv
  double rng_uniform();             // D2

  // sizes / flags
  int nop1, nop2, chnln1, chnln2, numbeads1, numbeads2, noptotal, nch1, nch2;
  int n_wrap, n_nab_cell, numbin = 2000, maxnbs = 256;
  bool canon, no_hbs;
  uint64_t seed, rng_ctr = 0;

  // tables (1-based; [0] unused)
  double sigma[29], welldia[29], epsilon[29], bmass[29];
  double sigma_sq[29][29], sigma_2b[29][29], welldia_sq[29][29], ep_sqrt[29][29], shlddia_sq[29][29];
  double ep[29][29], bds[29][29], wel[29][29];
  double ev_param[4][51], sqz610[6][29], rlsq[51];
  std::vector<double> bdln, bl_rn, bl_rc, del_bdln, del_blrn, del_blrc;
  double shder_dist1, shder_dist2, shder_dist3, shder_dist4;
  double boxl, boxl_orig, setemp, sig_max_all, hdelr, width, half;
  double t = 0, tfalse = 0, old_tfalse = 0, interval, t_fact, interval_max, sortsize, tbin_off = 0, n_forced;
  double avegtime;
  int num_cell = 0, nbin = 1;
  long long coll = 0, numghosts = 0, nupdates = 0, nforcedupdate = 0;
  long long nevents[32] = {0};
  long long n_pair_predictions = 0, n_nbr_visits = 0;

  std::vector<int> aa, hp, hp1, hp2, fside1, fside2;
  std::vector<int> identity, chnnum, bptnr, coltype, nptnr, extra_repuls;  // extra_repuls[(k)*5 + slot]
  std::vector<double> sv, bm, tim, old_rx, old_ry, old_rz;                 // sv[k*6 + c-1]
  std::vector<int8_t> ev_code;
  std::vector<int> npt, nb, na_npt, npt_dn, dnnab, nnabdn;
  std::vector<int> tlinks, tlinks2, bin, cell, wrap_map, clinks, map;

  std::vector<EventRec> log;
  size_t log_capacity = 0;
  std::vector<EnergyRec> energy_log;

 private:
  void inputinfo(const dmdb_topology& topo, const dmdb_tables& tab);  // inputinfo.f:105-411
  void scale_down();                                                   // scale_down.f:27-79
  void make_code();                                                    // make_code.f:18-566
  void setec(int k, int l, int v);
  int local_index(int i) const;  // the reference's ii / jj (main.F90:1468-1481)
  bool terminal_ok(int i, int ii, int j, int jj) const;  // "didn't involve an end bead" tests
  void redo_full(int l);
  void pair_event(int i);
  void ghost_event();
  void interval_event();
  void output_event();
  inline double& S(int c, int k) { return sv[(size_t)k * 6 + (c - 1)]; }
  inline double S(int c, int k) const { return sv[(size_t)k * 6 + (c - 1)]; }
  inline int& ER(int k, int s) { return extra_repuls[(size_t)k * 5 + s]; }
  inline int ER(int k, int s) const { return extra_repuls[(size_t)k * 5 + s]; }
  mutable int ncim1 = 0, ncai = 0, ncaj = 0, nnjp1 = 0;
};

double fdlibm_log(double x);  // D4

}  // namespace dmdo
