// oracle_capi.cpp -- extern "C" face of the CPU ORACLE (test infrastructure, NOT product code) so that
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can drive it through
// ctypes.  Mirrors the getters of include/dmdb200.h so the parity tests read symmetrically.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <string>

#include "dmd_oracle.hpp"

using dmdo::Oracle;

static thread_local std::string g_err;

#define GUARD(body)                 \
  try {                             \
    body;                           \
    return 0;                       \
  } catch (const std::exception& e) { \
    g_err = e.what();               \
    return 1;                       \
  }

extern "C" {

const char* dmdo_last_error() { return g_err.c_str(); }

int dmdo_create(const dmdb_params* p, const dmdb_topology* topo, const dmdb_tables* tab, void** out) {
  GUARD(*out = new Oracle(*p, *topo, *tab))
}
void dmdo_destroy(void* h) { delete (Oracle*)h; }
int dmdo_num_beads(void* h) { return ((Oracle*)h)->N(); }
int dmdo_num_cells(void* h) { return ((Oracle*)h)->num_cell; }

int dmdo_set_state(void* h, const double* sv, const int32_t* bptnr) { GUARD(((Oracle*)h)->set_state(sv, bptnr)) }
int dmdo_set_temperature(void* h, double tstar) { GUARD(((Oracle*)h)->set_temperature(tstar)) }
int dmdo_retemp(void* h, double tstar) { GUARD(((Oracle*)h)->retemp(tstar)) }
int dmdo_nbor(void* h) { GUARD(((Oracle*)h)->nbor()) }
int dmdo_predict_all(void* h) {
  Oracle* o = (Oracle*)h;
  GUARD({
    // events() is only ever called by the reference right after all times were reset
    // (main.F90:1172-1177) with tfalse == 0
    for (int k = 1; k <= o->N(); k++) {
      o->tim[k] = o->interval_max + 1e-10;
      o->coltype[k] = -1;
      o->nptnr[k] = -1;
    }
    o->nbin = 1;
    o->events();
  })
}
int dmdo_run(void* h, int64_t n_events, double* seconds) {
  Oracle* o = (Oracle*)h;
  GUARD({
    auto t0 = std::chrono::steady_clock::now();
    o->run(n_events);
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  })
}
int dmdo_sync_positions(void* h) { GUARD(((Oracle*)h)->sync_positions()) }

int dmdo_get_cells(void* h, int32_t* cell_of_bead) {
  Oracle* o = (Oracle*)h;
  for (int k = 1; k <= o->N(); k++) cell_of_bead[k - 1] = o->cell_of(k);
  return 0;
}

int dmdo_get_nbors(void* h, int down, int32_t* offsets, int32_t* nbv) {
  Oracle* o = (Oracle*)h;
  int n = 0;
  for (int i = 1; i <= o->N(); i++) {
    offsets[i - 1] = n;
    int cnt = down ? o->nnabdn[i] : o->na_npt[i];
    int start = (i - 1) * o->maxnbs + 1;
    if (nbv) {
      for (int k = 0; k < cnt; k++) nbv[n + k] = down ? o->dnnab[start + k] : o->nb[start + k];
      std::sort(nbv + n, nbv + n + cnt);
    }
    n += cnt;
  }
  offsets[o->N()] = n;
  return 0;
}

int dmdo_get_calendar(void* h, double* tim, int32_t* nptnr, int32_t* coltype) {
  Oracle* o = (Oracle*)h;
  for (int k = 1; k <= o->N() + 3; k++) {
    tim[k - 1] = o->tim[k];
    nptnr[k - 1] = o->nptnr[k];
    coltype[k - 1] = o->coltype[k];
  }
  return 0;
}

int dmdo_get_state(void* h, double* sv, int32_t* bptnr, int32_t* identity, int32_t* extra_repuls, double* t,
                   double* tfalse, int64_t* coll) {
  Oracle* o = (Oracle*)h;
  const int N = o->N();
  for (int k = 1; k <= N; k++) {
    if (sv) for (int c = 0; c < 6; c++) sv[(size_t)(k - 1) * 6 + c] = o->sv[(size_t)k * 6 + c];
    if (bptnr) bptnr[k - 1] = o->bptnr[k];
    if (identity) identity[k - 1] = o->identity[k];
    if (extra_repuls)
      for (int s = 1; s <= 4; s++) extra_repuls[(size_t)(s - 1) * N + (k - 1)] = o->extra_repuls[(size_t)k * 5 + s];
  }
  if (t) *t = o->t;
  if (tfalse) *tfalse = o->tfalse;
  if (coll) *coll = o->coll;
  return 0;
}

int dmdo_get_evcode(void* h, int n_pairs, const int32_t* i, const int32_t* j, int32_t* code) {
  Oracle* o = (Oracle*)h;
  for (int k = 0; k < n_pairs; k++) code[k] = o->ev(i[k], j[k]);
  return 0;
}

int dmdo_get_evcode_matrix(void* h, int8_t* m /* N*N row-major: m[(i-1)*N + (j-1)] = ev_code(i,j) */) {
  Oracle* o = (Oracle*)h;
  const int N = o->N();
  for (int i = 1; i <= N; i++)
    for (int j = 1; j <= N; j++) m[(size_t)(i - 1) * N + (j - 1)] = (int8_t)o->ev(i, j);
  return 0;
}

int dmdo_energy(void* h, dmdb_energy* e) {
  Oracle* o = (Oracle*)h;
  dmdo::EnergyRec r = o->energy();
  e->ered = r.ered; e->tred = r.tred; e->sumvel = r.sumvel; e->ehh_ii = r.ehh_ii; e->ehh_ij = r.ehh_ij;
  e->hb_alpha = r.hb_alpha; e->hb_ii = r.hb_ii; e->hb_ij = r.hb_ij; e->reserved = 0;
  return 0;
}

int dmdo_checkover(void* h, char* why, int why_len) {
  Oracle* o = (Oracle*)h;
  std::string w;
  bool over = o->checkover(&w);
  if (why && why_len > 0) {
    std::strncpy(why, w.c_str(), why_len - 1);
    why[why_len - 1] = 0;
  }
  return over ? 1 : 0;
}

int dmdo_check_nc_int(void* h, int32_t* out /* boundbad, unboundbad, m_ss, n_ss, no_ss, pairs15 */) {
  Oracle* o = (Oracle*)h;
  const Oracle::NcAudit a = o->check_nc_int();
  out[0] = a.boundbad; out[1] = a.unboundbad; out[2] = a.m_ss; out[3] = a.n_ss; out[4] = a.no_ss; out[5] = a.pairs15;
  return 0;
}

int dmdo_adopt_state(void* h, const double* sv, double tfalse, const int32_t* bptnr, const int32_t* identity,
                     const int32_t* extra_repuls, const int32_t* nb_offsets, const int32_t* nb) {
  GUARD(((Oracle*)h)->adopt_state(sv, tfalse, bptnr, identity, extra_repuls, nb_offsets, nb))
}

int dmdo_get_event_log(void* h, int64_t first, int64_t n, dmdb_event* out, int64_t* n_out) {
  Oracle* o = (Oracle*)h;
  int64_t m = 0;
  for (int64_t k = first; k < first + n && k < (int64_t)o->log.size(); k++, m++) {
    const dmdo::EventRec& r = o->log[k];
    out[m].t = r.t; out[m].i = r.i; out[m].j = r.j; out[m].type = r.type; out[m].evcode = r.evcode;
  }
  *n_out = m;
  return 0;
}

int dmdo_get_stats(void* h, dmdb_stats* s) {
  Oracle* o = (Oracle*)h;
  std::memset(s, 0, sizeof(*s));
  s->events = o->coll;
  for (int k = 0; k < 32; k++) { s->nevents[k] = o->nevents[k]; s->pair_events += o->nevents[k]; }
  s->ghosts = o->numghosts;
  s->updates = o->nupdates - o->nforcedupdate;
  s->forced_updates = o->nforcedupdate;
  s->pair_predictions = o->n_pair_predictions;
  s->nbr_visits = o->n_nbr_visits;
  return 0;
}

// derived constants for the known-answer tests
int dmdo_get_constants(void* h, double* out /* 16 */, double* rlsq /* 50 */) {
  Oracle* o = (Oracle*)h;
  out[0] = o->sig_max_all; out[1] = o->hdelr; out[2] = o->width; out[3] = o->half;
  out[4] = o->setemp; out[5] = o->interval; out[6] = o->interval_max; out[7] = o->sortsize;
  out[8] = o->boxl_orig; out[9] = o->tim[o->N() + 3]; out[10] = o->ev_param[1][15]; out[11] = o->t_fact;
  out[12] = o->n_forced; out[13] = o->avegtime; out[14] = 0; out[15] = 0;
  for (int i = 1; i <= 50; i++) rlsq[i - 1] = o->rlsq[i];
  return 0;
}

int dmdo_get_masses(void* h, double* bm) {
  Oracle* o = (Oracle*)h;
  for (int k = 1; k <= o->N(); k++) bm[k - 1] = o->bm[k];
  return 0;
}

double dmdo_log(double x) { return dmdo::fdlibm_log(x); }
double dmdo_rng(void* h) { return ((Oracle*)h)->rng_uniform(); }

}  // extern "C"
