// dmd_types.h -- data layout of the B200 engine (host + device).
//
// Design: ONE WARP OWNS ONE REPLICA (an independent PRIME20 trajectory).  All per-replica state lives in
// HBM in structure-of-arrays blocks with a fixed replica stride; a warp streams its replica's data with
// coalesced 32-wide accesses (calendar groups, position advance) and 64-byte bead-record gathers (pair
// prediction).  Nothing is shared between warps, so the event loop needs no atomics and no block barriers.
//
// Reference state this replaces: module `global`, code/header.f:14-63 (sv, tim, nptnr, coltype, identity,
// bptnr, extra_repuls, ev_code, nb/dnnab, bin/tlinks, cell/clinks, scalars).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DMD_HD __host__ __device__ __forceinline__
#else
#define DMD_HD inline
#endif

namespace dmd {

// ---- one bead, 64 bytes, 64-byte aligned: a partner gather is exactly two 32-byte sectors ----------------
// sv(1:6,k) of header.f:42 + identity (header.f:22) + bptnr + the bead's two auxiliary partners
// extra_repuls(k,1:2) (header.f:20) + the ev_code overlay (40/50/1) of those two pairs as seen from row k.
struct alignas(64) BeadRec {
  double x, y, z, vx, vy, vz;
  int32_t bptnr;   // 0-based partner or -1
  int32_t er1;     // extra_repuls(k,1), 0-based or -1
  int32_t er2;     // extra_repuls(k,2)
  uint8_t ident;   // identity 1..28 (changes by +-4 on H-bond formation, main.F90:1643)
  uint8_t ov1;     // ev_code(k, er1): 1 (off), 40 or 50
  uint8_t ov2;     // ev_code(k, er2)
  uint8_t pad;
};
static_assert(sizeof(BeadRec) == 64, "BeadRec must be 64 bytes");

// per-bead static topology word (shared by all replicas): what make_code.f derives its N x N matrix from
//   bits 0-1 class (0 Ca, 1 N, 2 C, 3 R) | 2-11 residue index 1..L | 12 hp flag | 13-17 initial identity |
//   18 species | 19 proline-exclusion flag for N beads (make_code.f:146-168) | 20-29 bead index local to chain
DMD_HD uint32_t meta_pack(int cls, int res, int hp, int id0, int sp, int proex, int local) {
  return (uint32_t)cls | ((uint32_t)res << 2) | ((uint32_t)hp << 12) | ((uint32_t)id0 << 13) | ((uint32_t)sp << 18) |
         ((uint32_t)proex << 19) | ((uint32_t)local << 20);
}
DMD_HD int meta_cls(uint32_t m) { return m & 3; }
DMD_HD int meta_res(uint32_t m) { return (m >> 2) & 1023; }
DMD_HD int meta_hp(uint32_t m) { return (m >> 12) & 1; }
DMD_HD int meta_id0(uint32_t m) { return (m >> 13) & 31; }
DMD_HD int meta_sp(uint32_t m) { return (m >> 18) & 1; }
DMD_HD int meta_proex(uint32_t m) { return (m >> 19) & 1; }
DMD_HD int meta_local(uint32_t m) { return (m >> 20) & 1023; }

// neighbour-list entry: static ev_code class in the top 5 bits, partner index in the low 27
constexpr int NB_SHIFT = 27;
constexpr uint32_t NB_MASK = (1u << NB_SHIFT) - 1;

constexpr int MAX_RES = 512;   // residues over both species (bond-limit tables)

// 28 x 28 pair tables of scale_down.f:37-67, index (id_i-1)*28 + (id_j-1)
struct PairTables {
  double sigma_sq[784];
  double welldia_sq[784];
  double shlddia_sq[784];
  double ep_sqrt[784];
};
// the two tables every prediction reads (a prefix of PairTables): staged into shared memory.  Shoulder diameters and
// well depths are needed by the rare H-bond / side-chain-well events only and stay in global memory (SysConst), so
// that the event-loop CTA needs less than 32 KB of shared memory and the SM keeps 224 KB of L1
struct HotTables {
  double sigma_sq[784];
  double welldia_sq[784];
};

// the constants every pair prediction / collision reads (a copy of the SysConst fields of the same name): staged
// into shared memory next to the pair tables so that the hot loop never goes to L2 for a squeeze factor or a mass
constexpr int HOT_MAX_RES = 64;  // per-residue bond windows staged in shared memory up to this many residues
struct HotConst {
  double ev_param1[51];
  double ev_param2[51];
  double ev_param3[51];
  double sqz610[5 * 29];
  double bmass[29];
  double binv[29];  // RN(1 / bmass): lets a division by a mass be done exactly with 5 multiply-adds (dmd_physics.h)
  int32_t chnln0, nres;
};

// everything that is constant during a run and identical for all replicas
struct SysConst {
  int32_t N;                 // beads per replica
  int32_t n_species, nop1;   // nop1 = beads of species 1
  int32_t nch[2], chnln[2], numbeads[2];
  int32_t n_wrap, num_cell, ncr;  // num_cell per dim incl. ghosts (main.F90:390), ncr = real cells per dim
  int32_t canon, no_hbs;
  int32_t cap;               // per-bead capacity of each neighbour list
  int32_t ngroups;           // ceil((N+3)/32) calendar groups
  int32_t log_cap, out_cap;
  int32_t sct_off[2];        // offsets of the species' rows in the same-chain class table (DevArrays.sctab), -1: none
  double ev_param1[51];      // make_code.f:18-39 (squeeze factors), index = ev_code
  double ev_param2[51];      // blmin of codes 4-9 (make_code.f:49-57)
  double ev_param3[51];      // blmax
  double sqz610[5 * 29];     // inputinfo.f:382-389, [(code-22)*29 + identity]
  double rlsq[51];           // nbor_setup.f:112-115
  double bmass[29];          // inputinfo.f:395-404
  double shder[4];           // scale_down.f:32-35
  double eps1;               // epsilon(1)
  double eps_hb;             // ep_sqrt(5,8), energy.f:74
  double hdelr, width, half, sig_max_all, boxl_orig;
  double rl_max;             // largest list cut-off, sqrt(max rlsq): reach of the chain-wise list rebuild
  int32_t chainwise;         // 1: the chain-wise list rebuild applies (<= 64 chains of <= 64 beads, rlsq(40..50) == rlsq(1))
  int32_t pad_cw;
  double shlddia_sq[784];    // copies of PairTables.shlddia_sq / ep_sqrt (cold: shoulder events, well depths, energy)
  double ep_sqrt[784];
  // per-residue side-chain bond limits, already multiplied out like bond.f:82-91:
  // [kind 0:R-Ca(10) 1:R-N(11) 2:R-C(12)][species residue index: species 0 -> r-1, species 1 -> chnln[0]+r-1]
  double blmin_sc[3][MAX_RES];
  double blmax_sc[3][MAX_RES];
};

// per-replica scalars (main.F90 locals + module scalars header.f:47-48)
struct alignas(128) RepScalars {
  double t, tfalse, old_tfalse;
  double setemp, interval, t_fact, interval_max, n_forced, avegtime;
  int64_t coll;
  uint64_t rng_seed, rng_ctr;
  int64_t nevents[32];
  int64_t numghosts, nupdates, nforcedupdate;
  int64_t n_pair_pred, n_nbr_visits;
  int32_t n_log, n_out;
  int32_t error;        // 0 ok; see DMD_E_* below
  int32_t error_info;
  int32_t pad[4];
};

constexpr int DMD_E_NBR_CAP = 1;    // neighbour list capacity exceeded
constexpr int DMD_E_CAL_EMPTY = 2;  // calendar has no finite entry
constexpr int DMD_E_NEG_TIME = 3;   // tij < -1e-10 (events.f:59-73 debugging guard)
constexpr int DMD_E_GRID = 4;       // bead outside the cell grid
constexpr int DMD_E_BAD_INPUT = 5;  // bptnr entry out of range (run start)
constexpr int DMD_E_SERVICE = 6;    // the list-rebuild service did not answer (event loop, dmd_cuda.cu)

// one calendar entry: tim(k), nptnr(k), coltype(k) of header.f:20-22,47 side by side so that popping the
// minimum delivers the whole event in one access.  type: low 8 bits coltype (as int8), bits 8-15 the static
// ev_code class of the (owner, partner) pair.
struct alignas(16) CalEnt {
  double t;
  int32_t ptnr;  // 0-based partner, -1 none, -2 pseudo-event
  int32_t type;
};
static_assert(sizeof(CalEnt) == 16, "CalEnt must be 16 bytes");

struct EventLogRec {  // == dmdb_event
  double t;
  int32_t i, j, type, evcode;
};

struct OutRec {  // one output pseudo-event (main.F90:1207-1210)
  int64_t coll;
  double t, ered, tred, sumvel, ehh_ii, ehh_ij;
  int32_t hb_alpha, hb_ii, hb_ij, pad;
};

// device pointers; element (replica r, index k) of array A with per-replica length n is A[r*n + k]
struct DevArrays {
  // shared by all replicas
  const SysConst* sys;
  const PairTables* tables;
  const HotConst* hot;
  const int32_t* nc_beads;  // ascending indices of the N and C beads (run-start fix-up, main.F90:249-321)
  const double* bl;       // nres x 6: per-residue side-chain bond windows (min,max of codes 10, 11, 12), bond.f:82-91
  const uint32_t* meta;   // N
  const int32_t* chain;   // N (global chain index)
  const uint8_t* sctab;   // static_code() of every bead pair of ONE chain of each species, [local_a * numbeads + local_b]
  // per replica
  BeadRec* rec;           // N
  CalEnt* cal;            // N+3 entries padded to ngroups*32 (padding t = 1e300)
  int32_t* er34;          // 2N: extra_repuls(k,3), extra_repuls(k,4), 0-based or -1
  uint32_t* up;           // N*cap
  uint32_t* dn;           // N*cap
  uint16_t* nup;          // N
  uint16_t* ndn;          // N
  double* oldr;           // 3N
  int32_t* cellhead;      // ceil(ncr/2)^3: list heads of the coarse grid (2 x 2 x 2 fine cells)
  uint32_t* cpk;          // N: packed fine cell coordinates (10 bits per dimension)
  int32_t* cnext;         // N
  int32_t* cellof;        // N (reference cell id of cell_add.f:25, for parity read-back)
  double* tmin1;          // ngroups
  RepScalars* scal;       // 1
  EventLogRec* log;       // log_cap
  OutRec* out;            // out_cap
  long long* blkstat;     // 8: batching statistics of the CTA-per-replica engine (dmd_block.h)
  int32_t n_replicas;
  int32_t cal_stride;     // ngroups*32
  int32_t n_beads;        // N (host-side copy of sys->N)
  int32_t nres;           // residues over both species (rows of bl)
  int32_t n_nc;           // entries of nc_beads
  // copies of the SysConst fields the per-replica address arithmetic needs: as kernel parameters they sit in the
  // constant bank, so a replica's array bases can be recomputed instead of being held (or spilled) in registers
  int32_t cap, ngroups, log_cap, out_cap;
  int32_t ncc3;           // coarse cells per replica (entries of cellhead)
  int32_t n_chains;       // chains over both species
  int32_t chainwise;      // copy of SysConst.chainwise
  // list-rebuild service of the warp-per-replica engine (dmd_cuda.cu: a few CTAs of the event-loop kernel do the
  // neighbour-list + calendar rebuilds for the warps of all other CTAs, so that the event-loop SMs keep only the
  // hot loop in their 32 KB instruction caches)
  int32_t* svc_flag;      // per replica: 0 idle, 1 rebuild requested, 2 being served, 3 taken back by its own warp
  unsigned long long* svc_ctl;  // [0] finished worker warps [1] rebuilds served [2] requests taken back [3] wait cycles [4] service cycles;
                                // from SVC_Q_HEAD on: the request queue (below)
};
constexpr int SVC_CTL_WORDS = 8;  // counters, cleared before every launch
// behind them, never cleared: the FIFO of pending requests.  A requester takes a ticket (SVC_Q_TAIL), writes
// (ticket + 1) << 24 | replica into slot ticket % capacity; a free service group claims the next ticket (SVC_Q_HEAD) and
// reads the slot once its tag says it is written.  A replica has one request outstanding at most and the capacity is
// twice the replica count, so a slot normally is read long before it comes round again; only the tickets of requests
// that were taken back (they stay queued and are skipped when claimed) can pile up behind a service that is far too
// slow -- a group that finds a LATER tag in its slot skips the ticket (tests/test_service_queue_model.py).
constexpr int SVC_Q_HEAD = 8, SVC_Q_TAIL = 9, SVC_Q_CAP = 10, SVC_Q_RING = 16;

}  // namespace dmd
