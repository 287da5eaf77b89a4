/*
 * dmdb200.h -- C ABI of libdmdb200.so, the B200-native PRIME20 DMD engine.
 *
 * Drop-in boundary for the simulation loop of
 * HallandSantiso-NCSU/Parallel-DMD-for-biomolecules (reference paths below are relative to
 * /root/reference/parallel-dmd-PRIME20/).  The reference has no FFI: its "operator API" is the set of
 * Fortran subroutines that code/main.F90 calls on module `global` (code/header.f:2-65).  Every entry
 * point here names the reference routine(s) it replaces.  A Fortran host binds these symbols through
 * ISO_C_BINDING (see parallel_dmd_for_biomolecules_b200/fortran/dmdb200_iso_c.f90 and INTEGRATION.md);
 * the Python/ctypes harness in parallel_dmd_for_biomolecules_b200/dmd.py binds the same symbols.
 *
 * Conventions (the reference's own, header.f / events.f):
 *   - bead indices crossing this boundary are 1-based; 0 means "none" (bptnr, extra_repuls),
 *     -1 means "no event" (nptnr, coltype), pseudo-event owners are N+1 (ghost), N+2 (interval),
 *     N+3 (output) with coltype -2 (main.F90:231-234);
 *   - sv is column-major 6 x N (x,y,z,vx,vy,vz per bead) in box units (box length 1), exactly the
 *     reference's sv(6,nop) (header.f:42);
 *   - all floating point is IEEE fp64 (the reference is built with -r8, qfile/script.sh:7);
 *   - every function returns 0 on success, non-zero on error (never exit()); the text of the last error
 *     is available from dmdb_last_error().  The caller owns every host buffer it passes in.
 *   - one host thread per handle; the library owns its CUDA stream(s).
 *   - there is NO CPU fallback: dmdb_create fails with DMDB_ERR_NO_DEVICE when no CUDA device exists.
 */
#ifndef DMDB200_H
#define DMDB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMDB_OK 0
#define DMDB_ERR_ARG 1
#define DMDB_ERR_NO_DEVICE 2
#define DMDB_ERR_CUDA 3
#define DMDB_ERR_STATE 4
#define DMDB_ERR_CAPACITY 5 /* a neighbour list or ring buffer overflowed; see dmdb_last_error */
#define DMDB_ERR_PHYSICS 6  /* device-side invariant tripped (e.g. tij < -1e-10, events.f:59-73) */

#define DMDB_MAX_SPECIES 2 /* the reference supports two peptide species (nop1/nop2, header.f:14) */

/* Raw contents of the reference's parameter files, exactly as read by code/inputinfo.f:162-404.
 * The library performs scale_down (scale_down.f:27-79), make_code (make_code.f:18-566, as a function of
 * topology instead of an N x N matrix) and nbor_setup (nbor_setup.f:13-118) itself. */
typedef struct dmdb_tables {
  double protein[12];   /* parameters/protein.data: sigma(N,Ca,R,C) well(N,Ca,R,C) eps(N,Ca,R,C) */
  double ep[400];       /* parametersep/ep19p_ha55a_weakhp.data, value of row (i,j) at [(i-9)*20+(j-9)],
                           file sign (the library stores -value like inputinfo.f:287) */
  double bds[400];      /* parameters/beadwell_ha55a.data bead diameter, same indexing */
  double wel[400];      /* parameters/beadwell_ha55a.data well diameter, same indexing */
  double mass[28];      /* parameters/mass.data reduced mass by identity id (index id-1); unused ids 0 */
  double rcarnrco[120]; /* parameters/rcarnrco.data 20 rows x 6: R-Ca, R-N, R-C length and tolerances */
  double sqz6to10[100]; /* parameters/sqz6to10.data 20 rows x 5 in FILE column order (sz8,sz6,sz7,sz9,sz10,
                           inputinfo.f:383) */
} dmdb_tables;

/* Chain topology: what the reference fixes at compile time with -Dnop1 -Dnop2 -Dchnln1 -Dchnln2
 * -Dnumbeads1 -Dnumbeads2 (qfile/script.sh:7) plus parameters/identity.inp, hp1/hp2.inp and
 * firstside1/2.data (inputinfo.f:105-132, 209-251). */
typedef struct dmdb_topology {
  int32_t n_species;                    /* 1 or 2 */
  int32_t n_chains[DMDB_MAX_SPECIES];   /* nop_s / numbeads_s */
  int32_t chnln[DMDB_MAX_SPECIES];      /* residues per chain */
  int32_t numbeads[DMDB_MAX_SPECIES];   /* beads per chain = 3*chnln + #side chains (Gly has none) */
  const int32_t* identity[DMDB_MAX_SPECIES]; /* numbeads ids: Ca x L (2), N x L (1), C x L (4), side chains 10..28 */
  const int32_t* hp[DMDB_MAX_SPECIES];       /* numbeads flags (hp1.inp / hp2.inp) */
  const int32_t* firstside[DMDB_MAX_SPECIES];/* chnln flags, 1 = residue has a side-chain bead */
} dmdb_topology;

/* Run parameters: the two stdin numbers (main.F90:126-128), the hard-coded box length
 * (inputinfo.f:78), and the compile-time behaviour flags (-Dcanon, -Dno_hbs, -Dn_wrap). */
typedef struct dmdb_params {
  double boxl;          /* box length in Angstrom (reference default 158.540) */
  double tstar;         /* reduced temperature T*; setemp = 12 T* (main.F90:127) */
  int32_t canon;        /* 1 = Andersen ghost collisions on (-Dcanon), 0 = NVE */
  int32_t no_hbs;       /* 1 = -Dno_hbs (no backbone hydrogen bonding) */
  int32_t n_wrap;       /* ghost-cell layers; the shipped build uses 2 (62-cell half stencil) */
  int32_t n_replicas;   /* independent trajectories held by this handle (one warp each on the device) */
  int32_t device;       /* CUDA device ordinal */
  int32_t nbr_capacity; /* per-bead capacity of the up and of the down neighbour list (0 = default 64) */
  int32_t log_capacity; /* per-replica event-log capacity in events (0 = no log); logging stops when it is full */
  int32_t engine;       /* event-loop engine: 0 = automatic, 1 = a group of lanes per replica, four replicas per warp (throughput: thousands of
                           replicas), 2 = one CTA per replica with batched conservative commit (latency: a few
                           trajectories; state resident in shared memory), 3 = the same batched commit with every round spread
                           over the whole GPU (one large system, e.g. 10^6 beads).  Results are bit-identical. */
  uint64_t seed;        /* replica r draws from the counter RNG stream seed + r (replaces Intel drandm) */
} dmdb_params;

typedef struct dmdb_event { /* one committed calendar event, for the event-sequence parity check */
  double t;      /* global time t + tfalse at which it was processed (box units) */
  int32_t i;     /* owner bead (1-based), or N+1 / N+2 / N+3 */
  int32_t j;     /* partner bead (1-based); ghost: the bead that was thermalised; else 0 */
  int32_t type;  /* executed coltype (1-3, 14-16, 20-27; -2 for pseudo-events) */
  int32_t evcode;/* ev_code(i,j) at dispatch (main.F90:587) */
} dmdb_event;

typedef struct dmdb_stats { /* main.F90:1356-1363 tallies, summed over the replicas of the handle */
  int64_t events;          /* calendar events processed (the reference's coll, main.F90:639) */
  int64_t pair_events;     /* events whose owner is a bead */
  int64_t nevents[32];     /* by executed type (main.F90:926) */
  int64_t ghosts;          /* numghosts */
  int64_t updates;         /* nupdates - nforcedupdate */
  int64_t forced_updates;  /* nforcedupdate */
  int64_t pair_predictions;/* pair-time evaluations made inside partial_events (for the roofline) */
  int64_t nbr_visits;      /* neighbour-list entries visited (up + down) */
  double device_ms;        /* CUDA-event time of the event-loop kernel(s) of the last dmdb_run */
  int32_t kernel_launches; /* kernels launched by the last dmdb_run */
  int32_t reserved;
} dmdb_stats;

typedef struct dmdb_energy { /* energy.f:25-101 outputs */
  double ered, tred, sumvel, ehh_ii, ehh_ij;
  int32_t hb_alpha, hb_ii, hb_ij, reserved;
} dmdb_energy;

typedef struct dmdb_handle dmdb_handle;

/* Replaces program start-up main.F90:117-234 (sizes, T*, flags) for n_replicas trajectories. */
int dmdb_create(const dmdb_params* p, const dmdb_topology* topo, const dmdb_tables* tab, dmdb_handle** out);
void dmdb_destroy(dmdb_handle* h);
const char* dmdb_last_error(const dmdb_handle* h); /* h may be NULL: error of the last failed dmdb_create */

int dmdb_num_beads(const dmdb_handle* h);
int dmdb_num_cells(const dmdb_handle* h); /* num_cell per dimension, main.F90:390 */

/* Restart path: inputinfo.f:76-101 + main.F90:205-321.  sv is 6 x N column-major (box units, wrapped by the
 * library like main.F90:206-208), bptnr has N entries (1-based partner or 0; NULL = all 0).  The library
 * rebuilds identity +4, extra_repuls and the 40/50 overlay geometrically (main.F90:249-321), sets the
 * pseudo-event times (main.F90:408-423), then runs nbor() and events().  replica = -1 loads every replica. */
int dmdb_set_state(dmdb_handle* h, int replica, const double* sv, const int32_t* bptnr);
/* Same for every replica at once from distinct configurations: sv_all is n_replicas x (6 x N), bptnr_all
 * n_replicas x N or NULL.  One host->device copy per array (this is the end-to-end path bench.py times). */
int dmdb_set_state_all(dmdb_handle* h, const double* sv_all, const int32_t* bptnr_all);
int dmdb_get_state_all(dmdb_handle* h, double* sv_all, int32_t* bptnr_all);
/* Temperature change on resident state (what a new `./dmd < temp_0xx` run does through a restart). */
int dmdb_set_temperature(dmdb_handle* h, int replica, double tstar);

/* = nbor()  (nbor.f:33-137, cell_add.f:12-28): cell binning + up/down neighbour lists, all replicas. */
int dmdb_nbor(dmdb_handle* h);
/* = events() (events.f:23-123): (tim, nptnr, coltype) of every bead from its up-list + aux slots. */
int dmdb_predict_all(dmdb_handle* h);
/* = the main loop main.F90:484-1258 with serial semantics: every replica processes n_events calendar events. */
int dmdb_run(dmdb_handle* h, int64_t n_events, dmdb_stats* stats);
/* Same, but every replica returns right after it has processed its next output pseudo-event (main.F90:1191-1246;
 * every 3.3/sqrt(setemp)+5 time units) or after max_events, whichever comes first: the host then writes the .energy
 * line (dmdb_energy_of) and the .config / .bptnr / .lastvel records (dmdb_get_state) exactly where the reference does. */
int dmdb_run_until_output(dmdb_handle* h, int64_t max_events, dmdb_stats* stats);
/* = main.F90:1288-1295: advance false positions to real positions and wrap (end of run). */
int dmdb_sync_positions(dmdb_handle* h);

/* Parity read-back. */
int dmdb_get_cells(dmdb_handle* h, int replica, int32_t* cell_of_bead /* N, 1-based cell id of cell_add.f:25 */);
/* offsets has N+1 entries; nb may be NULL to query sizes.  Lists are sorted ascending per bead (the
 * reference's order within a list is cell-scan order; only the SET is the parity surface). */
int dmdb_get_nbors(dmdb_handle* h, int replica, int down, int32_t* offsets, int32_t* nb);
int dmdb_get_calendar(dmdb_handle* h, int replica, double* tim, int32_t* nptnr, int32_t* coltype /* N+3 each */);
int dmdb_get_state(dmdb_handle* h, int replica, double* sv, int32_t* bptnr, int32_t* identity,
                   int32_t* extra_repuls /* N x 4 column-major like header.f:20 */, double* t, double* tfalse,
                   int64_t* coll);
int dmdb_get_evcode(dmdb_handle* h, int replica, int n_pairs, const int32_t* i, const int32_t* j,
                    int32_t* code /* ev_code(i,j) incl. the 40/50 overlay */);
int dmdb_energy_of(dmdb_handle* h, int replica, dmdb_energy* e); /* = energy() */
int dmdb_get_event_log(dmdb_handle* h, int replica, int64_t first, int64_t n, dmdb_event* out, int64_t* n_out);
int dmdb_get_replica_stats(dmdb_handle* h, int replica, dmdb_stats* s);
/* Batching statistics of the CTA-per-replica engine (summed over replicas when replica = -1): out[0] rounds,
 * [1] events executed speculatively, [2] of those rolled back (main.F90:970-993 rule), [3] candidates dropped by a
 * footprint conflict, [4] events processed serially at the head (H-bond events, pseudo-events), [5..7] reserved,
 * [8..15] SM clock cycles spent in the phases scan / select+rank / claim / check / exec / commit / serial head
 * events / of which list rebuilds. */
int dmdb_get_batch_stats(dmdb_handle* h, int replica, int64_t out[16]);

/* Replica exchange (new functionality: the reference runs its temp_0xx files one after the other by hand,
 * qfile/script.sh:11-18; SURVEY.md 8b "dmdb_exchange(h, ncclComm_t)" and 8e).
 *
 * dmdb_exchange is one exchange step, entirely on the device and on the library's stream: energy.f reduction of the
 * local replicas -> (E_pot, T*) -> ncclAllGather over the communicator -> the same Metropolis decision on every rank
 * (dmd_exchange.h: ladders of `ladder_size` replicas cut from the slot order r * world + rank, temperature
 * neighbours swap TEMPERATURES with probability min(1, exp((beta_a - beta_b)(E_a - E_b))), counter RNG shared by all
 * ranks) -> velocities of the replicas whose temperature changed rescaled by sqrt(T_new/T_old), time constants reset
 * (main.F90:143-156), lists + calendar rebuilt.  Nothing but the 16 bytes per replica of the gather crosses NVLink,
 * and nothing comes back to the host except the counters below.
 *   nccl_comm   an ncclComm_t the host created (a Fortran / C++ host linked against NCCL), or NULL: the communicator
 *               of dmdb_comm_init, or no collective at all when the handle is the only rank.
 *   ladder_size 2..32; 0 = one ladder over everything (at most 32 replicas).  Replicas beyond the last whole ladder
 *               keep their temperature.
 * dmdb_nccl_unique_id / dmdb_comm_init: create the communicator inside the library (rank 0 makes the id, the host
 * broadcasts the 128 bytes by whatever means it has, every rank calls dmdb_comm_init).  NCCL is bound at run time
 * (libnccl.so.2; the instance the host process has already loaded, if any).
 * dmdb_exchange_gathered: the same decision + temperature change for a host that gathers (E_pot, T*) itself
 * (MPI in a Fortran host, gloo in the CPU tests): `gathered` holds world x n_replicas x 2 doubles, rank-major. */
typedef struct dmdb_exchange_stats {
  int32_t ladders;        /* ladders in the gathered replica set */
  int32_t attempted;      /* neighbour pairs that attempted a swap (all ranks) */
  int32_t accepted;       /* ... and swapped */
  int32_t changed_local;  /* local replicas whose temperature changed */
  double device_ms;       /* CUDA-event time of the whole step on the library's stream */
  int32_t kernel_launches;
  int32_t reserved;
} dmdb_exchange_stats;
int dmdb_nccl_unique_id(char id[128]);
int dmdb_comm_init(dmdb_handle* h, const char id[128], int world, int rank);
int dmdb_exchange(dmdb_handle* h, void* nccl_comm, int64_t step, uint64_t seed, int32_t ladder_size, dmdb_exchange_stats* out);
int dmdb_exchange_gathered(dmdb_handle* h, const double* gathered, int world, int rank, int64_t step, uint64_t seed,
                           int32_t ladder_size, dmdb_exchange_stats* out);
/* The pieces, for a host that wants to decide itself: potential energies of the local replicas ... */
int dmdb_potential_energies(dmdb_handle* h, double* epot /* n_replicas */, double* tstar /* n_replicas */);
/* ... and the temperature change: replicas whose entry differs from their current T* have their velocities
 * rescaled by sqrt(T_new/T_old), their time constants reset (main.F90:143-156) and their calendar re-derived, all
 * on the device; neighbour lists (still valid: positions do not change) and H-bond state are kept.  Entries <= 0
 * leave the replica untouched.  (dmdb_set_temperature is the other temperature change: a restart as the reference
 * would do it between two runs.) */
int dmdb_apply_temperatures(dmdb_handle* h, const double* tstar_new /* n_replicas */);


/* beta-sheet observables of every replica, computed on the device from resident state (no trajectory download): the
 * definitions of the reference's post-processing program results/r/fibril_list_assign.f:51-60,89,228 (one peptide
 * species).  out: n_replicas x 8 int32 -- [0] inter-chain backbone H-bonds, [1] sheet-partner pairs (hb_contact >=
 * chnln/2 + 1), [2] sheets (connected components of >= 2 peptides), [3] peptides in the largest sheet, [4] peptides in
 * sheets, [5] intra-chain H-bonds, [6..7] reserved. */
int dmdb_sheet_observables(dmdb_handle* h, int32_t* out);

/* Engine 1 tuning (no reference counterpart).  The event-loop kernel gives a few of its CTAs the role of a LIST-REBUILD
 * SERVICE: they run nbor() + events() (nbor.f:33-137, events.f:23-107) for the warps of all other CTAs, so that the
 * SMs running the event loop keep only the hot loop in their 32 KB instruction caches.
 * dmdb_device_fill: the replica count that fills `device` in one wave with the default split, and that split.
 * dmdb_set_service_ctas: n = -1 automatic (default: service CTAs whenever they and every event-loop CTA can be
 * resident at once), 0 = every warp rebuilds its own lists, n > 0 = that many service CTAs.  Results do not depend
 * on the setting. */
int dmdb_device_fill(int device, int32_t* n_replicas, int32_t* n_service_ctas);
int dmdb_set_service_ctas(dmdb_handle* h, int n);

#ifdef __cplusplus
}
#endif
#endif /* DMDB200_H */
