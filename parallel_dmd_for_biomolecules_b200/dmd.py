"""ctypes binding of libdmdb200.so -- the host-side mirror of the reference's operator interface.

The reference's "API" for the hot path is the set of Fortran subroutines main.F90 calls on module ``global``
(SURVEY.md 8b): ``nbor()``, ``events()``, the main loop, ``energy(...)``.  :class:`DMD` exposes the same
operations with the same names and argument meaning (1-based bead indices, ``sv`` as 6 x N column-major,
``coltype`` / ``nptnr`` conventions of header.f) on top of the C ABI in ``include/dmdb200.h``.

There is no CPU fallback: constructing :class:`DMD` raises if ``libdmdb200.so`` is missing or no CUDA device
is present.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from .tables import EVENT_DTYPE, Energy, Event, Params, Stats, Tables, Topology

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdmdb200.so")

#: every symbol include/dmdb200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "dmdb_create", "dmdb_destroy", "dmdb_last_error", "dmdb_num_beads", "dmdb_num_cells", "dmdb_set_state",
    "dmdb_set_temperature", "dmdb_nbor", "dmdb_predict_all", "dmdb_run", "dmdb_sync_positions", "dmdb_get_cells",
    "dmdb_get_nbors", "dmdb_get_calendar", "dmdb_get_state", "dmdb_get_evcode", "dmdb_energy_of",
    "dmdb_get_event_log", "dmdb_get_replica_stats", "dmdb_potential_energies", "dmdb_set_state_all",
    "dmdb_get_state_all", "dmdb_apply_temperatures", "dmdb_get_batch_stats", "dmdb_run_until_output",
    "dmdb_device_fill", "dmdb_set_service_ctas", "dmdb_nccl_unique_id", "dmdb_comm_init", "dmdb_exchange",
    "dmdb_exchange_gathered", "dmdb_sheet_observables",
]


class ExchangeStats(C.Structure):
    """dmdb_exchange_stats"""
    _fields_ = [("ladders", C.c_int32), ("attempted", C.c_int32), ("accepted", C.c_int32), ("changed_local", C.c_int32),
                ("device_ms", C.c_double), ("kernel_launches", C.c_int32), ("reserved", C.c_int32)]


def device_fill(device: int = 0, lib_path: Optional[str] = None):
    """(replicas, service CTAs): the replica count that fills `device` in one wave of the warp-per-replica engine
    with the default list-rebuild service split (dmdb_device_fill)"""
    lib = load_library(lib_path)
    nr, ns = C.c_int32(0), C.c_int32(0)
    rc = lib.dmdb_device_fill(int(device), C.byref(nr), C.byref(ns))
    if rc != 0:
        raise DMDError(rc, (lib.dmdb_last_error(None) or b"").decode())
    return nr.value, ns.value


class DMDError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dmdb200 error {code}: {msg}")
        self.code = code


def load_library(path: Optional[str] = None):
    path = path or os.environ.get("DMDB_LIB") or LIB_PATH
    if not os.path.exists(path):
        raise DMDError(-1, f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(path)
    lib.dmdb_last_error.restype = C.c_char_p
    lib.dmdb_last_error.argtypes = [C.c_void_p]
    lib.dmdb_destroy.restype = None
    lib.dmdb_destroy.argtypes = [C.c_void_p]
    return lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class DMD:
    """One handle = ``n_replicas`` independent PRIME20 trajectories resident on one GPU."""

    def __init__(self, params: Params, topo: Topology, tables: Tables, lib_path: Optional[str] = None):
        self._l = load_library(lib_path)
        self._topo = topo
        self._tc = topo.to_c()
        self._h = C.c_void_p()
        rc = self._l.dmdb_create(C.byref(params), C.byref(self._tc), C.byref(tables), C.byref(self._h))
        if rc != 0:
            raise DMDError(rc, self._l.dmdb_last_error(None).decode())
        self.N = self._l.dmdb_num_beads(self._h)
        self.n_replicas = params.n_replicas

    # -- plumbing ---------------------------------------------------------------------------------------
    def _chk(self, rc):
        if rc != 0:
            raise DMDError(rc, self._l.dmdb_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._l.dmdb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- run start (inputinfo.f:76-101 + main.F90:205-424) ----------------------------------------------
    def set_state(self, sv: np.ndarray, bptnr: Optional[np.ndarray] = None, replica: int = -1):
        sv = np.ascontiguousarray(sv, dtype=np.float64)
        if sv.shape != (self.N, 6):
            raise ValueError(f"sv must have shape ({self.N}, 6) (Fortran sv(6,N))")
        bp = None if bptnr is None else np.ascontiguousarray(bptnr, dtype=np.int32)
        self._chk(self._l.dmdb_set_state(self._h, replica, _p(sv, C.c_double), None if bp is None else _p(bp, C.c_int32)))

    def set_state_all(self, sv_all: np.ndarray, bptnr_all: Optional[np.ndarray] = None):
        """distinct configuration per replica: sv_all (n_replicas, N, 6); one H2D copy per device array"""
        if sv_all.dtype != np.float64 or not sv_all.flags.c_contiguous or sv_all.shape != (self.n_replicas, self.N, 6):
            raise ValueError(f"sv_all must be C-contiguous float64 of shape ({self.n_replicas}, {self.N}, 6)")
        bp = None if bptnr_all is None else np.ascontiguousarray(bptnr_all, dtype=np.int32)
        self._chk(self._l.dmdb_set_state_all(self._h, _p(sv_all, C.c_double), None if bp is None else _p(bp, C.c_int32)))

    def get_state_all(self, out: Optional[np.ndarray] = None, out_bptnr: Optional[np.ndarray] = None) -> np.ndarray:
        """every replica's sv (and optionally bptnr) into caller-owned host buffers (may be pinned)"""
        if out is None:
            out = np.empty((self.n_replicas, self.N, 6))
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape == (self.n_replicas, self.N, 6)
        if out_bptnr is not None:
            assert out_bptnr.dtype == np.int32 and out_bptnr.flags.c_contiguous and out_bptnr.shape == (self.n_replicas, self.N)
        self._chk(self._l.dmdb_get_state_all(self._h, _p(out, C.c_double),
                                             None if out_bptnr is None else _p(out_bptnr, C.c_int32)))
        return out

    def apply_temperatures(self, tstar_new):
        """replica-exchange outcome: per-replica new T* (entries equal to the current T* are left alone)"""
        t = np.ascontiguousarray(tstar_new, dtype=np.float64)
        assert t.shape == (self.n_replicas,)
        self._chk(self._l.dmdb_apply_temperatures(self._h, _p(t, C.c_double)))

    # -- replica exchange (dmdb_exchange: energies, NCCL all-gather, decision and temperature change on the device) ----
    def comm_init(self, world: int, rank: int, broadcast=None):
        """create the NCCL communicator inside the library.  `broadcast(bytes_or_None) -> bytes` must return rank 0's
        128-byte id on every rank (e.g. torch.distributed.broadcast_object_list); world == 1 needs none."""
        buf = (C.c_char * 128)()
        if world > 1:
            if rank == 0:
                rc = self._l.dmdb_nccl_unique_id(buf)
                if rc != 0:
                    raise DMDError(rc, (self._l.dmdb_last_error(None) or b"").decode())
            data = broadcast(bytes(buf.raw) if rank == 0 else None)
            buf = (C.c_char * 128).from_buffer_copy(data)
        self._chk(self._l.dmdb_comm_init(self._h, buf, int(world), int(rank)))

    def exchange(self, step: int, seed: int = 12345, ladder_size: int = 0, nccl_comm: Optional[int] = None) -> ExchangeStats:
        """one exchange step over the communicator of comm_init (or `nccl_comm`, an ncclComm_t as integer)"""
        st = ExchangeStats()
        self._chk(self._l.dmdb_exchange(self._h, C.c_void_p(nccl_comm), C.c_int64(step), C.c_uint64(seed),
                                        C.c_int32(ladder_size), C.byref(st)))
        return st

    def exchange_gathered(self, gathered: np.ndarray, world: int, rank: int, step: int, seed: int = 12345,
                          ladder_size: int = 0) -> ExchangeStats:
        """the same decision + temperature change from (E_pot, T*) the host gathered itself: (world * R, 2) float64"""
        g = np.ascontiguousarray(gathered, dtype=np.float64)
        assert g.shape == (world * self.n_replicas, 2)
        st = ExchangeStats()
        self._chk(self._l.dmdb_exchange_gathered(self._h, _p(g, C.c_double), int(world), int(rank), C.c_int64(step),
                                                 C.c_uint64(seed), C.c_int32(ladder_size), C.byref(st)))
        return st

    def set_temperature(self, tstar: float, replica: int = -1):
        self._chk(self._l.dmdb_set_temperature(self._h, replica, C.c_double(tstar)))

    # -- the reference's operators ---------------------------------------------------------------------
    def nbor(self):
        """nbor() -- nbor.f:33-137"""
        self._chk(self._l.dmdb_nbor(self._h))

    def events(self):
        """events() -- events.f:23-123"""
        self._chk(self._l.dmdb_predict_all(self._h))

    predict_all = events

    def run(self, n_events: int) -> Stats:
        """the main loop main.F90:484-1258: every replica processes ``n_events`` calendar events"""
        s = Stats()
        self._chk(self._l.dmdb_run(self._h, C.c_int64(n_events), C.byref(s)))
        return s

    def run_until_output(self, max_events: int) -> Stats:
        """like run(), but returns right after the next output pseudo-event (main.F90:1191-1246)"""
        s = Stats()
        self._chk(self._l.dmdb_run_until_output(self._h, C.c_int64(max_events), C.byref(s)))
        return s

    def sync_positions(self):
        self._chk(self._l.dmdb_sync_positions(self._h))

    def energy(self, replica: int = 0) -> Energy:
        """energy(ered,tred,sumvel,hb_alpha,hb_ii,hb_ij,ehh_ii,ehh_ij) -- energy.f:25-101"""
        e = Energy()
        self._chk(self._l.dmdb_energy_of(self._h, replica, C.byref(e)))
        return e

    def sheet_observables(self) -> np.ndarray:
        """(n_replicas, 8) int32 from the device: inter-chain H-bonds, sheet-partner pairs, sheets, largest sheet, peptides
        in sheets, intra-chain H-bonds (fibril_list_assign.f definitions; observables.py is the numpy restatement)"""
        out = np.zeros((self.n_replicas, 8), dtype=np.int32)
        self._chk(self._l.dmdb_sheet_observables(self._h, _p(out, C.c_int32)))
        return out

    def potential_energies(self):
        ep = np.zeros(self.n_replicas)
        ts = np.zeros(self.n_replicas)
        self._chk(self._l.dmdb_potential_energies(self._h, _p(ep, C.c_double), _p(ts, C.c_double)))
        return ep, ts

    # -- parity read-back ------------------------------------------------------------------------------
    @property
    def num_cell(self):
        return self._l.dmdb_num_cells(self._h)

    def cells(self, replica=0):
        out = np.zeros(self.N, dtype=np.int32)
        self._chk(self._l.dmdb_get_cells(self._h, replica, _p(out, C.c_int32)))
        return out

    def nbors(self, replica=0, down=False):
        off = np.zeros(self.N + 1, dtype=np.int32)
        self._chk(self._l.dmdb_get_nbors(self._h, replica, int(down), _p(off, C.c_int32), None))
        nb = np.zeros(max(int(off[-1]), 1), dtype=np.int32)
        self._chk(self._l.dmdb_get_nbors(self._h, replica, int(down), _p(off, C.c_int32), _p(nb, C.c_int32)))
        return off, nb[: off[-1]]

    def calendar(self, replica=0):
        tim = np.zeros(self.N + 3)
        nptnr = np.zeros(self.N + 3, dtype=np.int32)
        coltype = np.zeros(self.N + 3, dtype=np.int32)
        self._chk(self._l.dmdb_get_calendar(self._h, replica, _p(tim, C.c_double), _p(nptnr, C.c_int32),
                                            _p(coltype, C.c_int32)))
        return tim, nptnr, coltype

    def state(self, replica=0):
        sv = np.zeros((self.N, 6))
        bptnr = np.zeros(self.N, dtype=np.int32)
        ident = np.zeros(self.N, dtype=np.int32)
        er = np.zeros((4, self.N), dtype=np.int32)
        t, tf, coll = C.c_double(), C.c_double(), C.c_int64()
        self._chk(self._l.dmdb_get_state(self._h, replica, _p(sv, C.c_double), _p(bptnr, C.c_int32),
                                         _p(ident, C.c_int32), _p(er, C.c_int32), C.byref(t), C.byref(tf),
                                         C.byref(coll)))
        return dict(sv=sv, bptnr=bptnr, identity=ident, extra_repuls=er, t=t.value, tfalse=tf.value, coll=coll.value)

    def evcode(self, i, j, replica=0):
        i = np.ascontiguousarray(i, dtype=np.int32)
        j = np.ascontiguousarray(j, dtype=np.int32)
        out = np.zeros(len(i), dtype=np.int32)
        self._chk(self._l.dmdb_get_evcode(self._h, replica, len(i), _p(i, C.c_int32), _p(j, C.c_int32),
                                          _p(out, C.c_int32)))
        return out

    def event_log(self, replica=0, first=0, n=1 << 20):
        out = np.zeros(n, dtype=EVENT_DTYPE)
        n_out = C.c_int64()
        self._chk(self._l.dmdb_get_event_log(self._h, replica, C.c_int64(first), C.c_int64(n),
                                             out.ctypes.data_as(C.POINTER(Event)), C.byref(n_out)))
        return out[: n_out.value]

    def set_service_ctas(self, n: int = -1):
        """engine 1: CTAs of the event-loop kernel that only rebuild neighbour lists + calendars for the others
        (-1 automatic, 0 none); results do not depend on it"""
        self._chk(self._l.dmdb_set_service_ctas(self._h, int(n)))

    def batch_stats(self, replica=-1) -> dict:
        """batching statistics of the CTA-per-replica engine (engine=2)"""
        out = (C.c_int64 * 16)()
        self._chk(self._l.dmdb_get_batch_stats(self._h, replica, out))
        return dict(rounds=out[0], executed=out[1], rolled_back=out[2], conflicts=out[3], serial=out[4],
                    cycles=dict(zip(("scan", "select_rank", "claim", "check", "exec", "commit", "serial", "rebuild"), list(out)[8:16])))

    def stats(self, replica=-1) -> Stats:
        s = Stats()
        self._chk(self._l.dmdb_get_replica_stats(self._h, replica, C.byref(s)))
        return s
