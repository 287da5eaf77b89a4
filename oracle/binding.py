"""ctypes binding of the CPU ORACLE (oracle/_build/liboracle.so) -- test infrastructure, NOT product code.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from parallel_dmd_for_biomolecules_b200.tables import (EVENT_DTYPE, Energy, Event, Params, Stats, Tables, Topology,
                                                      TopologyC)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_LIB_FAST = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "_build", "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("dmd_oracle.cpp", "oracle_capi.cpp", "dmd_oracle.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.dmdo_last_error.restype = C.c_char_p
        _LIB.dmdo_log.restype = C.c_double
        _LIB.dmdo_log.argtypes = [C.c_double]
        _LIB.dmdo_rng.restype = C.c_double
        _LIB.dmdo_rng.argtypes = [C.c_void_p]
    return _LIB


def lib_fast():
    """the -O3 / AVX2 / FMA build of the same source: bench.py's CPU baseline only, never a checker"""
    global _LIB_FAST
    if _LIB_FAST is None:
        build()
        _LIB_FAST = C.CDLL(os.path.join(_HERE, "_build", "liboracle_fast.so"))
        _LIB_FAST.dmdo_last_error.restype = C.c_char_p
    return _LIB_FAST


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class OracleDMD:
    def __init__(self, params: Params, topo: Topology, tables: Tables, fast: bool = False):
        self._l = lib_fast() if fast else lib()
        self._topo = topo
        self._tc = topo.to_c()
        self._h = C.c_void_p()
        self._chk(self._l.dmdo_create(C.byref(params), C.byref(self._tc), C.byref(tables), C.byref(self._h)))
        self.N = self._l.dmdo_num_beads(self._h)

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: " + self._l.dmdo_last_error().decode())

    def close(self):
        if self._h:
            self._l.dmdo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, sv: np.ndarray, bptnr=None):
        sv = np.ascontiguousarray(sv, dtype=np.float64)
        assert sv.shape == (self.N, 6)
        bp = None if bptnr is None else np.ascontiguousarray(bptnr, dtype=np.int32)
        self._chk(self._l.dmdo_set_state(self._h, _p(sv, C.c_double), None if bp is None else _p(bp, C.c_int32)))

    def set_temperature(self, tstar):
        self._chk(self._l.dmdo_set_temperature(self._h, C.c_double(tstar)))

    def retemp(self, tstar):
        self._chk(self._l.dmdo_retemp(self._h, C.c_double(tstar)))

    def nbor(self):
        self._chk(self._l.dmdo_nbor(self._h))

    def predict_all(self):
        self._chk(self._l.dmdo_predict_all(self._h))

    def run(self, n_events: int) -> float:
        sec = C.c_double()
        self._chk(self._l.dmdo_run(self._h, C.c_int64(n_events), C.byref(sec)))
        return sec.value

    def sync_positions(self):
        self._chk(self._l.dmdo_sync_positions(self._h))

    @property
    def num_cell(self):
        return self._l.dmdo_num_cells(self._h)

    def cells(self):
        out = np.zeros(self.N, dtype=np.int32)
        self._l.dmdo_get_cells(self._h, _p(out, C.c_int32))
        return out

    def nbors(self, down=False):
        off = np.zeros(self.N + 1, dtype=np.int32)
        self._l.dmdo_get_nbors(self._h, int(down), _p(off, C.c_int32), None)
        nb = np.zeros(max(int(off[-1]), 1), dtype=np.int32)
        self._l.dmdo_get_nbors(self._h, int(down), _p(off, C.c_int32), _p(nb, C.c_int32))
        return off, nb[: off[-1]]

    def calendar(self):
        tim = np.zeros(self.N + 3)
        nptnr = np.zeros(self.N + 3, dtype=np.int32)
        coltype = np.zeros(self.N + 3, dtype=np.int32)
        self._l.dmdo_get_calendar(self._h, _p(tim, C.c_double), _p(nptnr, C.c_int32), _p(coltype, C.c_int32))
        return tim, nptnr, coltype

    def state(self):
        sv = np.zeros((self.N, 6))
        bptnr = np.zeros(self.N, dtype=np.int32)
        ident = np.zeros(self.N, dtype=np.int32)
        er = np.zeros((4, self.N), dtype=np.int32)
        t, tf, coll = C.c_double(), C.c_double(), C.c_int64()
        self._l.dmdo_get_state(self._h, _p(sv, C.c_double), _p(bptnr, C.c_int32), _p(ident, C.c_int32),
                               _p(er, C.c_int32), C.byref(t), C.byref(tf), C.byref(coll))
        return dict(sv=sv, bptnr=bptnr, identity=ident, extra_repuls=er, t=t.value, tfalse=tf.value, coll=coll.value)

    def evcode(self, i, j):
        i = np.ascontiguousarray(i, dtype=np.int32)
        j = np.ascontiguousarray(j, dtype=np.int32)
        out = np.zeros(len(i), dtype=np.int32)
        self._l.dmdo_get_evcode(self._h, len(i), _p(i, C.c_int32), _p(j, C.c_int32), _p(out, C.c_int32))
        return out

    def evcode_matrix(self):
        m = np.zeros((self.N, self.N), dtype=np.int8)
        self._l.dmdo_get_evcode_matrix(self._h, _p(m, C.c_int8))
        return m

    def energy(self) -> Energy:
        e = Energy()
        self._l.dmdo_energy(self._h, C.byref(e))
        return e

    def checkover(self):
        buf = C.create_string_buffer(4096)
        over = self._l.dmdo_checkover(self._h, buf, 4096)
        return bool(over), buf.value.decode()

    def check_nc_int(self) -> dict:
        """check_nc_int.f:21-360: the reference's own audit of the H-bond / auxiliary-shoulder state (m_ss != n_ss makes
        the Fortran exit)"""
        out = np.zeros(6, dtype=np.int32)
        self._l.dmdo_check_nc_int(self._h, _p(out, C.c_int32))
        return dict(zip(("boundbad", "unboundbad", "m_ss", "n_ss", "no_ss", "pairs15"), (int(x) for x in out)))

    def adopt_state(self, state: dict, nbors):
        """overwrite the live state with one read back from another engine (DMD.state() + DMD.nbors()), without the
        restart reconstruction, so that checkover / check_nc_int audit THAT state"""
        off, nb = nbors
        sv = np.ascontiguousarray(state["sv"], dtype=np.float64)
        bp = np.ascontiguousarray(state["bptnr"], dtype=np.int32)
        idn = np.ascontiguousarray(state["identity"], dtype=np.int32)
        er = np.ascontiguousarray(state["extra_repuls"], dtype=np.int32)
        off = np.ascontiguousarray(off, dtype=np.int32)
        nb = np.ascontiguousarray(nb if len(nb) else np.zeros(1), dtype=np.int32)
        rc = self._l.dmdo_adopt_state(self._h, _p(sv, C.c_double), C.c_double(state["tfalse"]), _p(bp, C.c_int32),
                                      _p(idn, C.c_int32), _p(er, C.c_int32), _p(off, C.c_int32), _p(nb, C.c_int32))
        if rc != 0:
            raise RuntimeError(self._l.dmdo_last_error().decode())

    def event_log(self, first=0, n=1 << 20):
        out = np.zeros(n, dtype=EVENT_DTYPE)
        n_out = C.c_int64()
        self._l.dmdo_get_event_log(self._h, C.c_int64(first), C.c_int64(n), out.ctypes.data_as(C.POINTER(Event)),
                                   C.byref(n_out))
        return out[: n_out.value]

    def stats(self) -> Stats:
        s = Stats()
        self._l.dmdo_get_stats(self._h, C.byref(s))
        return s

    def constants(self):
        c = np.zeros(16)
        rlsq = np.zeros(50)
        self._l.dmdo_get_constants(self._h, _p(c, C.c_double), _p(rlsq, C.c_double))
        names = ["sig_max_all", "hdelr", "width", "half", "setemp", "interval", "interval_max", "sortsize", "boxl_orig",
                 "t_output", "ev_param_1_15", "t_fact", "n_forced", "avegtime"]
        d = dict(zip(names, c))
        d["rlsq"] = rlsq
        return d

    def masses(self):
        bm = np.zeros(self.N)
        self._l.dmdo_get_masses(self._h, _p(bm, C.c_double))
        return bm
