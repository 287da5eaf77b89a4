#!/usr/bin/env python
"""Freezes the oracle's first 10^4 committed events (NVE, fixed snapshot) of systems A and B as fixtures
(the reference ships no golden vectors for event times / partners / sequences, so the oracle's outputs
are pinned here; a change of the oracle or of the engines that alters a single event shows up against these files).
usage: python tests/golden/make_event_fixtures.py   (needs oracle/_build/liboracle.so: __graft_entry__.build())"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
from oracle.binding import OracleDMD  # noqa: E402
from parallel_dmd_for_biomolecules_b200 import fileio, genconfig, tables  # noqa: E402

N_EVENTS = 10000


def systems(tab):
    seq = "GVAYVGSKTKEGVVHGVATVAE"
    topo_a = tables.Topology([tables.Species.from_sequence(seq, 4), tables.Species.from_sequence(seq, 4)])
    sv_a = fileio.sv_from_files(os.path.join(HERE, "systemA_run0000.config"), os.path.join(HERE, "systemA_run0000.lastvel"))
    yield "A", topo_a, sv_a, 110.0, 0.5
    topo_b, sv_b = genconfig.system_b(tab, 0.18, seed=1)
    yield "B", topo_b, sv_b, 158.54, 0.18


def main():
    tab = tables.load_default_tables()
    for name, topo, sv, boxl, tstar in systems(tab):
        p = tables.make_params(boxl=boxl, tstar=tstar, canon=False, log_capacity=N_EVENTS)
        o = OracleDMD(p, topo, tab)
        o.set_state(sv)
        tim, nptnr, coltype = o.calendar()
        o.run(N_EVENTS)
        log = o.event_log()
        e = o.energy()
        out = os.path.join(HERE, "events_system%s_nve.npz" % name)
        np.savez_compressed(out, sv0=np.asarray(sv, dtype=np.float64), boxl=boxl, tstar=tstar,
                            i=log["i"].astype(np.int32), j=log["j"].astype(np.int32), type=log["type"].astype(np.int8),
                            evcode=log["evcode"].astype(np.int8), t=log["t"].astype(np.float64),
                            cal_t0=tim, cal_ptnr0=nptnr.astype(np.int32), cal_type0=coltype.astype(np.int8),
                            ered=e.ered, hb=np.array([e.hb_alpha, e.hb_ii, e.hb_ij]))
        print(name, out, os.path.getsize(out), "bytes; last event time", log["t"][-1])


if __name__ == "__main__":
    main()
