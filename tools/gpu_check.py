#!/usr/bin/env python
"""GPU-side diagnostic (run under gpurun): device engine vs CPU oracle on system A, with first-mismatch dumps,
then a quick throughput sweep over replica counts.  Not a test -- tests/ holds the asserts."""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle.binding import OracleDMD  # noqa: E402
from parallel_dmd_for_biomolecules_b200 import fileio, tables  # noqa: E402
from parallel_dmd_for_biomolecules_b200.dmd import DMD  # noqa: E402


def main():
    seq = "GVAYVGSKTKEGVVHGVATVAE"
    topo = tables.Topology([tables.Species.from_sequence(seq, 4), tables.Species.from_sequence(seq, 4)])
    tab = tables.load_default_tables()
    g = os.path.join(ROOT, "tests", "golden")
    sv = fileio.sv_from_files(os.path.join(g, "systemA_run0000.config"), os.path.join(g, "systemA_run0000.lastvel"))
    for canon in (False, True):
        nev = 100000
        p = tables.make_params(boxl=110.0, tstar=0.5, canon=canon, n_replicas=3, log_capacity=nev)
        o = OracleDMD(p, topo, tab)
        o.set_state(sv)
        d = DMD(p, topo, tab)
        d.set_state(sv)
        print("canon", canon, "cells", np.array_equal(o.cells(), d.cells(0)))
        for down in (False, True):
            a, b = o.nbors(down), d.nbors(0, down)
            print(" nbors down=%s equal=%s %d %d" % (down, np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0]), len(a[1]), len(b[1])))
        ta, na, ca = o.calendar()
        tb, nb, cb = d.calendar(0)
        print(" calendar nptnr", np.array_equal(na, nb), "coltype", np.array_equal(ca, cb), "tim bit-equal", np.array_equal(ta, tb),
              "max rel", np.max(np.abs(ta - tb) / np.abs(ta)))
        N = o.N
        ii, jj = np.meshgrid(np.arange(1, N + 1), np.arange(1, N + 1), indexing="ij")
        m = ii != jj
        print(" evcode mismatches", int((o.evcode(ii[m], jj[m]) != d.evcode(ii[m], jj[m])).sum()))
        t0 = time.time()
        o.run(nev)
        t1 = time.time()
        st = d.run(nev)
        t2 = time.time()
        print(" oracle %.3fs  gpu(3 replicas) %.3fs device_ms %.1f" % (t1 - t0, t2 - t1, st.device_ms))
        la, lb = o.event_log(), d.event_log(0)
        n = min(len(la), len(lb))
        same = (la["i"][:n] == lb["i"][:n]) & (la["j"][:n] == lb["j"][:n]) & (la["type"][:n] == lb["type"][:n]) & (la["evcode"][:n] == lb["evcode"][:n])
        print(" log lens", len(la), len(lb), "seq same", bool(same.all()), "t bit-equal", np.array_equal(la["t"][:n], lb["t"][:n]))
        if not same.all():
            k = int(np.argmin(same))
            print("  first diff at", k)
            print(la[max(0, k - 3):k + 3])
            print(lb[max(0, k - 3):k + 3])
        sa, sb = o.state(), d.state(0)
        print(" state sv", np.array_equal(sa["sv"], sb["sv"]), "bptnr", np.array_equal(sa["bptnr"], sb["bptnr"]), "ident",
              np.array_equal(sa["identity"], sb["identity"]), "er", np.array_equal(sa["extra_repuls"], sb["extra_repuls"]))
        ea, eb = o.energy(), d.energy(0)
        print(" energy", ea.ered, eb.ered, ea.ehh_ij, eb.ehh_ij, ea.hb_ij, eb.hb_ij)
        s1, s2 = o.stats(), d.stats(0)
        print(" events", [(k, s1.nevents[k], s2.nevents[k]) for k in range(32) if s1.nevents[k] or s2.nevents[k]], s1.updates, s2.updates)
        if canon:
            lc = d.event_log(1)
            print(" replica 1 differs from replica 0 (different RNG stream):", not np.array_equal(lc["i"][:5000], lb["i"][:5000]))
        d.close()
    # ---- throughput sweep
    for R in (148, 592, 1184, 4144, 8288):
        p = tables.make_params(boxl=110.0, tstar=0.5, canon=True, n_replicas=R, log_capacity=0)
        d = DMD(p, topo, tab)
        t0 = time.time()
        d.set_state(sv)
        t1 = time.time()
        d.run(2000)
        nev = 20000
        st = d.run(nev)
        s = d.stats()
        print("R=%d set_state %.2fs  run %d ev/replica: %.1f ms -> %.3e events/s  (pairpred/ev %.1f)" % (
            R, t1 - t0, nev, st.device_ms, R * nev / (st.device_ms * 1e-3), s.pair_predictions / max(s.events, 1)))
        d.close()


if __name__ == "__main__":
    main()
