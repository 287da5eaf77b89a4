#!/usr/bin/env python
"""GPU-side check of the large boxes of BASELINE.json: config 4 (192 chains x 16 residues, 12 288 beads) against
the oracle, and the bulk kernels (run start, nbor, events) on a ~10^6-bead box (config 5) with their device times."""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle.binding import OracleDMD  # noqa: E402
from parallel_dmd_for_biomolecules_b200 import genconfig, tables  # noqa: E402
from parallel_dmd_for_biomolecules_b200.dmd import DMD  # noqa: E402

tab = tables.load_default_tables()
topo, sv = genconfig.generate_box(["KLVFFAEKLVFFAEKL"], [192], 200.0, 0.3, tab, seed=3)
print("config 4: N =", topo.n_beads)
n = 20000
for engine in (() if os.environ.get("SKIP4") else (2, 1)):
    p = tables.make_params(boxl=200.0, tstar=0.3, canon=True, n_replicas=2, log_capacity=n, engine=engine)
    o = OracleDMD(p, topo, tab)
    o.set_state(sv)
    d = DMD(p, topo, tab)
    d.set_state(sv)
    ok_cells = np.array_equal(o.cells(), d.cells(0))
    ok_nb = all(np.array_equal(a, b) for a, b in zip(o.nbors(), d.nbors(0))) and all(np.array_equal(a, b) for a, b in zip(o.nbors(True), d.nbors(0, True)))
    ta, na, ca = o.calendar()
    tb, nb, cb = d.calendar(0)
    print(" engine", engine, "cells", ok_cells, "nbors", ok_nb, "calendar", np.array_equal(na, nb), np.array_equal(ca, cb), np.array_equal(ta, tb))
    t0 = time.time()
    o.run(n)
    t1 = time.time()
    st = d.run(n)
    la, lb = o.event_log(), d.event_log(0)
    same = all(np.array_equal(la[f], lb[f]) for f in ("i", "j", "type", "t"))
    print("  %d events: sequence identical %s; oracle %.3f s, device %.1f ms (%.3e ev/s per trajectory) %s" % (
        n, same, t1 - t0, st.device_ms, n / (st.device_ms * 1e-3), d.batch_stats(0) if engine == 2 else ""))
    d.close()

nch = int(os.environ.get("BIG_CHAINS", "35715"))
boxl = 158.54 * (nch / 48.0) ** (1.0 / 3.0)
t0 = time.time()
topo, sv = genconfig.generate_box(["KLVFFAE"], [nch], boxl, 0.5, tab, seed=5)
print("config 5: N = %d, L = %.1f A, generated in %.1f s" % (topo.n_beads, boxl, time.time() - t0))
p = tables.make_params(boxl=boxl, tstar=0.5, canon=True, n_replicas=1, engine=1, nbr_capacity=32)
d = DMD(p, topo, tab)
t0 = time.time()
d.set_state(sv)
print(" set_state (H2D + run start + nbor + events): %.3f s, num_cell %d" % (time.time() - t0, d.num_cell))
for name, fn in (("nbor", d.nbor), ("events", d.events)):
    for rep in range(3):
        t0 = time.time()
        fn()
        dt = time.time() - t0
    print(" %s: device %.3f ms (wall %.3f ms)" % (name, d.stats().device_ms, dt * 1e3))
off, nb = d.nbors(0)
offd, nbd = d.nbors(0, True)
tim, nptnr, coltype = d.calendar(0)
print(" up pairs %d (%.2f per bead, max %d), down pairs %d; beads with an event %d; earliest %.3e" % (
    len(nb), len(nb) / topo.n_beads, np.diff(off).max(), len(nbd), int((nptnr[:-3] > 0).sum()), tim[:-3].min()))
assert len(nb) == len(nbd)
# symmetric: j in up(i) <=> i in dn(j)
i_of = np.repeat(np.arange(1, topo.n_beads + 1), np.diff(off))
j_of = np.repeat(np.arange(1, topo.n_beads + 1), np.diff(offd))
a = np.stack([i_of, nb], 1)
b = np.stack([nbd, j_of], 1)
a = a[np.lexsort((a[:, 1], a[:, 0]))]
b = b[np.lexsort((b[:, 1], b[:, 0]))]
print(" up/down lists are transposes of each other:", np.array_equal(a, b))
e = d.energy(0)
print(" energy: T %.4f (expect %.4f) hb %d" % (e.tred, 6.0, e.hb_ij + e.hb_ii))
