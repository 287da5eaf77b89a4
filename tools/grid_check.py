#!/usr/bin/env python
"""GPU-side check of the whole-GPU engine (engine=3): parity against the oracle on config 2 and config 4, then
throughput on config 4 and on the ~10^6-bead box of config 5."""
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
cases = [("config2", genconfig.system_b(tab, 0.18, seed=1), 158.54, 0.18, 30000),
         ("config4", genconfig.generate_box(["KLVFFAEKLVFFAEKL"], [192], 200.0, 0.3, tab, seed=3), 200.0, 0.3, 30000)]
for name, (topo, sv), boxl, ts, n in cases:
    p = tables.make_params(boxl=boxl, tstar=ts, canon=True, n_replicas=1, log_capacity=n, engine=3)
    o = OracleDMD(p, topo, tab)
    o.set_state(sv)
    d = DMD(p, topo, tab)
    d.set_state(sv)
    o.run(n)
    t0 = time.time()
    st = d.run(n)
    wall = time.time() - t0
    la, lb = o.event_log(), d.event_log(0)
    m = min(len(la), len(lb))
    same = (la["i"][:m] == lb["i"][:m]) & (la["j"][:m] == lb["j"][:m]) & (la["type"][:m] == lb["type"][:m]) & (la["t"][:m] == lb["t"][:m])
    bs = d.batch_stats(0)
    print("%s N=%d: lens %d %d identical %s sv equal %s | wall %.1f ms, %d launches, %.3e ev/s, events/round %.1f, rolled back %d, conflicts %d" % (
        name, topo.n_beads, len(la), len(lb), bool(same.all()) and len(la) == len(lb), np.array_equal(o.state()["sv"], d.state(0)["sv"]),
        wall * 1e3, st.kernel_launches, n / wall, (bs["executed"] - bs["rolled_back"]) / max(bs["rounds"], 1), bs["rolled_back"], bs["conflicts"]))
    if not same.all():
        k = int(np.argmin(same))
        print(" first diff at", k)
        print(la[max(0, k - 3):k + 3])
        print(lb[max(0, k - 3):k + 3])
    d.close()
if "--mid" in sys.argv:
    # the largest box the oracle can hold (it keeps the reference's N x N ev_code matrix: 2.5 GB at 50 400 beads)
    nch = 1800
    boxl = 158.54 * (nch / 48.0) ** (1.0 / 3.0)
    topo, sv = genconfig.generate_box(["KLVFFAE"], [nch], boxl, 0.5, tab, seed=9)
    n = 100000
    p = tables.make_params(boxl=boxl, tstar=0.5, canon=True, n_replicas=1, log_capacity=n, engine=3, nbr_capacity=32)
    t0 = time.time()
    o = OracleDMD(p, topo, tab)
    o.set_state(sv)
    t1 = time.time()
    o.run(n)
    t2 = time.time()
    d = DMD(p, topo, tab)
    d.set_state(sv)
    t3 = time.time()
    d.run(n)
    t4 = time.time()
    la, lb = o.event_log(), d.event_log(0)
    same = all(np.array_equal(la[f], lb[f]) for f in ("i", "j", "type", "t"))
    bs = d.batch_stats(0)
    print("mid N=%d: sequence identical %s, sv equal %s | oracle start %.1f s, %d events %.2f s (%.3e ev/s); device %.1f ms (%.3e ev/s), events/round %.1f" % (
        topo.n_beads, same, np.array_equal(o.state()["sv"], d.state(0)["sv"]), t1 - t0, n, t2 - t1, n / (t2 - t1), (t4 - t3) * 1e3, n / (t4 - t3),
        (bs["executed"] - bs["rolled_back"]) / max(bs["rounds"], 1)))
    d.close()
if "--big" in sys.argv:
    nch = int(os.environ.get("BIG_CHAINS", "35715"))
    boxl = 158.54 * (nch / 48.0) ** (1.0 / 3.0)
    topo, sv = genconfig.generate_box(["KLVFFAE"], [nch], boxl, 0.5, tab, seed=5)
    for canon in (False, True):
        d = DMD(tables.make_params(boxl=boxl, tstar=0.5, canon=canon, n_replicas=1, engine=3, nbr_capacity=32), topo, tab)
        d.set_state(sv)
        e0 = d.energy(0)
        d.run(200000)
        b0 = d.batch_stats(0)
        n = 2000000
        t0 = time.time()
        st = d.run(n)
        wall = time.time() - t0
        b1 = d.batch_stats(0)
        e1 = d.energy(0)
        print("config5 N=%d canon=%s: %d events in %.1f ms (%d launches) -> %.3e events/s; events/round %.1f; E %.6f -> %.6f, T %.4f" % (
            topo.n_beads, canon, n, wall * 1e3, st.kernel_launches, n / wall,
            (b1["executed"] - b1["rolled_back"] - b0["executed"] + b0["rolled_back"]) / max(b1["rounds"] - b0["rounds"], 1), e0.ered, e1.ered, e1.tred))
        d.close()
