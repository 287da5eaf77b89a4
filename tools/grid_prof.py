#!/usr/bin/env python
"""config 5 (10^6 beads) on the whole-GPU engine: a short run for ncu launch lists"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from parallel_dmd_for_biomolecules_b200 import genconfig, tables
from parallel_dmd_for_biomolecules_b200.dmd import DMD
tab = tables.load_default_tables()
nch = int(os.environ.get("BIG_CHAINS", "35715"))
boxl = 158.54 * (nch / 48.0) ** (1.0 / 3.0)
topo, sv = genconfig.generate_box(["KLVFFAE"], [nch], boxl, 0.5, tab, seed=5)
d = DMD(tables.make_params(boxl=boxl, tstar=0.5, canon=True, n_replicas=1, engine=3, nbr_capacity=32), topo, tab)
d.set_state(sv)
n = int(os.environ.get("GRID_EVENTS", "400000"))
t0 = time.time(); st = d.run(n); wall = time.time() - t0
bs = d.batch_stats(0)
print("N=%d: %d events in %.1f ms (%d launches) -> %.3e ev/s; rounds %d events/round %.1f" % (topo.n_beads, n, wall * 1e3, st.kernel_launches, n / wall, bs["rounds"], (bs["executed"] - bs["rolled_back"]) / max(bs["rounds"], 1)))
r = max(bs["rounds"], 1)
print("cycles per round (grid engine: scan, rank, claim, check, exec, commit):", [round(v / r) for v in list(bs["cycles"].values())[:6]])
