#!/usr/bin/env python
"""per-phase SM cycles per round of the CTA-per-replica engine on one trajectory of config 2"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from parallel_dmd_for_biomolecules_b200 import genconfig, tables
from parallel_dmd_for_biomolecules_b200.dmd import DMD
tab = tables.load_default_tables(); topo, sv = genconfig.system_b(tab, 0.18, seed=1)
d = DMD(tables.make_params(boxl=158.54, tstar=0.18, canon=True, n_replicas=1, engine=2), topo, tab); d.set_state(sv); d.run(20000)
b0 = d.batch_stats(); s0 = d.stats(); st = d.run(200000); b1 = d.batch_stats(); s1 = d.stats()
r = b1["rounds"] - b0["rounds"]
print("events/s %.3e  rounds %d  events/round %.2f  rebuilds %d" % (200000 / (st.device_ms * 1e-3), r, (b1["executed"] - b1["rolled_back"] - b0["executed"] + b0["rolled_back"]) / r,
      s1.updates + s1.forced_updates - s0.updates - s0.forced_updates))
print({k: round((b1["cycles"][k] - b0["cycles"][k]) / r) for k in b1["cycles"]})
