#!/usr/bin/env python
"""Throughput of the event loop from the pre-aggregated fixture (tests/golden/aggregated_L80.npz): usage aggr_run.py R warm events [service CTAs: R is then per event-loop CTA]"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from parallel_dmd_for_biomolecules_b200 import genconfig, tables  # noqa: E402
from parallel_dmd_for_biomolecules_b200.dmd import DMD  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 7168
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
nev = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
fx = np.load(os.path.join(ROOT, "tests", "golden", "aggregated_L80.npz"))
tab = tables.load_default_tables()
boxl = float(fx["boxl"])
topo, _ = genconfig.system_b(tab, 0.18, seed=1, boxl=boxl)
svc = int(sys.argv[4]) if len(sys.argv) > 4 else -1
if svc >= 0:
    import torch
    R = R * (torch.cuda.get_device_properties(0).multi_processor_count - svc)
d = DMD(tables.make_params(boxl=boxl, tstar=0.18, canon=True, n_replicas=R, seed=77001), topo, tab, lib_path=os.environ.get("DMDB_LIB"))
if svc >= 0:
    d.set_service_ctas(svc)
d.set_state(np.ascontiguousarray(fx["sv"]), np.ascontiguousarray(fx["bptnr"]))
d.run(warm)
st = d.run(nev)
print("aggregated box: service=%d R=%d events/replica=%d device_ms=%.2f events/s=%.3e" % (svc, R, nev, st.device_ms, R * nev / (st.device_ms * 1e-3)))
