#!/usr/bin/env python
"""Event-loop / list-rebuild-service split sweep on system B: for each service CTA count S the device is filled with
(SMs - S) event-loop CTAs of replicas; prints events/s.  usage: svc_sweep.py S1,S2,... [events] [lib] [replicas per CTA]"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from parallel_dmd_for_biomolecules_b200 import genconfig, tables  # noqa: E402
from parallel_dmd_for_biomolecules_b200.dmd import DMD, device_fill  # noqa: E402

svc = [int(x) for x in sys.argv[1].split(",")]
nev = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
lib = sys.argv[3] if len(sys.argv) > 3 else None
sms = torch.cuda.get_device_properties(0).multi_processor_count
fr, fs = device_fill(0)
per_cta = int(sys.argv[4]) if len(sys.argv) > 4 else fr // (sms - fs)
tab = tables.load_default_tables()
topo, sv = genconfig.system_b(tab, 0.18, seed=1)
for S in svc:
    R = (sms - S) * per_cta
    d = DMD(tables.make_params(boxl=158.54, tstar=0.18, canon=True, n_replicas=R), topo, tab, lib_path=lib)
    d.set_service_ctas(S)
    d.set_state(sv)
    d.run(nev)
    best = 0.0
    for _ in range(2):
        st = d.run(nev)
        best = max(best, R * nev / (st.device_ms * 1e-3))
    print("service=%d replicas=%d events/s=%.3e" % (S, R, best), flush=True)
    d.close() if hasattr(d, "close") else None
    del d
