#!/usr/bin/env python
"""Small driver for ncu captures: system B (48 x KLVFFAE), R replicas, a warm-up launch then one profiled
launch of the event loop.  usage: prof_run.py R warm_events events"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from parallel_dmd_for_biomolecules_b200 import genconfig, tables  # noqa: E402
from parallel_dmd_for_biomolecules_b200.dmd import DMD  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 592
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
nev = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
tab = tables.load_default_tables()
topo, sv = genconfig.system_b(tab, 0.18, seed=1)
p = tables.make_params(boxl=158.54, tstar=0.18, canon=True, n_replicas=R, engine=int(os.environ.get('DMDB_ENGINE', '1')))
d = DMD(p, topo, tab, lib_path=os.environ.get('DMDB_LIB'))
d.set_state(sv)
if os.environ.get('DMDB_SVC'):
    d.set_service_ctas(int(os.environ['DMDB_SVC']))
d.run(warm)
st = d.run(nev)
print("R=%d events/replica=%d device_ms=%.2f events/s=%.3e" % (R, nev, st.device_ms, R * nev / (st.device_ms * 1e-3)))
