#!/usr/bin/env python
"""GPU-side diagnostic of the CTA-per-replica engine (engine=2): parity against the oracle with first-mismatch
dump, then single-trajectory / small-ensemble throughput beside the warp-per-replica engine (engine=1)."""
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
topo, sv = genconfig.system_b(tab, 0.18, seed=1)
if "--skip-parity" not in sys.argv:
    n = 50000
    p = tables.make_params(boxl=158.54, tstar=0.18, canon=True, n_replicas=2, log_capacity=n, engine=2)
    o = OracleDMD(p, topo, tab)
    o.set_state(sv)
    d = DMD(p, topo, tab)
    d.set_state(sv)
    o.run(n)
    st = d.run(n)
    la, lb = o.event_log(), d.event_log(0)
    m = min(len(la), len(lb))
    same = (la["i"][:m] == lb["i"][:m]) & (la["j"][:m] == lb["j"][:m]) & (la["type"][:m] == lb["type"][:m]) & (la["t"][:m] == lb["t"][:m])
    print("parity: lens", len(la), len(lb), "identical", bool(same.all()), "device_ms %.1f" % st.device_ms, d.batch_stats(0))
    if not same.all():
        k = int(np.argmin(same))
        print(" first diff at", k)
        print(la[max(0, k - 3):k + 3])
        print(lb[max(0, k - 3):k + 3])
    print(" sv equal", np.array_equal(o.state()["sv"], d.state(0)["sv"]))
    d.close()
nev = int(os.environ.get("BLK_EVENTS", "200000"))
for engine, Rs in ((2, (1, 8, 148, 296, 592)), (1, (1, 148))):
    for R in Rs:
        p = tables.make_params(boxl=158.54, tstar=0.18, canon=True, n_replicas=R, engine=engine)
        d = DMD(p, topo, tab)
        d.set_state(sv)
        d.run(20000)
        e = nev if engine == 2 else nev // 4
        st = d.run(e)
        bs = d.batch_stats()
        per_round = (bs["executed"] - bs["rolled_back"]) / max(bs["rounds"], 1)
        print("engine %d R=%4d: %7d ev/replica in %8.2f ms -> %.3e events/s total, %.3e per trajectory  (events/round %.2f, "
              "rolled back %.1f%%)" % (engine, R, e, st.device_ms, R * e / (st.device_ms * 1e-3), e / (st.device_ms * 1e-3),
                                       per_round, 100.0 * bs["rolled_back"] / max(bs["executed"], 1)))
        d.close()
