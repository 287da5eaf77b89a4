#!/usr/bin/env python
"""Pre-aggregated starting states for the aggregated-regime measurement (SURVEY.md 8d expects
n_bar = 20-40 once beta-sheets form -- the regime the reference spends its 2e10 events in).

Runs the reference's annealing schedule (qfile/script.sh:11-14: T* = 0.50 0.45 0.40 0.35 0.30 0.28 0.26 0.24 0.22, then
0.18) on RESIDENT device state (dmdb_set_temperature) for an ensemble of 48-peptide KLVFFAE boxes with the CTA-per-replica
engine, and saves the replica with the most inter-chain hydrogen bonds (positions, velocities, bptnr) plus the
ensemble's observables along the way.  Aggregation at the reference's 20 mM (L = 158.54 A) needs ~1e10 events per
trajectory -- hours per trajectory on any hardware -- so the box is 80 A (8 x the concentration), where sheets nucleate
within ~1e7-1e8 events.  usage (GPU box): make_aggregated_fixture.py [events_per_anneal_T] [events_at_018] [replicas]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from parallel_dmd_for_biomolecules_b200 import genconfig, tables  # noqa: E402
from parallel_dmd_for_biomolecules_b200.dmd import DMD  # noqa: E402

BOXL = 80.0
SCHEDULE = (0.50, 0.45, 0.40, 0.35, 0.30, 0.28, 0.26, 0.24, 0.22)


def main():
    n_anneal = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
    n_cold = int(float(sys.argv[2])) if len(sys.argv) > 2 else 30_000_000
    R = int(sys.argv[3]) if len(sys.argv) > 3 else 148
    out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "gpurun_out", "aggregated_L80.npz")
    tab = tables.load_default_tables()
    topo, sv = genconfig.system_b(tab, 0.5, seed=1, boxl=BOXL)
    d = DMD(tables.make_params(boxl=BOXL, tstar=0.5, canon=True, n_replicas=R, engine=2, seed=20261018), topo, tab)
    d.set_state(sv)
    hist = []
    t0 = time.time()

    def sample(tag, tstar):
        so = d.sheet_observables()
        ep = d.potential_energies()[0]
        hist.append(dict(stage=tag, tstar=tstar, wall_s=time.time() - t0, hb_inter_mean=float(so[:, 0].mean()), hb_inter_max=int(so[:, 0].max()),
                         largest_sheet_mean=float(so[:, 3].mean()), largest_sheet_max=int(so[:, 3].max()),
                         peptides_in_sheets_mean=float(so[:, 4].mean()), epot_mean=float(ep.mean())))
        print(json.dumps(hist[-1]), flush=True)

    for k, T in enumerate(SCHEDULE):
        if k:
            d.set_temperature(T)
        d.run(n_anneal)
        sample("anneal", T)
    d.set_temperature(0.18)
    chunk = max(n_cold // 6, 1)
    for k in range(6):
        d.run(chunk)
        sample("cold", 0.18)
    so = d.sheet_observables()
    best = int(np.argmax(so[:, 0] * 100 + so[:, 3]))
    d.sync_positions()
    st = d.state(best)
    up, dn = d.nbors(best, False), d.nbors(best, True)
    nbar = (len(up[1]) + len(dn[1])) / topo.n_beads
    np.savez_compressed(out, sv=st["sv"], bptnr=st["bptnr"], boxl=BOXL, sheet_observables=so[best], nbar=nbar,
                        history=json.dumps(hist), schedule=np.array(SCHEDULE), events_per_anneal_T=n_anneal, events_at_018=n_cold)
    print("saved replica %d: hb_inter %d, largest sheet %d, peptides in sheets %d, n_bar %.1f -> %s" % (
        best, so[best, 0], so[best, 3], so[best, 4], nbar, out))


if __name__ == "__main__":
    main()
