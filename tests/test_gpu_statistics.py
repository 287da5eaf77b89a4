"""`-m gpu`: long-run STATISTICAL equivalence with the oracle on BASELINE config 2 (north_star: "longer runs must show
statistically equivalent energy conservation, temperature, potential-energy histograms and H-bond / beta-sheet
content").

64 device replicas and 16 oracle trajectories -- all with DIFFERENT random-number streams, so nothing here is
bit-comparable -- go through the reference's annealing schedule (qfile/script.sh:11-14: T* = 0.50 ... 0.22) and then
sample at T* = 0.18.  For every observable the two ensembles must pass a two-sample Kolmogorov-Smirnov test; the
device-side observables (dmdb_potential_energies, dmdb_sheet_observables: no trajectory leaves the GPU) are
additionally held against the host restatements on downloaded states."""
import threading

import numpy as np
import pytest
from scipy import stats

from oracle.binding import OracleDMD
from parallel_dmd_for_biomolecules_b200 import genconfig, observables, tables
from parallel_dmd_for_biomolecules_b200.dmd import DMD
from test_observables import check_device_sheet_observables

pytestmark = pytest.mark.gpu

SCHEDULE = (0.50, 0.45, 0.40, 0.35, 0.30, 0.28, 0.26, 0.24, 0.22)
N_ANNEAL, N_SAMPLES, N_BETWEEN = 50000, 8, 100000
P_MIN = 1e-3  # a test this strict fails one time in a thousand for identical distributions


def _oracle_series(topo, tab, sv, boxl, seed, out, k):
    o = OracleDMD(tables.make_params(boxl=boxl, tstar=SCHEDULE[0], canon=True, seed=seed), topo, tab)
    o.set_state(sv)
    for i, T in enumerate(SCHEDULE):
        if i:
            o.set_temperature(T)
        o.run(N_ANNEAL)
    o.set_temperature(0.18)
    rows = []
    for _ in range(N_SAMPLES):
        o.run(N_BETWEEN)
        e = o.energy()
        st = o.state()
        xyz = st["sv"][:, :3] + st["sv"][:, 3:] * st["tfalse"]
        sh = observables.sheets_and_fibrils(topo, tab, xyz, st["bptnr"], boxl)
        rows.append([e.ered - 0.5 * e.sumvel, e.tred / 12.0, e.hb_ii + e.hb_ij, e.ehh_ii + e.ehh_ij, sh["largest_sheet"],
                     int((sh["hb_contact"] > 0).sum() // 2)])
    out[k] = np.array(rows)
    o.close()


@pytest.mark.parametrize("boxl", [158.54, 80.0])  # config 2, and the same box at 8 x the concentration (more H-bonds)
def test_ensemble_statistics_match_the_oracle(tab, boxl):
    topo, sv = genconfig.system_b(tab, SCHEDULE[0], seed=1, boxl=boxl)
    R, K = 64, 16
    # ---- oracle ensemble on host threads (ctypes releases the GIL)
    res = {}
    th = [threading.Thread(target=_oracle_series, args=(topo, tab, sv, boxl, 500000 + 977 * k, res, k)) for k in range(K)]
    for t in th:
        t.start()
    # ---- device ensemble, same protocol on resident state
    d = DMD(tables.make_params(boxl=boxl, tstar=SCHEDULE[0], canon=True, n_replicas=R, seed=900000), topo, tab)
    d.set_state(sv)
    for i, T in enumerate(SCHEDULE):
        if i:
            d.set_temperature(T)
        d.run(N_ANNEAL)
    d.set_temperature(0.18)
    dev = np.zeros((N_SAMPLES, R, 6))
    for s in range(N_SAMPLES):
        d.run(N_BETWEEN)
        ep, _ = d.potential_energies()
        so = d.sheet_observables()
        dev[s, :, 0] = ep
        dev[s, :, 2] = so[:, 0] + so[:, 5]
        dev[s, :, 4] = so[:, 3]
        for r in range(R):
            e = d.energy(r)
            dev[s, r, 1] = e.tred / 12.0
            dev[s, r, 3] = e.ehh_ii + e.ehh_ij
            assert e.hb_ii + e.hb_ij == so[r, 0] + so[r, 5] and abs((e.ered - 0.5 * e.sumvel) - ep[r]) < 1e-9
    # the device-side sheet reduction against the numpy restatement on a downloaded state
    for r in (0, R // 2, R - 1):
        st = d.state(r)
        xyz = st["sv"][:, :3] + st["sv"][:, 3:] * st["tfalse"]
        sh = observables.sheets_and_fibrils(topo, tab, xyz, st["bptnr"], boxl)
        so = d.sheet_observables()[r]
        assert so[3] == sh["largest_sheet"] and so[4] == sh["peptides_in_sheets"] and so[2] == len(sh["sheets"])
    d.close()
    for t in th:
        t.join()
    ora = np.stack([res[k] for k in range(K)], axis=1)  # (samples, K, 6)
    names = ("E_pot", "T*", "H-bonds", "E_hydrophobic", "largest sheet")
    report = {}
    for c, name in enumerate(names):
        a, b = dev[:, :, c].mean(axis=0), ora[:, :, c].mean(axis=0)  # per-trajectory means: independent samples
        p = stats.ks_2samp(a, b).pvalue
        report[name] = (float(a.mean()), float(b.mean()), float(p))
        assert p > P_MIN, (name, report[name])
        sem = np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))  # ensemble means agree within 4 standard errors
        assert abs(a.mean() - b.mean()) <= 4.0 * sem + 1e-12, (name, report[name], sem)
        # the pooled histograms (every sample of every trajectory), north_star "potential-energy histograms"
        if name in ("E_pot", "T*"):
            assert stats.ks_2samp(dev[:, :, c].ravel(), ora[:, :, c].ravel()).pvalue > P_MIN / 10, name
    # the Andersen thermostat (one ghost collision per ~250 events) is still pulling the boxes down from the annealing
    # temperatures -- a restart keeps the velocities, like the reference's .lastvel -- and both ensembles are equally far
    assert 0.18 < report["T*"][0] < 0.50 and abs(report["T*"][0] / report["T*"][1] - 1) < 0.03
    print("ensemble means (device, oracle, KS p):", report)


def test_device_sheet_observables(tab):
    check_device_sheet_observables(tab, None)
