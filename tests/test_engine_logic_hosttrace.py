"""Event-loop LOGIC of the engine source (csrc/dmd_engine.h) against the oracle, on the CPU.

The library under test here is tests/host_trace/libdmdb_hosttrace.so: the SAME engine source compiled with a
1-lane warp (DMD_HOST_TRACE).  It is test scaffolding -- never shipped, never loaded by the product path -- and
exercises none of the 32-lane collectives; the real parity tests are the `-m gpu` ones in test_gpu_parity.py.
What it buys: the calendar, cascade, bookkeeping, rebuild and thermostat logic is diffed event by event in the
GPU-less build container."""
import numpy as np
import pytest

from conftest import compare_engines
from oracle.binding import OracleDMD
from parallel_dmd_for_biomolecules_b200 import genconfig, tables
from parallel_dmd_for_biomolecules_b200.dmd import DMD


def _pair(p, topo, tab, sv, lib, bptnr=None):
    ora = OracleDMD(p, topo, tab)
    ora.set_state(sv, bptnr)
    dev = DMD(p, topo, tab, lib_path=lib)
    dev.set_state(sv, bptnr)
    return ora, dev


@pytest.mark.parametrize("engine", [1, 2])  # 2: the CTA-per-replica engine, emulated with one host thread per virtual warp
@pytest.mark.parametrize("which,canon,n_events", [("A", False, 30000), ("A", True, 30000), ("B", False, 30000), ("B", True, 60000)])
def test_event_sequence_matches_oracle(tab, system_a, system_b, hosttrace_lib, which, canon, n_events, engine):
    topo, sv, boxl = system_a if which == "A" else system_b
    p = tables.make_params(boxl=boxl, tstar=0.5 if which == "A" else 0.18, canon=canon, n_replicas=2, log_capacity=n_events,
                           engine=engine)
    ora, dev = _pair(p, topo, tab, sv, hosttrace_lib)
    compare_engines(ora, dev, replica=0, n_events=n_events)
    ea, eb = ora.energy(), dev.energy(0)
    assert (ea.hb_ii, ea.hb_ij, ea.hb_alpha) == (eb.hb_ii, eb.hb_ij, eb.hb_alpha)
    np.testing.assert_allclose([eb.ered, eb.tred, eb.ehh_ii, eb.ehh_ij], [ea.ered, ea.tred, ea.ehh_ii, ea.ehh_ij], rtol=1e-12, atol=1e-12)
    sa, sb = ora.stats(), dev.stats(0)
    assert list(sa.nevents) == list(sb.nevents)
    assert (sa.ghosts, sa.updates, sa.forced_updates) == (sb.ghosts, sb.updates, sb.forced_updates)
    assert abs(sb.nbr_visits / sa.nbr_visits - 1.0) < 0.05  # work counters (roofline input): the oracle counts a cascaded bead per request


@pytest.mark.parametrize("engine,warps", [(1, 0), (2, 3), (2, 16)])
def test_hbond_rich_trajectory_and_restart(tab, hosttrace_lib, engine, warps, monkeypatch):
    """A dense, cold box forms hydrogen bonds quickly: exercises types 7/10/12 resolution, the 40/50 overlay and
    the restart reconstruction of main.F90:249-321 from (sv, bptnr)."""
    topo, sv = genconfig.generate_box(["AAAAAAAAAAAA"], [8], 45.0, 0.10, tab, seed=1)
    n = 600000 if engine == 1 else 250000
    if warps:
        monkeypatch.setenv("DMDB_TRACE_WARPS", str(warps))
    p = tables.make_params(boxl=45.0, tstar=0.10, canon=True, n_replicas=1, log_capacity=n, seed=11, engine=engine)
    ora, dev = _pair(p, topo, tab, sv, hosttrace_lib)
    compare_engines(ora, dev, n_events=n)
    st = ora.stats()
    assert st.nevents[20] > 0 and min(st.nevents[14], st.nevents[15], st.nevents[16], st.nevents[26]) > 0
    assert engine == 2 or st.nevents[24] > 0
    # restart both from the oracle's end state (true positions) with its bptnr
    ora.sync_positions()
    s = ora.state()
    assert (s["bptnr"] > 0).sum() >= 2, "expected at least one hydrogen bond to exercise the restart fix-up"
    p2 = tables.make_params(boxl=45.0, tstar=0.10, canon=True, n_replicas=1, log_capacity=20000, seed=12, engine=engine)
    ora2, dev2 = _pair(p2, topo, tab, s["sv"], hosttrace_lib, bptnr=s["bptnr"])
    assert np.array_equal(ora2.state()["identity"], dev2.state()["identity"])
    assert np.array_equal(ora2.state()["extra_repuls"], dev2.state()["extra_repuls"])
    compare_engines(ora2, dev2, n_events=20000)


def test_static_evcode_function_equals_literal_matrix(tab, hosttrace_lib):
    """csrc/dmd_topology.h (closed form) against the oracle's literal make_code.f matrix, incl. Gly and Pro."""
    for seqs, counts in ((["KLVFFAE", "GVAYVGSKTKEGVVHGVATVAE"], [3, 2]), (["APGLPVAEKG"], [4]), (["GGPKAG", "PAAPG"], [2, 3])):
        topo = tables.Topology([tables.Species.from_sequence(s, c) for s, c in zip(seqs, counts)])
        p = tables.make_params(boxl=200.0, tstar=0.3, canon=False)
        rng = np.random.default_rng(0)
        sv = np.concatenate([rng.random((topo.n_beads, 3)) - 0.5, rng.normal(size=(topo.n_beads, 3))], axis=1)
        ora = OracleDMD(p, topo, tab)
        m = ora.evcode_matrix()
        dev = DMD(p, topo, tab, lib_path=hosttrace_lib)
        dev.set_state(sv)  # positions are irrelevant for the static classes (no pair is inside a well by luck: checked)
        N = topo.n_beads
        ii, jj = np.meshgrid(np.arange(1, N + 1), np.arange(1, N + 1), indexing="ij")
        mask = ii != jj
        got = dev.evcode(ii[mask], jj[mask])
        exp = m[mask]
        overlay = got >= 40
        assert np.array_equal(got[~overlay], exp[~overlay])


def test_retemp_on_resident_state(tab, system_b, hosttrace_lib):
    topo, sv, boxl = system_b
    p = tables.make_params(boxl=boxl, tstar=0.18, canon=True, n_replicas=3, log_capacity=40000)
    ora, dev = _pair(p, topo, tab, sv, hosttrace_lib)
    ora.run(10000)
    dev.run(10000)
    ora.retemp(0.26)
    dev.apply_temperatures([0.26, 0.18, 0.30])
    compare_engines(ora, dev, replica=0, n_events=0, cells=False)  # a temperature change keeps the lists: no new cells
    ora.run(20000)
    dev.run(20000)
    la, lb = ora.event_log(), dev.event_log(0)
    assert np.array_equal(la["i"], lb["i"]) and np.array_equal(la["type"], lb["type"]) and np.array_equal(la["t"], lb["t"])
    assert abs(dev.energy(0).tred - ora.energy().tred) < 1e-12
    ep, ts = dev.potential_energies()
    assert list(ts) == [0.26, 0.18, 0.30]


def test_distinct_states_per_replica(tab, hosttrace_lib):
    topo, sv0 = genconfig.system_b(tab, 0.18, seed=5)
    _, sv1 = genconfig.system_b(tab, 0.18, seed=6)
    p = tables.make_params(boxl=158.54, tstar=0.18, canon=False, n_replicas=2, log_capacity=5000)
    dev = DMD(p, topo, tab, lib_path=hosttrace_lib)
    dev.set_state_all(np.ascontiguousarray(np.stack([sv0, sv1])))
    for r, sv in ((0, sv0), (1, sv1)):
        ora = OracleDMD(p, topo, tab)
        ora.set_state(sv)
        compare_engines(ora, dev, replica=r)
    dev.run(5000)
    out = dev.get_state_all()
    assert out.shape == (2, topo.n_beads, 6) and not np.array_equal(out[0], out[1])
    np.testing.assert_array_equal(out[1], dev.state(1)["sv"])


def test_capacity_error_is_reported(tab, system_b, hosttrace_lib):
    from parallel_dmd_for_biomolecules_b200.dmd import DMDError
    topo, sv, boxl = system_b
    dev = DMD(tables.make_params(boxl=boxl, tstar=0.18, nbr_capacity=4), topo, tab, lib_path=hosttrace_lib)
    with pytest.raises(DMDError) as e:
        dev.set_state(sv)
    assert e.value.code == 5 and "capacity" in str(e.value)


@pytest.mark.parametrize("engine", [1, 2])
def test_ragged_two_species_box_with_gly(tab, hosttrace_lib, engine):
    """Two species of different length with glycines (no side-chain bead) in one box, with and without the -Dno_hbs
    build flag: same committed-event sequence as the oracle.  (Proline topologies: test_static_evcode_*.)"""
    for no_hbs in (False, True):
        topo, sv = genconfig.generate_box(["GAKLGVFE", "GAAKGS"], [3, 4], 60.0, 0.35, tab, seed=4)
        n = 40000
        p = tables.make_params(boxl=60.0, tstar=0.35, canon=True, no_hbs=no_hbs, n_replicas=1, log_capacity=n, seed=21,
                               engine=engine)
        ora, dev = _pair(p, topo, tab, sv, hosttrace_lib)
        compare_engines(ora, dev, n_events=n)
        assert not ora.checkover()[0]


def test_box_too_small_and_bad_bptnr_are_errors(tab, hosttrace_lib):
    from parallel_dmd_for_biomolecules_b200.dmd import DMDError
    topo, sv = genconfig.generate_box(["AAAA"], [2], 40.0, 0.5, tab, seed=2)
    with pytest.raises(DMDError):  # fewer than 5 cells per dimension (main.F90:390)
        DMD(tables.make_params(boxl=7.0, tstar=0.5), topo, tab, lib_path=hosttrace_lib)
    d = DMD(tables.make_params(boxl=40.0, tstar=0.5), topo, tab, lib_path=hosttrace_lib)
    bad = np.zeros(topo.n_beads, dtype=np.int32)
    bad[3] = topo.n_beads + 5
    with pytest.raises(DMDError) as e:
        d.set_state(sv, bad)
    assert e.value.code == 1  # DMDB_ERR_ARG


@pytest.mark.parametrize("engine", [1, 2])
@pytest.mark.parametrize("which", ["A", "B"])
def test_frozen_event_sequence(tab, system_a, system_b, hosttrace_lib, which, engine):
    """the engine source (1-lane trace build) against the frozen oracle events of tests/golden/"""
    from conftest import check_against_frozen_events, frozen_events
    fx = frozen_events(which)
    topo = (system_a if which == "A" else system_b)[0]
    p = tables.make_params(boxl=float(fx["boxl"]), tstar=float(fx["tstar"]), canon=False, n_replicas=1, log_capacity=len(fx["t"]),
                           engine=engine)
    dev = DMD(p, topo, tab, lib_path=hosttrace_lib)
    dev.set_state(fx["sv0"])
    check_against_frozen_events(dev, fx, replica=0)
