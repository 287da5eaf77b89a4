"""`-m gpu` parity tests: libdmdb200.so (CUDA, sm_100a) through the C ABI against the CPU oracle on identical
snapshots.  Bars (BASELINE.json north_star): cell / neighbour assignment and per-bead next-event partner and
type bit-exact, event times within 1e-12 relative (they are in fact bit-equal), first 1e4+ committed events in
sequence; at full size (thousands of replicas) size-independent properties: NVE energy conservation, no
overlaps (checkover.f) and run-to-run determinism."""
import numpy as np
import pytest

from conftest import audit_nc, compare_engines
from oracle.binding import OracleDMD
from parallel_dmd_for_biomolecules_b200 import genconfig, tables
from parallel_dmd_for_biomolecules_b200.dmd import DMD, DMDError, device_fill

pytestmark = pytest.mark.gpu


def _pair(p, topo, tab, sv, bptnr=None):
    ora = OracleDMD(p, topo, tab)
    ora.set_state(sv, bptnr)
    dev = DMD(p, topo, tab)  # the product library; raises without a CUDA device
    dev.set_state(sv, bptnr)
    return ora, dev


@pytest.mark.parametrize("engine", [1, 2])  # 1 = warp per replica, 2 = CTA per replica with batched commit
@pytest.mark.parametrize("which,canon,n_events", [("A", False, 20000), ("A", True, 50000), ("B", False, 20000), ("B", True, 100000),
                                                  ("A018", True, 60000)])
def test_event_sequence_matches_oracle(tab, system_a, system_b, which, canon, n_events, engine):
    """A = BASELINE config 1, the shipped snapshot (at the T* = 0.5 it was generated for, and -- "A018" -- at the
    temp_018 setup the config names: T* = 0.18, canon); B = config 2."""
    topo, sv, boxl = system_b if which == "B" else system_a
    tstar = {"A": 0.5, "B": 0.18, "A018": 0.18}[which]
    p = tables.make_params(boxl=boxl, tstar=tstar, canon=canon, n_replicas=5, log_capacity=n_events, engine=engine)
    ora, dev = _pair(p, topo, tab, sv)
    compare_engines(ora, dev, replica=0, n_events=n_events)
    ea, eb = ora.energy(), dev.energy(0)
    assert (ea.hb_ii, ea.hb_ij, ea.hb_alpha) == (eb.hb_ii, eb.hb_ij, eb.hb_alpha)
    np.testing.assert_allclose([eb.ered, eb.tred, eb.ehh_ii, eb.ehh_ij], [ea.ered, ea.tred, ea.ehh_ii, ea.ehh_ij], rtol=1e-12, atol=1e-12)
    sa, sb = ora.stats(), dev.stats(0)
    assert list(sa.nevents) == list(sb.nevents)
    assert (sa.ghosts, sa.updates, sa.forced_updates) == (sb.ghosts, sb.updates, sb.forced_updates)
    assert abs(sb.nbr_visits / sa.nbr_visits - 1.0) < 0.05  # work counters (roofline input): the oracle counts a cascaded bead per request
    if canon:  # replica 3 draws from another RNG stream: compare it with its own oracle
        p3 = tables.make_params(boxl=boxl, tstar=tstar, canon=True, log_capacity=n_events, seed=p.seed + 3)
        o3 = OracleDMD(p3, topo, tab)
        o3.set_state(sv)
        o3.run(n_events)
        l3, d3 = o3.event_log(), dev.event_log(3)
        assert np.array_equal(l3["i"], d3["i"]) and np.array_equal(l3["j"], d3["j"]) and np.array_equal(l3["type"], d3["type"])
    else:  # NVE replicas started from the same snapshot are bit-identical
        assert np.array_equal(dev.state(0)["sv"], dev.state(4)["sv"])


@pytest.mark.parametrize("engine", [1, 2])
def test_hbond_rich_trajectory_and_restart(tab, engine):
    topo, sv = genconfig.generate_box(["AAAAAAAAAAAA"], [8], 45.0, 0.10, tab, seed=1)
    n = 600000
    p = tables.make_params(boxl=45.0, tstar=0.10, canon=True, n_replicas=2, log_capacity=n, seed=11, engine=engine)
    ora, dev = _pair(p, topo, tab, sv)
    compare_engines(ora, dev, n_events=n)
    st = ora.stats()
    assert min(st.nevents[14], st.nevents[15], st.nevents[16], st.nevents[20], st.nevents[24], st.nevents[26]) > 0
    a = audit_nc(ora, dev, make_oracle=lambda: OracleDMD(p, topo, tab))  # check_nc_int.f on oracle AND device state
    assert a["m_ss"] >= 2 and a["pairs15"] > a["m_ss"]
    ora.sync_positions()
    s = ora.state()
    assert (s["bptnr"] > 0).sum() >= 2
    p2 = tables.make_params(boxl=45.0, tstar=0.10, canon=True, n_replicas=1, log_capacity=20000, seed=12, engine=engine)
    ora2, dev2 = _pair(p2, topo, tab, s["sv"], bptnr=s["bptnr"])
    assert np.array_equal(ora2.state()["identity"], dev2.state()["identity"])
    assert np.array_equal(ora2.state()["extra_repuls"], dev2.state()["extra_repuls"])
    compare_engines(ora2, dev2, n_events=20000)


def test_static_evcode_function_equals_literal_matrix(tab):
    for seqs, counts in ((["KLVFFAE", "GVAYVGSKTKEGVVHGVATVAE"], [3, 2]), (["APGLPVAEKG"], [4]), (["GGPKAG", "PAAPG"], [2, 3])):
        topo = tables.Topology([tables.Species.from_sequence(s, c) for s, c in zip(seqs, counts)])
        p = tables.make_params(boxl=200.0, tstar=0.3, canon=False)
        rng = np.random.default_rng(0)
        sv = np.concatenate([rng.random((topo.n_beads, 3)) - 0.5, rng.normal(size=(topo.n_beads, 3))], axis=1)
        m = OracleDMD(p, topo, tab).evcode_matrix()
        dev = DMD(p, topo, tab)
        dev.set_state(sv)
        N = topo.n_beads
        ii, jj = np.meshgrid(np.arange(1, N + 1), np.arange(1, N + 1), indexing="ij")
        mask = ii != jj
        got = dev.evcode(ii[mask], jj[mask])
        overlay = got >= 40
        assert np.array_equal(got[~overlay], m[mask][~overlay])


def test_engines_alternate_on_one_handle(tab, system_b):
    """Both engines work on the same resident state: chunks run alternately by the CTA-per-replica engine and the
    warp-per-replica engine give the oracle's sequence, and the batching statistics show real batches."""
    topo, sv, boxl = system_b
    n = 60000
    mk = lambda e: tables.make_params(boxl=boxl, tstar=0.18, canon=True, n_replicas=2, log_capacity=n, engine=e)
    ora = OracleDMD(mk(0), topo, tab)
    ora.set_state(sv)
    ora.run(n)
    dev = DMD(mk(2), topo, tab)
    dev.set_state(sv)
    dev.run(n)
    st = dev.batch_stats(0)
    assert st["rounds"] > 0 and (n - st["serial"]) / st["rounds"] > 3.0, st  # several events committed per round
    assert st["executed"] - st["rolled_back"] + st["serial"] == n
    la, lb = ora.event_log(), dev.event_log(0)
    for f in ("i", "j", "type", "evcode", "t"):
        assert np.array_equal(la[f], lb[f]), f
    assert np.array_equal(ora.state()["sv"], dev.state(0)["sv"])


@pytest.mark.parametrize("engine", [1, 2])
def test_ragged_two_species_box_with_gly(tab, engine):
    """Two species of different length with glycines, with and without -Dno_hbs: the oracle's sequence; and the run
    is deterministic (a second handle started from the same state repeats it bit for bit)."""
    for no_hbs in (False, True):
        topo, sv = genconfig.generate_box(["GAKLGVFE", "GAAKGS"], [3, 4], 60.0, 0.35, tab, seed=4)
        n = 60000
        p = tables.make_params(boxl=60.0, tstar=0.35, canon=True, no_hbs=no_hbs, n_replicas=3, log_capacity=n, seed=21,
                               engine=engine)
        ora, dev = _pair(p, topo, tab, sv)
        compare_engines(ora, dev, n_events=n)
        dev2 = DMD(p, topo, tab)
        dev2.set_state(sv)
        dev2.run(n)
        for r in range(3):
            assert np.array_equal(dev.event_log(r), dev2.event_log(r))
            assert np.array_equal(dev.state(r)["sv"], dev2.state(r)["sv"])


def test_run_until_output_both_engines(tab):
    """dmdb_run_until_output: both engines stop right after the first output pseudo-event, at the same event."""
    from conftest import check_run_until_output
    coll, t = check_run_until_output(tab, [1, 2])
    assert coll > 1000


def test_retemp_and_distinct_states(tab, system_b):
    topo, sv, boxl = system_b
    p = tables.make_params(boxl=boxl, tstar=0.18, canon=True, n_replicas=3, log_capacity=40000)
    ora, dev = _pair(p, topo, tab, sv)
    ora.run(10000)
    dev.run(10000)
    ora.retemp(0.26)
    dev.apply_temperatures([0.26, 0.18, 0.30])
    compare_engines(ora, dev, replica=0, n_events=0, cells=False)  # a temperature change keeps the lists: no new cells
    ora.run(20000)
    dev.run(20000)
    la, lb = ora.event_log(), dev.event_log(0)
    assert np.array_equal(la["i"], lb["i"]) and np.array_equal(la["type"], lb["type"]) and np.array_equal(la["t"], lb["t"])
    # distinct configurations per replica through the batched upload
    _, sv1 = genconfig.system_b(tab, 0.18, seed=6)
    p2 = tables.make_params(boxl=boxl, tstar=0.18, canon=False, n_replicas=2, log_capacity=5000)
    dev2 = DMD(p2, topo, tab)
    dev2.set_state_all(np.ascontiguousarray(np.stack([sv, sv1])))
    for r, s in ((0, sv), (1, sv1)):
        o = OracleDMD(p2, topo, tab)
        o.set_state(s)
        compare_engines(o, dev2, replica=r, n_events=5000 if r == 1 else 0)


@pytest.mark.parametrize("which,canon,tstar,n_events", [("B", True, 0.18, 100000), ("A", True, 0.5, 60000), ("H", True, 0.10, 300000)])
def test_list_rebuild_service_does_not_change_results(tab, system_a, system_b, which, canon, tstar, n_events):
    """Engine 1 with list-rebuild service CTAs (nbor() + events() done by other CTAs of the event-loop kernel, on
    other SMs) against the oracle, and against the same run with every warp rebuilding its own lists: identical
    committed-event sequences, times, final state and tallies -- also when the lists are rebuilt many times and
    H-bonds form (H: the 8 x A12 box at T* = 0.10)."""
    if which == "H":
        topo, sv = genconfig.generate_box(["AAAAAAAAAAAA"], [8], 45.0, 0.10, tab, seed=1)
        boxl = 45.0
    else:
        topo, sv, boxl = system_a if which == "A" else system_b
    p = tables.make_params(boxl=boxl, tstar=tstar, canon=canon, n_replicas=3, log_capacity=n_events, engine=1, seed=7)
    ora, dev = _pair(p, topo, tab, sv)
    dev.set_service_ctas(3)
    compare_engines(ora, dev, replica=0, n_events=n_events)
    assert dev.stats(0).updates + dev.stats(0).forced_updates >= 10  # the lists were rebuilt by the service
    for down in (False, True):  # ... and the lists of the last rebuild are the oracle's (sets; entry order is free)
        a, b = ora.nbors(down), dev.nbors(0, down)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), "neighbour lists after service rebuilds (down=%s)" % down
    ref = DMD(p, topo, tab)
    ref.set_state(sv)
    ref.set_service_ctas(0)
    ref.run(n_events)
    for r in range(3):
        a, b = dev.event_log(r), ref.event_log(r)
        for f in ("i", "j", "type", "t"):
            assert np.array_equal(a[f], b[f]), (r, f)
        assert np.array_equal(dev.state(r)["sv"], ref.state(r)["sv"])
        assert np.array_equal(dev.calendar(r)[0], ref.calendar(r)[0])
    with pytest.raises(DMDError):
        dev.set_service_ctas(-2)


def test_config3_temperature_ladder_with_exchange(tab, system_b):
    """BASELINE config 3 on one GPU: the reference's 11 temperatures (temp_018 ... temp_050) as two ladders of 48-peptide
    boxes with replica exchange between the runs.  The set of temperatures of each ladder is conserved, swaps are
    accepted, every replica's kinetic temperature follows the temperature it currently holds (Andersen thermostat),
    hot replicas lose their order faster than cold ones, and a sampled final state passes checkover.f."""
    from parallel_dmd_for_biomolecules_b200 import replica_exchange as rx
    topo, sv, boxl = system_b
    L = len(rx.LADDER)
    R = 2 * L
    dev = DMD(tables.make_params(boxl=boxl, tstar=0.18, canon=True, n_replicas=R, seed=5), topo, tab)
    dev.set_state(sv)
    dev.apply_temperatures(list(rx.LADDER) * 2)
    swaps = 0
    for step in range(8):
        dev.run(30000)
        epot, before = dev.potential_energies()
        st = dev.exchange(step, seed=99, ladder_size=L)  # dmdb_exchange: decision + temperature change on the device
        new_t = dev.potential_energies()[1]
        assert np.array_equal(new_t, rx.decide_swaps(epot, before, step, seed=99, ladder_size=L))  # numpy restatement
        assert (st.ladders, st.attempted) == (2, 2 * 5) and st.changed_local == 2 * st.accepted
        swaps += st.changed_local
        for lad in range(2):
            assert sorted(new_t[lad * L:(lad + 1) * L]) == sorted(rx.LADDER)
        assert st.changed_local == int((new_t != before).sum())
    assert swaps > 0
    dev.run(60000)
    _, tnow = dev.potential_energies()
    tred = np.array([dev.energy(r).tred for r in range(R)])
    assert np.all(np.abs(tred / (12.0 * tnow) - 1.0) < 0.12)  # 1344 beads: sigma of the kinetic temperature ~ 2 %
    assert dev.stats().forced_updates >= 0 and dev.stats().events == R * (8 * 30000 + 60000)
    dev.sync_positions()
    hot = int(np.argmax(tnow))
    o = OracleDMD(tables.make_params(boxl=boxl, tstar=float(tnow[hot]), canon=True), topo, tab)
    o.set_state(dev.state(hot)["sv"], dev.state(hot)["bptnr"])
    assert not o.checkover()[0]


def test_more_replicas_than_one_wave(tab, system_b):
    """Twice the replicas that fill the device: the event-loop CTAs take turns on the SMs the list-rebuild service CTAs
    leave free (several waves of one kernel).  NVE replicas started from one snapshot must stay bit-identical to
    each other and to a small handle that rebuilds its lists in place."""
    topo, sv, boxl = system_b
    R = 2 * device_fill(0)[0]
    n = 6000
    big = DMD(tables.make_params(boxl=boxl, tstar=0.18, canon=False, n_replicas=R, engine=1), topo, tab)
    big.set_state(sv)
    st = big.run(n)
    assert st.events == R * n
    small = DMD(tables.make_params(boxl=boxl, tstar=0.18, canon=False, n_replicas=2, engine=1), topo, tab)
    small.set_state(sv)
    small.set_service_ctas(0)
    small.run(n)
    assert small.stats(0).updates + small.stats(0).forced_updates >= 1  # the window contains list rebuilds
    ref = small.state(0)["sv"]
    for r in (0, 1, R // 2 - 1, R // 2, R - 29, R - 1):
        assert np.array_equal(big.state(r)["sv"], ref), r
        assert np.array_equal(big.calendar(r)[0], small.calendar(0)[0]), r


@pytest.mark.parametrize("engine", [1, 2, 3])
@pytest.mark.parametrize("which", ["A", "B"])
def test_frozen_event_sequence(tab, system_a, system_b, which, engine):
    """all three engines against the committed fixtures tests/golden/events_system{A,B}_nve.npz (the oracle's first
    10^4 events, frozen): no oracle in the loop, the files are the reference"""
    from conftest import check_against_frozen_events, frozen_events
    fx = frozen_events(which)
    topo = (system_a if which == "A" else system_b)[0]
    R = 1 if engine == 3 else 2
    p = tables.make_params(boxl=float(fx["boxl"]), tstar=float(fx["tstar"]), canon=False, n_replicas=R, log_capacity=len(fx["t"]),
                           engine=engine)
    dev = DMD(p, topo, tab)
    dev.set_state(fx["sv0"])
    if engine == 1:
        dev.set_service_ctas(2)
    check_against_frozen_events(dev, fx, replica=R - 1)


def test_device_fill_matches_the_launch(tab):
    replicas, service = device_fill(0)
    per_cta = 112  # 28 hardware warps per event-loop CTA, four replicas (8 lanes each) per warp
    assert replicas > 0 and replicas % per_cta == 0 and service >= 0
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    assert replicas // per_cta + service <= sms


def test_full_size_properties(tab, system_b):
    """BASELINE config 2 at bench size: the replica count that fills the device (dmdb_device_fill: 3584 replicas of the
    48-peptide box on a 148-SM B200, next to 20 list-rebuild service CTAs).  Size-independent properties:
    NVE energy is conserved in every replica, every replica's final state passes checkover.f (sampled), identical
    replicas stay bit-identical, and a second run of the same handle state is deterministic."""
    topo, sv, boxl = system_b
    R = device_fill(0)[0]
    p = tables.make_params(boxl=boxl, tstar=0.18, canon=False, n_replicas=R)
    dev = DMD(p, topo, tab)
    dev.set_state(sv)
    e0 = dev.energy(0).ered
    st = dev.run(20000)
    assert st.events == R * 20000
    ep, _ = dev.potential_energies()
    out = dev.get_state_all()
    assert np.array_equal(out[0], out[R - 1]) and np.array_equal(out[0], out[R // 2])
    for r in (0, 777, R - 1):
        assert abs(dev.energy(r).ered - e0) < 1e-6
    o = OracleDMD(tables.make_params(boxl=boxl, tstar=0.18, canon=False), topo, tab)
    dev.sync_positions()
    o.set_state(dev.state(R - 1)["sv"], dev.state(R - 1)["bptnr"])
    assert not o.checkover()[0]
    # thermostatted ensemble: replicas decorrelate, temperature is held, no replica reports a device error
    p = tables.make_params(boxl=boxl, tstar=0.18, canon=True, n_replicas=R)
    dev = DMD(p, topo, tab)
    dev.set_state(sv)
    dev.run(30000)
    tr = np.array([dev.energy(r).tred for r in range(0, R, 37)])
    assert abs(tr.mean() - 2.16) < 0.05 and tr.std() > 1e-4
    dev.sync_positions()
    o.set_state(dev.state(5)["sv"], dev.state(5)["bptnr"])
    assert not o.checkover()[0]


def test_errors_are_reported_not_fatal(tab, system_b):
    topo, sv, boxl = system_b
    dev = DMD(tables.make_params(boxl=boxl, tstar=0.18, nbr_capacity=4), topo, tab)
    with pytest.raises(DMDError) as e:
        dev.set_state(sv)
    assert e.value.code == 5
    dev2 = DMD(tables.make_params(boxl=boxl, tstar=0.18), topo, tab)
    with pytest.raises(DMDError):
        dev2.run(10)  # no state loaded
    with pytest.raises(ValueError):
        dev2.set_state(sv[:10])


@pytest.mark.parametrize("engine", [1, 2, 3])
def test_config4_12288_beads_matches_oracle(tab, engine):
    """BASELINE config 4 (192 chains x 16 residues = 12 288 beads): too large for the shared-memory state of the
    CTA-per-replica engine, which then works on the global arrays; cells, lists, calendar and 20 000 committed events
    equal the oracle's."""
    topo, sv = genconfig.generate_box(["KLVFFAEKLVFFAEKL"], [192], 200.0, 0.3, tab, seed=3)
    assert topo.n_beads == 12288
    p = tables.make_params(boxl=200.0, tstar=0.3, canon=True, n_replicas=2, log_capacity=20000, engine=engine)
    ora, dev = _pair(p, topo, tab, sv)
    compare_engines(ora, dev, n_events=20000)
    if engine >= 2:
        st = dev.batch_stats(0)
        assert (st["executed"] - st["rolled_back"]) / st["rounds"] > 8.0  # larger box, larger batches


def test_config5_million_bead_box_bulk_kernels(tab):
    """BASELINE config 5 (35 715 x KLVFFAE = 1 000 020 beads at the density of config 2): run start, nbor() and
    events() on the whole GPU.  The oracle's run start is O(N^2), so the check is by size-independent properties:
    up and down lists are transposes of each other, every listed partner of a bead with an event is on its up list,
    bonded-class neighbours are complete (each backbone bead lists its covalent partners), T* is the generated one."""
    nch = 35715
    boxl = 158.54 * (nch / 48.0) ** (1.0 / 3.0)
    topo, sv = genconfig.generate_box(["KLVFFAE"], [nch], boxl, 0.5, tab, seed=5)
    N = topo.n_beads
    assert N == 1000020
    dev = DMD(tables.make_params(boxl=boxl, tstar=0.5, canon=True, n_replicas=1, engine=1, nbr_capacity=32), topo, tab)
    dev.set_state(sv)
    off, nb = dev.nbors(0)
    offd, nbd = dev.nbors(0, True)
    assert len(nb) == len(nbd) > 4 * N
    i_of = np.repeat(np.arange(1, N + 1), np.diff(off))
    j_of = np.repeat(np.arange(1, N + 1), np.diff(offd))
    a = np.stack([i_of, nb], 1)
    b = np.stack([nbd, j_of], 1)
    assert np.array_equal(a[np.lexsort((a[:, 1], a[:, 0]))], b[np.lexsort((b[:, 1], b[:, 0]))])
    assert (nb > i_of).all() and (nbd < j_of).all()
    tim, nptnr, coltype = dev.calendar(0)
    has = nptnr[:N] > 0
    assert has.sum() > N // 2
    # the predicted partner of bead i is one of its up-list entries
    key = set(zip(i_of[:200000].tolist(), nb[:200000].tolist()))
    idx = np.nonzero(has[: i_of[199999] - 1])[0][:20000]
    assert all((int(k) + 1, int(nptnr[k])) in key for k in idx)
    # covalent partners: Ca_r lists N_r (code 4) and C_r (code 5) -- bead order per chain: Ca x7, N x7, C x7, R x7
    first = np.arange(0, 1000) * 28 + 1  # Ca of residue 1 of the first 1000 chains (1-based)
    for ca in first[:200]:
        ups = set(nb[off[ca - 1]:off[ca]].tolist())
        assert ca + 7 in ups and ca + 14 in ups
    e = dev.energy(0)
    assert abs(e.tred - 6.0) < 1e-6 and e.hb_ii + e.hb_ij == 0
    # lists rebuilt a second time are identical as sets; events() is idempotent
    dev.nbor()
    off2, nb2 = dev.nbors(0)
    assert np.array_equal(off, off2) and np.array_equal(nb, nb2)
    dev.events()
    tim2, nptnr2, coltype2 = dev.calendar(0)
    assert np.array_equal(tim, tim2) and np.array_equal(nptnr, nptnr2) and np.array_equal(coltype, coltype2)


def test_whole_gpu_engine_on_small_and_million_bead_boxes(tab, system_b):
    """Engine 3 (every round of the batched commit spread over the whole GPU; interval events through the bulk
    kernels): the oracle's sequence on config 2 (thermostat, interval events, list rebuilds on the way), and on the
    1 000 020-bead box of config 5 -- where the oracle cannot start -- exact NVE energy conservation over 10^6
    events with hundreds of events committed per round, and no overlap / broken bond in a sampled region."""
    topo, sv, boxl = system_b
    n = 60000
    p = tables.make_params(boxl=boxl, tstar=0.18, canon=True, n_replicas=1, log_capacity=n, engine=3)
    ora, dev = _pair(p, topo, tab, sv)
    compare_engines(ora, dev, n_events=n)
    assert dev.stats(0).updates + dev.stats(0).forced_updates > 10  # list rebuilds happened
    nch = 35715
    boxl5 = 158.54 * (nch / 48.0) ** (1.0 / 3.0)
    topo5, sv5 = genconfig.generate_box(["KLVFFAE"], [nch], boxl5, 0.5, tab, seed=5)
    d = DMD(tables.make_params(boxl=boxl5, tstar=0.5, canon=False, n_replicas=1, engine=3, nbr_capacity=32), topo5, tab)
    d.set_state(sv5)
    e0 = d.energy(0)
    st = d.run(1000000)
    e1 = d.energy(0)
    assert st.events == 1000000
    assert abs(e1.ered - e0.ered) < 1e-6 * abs(e0.ered) and abs(e1.ered - e0.ered) < 1e-3
    bs = d.batch_stats(0)
    assert (bs["executed"] - bs["rolled_back"]) / bs["rounds"] > 100
    # bonds of the first chains are still inside their windows (positions brought to the current time)
    d.sync_positions()
    x = d.state(0)["sv"][:28 * 50, :3].reshape(50, 28, 3)
    dca = x[:, 1:7] - x[:, 0:6]
    dca -= np.round(dca)
    dist = np.sqrt((dca ** 2).sum(-1)) * boxl5
    assert np.all(np.abs(dist - 3.8) <= 3.8 * 0.02375 + 1e-9)  # Ca-Ca pseudo-bond window (def.h:10,39)


def test_long_run_statistics_match_the_oracle_ensemble(tab):
    """North-star 'longer runs': an H-bond forming box (8 x A12, T* = 0.10) run for 1.5e6 events.  Beyond the
    bit-identical replica (same seed), an ensemble of 64 device replicas with OTHER seeds is statistically
    indistinguishable from an ensemble of 6 oracle trajectories: thermostat temperature, potential energy and
    H-bond count agree within the ensembles' standard errors."""
    topo, sv = genconfig.generate_box(["AAAAAAAAAAAA"], [8], 45.0, 0.10, tab, seed=1)
    n = 1500000
    R = 64
    dev = DMD(tables.make_params(boxl=45.0, tstar=0.10, canon=True, n_replicas=R, seed=1000, engine=1), topo, tab)
    dev.set_state(sv)
    dev.run(n)
    ep_d, _ = dev.potential_energies()
    hb_d = np.array([dev.energy(r).hb_ii + dev.energy(r).hb_ij for r in range(R)])
    t_d = np.array([dev.energy(r).tred for r in range(R)])
    ep_o, hb_o, t_o = [], [], []
    for seed in range(6):
        o = OracleDMD(tables.make_params(boxl=45.0, tstar=0.10, canon=True, seed=seed + 1), topo, tab)
        o.set_state(sv)
        o.run(n)
        e = o.energy()
        ep_o.append(e.ered - 0.5 * e.sumvel)
        hb_o.append(e.hb_ii + e.hb_ij)
        t_o.append(e.tred)
    ep_o, hb_o, t_o = np.array(ep_o), np.array(hb_o), np.array(t_o)
    assert abs(t_d.mean() - 1.2) < 0.03 and abs(t_o.mean() - 1.2) < 0.08  # setemp = 12 T*
    for a, b in ((ep_d, ep_o), (hb_d.astype(float), hb_o.astype(float))):
        se = np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
        assert abs(a.mean() - b.mean()) < 4.0 * se + 1e-9, (a.mean(), b.mean(), se)
    assert hb_d.mean() > 1.0  # hydrogen bonds did form


def test_whole_gpu_engine_25200_beads_matches_oracle(tab):
    """Engine 3 on 900 x KLVFFAE (25 200 beads, towards the largest box the oracle -- with the reference's N x N
    ev_code matrix -- can hold): dozens of events per round, the oracle's committed sequence bit for bit."""
    nch = 900
    boxl = 158.54 * (nch / 48.0) ** (1.0 / 3.0)
    topo, sv = genconfig.generate_box(["KLVFFAE"], [nch], boxl, 0.5, tab, seed=9)
    n = 60000
    p = tables.make_params(boxl=boxl, tstar=0.5, canon=True, n_replicas=1, log_capacity=n, engine=3, nbr_capacity=32)
    ora, dev = _pair(p, topo, tab, sv)
    compare_engines(ora, dev, n_events=n)
    bs = dev.batch_stats(0)
    assert (bs["executed"] - bs["rolled_back"]) / bs["rounds"] > 25
