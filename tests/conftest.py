import os
import sys

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from parallel_dmd_for_biomolecules_b200 import fileio, genconfig, tables  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEQ_A = "GVAYVGSKTKEGVVHGVATVAE"  # recovered from genconfig/checks/identity.out
HOSTTRACE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_trace", "libdmdb_hosttrace.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def tab():
    return tables.load_default_tables()


@pytest.fixture(scope="session")
def system_a():
    """The reference's shipped snapshot: 8 x 22-mer, N=672, L=110 A, T*=0.5."""
    topo = tables.Topology([tables.Species.from_sequence(SEQ_A, 4), tables.Species.from_sequence(SEQ_A, 4)])
    sv = fileio.sv_from_files(os.path.join(GOLDEN, "systemA_run0000.config"), os.path.join(GOLDEN, "systemA_run0000.lastvel"))
    return topo, sv, 110.0


@pytest.fixture(scope="session")
def system_b(tab):
    """BASELINE config 2: 48 x KLVFFAE, N=1344, L=158.54 A."""
    topo, sv = genconfig.system_b(tab, 0.18, seed=1)
    return topo, sv, 158.54


@pytest.fixture(scope="session")
def hosttrace_lib():
    import __graft_entry__ as g
    g.build()
    return HOSTTRACE


def compare_engines(ora, dev, replica=0, n_events=0, time_rtol=1e-12, cells=True):
    """The parity surface of BASELINE.json's north_star: bit-exact cells / neighbour sets / next-event partner and
    type, event times within 1e-12 relative, identical committed-event sequence.  The engine is held to more than
    the stated tolerance: every time and every state double must be BIT-EQUAL to the oracle's (array_equal), so that
    a 1-ulp drift cannot hide behind the tolerance."""
    if cells:  # (the engine reports the cells of its last list rebuild, the oracle those of the current positions:
        #    comparable right after a rebuild only)
        assert np.array_equal(ora.cells(), dev.cells(replica)), "cell assignment"
    for down in (False, True):
        a, b = ora.nbors(down), dev.nbors(replica, down)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), "neighbour lists (down=%s)" % down
    ta, na, ca = ora.calendar()
    tb, nb, cb = dev.calendar(replica)
    assert np.array_equal(na, nb), "next-event partner"
    assert np.array_equal(ca, cb), "next-event type"
    np.testing.assert_allclose(tb, ta, rtol=time_rtol, atol=0)  # the stated bar ...
    assert np.array_equal(tb, ta), "calendar times are not bit-equal"  # ... and the one the engine is held to
    if n_events:
        ora.run(n_events)
        dev.run(n_events)
        la, lb = ora.event_log(), dev.event_log(replica)
        assert len(la) == len(lb) == n_events
        for f in ("i", "j", "type", "evcode"):
            bad = np.nonzero(la[f] != lb[f])[0]
            assert bad.size == 0, "event sequence differs in %s at event %d: %s vs %s" % (f, bad[0], la[bad[0]], lb[bad[0]])
        np.testing.assert_allclose(lb["t"], la["t"], rtol=time_rtol, atol=1e-300)
        assert np.array_equal(lb["t"], la["t"]), "event times are not bit-equal"
        sa, sb = ora.state(), dev.state(replica)
        assert np.array_equal(sb["sv"], sa["sv"]), "positions / velocities are not bit-equal"
        assert sa["t"] == sb["t"] and sa["tfalse"] == sb["tfalse"]
        for f in ("bptnr", "identity", "extra_repuls"):
            assert np.array_equal(sa[f], sb[f]), f
        assert sa["coll"] == sb["coll"]


def audit_nc(ora, dev=None, replica=0, make_oracle=None):
    """check_nc_int.f:21-360 (main.F90:425, 1199): the reference's own audit of the H-bond <-> auxiliary-shoulder state
    machine, on the oracle's state and -- through a second oracle instance that ADOPTS the device's read-back state and
    lists -- on the device's.  boundbad == unboundbad == 0, no "no ss" complaint, m_ss == n_ss (the Fortran exits
    otherwise)."""
    a = ora.check_nc_int()
    assert a["boundbad"] == 0 and a["unboundbad"] == 0 and a["no_ss"] == 0 and a["m_ss"] == a["n_ss"], a
    if dev is not None:
        aud = make_oracle()
        aud.adopt_state(dev.state(replica), dev.nbors(replica))
        b = aud.check_nc_int()
        assert b == a, (a, b)
        assert not aud.checkover()[0]
    return a


def check_run_until_output(tab, engines, lib_path=None):
    from parallel_dmd_for_biomolecules_b200.dmd import DMD
    topo, sv = genconfig.generate_box(["AAAA"], [2], 40.0, 0.5, tab, seed=2)
    N = topo.n_beads
    stops = []
    for engine in engines:
        p = tables.make_params(boxl=40.0, tstar=0.5, canon=True, n_replicas=1, log_capacity=4, engine=engine, seed=7)
        with DMD(p, topo, tab, lib_path=lib_path) as d:
            d.set_state(sv)
            budget = 20_000_000
            d.run_until_output(budget)
            st = d.state(0)
            assert 0 < st["coll"] < budget
            tim, nptnr, coltype = d.calendar(0)
            period = 3.3 / np.sqrt(6.0) + 5
            assert abs((tim[N + 2] - st["tfalse"]) - period) < 1e-9  # the output event has just re-armed itself
            assert abs(st["t"] + st["tfalse"] - period) < 1e-6       # ... at the first output time
            stops.append((st["coll"], st["t"] + st["tfalse"]))
            d.run(1000)
            assert d.state(0)["coll"] == st["coll"] + 1000
    assert all(s == stops[0] for s in stops)
    return stops[0]


def frozen_events(which):
    """the oracle's first 10^4 NVE events of system A / B, frozen by tests/golden/make_event_fixtures.py"""
    return np.load(os.path.join(GOLDEN, "events_system%s_nve.npz" % which))


def check_against_frozen_events(engine_obj, fx, replica=None):
    """engine_obj: an OracleDMD or DMD with the fixture's snapshot loaded and a log of >= 10^4 events"""
    kw = {} if replica is None else {"replica": replica}
    tim, nptnr, coltype = engine_obj.calendar(**kw) if kw else engine_obj.calendar()
    assert np.array_equal(nptnr, fx["cal_ptnr0"]) and np.array_equal(coltype, fx["cal_type0"])
    assert np.array_equal(tim, fx["cal_t0"])  # bit-equal (the bar is 1e-12 relative)
    n = len(fx["t"])
    engine_obj.run(n)
    log = engine_obj.event_log(**kw) if kw else engine_obj.event_log()
    assert len(log) == n
    for f in ("i", "j", "type", "evcode"):
        bad = np.nonzero(log[f] != fx[f])[0]
        assert bad.size == 0, "frozen event sequence differs in %s at event %d" % (f, bad[0])
    np.testing.assert_allclose(log["t"], fx["t"], rtol=1e-12, atol=0)
    assert np.array_equal(log["t"], fx["t"])
