"""Pins the CPU oracle against everything the reference ships that constrains this path:
the system-A snapshot + genconfig/checks known answers, the derived constants of SURVEY.md App. C, the static
ev_code histogram, and the reference's own invariants (checkover.f, NVE nint(E) conservation)."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN, SEQ_A
from oracle.binding import OracleDMD, lib as oracle_lib
from parallel_dmd_for_biomolecules_b200 import fileio, genconfig, tables


def test_snapshot_format_and_known_answers(tab, system_a):
    topo, sv, boxl = system_a
    coll, t, xyz = fileio.read_config(os.path.join(GOLDEN, "systemA_run0000.config"))
    assert coll == 0 and t == 0.0 and xyz.shape == (3, 672)
    assert np.abs(xyz).max() <= 0.5
    chk = np.load(os.path.join(GOLDEN, "systemA_checks.npz"))
    # bead identities incl. Gly (id 9) as genconfig/checks/identity.out lists them: 8 chains x 88
    ident = chk["identity_with_gly"]
    assert len(ident) == 704
    assert np.array_equal(ident[ident != 9], topo.bead_identity())
    # masses.out: column 3 = mass of each of the 704 slots
    o = OracleDMD(tables.make_params(boxl=boxl, tstar=0.5, canon=False), topo, tab)
    o.set_state(sv)
    assert np.allclose(o.masses(), chk["masses"][ident != 9, 2], rtol=0, atol=1e-12)
    # sumvelcheck.out: running sum of m v^2 over beads; the last value is 3 N setemp = 12096
    run = np.cumsum(o.masses() * (sv[:, 3:] ** 2).sum(axis=1))
    assert np.allclose(run, chk["sumvelcheck"][:, 1], rtol=1e-12)
    e = o.energy()
    assert abs(e.sumvel - 12096.0) < 1e-6 and abs(e.tred - 6.0) < 1e-9
    assert (e.hb_ii, e.hb_ij, e.hb_alpha) == (0, 0, 0)
    assert not o.checkover()[0]


@pytest.mark.parametrize("which", ["A", "B"])
def test_derived_constants_appendix_c(tab, system_a, system_b, which):
    topo, sv, boxl = system_a if which == "A" else system_b
    tstar = 0.18
    o = OracleDMD(tables.make_params(boxl=boxl, tstar=tstar, canon=False), topo, tab)
    o.set_state(sv)
    c = o.constants()
    exp = dict(A=(672, 6.953, 30, 4.2308, 5.3906, 5.8906, 8.3436, 0.5562),
               B=(1344, 6.887, 42, 4.1721, 5.3774, 5.8774, 8.2644, 0.5510))[which]
    assert o.N == exp[0]
    assert abs(c["sig_max_all"] * boxl - exp[1]) < 1e-9
    assert o.num_cell == exp[2]
    assert abs(c["width"] * boxl - exp[3]) < 5e-5
    rl = np.sqrt(c["rlsq"]) * boxl
    assert abs(rl[0] - exp[4]) < 5e-5 and abs(rl[39] - exp[4]) < 5e-5 and abs(rl[49] - exp[4]) < 5e-5
    assert abs(rl[14] - exp[5]) < 5e-5 and abs(rl[15] - exp[6]) < 5e-5
    assert abs(math.sqrt(c["hdelr"]) * boxl - exp[7]) < 5e-5
    # T*=0.18 row: setemp, interval, interval_max, sortsize, output period
    assert abs(c["setemp"] - 2.16) < 1e-12
    assert abs(c["interval"] - 3.40207e-5) < 1e-10
    assert abs(c["interval_max"] - 5.10310e-3) < 1e-8
    assert abs(c["sortsize"] - 2.55155e-6) < 1e-11
    assert abs(c["t_output"] - 7.24537) < 1e-5


def test_neighbour_statistics_of_snapshot(tab, system_a):
    """P_up = 3587 up-pairs, 5.34 per bead, max up/down list length 12/10 on the shipped snapshot."""
    topo, sv, boxl = system_a
    o = OracleDMD(tables.make_params(boxl=boxl, tstar=0.5, canon=False), topo, tab)
    o.set_state(sv)
    off, nb = o.nbors()
    assert len(nb) == 3587 and np.diff(off).max() == 12
    offd, nbd = o.nbors(down=True)
    assert len(nbd) == 3587 and np.diff(offd).max() == 10
    codes = o.evcode(np.repeat(np.arange(1, o.N + 1), np.diff(off)), nb)
    assert (codes == 1).sum() == 424 and (codes == 15).sum() == 0 and (codes == 16).sum() == 3
    assert ((codes >= 4) & (codes <= 26) & (codes != 15) & (codes != 16)).sum() == 3160


@pytest.mark.parametrize("seq,expect", [
    ("KLVFFAE", {4: 56, 5: 56, 6: 48, 7: 48, 8: 104, 9: 48, 10: 56, 11: 56, 12: 56, 15: 1272, 16: 636, 17: 48, 18: 48,
                 19: 40, 20: 48, 21: 48, 22: 48, 23: 48, 24: 48, 25: 48, 26: 40}),
    (SEQ_A, {4: 176, 5: 176, 6: 168, 7: 168, 8: 344, 9: 168, 10: 144, 11: 144, 12: 144, 15: 14352, 16: 4800, 17: 168,
             18: 168, 19: 160, 20: 168, 21: 168, 22: 144, 23: 136, 24: 144, 25: 136, 26: 136}),
])
def test_static_evcode_histogram_2plus2_chains(tab, seq, expect):
    """SURVEY.md App. C last row: off-diagonal (symmetric) counts of the literal make_code.f matrix, 2+2 chains."""
    topo = tables.Topology([tables.Species.from_sequence(seq, 2), tables.Species.from_sequence(seq, 2)])
    o = OracleDMD(tables.make_params(boxl=110.0, tstar=0.5, canon=False), topo, tab)
    m = o.evcode_matrix()
    assert np.array_equal(m, m.T)
    u, cnt = np.unique(m[~np.eye(o.N, dtype=bool)], return_counts=True)
    got = {int(k): int(v) for k, v in zip(u, cnt) if k != 1}
    assert got == expect


def test_fdlibm_log_and_rng(tab, system_a):
    l = oracle_lib()
    xs = np.concatenate([np.random.default_rng(0).random(2000), [1e-18, 0.5, 0.999999999, 1.0 - 2 ** -53, 2 ** -30]])
    for x in xs:
        assert abs(l.dmdo_log(float(x)) - math.log(x)) <= 5e-16 * max(1.0, abs(math.log(x)))  # fdlibm: < 1 ulp
    topo, sv, boxl = system_a
    o = OracleDMD(tables.make_params(boxl=boxl, tstar=0.5, seed=7), topo, tab)
    u = np.array([l.dmdo_rng(o._h) for _ in range(20000)])
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.01 and abs(u.var() - 1 / 12) < 0.005


@pytest.mark.parametrize("which", ["A", "B"])
def test_reference_invariants_after_events(tab, system_a, system_b, which):
    """checkover.f (no core overlap, bonds in range) and the NVE check of main.F90:928-942 (nint(E) constant)."""
    topo, sv, boxl = system_a if which == "A" else system_b
    tstar = 0.5 if which == "A" else 0.18
    o = OracleDMD(tables.make_params(boxl=boxl, tstar=tstar, canon=False), topo, tab)
    o.set_state(sv)
    e0 = o.energy().ered
    for _ in range(4):
        o.run(50000)
        assert not o.checkover()[0]
        assert round(o.energy().ered) == round(e0)
        assert abs(o.energy().ered - e0) < 1e-6
    s = o.stats()
    assert s.events == 200000 and s.ghosts == 0 and s.updates > 50


def test_thermostat_holds_temperature(tab, system_b):
    topo, sv, boxl = system_b
    o = OracleDMD(tables.make_params(boxl=boxl, tstar=0.18, canon=True), topo, tab)
    o.set_state(sv)
    tr = []
    for _ in range(10):
        o.run(100000)
        tr.append(o.energy().tred)
    assert abs(np.mean(tr[3:]) - 2.16) < 0.1  # setemp = 12 T*
    assert o.stats().ghosts > 1000
    assert not o.checkover()[0]


@pytest.mark.parametrize("which", ["A", "B"])
def test_oracle_reproduces_its_frozen_event_sequence(tab, system_a, system_b, which):
    """tests/golden/events_system{A,B}_nve.npz: the first 10^4 committed events (owner, partner, type, ev_code, time)
    and the initial calendar, frozen from the oracle.  The reference has no golden vectors for these,
    so this pins the oracle against its own history."""
    from conftest import check_against_frozen_events, frozen_events
    fx = frozen_events(which)
    topo = (system_a if which == "A" else system_b)[0]
    p = tables.make_params(boxl=float(fx["boxl"]), tstar=float(fx["tstar"]), canon=False, log_capacity=len(fx["t"]))
    o = OracleDMD(p, topo, tab)
    o.set_state(fx["sv0"])
    check_against_frozen_events(o, fx)
    e = o.energy()
    assert abs(e.ered - float(fx["ered"])) < 1e-9 and [e.hb_alpha, e.hb_ii, e.hb_ij] == list(fx["hb"])


def test_run_file_round_trip(tmp_path, system_a):
    topo, sv, boxl = system_a
    cfg, vel, bp = tmp_path / "run0001.config", tmp_path / "run0001.lastvel", tmp_path / "run0001.bptnr"
    for k in range(3):
        fileio.append_config(str(cfg), 100 * k, 0.5 * k, sv[:, :3].T)
        fileio.append_bptnr(str(bp), 100 * k, np.arange(len(sv), dtype=np.int32) * (k == 2))
    fileio.write_lastvel(str(vel), 200, sv[:, 3:].T)
    coll, t, xyz = fileio.read_config(str(cfg))
    assert (coll, t) == (200, 1.0) and np.array_equal(xyz.T, sv[:, :3])
    assert np.array_equal(fileio.read_lastvel(str(vel))[1].T, sv[:, 3:])
    assert np.array_equal(fileio.read_bptnr(str(bp), len(sv)), np.arange(len(sv)))
    # the reference's own snapshot is byte-identical after a read/append round trip
    src = os.path.join(GOLDEN, "systemA_run0000.config")
    c0, t0, x0 = fileio.read_config(src)
    out = tmp_path / "copy.config"
    fileio.append_config(str(out), c0, t0, x0)
    assert open(src, "rb").read() == open(out, "rb").read()
    assert fileio.energy_line(12, 0.5, 6048.0, 6.0, 0, 1, 2, 0.5, 1.5, 10.0, 20.0) == \
        "             12      0.5000   6048.0000      6.0000       0       1       2      0.5000      1.5000     10.0000     20.0000"


def test_check_nc_int_audit_holds_along_an_hbond_rich_run(tab):
    """check_nc_int.f:21-360 (an oracle invariant the reference itself defines): along a run in which
    hydrogen bonds form and break, every audit finds boundbad == unboundbad == 0, no missing shoulder set and
    m_ss == n_ss -- the condition whose violation makes the Fortran exit (check_nc_int.f:352-355)."""
    from conftest import audit_nc
    from oracle.binding import OracleDMD
    from parallel_dmd_for_biomolecules_b200 import genconfig
    topo, sv = genconfig.generate_box(["AAAAAAAAAAAA"], [8], 45.0, 0.10, tab, seed=1)
    o = OracleDMD(tables.make_params(boxl=45.0, tstar=0.10, canon=True, seed=11), topo, tab)
    o.set_state(sv)
    seen = 0
    for _ in range(10):
        o.run(60000)
        seen = max(seen, audit_nc(o)["m_ss"])
        assert not o.checkover()[0]
    assert seen >= 4  # shoulder sets were switched on (N-C pairs inside their well), so the audit had something to check
