"""`-m gpu`: the run driver (driver.py = what `./dmd < temp_0xx` and qfile/script.sh:11-18 do around the loop) on
libdmdb200.so.  SURVEY.md 8f rows 1 and 3: run numbering, restart chaining, .energy lines and unformatted records at
the output pseudo-events, the final PDB and the .rca audit -- identical to the same driver on the oracle-checked 1-lane
host-trace build of the engine source; and the annealing schedule on RESIDENT device state (one handle,
dmdb_set_temperature) against one process per temperature chained through the restart files."""
import os

import pytest

from test_driver_hosttrace import _seed_results, check_anneal_resident_equals_chained
from parallel_dmd_for_biomolecules_b200 import driver

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("engine", [1, 2])
def test_driver_files_equal_the_hosttrace_build(tmp_path, tab, system_a, hosttrace_lib, engine):
    topo, sv, boxl = system_a
    a, b = tmp_path / "gpu", tmp_path / "trace"
    a.mkdir()
    b.mkdir()
    ra, rb = _seed_results(a), _seed_results(b)
    sched = [(0.5, 40000), (0.45, 15000)]
    sa = driver.anneal(str(a), topo, tab, sched, resident=True, boxl=boxl, engine=engine)                     # the product library
    sb = driver.anneal(str(b), topo, tab, sched, resident=False, boxl=boxl, lib_path=hosttrace_lib)           # test scaffolding
    assert [s["events"] for s in sa] == [s["events"] for s in sb] == [40000, 15000]
    names = sorted(os.listdir(ra))
    assert names == sorted(os.listdir(rb))
    for n in names:
        assert (ra / n).read_bytes() == (rb / n).read_bytes(), n
    assert sa[0]["energy_lines"] >= 2 and (ra / "run0002.pdb").exists() and (ra / "run0001.rca").exists()


def test_anneal_on_resident_device_state_equals_chained_runs(tmp_path, tab, system_a):
    """qfile/script.sh:11-14 (0.50 -> 0.22, shortened) on one handle: no per-temperature host round trip (8f-3)"""
    check_anneal_resident_equals_chained(tmp_path, tab, system_a, None,
                                         [(0.50, 20000), (0.45, 10000), (0.40, 10000), (0.35, 10000), (0.30, 10000)])


def test_run_after_sync_positions_is_refused(tab, system_a):
    from parallel_dmd_for_biomolecules_b200 import tables
    from parallel_dmd_for_biomolecules_b200.dmd import DMD, DMDError
    topo, sv, boxl = system_a
    d = DMD(tables.make_params(boxl=boxl, tstar=0.5, canon=True, n_replicas=2), topo, tab)
    d.set_state(sv)
    d.run(2000)
    d.sync_positions()
    x = d.state(0)["sv"].copy()
    d.sync_positions()  # idempotent
    assert (d.state(0)["sv"] == x).all()
    with pytest.raises(DMDError):
        d.run(10)
    d.set_temperature(0.45)  # a restart on the resident state makes the handle runnable again
    d.run(10)
