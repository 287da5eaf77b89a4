"""Run driver (driver.py = what one `./dmd < temp_0xx` does around the loop): run numbering and restart chaining
(files_opn.f:19-46, inputinfo.f:79-101), record formats (config.f:16-40), the .energy line (main.F90:1380).
Runs on the CPU through the host-trace build of the engine (test scaffolding)."""
import os
import shutil

import numpy as np

from conftest import GOLDEN
from parallel_dmd_for_biomolecules_b200 import driver, fileio, tables


def test_two_chained_runs_write_reference_style_files(tmp_path, tab, system_a, hosttrace_lib):
    topo, sv, boxl = system_a
    res = tmp_path / "results"
    res.mkdir()
    # what genconfig leaves behind: run0000.config/.lastvel plus EMPTY run0000.energy/.bptnr
    shutil.copy(os.path.join(GOLDEN, "systemA_run0000.config"), res / "run0000.config")
    shutil.copy(os.path.join(GOLDEN, "systemA_run0000.lastvel"), res / "run0000.lastvel")
    (res / "run0000.energy").write_text("")
    (res / "run0000.bptnr").write_bytes(b"")
    # a short output period is not configurable (it is the reference's 3.3/sqrt(setemp)+5): events only
    s1 = driver.run_temperature(str(tmp_path), topo, tab, 0.5, 30000, boxl=boxl, lib_path=hosttrace_lib)
    assert s1["run"] == 1 and s1["events"] == 30000
    lines = (res / "run0001.energy").read_text().splitlines()
    assert len(lines) == s1["energy_lines"] >= 2
    assert all(len(ln) == 15 + 3 * 12 + 3 * 8 + 4 * 12 for ln in lines)  # (i15,3f12.4,3i8,4f12.4)
    assert int(lines[0][:15]) == 0 and int(lines[-1][:15]) == 29999
    coll, t, xyz = fileio.read_config(str(res / "run0001.config"))
    assert coll == 30000 and xyz.shape == (3, topo.n_beads) and np.abs(xyz).max() <= 0.5 and t > 0
    c2, vel = fileio.read_lastvel(str(res / "run0001.lastvel"))
    assert c2 == 30000 and vel.shape == (3, topo.n_beads)
    # temperature of the written velocities equals the last energy line's T column
    from parallel_dmd_for_biomolecules_b200 import genconfig
    m = genconfig.masses_of(topo, tab)
    tred = float((m * (vel ** 2).sum(axis=0)).sum() / 3.0 / topo.n_beads)
    assert abs(tred - float(lines[-1][15 + 24:15 + 36])) < 1e-3
    # second run restarts from run0001 and becomes run0002
    s2 = driver.run_temperature(str(tmp_path), topo, tab, 0.45, 10000, boxl=boxl, lib_path=hosttrace_lib)
    assert s2["run"] == 2
    coll0, t0, xyz0 = fileio.read_config(str(res / "run0001.config"))
    sv2 = fileio.sv_from_files(str(res / "run0001.config"), str(res / "run0001.lastvel"))
    assert np.array_equal(sv2[:, :3].T, xyz0)
    assert os.path.exists(res / "run0002.energy") and os.path.getsize(res / "run0002.config") > 0
    assert driver.first_unused_run(str(res)) == 3
    # final PDB of each run (write_rasmol-YM.f) and the bond audit of the restart structure under the PREVIOUS number
    _check_pdb(res / "run0001.pdb", topo, xyz0.T * boxl)
    assert os.path.exists(res / "run0002.pdb")
    rca = (res / "run0000.rca").read_text().splitlines()
    n_side = sum(sum(sp.firstside) * sp.n_chains for sp in topo.species)
    assert len([ln for ln in rca if ln[:3] in ("rca", "rnh", "rco")]) == 3 * n_side and os.path.exists(res / "run0001.rca")
    assert not any("overlap" in ln for ln in rca)  # genconfig places every side chain inside its bond windows


def _check_pdb(path, topo, xyz):
    import re
    lines = open(path).read().splitlines()
    assert len(lines) == sum(sp.n_chains * (sp.numbeads + sp.chnln) for sp in topo.species)
    pat = re.compile(r"^ATOM   [ \d*]{4}  (N |CA|C |O |CB)  [A-Z]{3} [AB][ \d]{4}    [ \-\d.]{24}$")
    assert all(pat.match(ln) for ln in lines), [ln for ln in lines if not pat.match(ln)][:3]
    pos = np.array([[float(ln[30:38]), float(ln[38:46]), float(ln[46:54])] for ln in lines])  # format 7: 3F8.3 from column 31
    name = [ln[12:16].strip() for ln in lines]
    sp = topo.species[0]
    L = sp.chnln
    # first chain: residue j has N = bead L+j, CA = bead j, C = bead 2L+j (1-based within the chain)
    k = 0
    nsc = 0
    for j in range(L):
        assert name[k:k + 4] == ["N", "CA", "C", "O"]
        np.testing.assert_allclose(pos[k], xyz[L + j], atol=6e-4)
        np.testing.assert_allclose(pos[k + 1], xyz[j], atol=6e-4)
        np.testing.assert_allclose(pos[k + 2], xyz[2 * L + j], atol=6e-4)
        assert abs(np.linalg.norm(pos[k + 3] - pos[k + 2]) - 1.231) < 3e-3  # the built carbonyl oxygen
        if sp.firstside[j]:
            assert name[k + 4] == "CB"
            np.testing.assert_allclose(pos[k + 4], xyz[3 * L + nsc], atol=6e-4)
            nsc += 1
            k += 5
        else:
            k += 4


def test_pdb_and_rca_of_the_shipped_snapshot(tmp_path, tab, system_a):
    topo, sv, boxl = system_a
    fileio.write_pdb(str(tmp_path / "a.pdb"), topo, sv[:, :3] * boxl)
    _check_pdb(tmp_path / "a.pdb", topo, sv[:, :3] * boxl)
    first = open(tmp_path / "a.pdb").readline().rstrip("\n")
    assert first == "ATOM      1  N   GLY A   1      -0.597  -4.345  49.828"
    lines = fileio.rca_lines(topo, tab, sv[:, :3], boxl)
    assert lines[0] == "rca    1    2  2  2.0029  2.0020 "  # format 7373; the integer column is the reference's aa(k) = 2
    # integers that do not fit their field print as asterisks, like the Fortran runtime does
    big = tables.Topology([tables.Species.from_sequence("KLVFFAE", 300)])
    xyz = np.zeros((big.n_beads, 3))
    xyz[:, 0] = np.arange(big.n_beads) * 0.01
    assert fileio.pdb_lines(big, xyz)[-1].startswith("ATOM   ****  CB  GLU A   7")


def test_observables_on_the_shipped_snapshot(tab, system_a):
    topo, sv, boxl = system_a
    rg = driver.radgyr(topo, sv[:, :3], boxl)
    e2e = driver.end_to_end(topo, sv[:, :3], boxl)
    # an extended 22-residue chain: Ca-Ca 3.8 A -> end-to-end of the order of 70 A, backbone Rg of the order of 20 A
    assert 55.0 < e2e < 85.0 and 15.0 < rg < 30.0


def test_run_until_output_stops_right_after_the_output_event(tab, hosttrace_lib):
    """A tiny box reaches the first output pseudo-event (3.3/sqrt(setemp)+5 time units, main.F90:423) within a few
    hundred thousand events: the run stops right after it and can go on.  (Both engines: tests/test_gpu_parity.py.)"""
    from conftest import check_run_until_output
    check_run_until_output(tab, [1], hosttrace_lib)


def test_command_line_mirrors_dmd_stdin(monkeypatch, capsys, tab):
    """`python -m parallel_dmd_for_biomolecules_b200 < temp_018`: T* and ncoll from stdin (main.F90:126-128, with the
    trailing comment temp_020 carries), sizes from options instead of -D macros."""
    import io
    from parallel_dmd_for_biomolecules_b200 import __main__ as cli
    seen = {}
    monkeypatch.setattr(cli.tables, "read_parameters_dir", lambda root: tab)
    monkeypatch.setattr(cli.tables, "read_topology_dir",
                        lambda root, n, L, nb: tables.Topology([tables.Species.from_sequence("KLVFFAE", k) for k in n]))
    monkeypatch.setattr(cli.driver, "run_temperature", lambda root, topo, t, tstar, ncoll, **kw: seen.update(
        tstar=tstar, ncoll=ncoll, n=topo.n_beads, **kw) or {"run": 1})
    monkeypatch.setattr("sys.stdin", io.StringIO("0.200D0 # temperature\n1000000000\n"))
    assert cli.main(["--root", "/nowhere", "--nve"]) == 0
    assert seen["tstar"] == 0.2 and seen["ncoll"] == 1000000000 and seen["n"] == 1344 and seen["canon"] is False
    assert "noptotal 1344" in capsys.readouterr().out


def _seed_results(root, golden=GOLDEN):
    res = root / "results"
    res.mkdir()
    shutil.copy(os.path.join(golden, "systemA_run0000.config"), res / "run0000.config")
    shutil.copy(os.path.join(golden, "systemA_run0000.lastvel"), res / "run0000.lastvel")
    (res / "run0000.energy").write_text("")
    (res / "run0000.bptnr").write_bytes(b"")
    return res


def check_anneal_resident_equals_chained(tmp_path, tab, system_a, lib_path, schedule):
    """qfile/script.sh:11-18 as ONE handle with restarts on resident state (dmdb_set_temperature) against one process per
    temperature chained through the restart files: every results file byte-identical"""
    topo, sv, boxl = system_a
    a, b = tmp_path / "resident", tmp_path / "chained"
    a.mkdir()
    b.mkdir()
    ra, rb = _seed_results(a), _seed_results(b)
    sa = driver.anneal(str(a), topo, tab, schedule, resident=True, boxl=boxl, lib_path=lib_path)
    sb = driver.anneal(str(b), topo, tab, schedule, resident=False, boxl=boxl, lib_path=lib_path)
    assert [s["run"] for s in sa] == [s["run"] for s in sb] == list(range(1, len(schedule) + 1))
    names = sorted(os.listdir(ra))
    assert names == sorted(os.listdir(rb)) and len(names) >= 4 + 6 * len(schedule)
    for n in names:
        assert (ra / n).read_bytes() == (rb / n).read_bytes(), n
    return sa


def test_anneal_on_resident_state_equals_chained_runs(tmp_path, tab, system_a, hosttrace_lib):
    check_anneal_resident_equals_chained(tmp_path, tab, system_a, hosttrace_lib, [(0.5, 6000), (0.45, 5000), (0.40, 5000)])


def test_truncated_record_keeps_the_complete_ones(tmp_path):
    """a run killed inside config() leaves a cut record: the reader keeps what is complete (read(7,end=120))"""
    src = os.path.join(GOLDEN, "systemA_run0000.config")
    data = open(src, "rb").read()
    p = tmp_path / "cut.config"
    p.write_bytes(data + data[: len(data) // 2])
    coll, t, xyz = fileio.read_config(str(p))
    c0, t0, x0 = fileio.read_config(src)
    assert coll == c0 and np.array_equal(xyz, x0)
