"""Drop-in run driver: what one ``./dmd < temp_0xx`` invocation of the reference does around its simulation loop
(code/main.F90:117-425 set-up, :1191-1246 output events, :1281-1374 wrap-up), with the loop itself on the GPU.

Files, all under ``<workdir>/results``:
  runNNNN.energy   text, one line per output event, format 22223 ``(i15,3f12.4,3i8,4f12.4)`` (main.F90:1380)
  runNNNN.config   ifort unformatted, appended records ``coll, t+tfalse, x[N], y[N], z[N]`` (config.f:16-24)
  runNNNN.bptnr    appended records ``coll, bptnr[N]`` (config.f:25-28)
  runNNNN.lastvel  one record ``coll, vx[N], vy[N], vz[N]``, rewritten at every output event (config.f:30-40)
  runNNNN.pdb      final structure with built carbonyl oxygens (write_rasmol-YM.f, main.F90:1343-1346)
  run(N-1).rca     side-chain bond audit of the restart structure, under the PREVIOUS run's number (inputinfo.f:413-505)
Run numbering and restart chaining follow files_opn.f:19-46 and inputinfo.f:79-101: the new run takes the first
unused number and restarts from the LAST record of the previous run's .config/.bptnr and its .lastvel.

The annealing schedule of qfile/script.sh:11-18 is ``anneal()``: one call per temperature on resident files.
Host-side observables that are not on the hot path (radius of gyration radgyr.f, end-to-end distance end2end.f)
are evaluated here in numpy from the downloaded configuration.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Sequence

import numpy as np

from . import fileio
from .dmd import DMD
from .tables import Tables, Topology, make_params

SIGMA_N = 3.3  # sigma(1) * boxl_orig in the reduced-time prefactor of main.F90:377,1210 (parameters/protein.data)


def first_unused_run(results_dir: str) -> int:
    """files_opn.f:19-46: first NNNN for which neither runNNNN.energy nor runNNNN.energy.gz exists."""
    i = 0
    while os.path.exists(os.path.join(results_dir, "run%04d.energy" % i)) or \
            os.path.exists(os.path.join(results_dir, "run%04d.energy.gz" % i)):
        i += 1
        if i > 9999:
            raise RuntimeError("file name error. too many output files.")
    return i


def _unwrap_backbone(xyz: np.ndarray, L: int) -> np.ndarray:
    """radgyr.f:22-56 for one chain: Ca chain made continuous, then N and C of each residue brought next to their
    Ca (box units, box length 1)."""
    p = xyz[: 3 * L].copy()
    for j in range(1, L):
        d = p[j] - p[j - 1]
        p[j] -= np.where(d > 0.5, 1.0, 0.0) - np.where(d < -0.5, 1.0, 0.0)
    for blk in (1, 2):
        d = p[blk * L:(blk + 1) * L] - p[:L]
        p[blk * L:(blk + 1) * L] -= np.where(d > 0.5, 1.0, 0.0) - np.where(d < -0.5, 1.0, 0.0)
    return p


def radgyr(topo: Topology, xyz: np.ndarray, boxl: float) -> float:
    """radgyr.f: backbone radius of gyration, pooled over chains (Angstrom).  xyz: (N,3) true positions, box units."""
    tot, nb, off = 0.0, 0, 0
    for sp in topo.species:
        L, nbd = sp.chnln, sp.numbeads
        for _ in range(sp.n_chains):
            p = _unwrap_backbone(xyz[off:off + nbd], L)
            tot += float(((p - p.mean(axis=0)) ** 2).sum())
            nb += 3 * L
            off += nbd
    return math.sqrt(tot / nb) * boxl


def end_to_end(topo: Topology, xyz: np.ndarray, boxl: float) -> float:
    """end2end.f: mean distance between the first N and the last C of each chain (Angstrom)."""
    tot, nch, off = 0.0, 0, 0
    for sp in topo.species:
        L, nbd = sp.chnln, sp.numbeads
        for _ in range(sp.n_chains):
            d = xyz[off + L] - xyz[off + 3 * L - 1]
            d = d - np.round(d)
            tot += float(np.sqrt((d * d).sum()))
            nch += 1
            off += nbd
    return tot / nch * boxl


class RunFiles:
    def __init__(self, workdir: str):
        self.results = os.path.join(workdir, "results")
        os.makedirs(self.results, exist_ok=True)
        self.run = first_unused_run(self.results)
        self.prev = self.run - 1
        self.energy_path = self.path(self.run, "energy")
        open(self.energy_path, "x").close()  # status='new' (files_opn.f:42)
        for ext in ("config", "bptnr"):
            open(self.path(self.run, ext), "ab").close()

    def path(self, run: int, ext: str) -> str:
        return os.path.join(self.results, "run%04d.%s" % (run, ext))

    def restart(self, n_beads: int):
        """inputinfo.f:79-101 + main.F90:241-246: last .config record, the .lastvel record, last .bptnr record of
        the previous run (an empty .bptnr leaves all zeros)."""
        if self.prev < 0:
            raise FileNotFoundError("no previous run to restart from: results/run0000.* (genconfig output) expected")
        sv = fileio.sv_from_files(self.path(self.prev, "config"), self.path(self.prev, "lastvel"))
        if sv.shape[0] != n_beads:
            raise ValueError("restart files hold %d beads, topology has %d" % (sv.shape[0], n_beads))
        bp = fileio.read_bptnr(self.path(self.prev, "bptnr"), n_beads) if os.path.exists(self.path(self.prev, "bptnr")) \
            else np.zeros(n_beads, dtype=np.int32)
        return sv, bp

    def energy_line(self, line: str) -> None:
        with open(self.energy_path, "a") as f:
            f.write(line + "\n")

    def config(self, coll: int, t: float, true_xyz: np.ndarray, vel: np.ndarray, bptnr: np.ndarray) -> None:
        """config.f:16-40"""
        fileio.append_config(self.path(self.run, "config"), coll, t, true_xyz.T)
        fileio.append_bptnr(self.path(self.run, "bptnr"), coll, bptnr)
        fileio.write_lastvel(self.path(self.run, "lastvel"), coll, vel.T)


def _run_on_handle(d: DMD, files: RunFiles, topo: Topology, tstar: float, ncoll: int, boxl: float) -> dict:
    """the body of one ``./dmd < temp_0xx`` run on a handle whose state has just been (re)started at ``tstar``:
    first .energy line (main.F90:347-378), the event loop with a record at every output pseudo-event
    (main.F90:1191-1246), the wrap-up records and the final PDB (main.F90:1288-1346)"""
    setemp = 12.0 * tstar
    period = 3.3 / math.sqrt(setemp) + 5.0  # main.F90:1244
    N = topo.n_beads
    lines = 0

    def observables():
        st = d.state(0)
        true_xyz = st["sv"][:, :3] + st["sv"][:, 3:] * st["tfalse"]  # config.f:19-23
        e = d.energy(0)
        rg, e2e = radgyr(topo, true_xyz, boxl), end_to_end(topo, true_xyz, boxl)
        return st, true_xyz, e, rg, e2e

    st, xyz, e, rg, e2e = observables()
    files.energy_line(fileio.energy_line(0, 0.0, e.ered, e.tred, e.hb_alpha, e.hb_ii, e.hb_ij, e.ehh_ii, e.ehh_ij, rg, e2e))
    lines += 1
    done = 0
    while done < ncoll:
        d.run_until_output(ncoll - done)
        st = d.state(0)
        done = st["coll"]
        # did the replica stop right after an output pseudo-event?  That event re-arms itself one period after its own
        # time (main.F90:1244), and tfalse moves on with every later event -- so the test also holds when the output
        # event was the very last event of the budget.
        tim = d.calendar(0)[0]
        if abs((tim[N + 2] - st["tfalse"]) - period) < 1e-9:
            st, xyz, e, rg, e2e = observables()
            tred_time = (st["t"] + st["tfalse"]) * math.sqrt(setemp) / SIGMA_N
            files.energy_line(fileio.energy_line(done, tred_time, e.ered, e.tred, e.hb_alpha, e.hb_ii, e.hb_ij,
                                                 e.ehh_ii, e.ehh_ij, rg, e2e))
            files.config(done, st["t"] + st["tfalse"], xyz, st["sv"][:, 3:], st["bptnr"])
            lines += 1
    # wrap-up, main.F90:1288-1330: true positions, wrapped; final line and records
    d.sync_positions()
    st = d.state(0)
    xyz = st["sv"][:, :3]
    e = d.energy(0)
    rg, e2e = radgyr(topo, xyz, boxl), end_to_end(topo, xyz, boxl)
    t_end = st["t"] + st["tfalse"]
    files.energy_line(fileio.energy_line(done - 1, t_end * math.sqrt(setemp) / SIGMA_N, e.ered, e.tred, e.hb_alpha,
                                         e.hb_ii, e.hb_ij, e.ehh_ii, e.ehh_ij, rg, e2e))
    files.config(done, t_end, xyz, st["sv"][:, 3:], st["bptnr"])
    lines += 1
    fileio.write_pdb(files.path(files.run, "pdb"), topo, xyz * boxl)  # after scale_up.f: Angstrom
    stats = d.stats(0)
    return dict(run=files.run, energy_lines=lines, events=done, ered=e.ered, tred=e.tred, hb=e.hb_ii + e.hb_ij,
                ghosts=stats.ghosts, updates=stats.updates + stats.forced_updates)


def run_temperature(workdir: str, topo: Topology, tables: Tables, tstar: float, ncoll: int, boxl: float = 158.54,
                    canon: bool = True, seed: int = 1058472402, engine: int = 0, lib_path: Optional[str] = None,
                    device: int = 0) -> dict:
    """One ``./dmd < temp_0xx`` run: restart from the previous run's files, ``ncoll`` calendar events at T*, results
    written like the reference.  Returns a summary (run number, lines written, final energy record)."""
    files = RunFiles(workdir)
    sv, bp = files.restart(topo.n_beads)
    fileio.write_rca(files.path(files.prev, "rca"), topo, tables, sv[:, :3], boxl)
    p = make_params(boxl=boxl, tstar=tstar, canon=canon, n_replicas=1, device=device, seed=seed, engine=engine)
    with DMD(p, topo, tables, lib_path=lib_path) as d:
        d.set_state(sv, bp)
        return _run_on_handle(d, files, topo, tstar, ncoll, boxl)


def anneal(workdir: str, topo: Topology, tables: Tables, schedule: Sequence[Sequence[float]], resident: bool = True,
           boxl: float = 158.54, canon: bool = True, seed: int = 1058472402, engine: int = 0,
           lib_path: Optional[str] = None, device: int = 0) -> List[dict]:
    """qfile/script.sh:11-18: ``for t in 050 045 ... 022: ./dmd < temp_$t`` then the long run at 0.18.

    resident=True (default): ONE handle for the whole schedule.  The first run restarts from the files in ``workdir``;
    every later run starts from the state resident on the device -- ``dmdb_set_temperature`` does on the device what
    the reference does through its restart files (true positions, wrap, start-up path with the new T*, RNG re-seeded)
    -- so the per-temperature host round trip is gone while every runNNNN.* file is byte-identical to the chained runs.
    resident=False: one run_temperature call per (T*, ncoll) pair, each restarting from the files of the previous."""
    kw = dict(boxl=boxl, canon=canon, seed=seed, engine=engine, lib_path=lib_path, device=device)
    if not resident:
        return [run_temperature(workdir, topo, tables, float(t), int(n), **kw) for t, n in schedule]
    out = []
    d = None
    try:
        for k, (t, n) in enumerate(schedule):
            files = RunFiles(workdir)
            if d is None:
                sv, bp = files.restart(topo.n_beads)
                p = make_params(boxl=boxl, tstar=float(t), canon=canon, n_replicas=1, device=device, seed=seed, engine=engine)
                d = DMD(p, topo, tables, lib_path=lib_path)
                d.set_state(sv, bp)
                start_xyz = sv[:, :3]
            else:
                d.set_temperature(float(t))  # restart on resident state
                start_xyz = d.state(0)["sv"][:, :3]
            fileio.write_rca(files.path(files.prev, "rca"), topo, tables, start_xyz, boxl)
            out.append(_run_on_handle(d, files, topo, float(t), int(n), boxl))
    finally:
        if d is not None:
            d.close()
    return out


def schedule_from_temp_files(root: str, names: Sequence[str]) -> List[List[float]]:
    """reads temp_0xx files (T*, ncoll) in the given order"""
    return [list(fileio.read_temp_file(os.path.join(root, n))) for n in names]
