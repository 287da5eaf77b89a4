"""β-sheet / fibril observables from a downloaded state, the definitions of the reference's
post-processing program ``results/r/fibril_list_assign.f`` (single species, as there):

* ``hb_contact(a,b)`` -- inter-chain backbone H-bonds between peptides a and b, counted from ``bptnr`` over the N and
  C beads of each chain except the first N and the last C (``:51-60``);
* ``hp_contact(a,b)`` -- side-chain pairs of a and b closer than their well diameter (``:61-78``);
* peptides a, b are SHEET partners (``dimer``) when ``hb_contact >= chnln/2 + 1`` (``:89``); a sheet is a connected
  component of that relation (``sheet_assign``);
* a, b are a hydrophobic STACK (``hp_dimer``) when ``hp_contact >= chnln/2`` and they share no H-bond (``:79-82``);
  two sheets joined by such a stack belong to one fibril (``fibril_assign``).

Host-side analysis in numpy: it reads ``bptnr`` and positions through the C ABI (``DMD.state``), nothing here is on
the hot path.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from .tables import Tables, Topology


def _components(n: int, adj: np.ndarray) -> List[List[int]]:
    seen, out = np.zeros(n, dtype=bool), []
    for s in range(n):
        if seen[s]:
            continue
        comp, stack = [], [s]
        seen[s] = True
        while stack:
            a = stack.pop()
            comp.append(a)
            for b in np.nonzero(adj[a])[0]:
                if not seen[b]:
                    seen[b] = True
                    stack.append(int(b))
        out.append(sorted(comp))
    return out


def contacts(topo: Topology, tables: Tables, xyz: np.ndarray, bptnr: np.ndarray, boxl: float):
    """(hb_contact, hp_contact) peptide x peptide matrices.  xyz: (N,3) box units; bptnr: 1-based partner or 0."""
    if any(sp.identity != topo.species[0].identity for sp in topo.species):
        raise ValueError("fibril_list_assign.f handles one peptide species")
    L, nbd = topo.species[0].chnln, topo.species[0].numbeads
    nc = sum(sp.n_chains for sp in topo.species)
    chain = np.arange(nc * nbd) // nbd
    hb = np.zeros((nc, nc), dtype=np.int32)
    local = np.arange(nc * nbd) % nbd
    inner = (local >= L + 1) & (local <= 3 * L - 2)  # N beads 2..L and C beads 1..L-1 (1-based L+2 .. 3L-1)
    for aa in np.nonzero(inner & (bptnr > 0))[0]:
        bb = int(bptnr[aa]) - 1
        if bb > aa and inner[bb] and chain[bb] != chain[aa]:
            hb[chain[aa], chain[bb]] += 1
            hb[chain[bb], chain[aa]] += 1
    ident = topo.bead_identity()
    sc = np.nonzero(local >= 3 * L)[0]
    wel = np.array(tables.wel).reshape(20, 20)
    d = xyz[sc][:, None, :] - xyz[sc][None, :, :]
    d -= np.round(d)
    dist = np.sqrt((d ** 2).sum(-1)) * boxl
    w = wel[ident[sc][:, None] - 9, ident[sc][None, :] - 9]
    close = (dist <= w) & (chain[sc][:, None] != chain[sc][None, :])
    hp = np.zeros((nc, nc), dtype=np.int32)
    np.add.at(hp, (chain[sc][:, None].repeat(len(sc), 1)[close], chain[sc][None, :].repeat(len(sc), 0)[close]), 1)
    return hb, hp


def sheets_and_fibrils(topo: Topology, tables: Tables, xyz: np.ndarray, bptnr: np.ndarray, boxl: float) -> Dict:
    L = topo.species[0].chnln
    hb, hp = contacts(topo, tables, xyz, bptnr, boxl)
    nc = hb.shape[0]
    dimer = hb >= L // 2 + 1
    hp_dimer = (hp >= L // 2) & (hb == 0)
    np.fill_diagonal(hp_dimer, False)
    sheets = [s for s in _components(nc, dimer) if len(s) > 1]
    sheet_of = -np.ones(nc, dtype=int)
    for k, s in enumerate(sheets):
        sheet_of[s] = k
    ns = len(sheets)
    sadj = np.zeros((ns, ns), dtype=bool)
    for a, b in zip(*np.nonzero(hp_dimer)):
        if sheet_of[a] >= 0 and sheet_of[b] >= 0 and sheet_of[a] != sheet_of[b]:
            sadj[sheet_of[a], sheet_of[b]] = True
    fibrils = [f for f in _components(ns, sadj) if len(f) > 1]
    return dict(hb_contact=hb, hp_contact=hp, sheets=sheets, fibrils=[[sheets[k] for k in f] for f in fibrils],
                peptides_in_sheets=int((sheet_of >= 0).sum()), largest_sheet=max([len(s) for s in sheets], default=0))
