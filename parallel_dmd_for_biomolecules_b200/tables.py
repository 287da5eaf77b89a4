"""ctypes mirrors of include/dmdb200.h plus loaders for the reference's parameter files.

The file formats parsed here are the reference's own (paths relative to
/root/reference/parallel-dmd-PRIME20/): ``parameters/protein.data`` (code/inputinfo.f:162-187),
``parametersep/ep19p_ha55a_weakhp.data`` (inputinfo.f:282-288), ``parameters/beadwell_ha55a.data``
(inputinfo.f:361-369), ``parameters/rcarnrco.data`` (:291-299), ``parameters/sqz6to10.data`` (:379-389),
``parameters/mass.data`` (:392-404), ``parameters/identity.inp`` (:209-228), ``hp1.inp``/``hp2.inp``
(:236-251), ``firstside1.data``/``firstside2.data`` (:105-132).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import re
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

DMDB_MAX_SPECIES = 2

#: one-letter residue codes in the reference's identity order, ids 9..28 (inputinfo.f:201)
RESIDUES = "GRNDQEHKPSTACILMFWYV"


class Tables(C.Structure):
    _fields_ = [
        ("protein", C.c_double * 12),
        ("ep", C.c_double * 400),
        ("bds", C.c_double * 400),
        ("wel", C.c_double * 400),
        ("mass", C.c_double * 28),
        ("rcarnrco", C.c_double * 120),
        ("sqz6to10", C.c_double * 100),
    ]


class TopologyC(C.Structure):
    _fields_ = [
        ("n_species", C.c_int32),
        ("n_chains", C.c_int32 * DMDB_MAX_SPECIES),
        ("chnln", C.c_int32 * DMDB_MAX_SPECIES),
        ("numbeads", C.c_int32 * DMDB_MAX_SPECIES),
        ("identity", C.POINTER(C.c_int32) * DMDB_MAX_SPECIES),
        ("hp", C.POINTER(C.c_int32) * DMDB_MAX_SPECIES),
        ("firstside", C.POINTER(C.c_int32) * DMDB_MAX_SPECIES),
    ]


class Params(C.Structure):
    _fields_ = [
        ("boxl", C.c_double),
        ("tstar", C.c_double),
        ("canon", C.c_int32),
        ("no_hbs", C.c_int32),
        ("n_wrap", C.c_int32),
        ("n_replicas", C.c_int32),
        ("device", C.c_int32),
        ("nbr_capacity", C.c_int32),
        ("log_capacity", C.c_int32),
        ("engine", C.c_int32),
        ("seed", C.c_uint64),
    ]


class Event(C.Structure):
    _fields_ = [("t", C.c_double), ("i", C.c_int32), ("j", C.c_int32), ("type", C.c_int32), ("evcode", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [
        ("events", C.c_int64),
        ("pair_events", C.c_int64),
        ("nevents", C.c_int64 * 32),
        ("ghosts", C.c_int64),
        ("updates", C.c_int64),
        ("forced_updates", C.c_int64),
        ("pair_predictions", C.c_int64),
        ("nbr_visits", C.c_int64),
        ("device_ms", C.c_double),
        ("kernel_launches", C.c_int32),
        ("reserved", C.c_int32),
    ]


class Energy(C.Structure):
    _fields_ = [
        ("ered", C.c_double),
        ("tred", C.c_double),
        ("sumvel", C.c_double),
        ("ehh_ii", C.c_double),
        ("ehh_ij", C.c_double),
        ("hb_alpha", C.c_int32),
        ("hb_ii", C.c_int32),
        ("hb_ij", C.c_int32),
        ("reserved", C.c_int32),
    ]


EVENT_DTYPE = np.dtype([("t", "<f8"), ("i", "<i4"), ("j", "<i4"), ("type", "<i4"), ("evcode", "<i4")])


@dataclass
class Species:
    """One peptide species: what -Dchnln/-Dnumbeads + identity.inp + hp.inp + firstside.data describe."""

    n_chains: int
    identity: List[int]  # numbeads ids: Ca x L (2), N x L (1), C x L (4), then side-chain ids (Gly omitted)
    hp: List[int]
    firstside: List[int]  # chnln flags

    @property
    def chnln(self) -> int:
        return len(self.firstside)

    @property
    def numbeads(self) -> int:
        return len(self.identity)

    @staticmethod
    def from_sequence(seq: str, n_chains: int, hydrophobic_all: bool = True) -> "Species":
        """Build a species from one-letter codes the way genconfig writes identity.inp / hp.inp /
        firstside.data (genconfig/gen_config_random-SQZ.f90:138-185): every non-Gly side chain gets hp=1."""
        L = len(seq)
        ids = [2] * L + [1] * L + [4] * L
        first = []
        for ch in seq.upper():
            rid = RESIDUES.index(ch) + 9
            if rid == 9:
                first.append(0)
            else:
                first.append(1)
                ids.append(rid)
        hp = [0] * (3 * L) + [1 if hydrophobic_all else 0] * (len(ids) - 3 * L)
        return Species(n_chains, ids, hp, first)

    def sequence(self) -> str:
        side = iter(self.identity[3 * self.chnln:])
        return "".join(RESIDUES[next(side) - 9] if f else "G" for f in self.firstside)


@dataclass
class Topology:
    species: List[Species]
    _keep: list = field(default_factory=list, repr=False)

    @property
    def n_beads(self) -> int:
        return sum(s.n_chains * s.numbeads for s in self.species)

    def bead_identity(self) -> np.ndarray:
        out = []
        for s in self.species:
            out += list(s.identity) * s.n_chains
        return np.asarray(out, dtype=np.int32)

    def bead_chain(self) -> np.ndarray:
        out, c = [], 0
        for s in self.species:
            for _ in range(s.n_chains):
                out += [c] * s.numbeads
                c += 1
        return np.asarray(out, dtype=np.int32)

    def to_c(self) -> TopologyC:
        t = TopologyC()
        t.n_species = len(self.species)
        self._keep.clear()
        for k, s in enumerate(self.species):
            t.n_chains[k] = s.n_chains
            t.chnln[k] = s.chnln
            t.numbeads[k] = s.numbeads
            for name, vals in (("identity", s.identity), ("hp", s.hp), ("firstside", s.firstside)):
                arr = (C.c_int32 * len(vals))(*vals)
                self._keep.append(arr)
                getattr(t, name)[k] = C.cast(arr, C.POINTER(C.c_int32))
        return t


_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _floats(path: str) -> List[float]:
    out = []
    with open(path) as f:
        for line in f:
            line = line.split("#")[0]
            for tok in line.replace(",", " ").split():
                try:
                    out.append(float(tok.replace("D", "E").replace("d", "e")))
                except ValueError:
                    pass
    return out


def _ints(path: str) -> List[int]:
    return [int(round(v)) for v in _floats(path)]


def read_parameters_dir(root: str, ep_file: str = "parametersep/ep19p_ha55a_weakhp.data") -> Tables:
    """Parse the reference's ``parameters/`` + ``parametersep/`` files found under ``root``."""
    t = Tables()
    prot = _floats(os.path.join(root, "parameters/protein.data"))
    for k in range(12):
        t.protein[k] = prot[k]
    pair_re = re.compile(r"^..(..)....(..)..(.*)$")  # (2(2x,i2,2x), ...)

    def pair_table(path, ncols):
        cols = [np.zeros(400) for _ in range(ncols)]
        with open(path) as f:
            for line in f:
                line = line.rstrip("\r\n")
                if not line.strip():
                    continue
                m = pair_re.match(line)
                i, j = int(m.group(1)), int(m.group(2))
                vals = m.group(3).split()
                for c in range(ncols):
                    cols[c][(i - 9) * 20 + (j - 9)] = float(vals[c])
        return cols

    (ep,) = pair_table(os.path.join(root, ep_file), 1)
    bds, wel = pair_table(os.path.join(root, "parameters/beadwell_ha55a.data"), 2)
    for k in range(400):
        t.ep[k], t.bds[k], t.wel[k] = ep[k], bds[k], wel[k]
    with open(os.path.join(root, "parameters/mass.data")) as f:
        for line in f:
            line = line.rstrip("\r\n")
            if len(line) < 16:
                continue
            ident = int(line[4:6])  # (4x,i2,2x,f8.3)
            t.mass[ident - 1] = float(line[8:16])
    rows = [l for l in open(os.path.join(root, "parameters/rcarnrco.data")).read().splitlines() if l.strip()]
    for r in range(20):
        vals = rows[r].split()
        for c in range(6):
            t.rcarnrco[r * 6 + c] = float(vals[c])
    rows = [l for l in open(os.path.join(root, "parameters/sqz6to10.data")).read().splitlines() if l.strip()]
    for r in range(20):
        vals = rows[r].split()
        for c in range(5):
            t.sqz6to10[r * 5 + c] = float(vals[c])
    return t


def read_topology_dir(root: str, n_chains: Sequence[int], chnln: Sequence[int], numbeads: Sequence[int]) -> Topology:
    """identity.inp / hp1.inp / hp2.inp / firstside{1,2}.data with the -D sizes given explicitly."""
    ids = _ints(os.path.join(root, "parameters/identity.inp"))
    species, off = [], 0
    for k in range(len(n_chains)):
        ident = ids[off:off + numbeads[k]]
        off += numbeads[k]
        hp = _ints(os.path.join(root, f"parameters/hp{k + 1}.inp"))[: numbeads[k]]
        first = _ints(os.path.join(root, f"parameters/firstside{k + 1}.data"))[: chnln[k]]
        species.append(Species(n_chains[k], ident, hp, first))
    return Topology(species)


def tables_to_dict(t: Tables) -> dict:
    return {name: [float(v) for v in getattr(t, name)] for name, _ in Tables._fields_}


def tables_from_dict(d: dict) -> Tables:
    t = Tables()
    for name, _ in Tables._fields_:
        arr = getattr(t, name)
        for k, v in enumerate(d[name]):
            arr[k] = v
    return t


def load_default_tables() -> Tables:
    """The PRIME20 parameter set the reference ships (ha55a bead/well table, ep19p weak-HP energies),
    imported once from the reference's data files by tools/import_reference_params.py."""
    with open(os.path.join(_DATA_DIR, "prime20_ha55a.json")) as f:
        return tables_from_dict(json.load(f)["tables"])


def make_params(boxl=158.54, tstar=0.18, canon=True, no_hbs=False, n_wrap=2, n_replicas=1, device=0,
                nbr_capacity=0, log_capacity=0, seed=1058472402, engine=0) -> Params:
    p = Params()
    p.boxl, p.tstar = boxl, tstar
    p.canon, p.no_hbs, p.n_wrap = int(canon), int(no_hbs), n_wrap
    p.n_replicas, p.device = n_replicas, device
    p.nbr_capacity, p.log_capacity, p.engine = nbr_capacity, log_capacity, engine
    p.seed = seed
    return p
