"""Reference run-file formats: ifort unformatted sequential records with 4-byte
little-endian length markers, as written by code/config.f:16-40 and read by code/inputinfo.f:79-101 and
code/main.F90:241-246; the ``.energy`` text line of main.F90:1210,1380; ``temp_0xx`` stdin files
(main.F90:126-128)."""
from __future__ import annotations

import struct
from typing import Iterator, List, Optional, Tuple

import numpy as np


def _records(path: str) -> Iterator[bytes]:
    with open(path, "rb") as f:
        data = f.read()
    off = 0
    while off + 4 <= len(data):
        (n,) = struct.unpack_from("<i", data, off)
        if n < 0 or off + 8 + n > len(data):
            break  # a record cut short (run killed inside config()): keep what is complete, like `read(7,end=120)`
        body = data[off + 4: off + 4 + n]
        (n2,) = struct.unpack_from("<i", data, off + 4 + n)
        if n2 != n:
            raise ValueError(f"{path}: corrupt record marker at {off}")
        yield body
        off += 8 + n


def _write_record(f, body: bytes) -> None:
    f.write(struct.pack("<i", len(body)))
    f.write(body)
    f.write(struct.pack("<i", len(body)))


def read_config(path: str, last: bool = True):
    """``runNNNN.config``: records ``coll(int64) t(f64) x[N] y[N] z[N]`` (config.f:22-24).
    Returns (coll, t, xyz[3,N]) of the last record (the restart rule of inputinfo.f:81-85) or a list."""
    out = []
    for body in _records(path):
        coll, t = struct.unpack_from("<qd", body, 0)
        xyz = np.frombuffer(body, dtype="<f8", offset=16).copy()
        out.append((coll, t, xyz.reshape(3, -1)))
    if not out:
        raise ValueError(f"{path}: no records")
    return out[-1] if last else out


def read_lastvel(path: str):
    """``runNNNN.lastvel``: one record ``coll vx[N] vy[N] vz[N]`` (config.f:30-40)."""
    body = next(_records(path))
    (coll,) = struct.unpack_from("<q", body, 0)
    v = np.frombuffer(body, dtype="<f8", offset=8).copy()
    return coll, v.reshape(3, -1)


def read_bptnr(path: str, n_beads: int) -> np.ndarray:
    """``runNNNN.bptnr``: records ``coll bptnr[N](int32)``; an empty file leaves all zeros
    (gen_config_random-SQZ.f90:934-938, main.F90:243-246)."""
    bp = np.zeros(n_beads, dtype=np.int32)
    for body in _records(path):
        bp = np.frombuffer(body, dtype="<i4", offset=8).copy()
    return bp


def append_config(path: str, coll: int, t: float, xyz: np.ndarray) -> None:
    with open(path, "ab") as f:
        _write_record(f, struct.pack("<qd", coll, t) + np.ascontiguousarray(xyz, dtype="<f8").tobytes())


def append_bptnr(path: str, coll: int, bptnr: np.ndarray) -> None:
    with open(path, "ab") as f:
        _write_record(f, struct.pack("<q", coll) + np.ascontiguousarray(bptnr, dtype="<i4").tobytes())


def write_lastvel(path: str, coll: int, vel: np.ndarray) -> None:
    with open(path, "wb") as f:
        _write_record(f, struct.pack("<q", coll) + np.ascontiguousarray(vel, dtype="<f8").tobytes())


def energy_line(coll, tstar_time, ered, tred, hb_alpha, hb_ii, hb_ij, ehh_ii, ehh_ij, rg, e2e) -> str:
    """format 22223: (i15,3f12.4,3i8,4f12.4), main.F90:1380."""
    return "%15d%12.4f%12.4f%12.4f%8d%8d%8d%12.4f%12.4f%12.4f%12.4f" % (
        coll, tstar_time, ered, tred, hb_alpha, hb_ii, hb_ij, ehh_ii, ehh_ij, rg, e2e)


def read_temp_file(path: str) -> Tuple[float, int]:
    """``temp_0xx``: line 1 T* (``0.180D0``), line 2 ncoll; trailing ``# comment`` tolerated."""
    vals = []
    with open(path) as f:
        for line in f:
            tok = line.split("#")[0].split()
            if tok:
                vals.append(tok[0].replace("D", "E").replace("d", "e"))
    return float(vals[0]), int(float(vals[1]))


def sv_from_files(config_path: str, lastvel_path: str) -> np.ndarray:
    """6 x N column-major state (returned as an (N,6) C array == Fortran sv(6,N))."""
    _, _, xyz = read_config(config_path)
    _, v = read_lastvel(lastvel_path)
    return np.ascontiguousarray(np.concatenate([xyz, v], axis=0).T)


# ---- final PDB (write_rasmol-YM.f:17-159) and the side-chain bond audit (inputinfo.f:413-505) ---------------------------
_AMINO = {10: "ARG", 11: "ASN", 12: "ASP", 13: "GLN", 14: "GLU", 15: "HIS", 16: "LYS", 17: "PRO", 18: "SER", 19: "THR",
          20: "ALA", 21: "CYS", 22: "ILE", 23: "LEU", 24: "MET", 25: "PHE", 26: "TRP", 27: "TYR", 28: "VAL"}


def _fortran_i(v: int, w: int) -> str:
    s = "%d" % v
    return s.rjust(w) if len(s) <= w else "*" * w  # an integer that does not fit its field prints as asterisks


def _fortran_f(v: float, w: int, d: int) -> str:
    s = "%.*f" % (d, v)
    if len(s) > w and s.startswith("0."):
        s = s[1:]
    elif len(s) > w and s.startswith("-0."):
        s = "-" + s[2:]
    return s.rjust(w) if len(s) <= w else "*" * w


def pdb_lines(topo, xyz_angstrom: np.ndarray) -> List[str]:
    """write_rasmol-YM.f: per residue N, CA, C, a carbonyl O built 1.231 A from C along the bisector direction
    (:72-86), and CB = the side-chain bead (none for Gly); serial numbers run over numbeads + chnln atoms per chain,
    chain letter A for species 1 and B for species 2, format 7 = (A4,3X,I4,1X,A4,1X,A3,1X,A1,I4,4X,3F8.3).
    xyz_angstrom: (N,3) wrapped true positions after scale_up.f."""
    out = []
    x = np.asarray(xyz_angstrom, dtype=np.float64)
    first, serial0 = 0, 1
    for s, sp in enumerate(topo.species):
        L, nbd = sp.chnln, sp.numbeads
        side_ids = iter(sp.identity[3 * L:])
        names = [_AMINO[next(side_ids)] if f else "GLY" for f in sp.firstside]
        col = "AB"[s]
        for c in range(sp.n_chains):
            n0 = first + c * nbd  # 0-based index of the chain's first bead (its first Ca)
            m = serial0 + c * (nbd + L)
            p = 0
            for j in range(1, L + 1):
                n = n0 + j - 1
                ca, nn, cc = x[n], x[n + L], x[n + 2 * L]
                d = cc - (ca + x[n + 1 + L]) / 2.0 if j != L else cc - ca
                r1 = float(np.sqrt(d[0] ** 2 + d[1] ** 2 + d[2] ** 2))
                o = cc + 1.231 / r1 * d
                atoms = [(" N  ", nn), (" CA ", ca), (" C  ", cc), (" O  ", o)]
                if sp.firstside[j - 1]:
                    atoms.append((" CB ", x[n - p + 3 * L]))
                else:
                    p += 1
                for k, (an, pos) in enumerate(atoms):
                    out.append("ATOM   %s %s %s %s%s    %s%s%s" % (_fortran_i(m + k, 4), an, names[j - 1], col, _fortran_i(j, 4),
                                                                   _fortran_f(pos[0], 8, 3), _fortran_f(pos[1], 8, 3),
                                                                   _fortran_f(pos[2], 8, 3)))
                m += len(atoms)
        first += sp.n_chains * nbd
        serial0 += sp.n_chains * (nbd + L)
    return out


def write_pdb(path: str, topo, xyz_angstrom: np.ndarray) -> None:
    with open(path, "w") as f:
        f.write("\n".join(pdb_lines(topo, xyz_angstrom)) + "\n")


def rca_lines(topo, tables, xyz_box: np.ndarray, boxl: float, delta: float = 0.02375) -> List[str]:
    """inputinfo.f:413-505: for every side chain its distances to Ca, N and C of its residue next to the nominal
    lengths of rcarnrco.data (format 7373 = a4,2(i4,1x),i2,1x,2(f7.4,1x)), and a list-directed ' overlap' line when a
    distance is outside length*(1 -+ tolerance).  xyz_box: (N,3) positions in box units (minimum image is applied).
    Reference quirks kept: l is the 1-based index of the chain's first bead WITHIN its species, the integer column
    prints aa(k) -- the identity of bead k of the species-1 template, i.e. 2 (a Ca) for every residue."""
    out = []
    x = np.asarray(xyz_box, dtype=np.float64)
    rc = np.asarray(tables.rcarnrco, dtype=np.float64).reshape(20, 6)
    first = 0
    for sp in topo.species:
        L, nbd = sp.chnln, sp.numbeads
        side_ids = iter(sp.identity[3 * L:])
        rows, l_side = [], 0
        for f in sp.firstside:
            if f:
                l_side += 1
                rows.append((3 * L + l_side, int(next(side_ids)) - 8))
            else:
                rows.append((0, 1))
        aa = topo.species[0].identity
        for c in range(sp.n_chains):
            l1 = c * nbd + 1
            base = first + c * nbd
            for k in range(1, L + 1):
                fs, iii = rows[k - 1]
                if not fs:
                    continue
                side = x[base + fs - 1]
                for tag, other, col in (("rca ", base + k - 1, 0), ("rnh ", base + k + L - 1, 1), ("rco ", base + k + 2 * L - 1, 2)):
                    d = side - x[other]
                    d = d - np.round(d)
                    dist = float(np.sqrt((d * d).sum())) * boxl
                    nominal = rc[iii - 1, col]
                    tol = max(rc[iii - 1, 3 + col], delta) * nominal
                    out.append("%s%s %s %s %s %s " % (tag, _fortran_i(l1, 4), _fortran_i(k, 4), _fortran_i(int(aa[k - 1]) if k - 1 < len(aa) else 0, 2),
                                                       _fortran_f(dist, 7, 4), _fortran_f(nominal, 7, 4)))
                    if dist < nominal - tol or dist > nominal + tol:
                        out.append(" overlap")
        first += sp.n_chains * nbd
    return out


def write_rca(path: str, topo, tables, xyz_box: np.ndarray, boxl: float) -> None:
    with open(path, "w") as f:
        f.write("\n".join(rca_lines(topo, tables, xyz_box, boxl)) + "\n")
