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
