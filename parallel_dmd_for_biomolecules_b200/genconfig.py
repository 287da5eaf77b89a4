"""Starting-configuration generator -- a restatement of what genconfig/gen_config_random-SQZ.f90 produces
(SURVEY.md 8f-2), with run-time sizes so the benchmark boxes of BASELINE.json can be synthesised.

Reference behaviour kept (paths relative to /root/reference/parallel-dmd-PRIME20/genconfig):
  * chain template = the first L residues of the 31-residue extended chain in inputs/peptide{x,y,z}.inp
    (gen_config_random-SQZ.f90:77-85), bead order Ca x L, N x L, C x L, side chains (Gly dropped, :354-365);
  * each side-chain bead is placed at the R-Ca / R-N / R-C distances of parameters/rcarnrco.data subject to
    the squeeze minimum distances of parameters/sqz6to10.data (:229-348).  The reference does this with a
    brute-force +-3.5 A grid search in 0.005 A steps; here it is closed-form trilateration followed by a
    small feasibility search inside the bond tolerance windows;
  * chains are placed at a uniformly random origin with the template orientation and rejected if any bead
    comes within 5 A of an earlier chain (:665-694);
  * momenta are drawn, the total momentum is removed and velocities are scaled so that sum m v^2 = 3 N (12 T*)
    exactly (:805-891).  numpy's generator replaces Intel drandm (not reproducible anyway).
"""
from __future__ import annotations

import json
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .tables import RESIDUES, Species, Tables, Topology

_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
DEL = 0.02375


def _template() -> np.ndarray:
    with open(os.path.join(_DATA_DIR, "chain_template31.json")) as f:
        xyz = np.asarray(json.load(f)["xyz"], dtype=np.float64)  # (3, 124)
    return xyz.T.copy()  # (124, 3): Ca 0..30, N 31..61, C 62..92, R 93..123


def _trilaterate(p1, p2, p3, r1, r2, r3, side_hint):
    """Point at distances r1,r2,r3 from p1,p2,p3 on the side of the p1-p2-p3 plane indicated by side_hint."""
    ex = (p2 - p1) / np.linalg.norm(p2 - p1)
    i = np.dot(ex, p3 - p1)
    ey = p3 - p1 - i * ex
    ey /= np.linalg.norm(ey)
    ez = np.cross(ex, ey)
    d = np.linalg.norm(p2 - p1)
    j = np.dot(ey, p3 - p1)
    x = (r1 * r1 - r2 * r2 + d * d) / (2 * d)
    y = (r1 * r1 - r3 * r3 + i * i + j * j) / (2 * j) - (i / j) * x
    z2 = r1 * r1 - x * x - y * y
    z = np.sqrt(max(z2, 0.0))
    if np.dot(ez, side_hint - p1) < 0:
        z = -z
    return p1 + x * ex + y * ey + z * ez, z2


def build_chain(seq: str, tables: Tables, rng: Optional[np.random.Generator] = None) -> np.ndarray:
    """Coordinates (Angstrom) of one extended chain in the reference's bead order (Gly side chains dropped)."""
    L = len(seq)
    if L > 31:
        raise ValueError("the reference's chain template has 31 residues")
    tpl = _template()
    ca, n, c, rt = tpl[0:31], tpl[31:62], tpl[62:93], tpl[93:124]
    rcar = np.asarray(tables.rcarnrco, dtype=np.float64).reshape(20, 6)
    sqz = np.asarray(tables.sqz6to10, dtype=np.float64).reshape(20, 5)  # file order sz8, sz6, sz7, sz9, sz10
    side = []
    rng = rng or np.random.default_rng(12345)
    for k, ch in enumerate(seq.upper()):
        rid = RESIDUES.index(ch)
        if rid == 0:
            continue  # glycine: no side-chain bead
        drca, drnh, drco = rcar[rid, 0:3]
        tol = np.maximum(rcar[rid, 3:6], DEL)

        def violations(pos, scale=1.0):
            v = 0.0
            sz8, sz6, sz7, sz9, sz10 = sqz[rid]
            if k >= 1:
                v += max(0.0, sz6 * scale - np.linalg.norm(pos - c[k - 1]))
                v += max(0.0, sz8 * scale - np.linalg.norm(pos - ca[k - 1]))
            if k + 1 < L:
                v += max(0.0, sz7 * scale - np.linalg.norm(pos - n[k + 1]))
                v += max(0.0, sz9 * scale - np.linalg.norm(pos - ca[k + 1]))
            if k >= 2:
                v += max(0.0, sz10 * scale - np.linalg.norm(pos - c[k - 2]))
            return v

        best, zz = _trilaterate(ca[k], n[k], c[k], drca, drnh, drco, rt[k])
        if zz < 0 or violations(best, 1.0005) > 0:
            # feasibility search inside the tolerance windows (what the reference's grid search achieves)
            cand_best, cand_score = None, 1e30
            for _ in range(4000):
                f = 1.0 + (rng.random(3) * 2 - 1) * tol * 0.9
                pos, z2 = _trilaterate(ca[k], n[k], c[k], drca * f[0], drnh * f[1], drco * f[2], rt[k])
                if z2 < 0:
                    continue
                score = violations(pos, 1.0005) * 100 + np.abs(f - 1).sum()
                if score < cand_score:
                    cand_best, cand_score = pos, score
            if cand_best is None or violations(cand_best, 1.0005) > 0:
                raise RuntimeError(f"no feasible side-chain position for residue {k + 1} ({ch})")
            best = cand_best
        side.append(best)
    pts = [ca[:L], n[:L], c[:L]]
    if side:
        pts.append(np.asarray(side))
    return np.concatenate(pts, axis=0)


def _random_rotation(rng) -> np.ndarray:
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    a, b, c, d = q
    return np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                     [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                     [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]])


def place_chains(chain_xyz: Sequence[np.ndarray], counts: Sequence[int], boxl: float, rng, min_dist=5.0,
                 rotate=False, max_tries=200000) -> np.ndarray:
    """gen_config_random-SQZ.f90:665-723 with a cell grid for the 5 A overlap test (so 1e6-bead boxes build)."""
    ncell = max(1, int(boxl // min_dist))
    w = boxl / ncell
    grid = {}
    placed: List[np.ndarray] = []

    def cells_of(p):
        return np.floor((p % boxl) / w).astype(int) % ncell

    def ok(pts):
        cc = cells_of(pts)
        for p, (cx, cy, cz) in zip(pts, cc):
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    for dz in (-1, 0, 1):
                        key = ((cx + dx) % ncell, (cy + dy) % ncell, (cz + dz) % ncell)
                        for q in grid.get(key, ()):
                            d = p - q
                            d -= boxl * np.round(d / boxl)
                            if d @ d * 1.0000000001 <= min_dist * min_dist:
                                return False
        return True

    out = []
    for xyz, cnt in zip(chain_xyz, counts):
        rel = xyz - xyz[0]
        for _ in range(cnt):
            for _try in range(max_tries):
                r = _random_rotation(rng) if rotate else np.eye(3)
                pts = rng.random(3) * boxl + rel @ r.T
                pts = pts - boxl * np.round(pts / boxl)
                if ok(pts):
                    break
            else:
                raise RuntimeError("could not place chain without overlap; box too dense")
            for p, key in zip(pts, map(tuple, cells_of(pts))):
                grid.setdefault(key, []).append(p)
            out.append(pts)
    return np.concatenate(out, axis=0)


def maxwell_velocities(masses: np.ndarray, tstar: float, rng) -> np.ndarray:
    """gen_config_random-SQZ.f90:805-891: zero total momentum, sum m v^2 = 3 N setemp exactly (setemp = 12 T*)."""
    setemp = 12.0 * tstar
    n = len(masses)
    p = rng.normal(size=(n, 3)) * np.sqrt(masses * setemp)[:, None]
    p -= p.mean(axis=0)
    sumvel = (p * p).sum(axis=1) / masses
    tred = sumvel.sum() / 3.0 / n
    const = np.sqrt(setemp / tred)
    return const * p / masses[:, None]


def masses_of(topo: Topology, tables: Tables) -> np.ndarray:
    bmass = np.asarray(tables.mass, dtype=np.float64).copy()
    bmass[2] = bmass[19]  # bmass(3) = bmass(20), inputinfo.f:400
    for i in range(4):
        bmass[i + 4] = bmass[i]
    return bmass[topo.bead_identity() - 1]


def generate_box(sequences: Sequence[str], counts: Sequence[int], boxl: float, tstar: float, tables: Tables,
                 seed: int = 1, rotate: bool = False) -> Tuple[Topology, np.ndarray]:
    """Returns (topology, sv) with sv of shape (N, 6) in BOX UNITS, i.e. what run0000.config / .lastvel hold."""
    if len(sequences) not in (1, 2):
        raise ValueError("one or two species")
    rng = np.random.default_rng(seed)
    species = [Species.from_sequence(s, c) for s, c in zip(sequences, counts)]
    topo = Topology(species)
    chains = [build_chain(s, tables, rng) for s in sequences]
    xyz = place_chains(chains, counts, boxl, rng, rotate=rotate)
    vel = maxwell_velocities(masses_of(topo, tables), tstar, rng)
    sv = np.concatenate([xyz / boxl, vel], axis=1)
    return topo, np.ascontiguousarray(sv)


def system_b(tables: Tables, tstar: float = 0.18, seed: int = 1, n_chains: int = 48, boxl: float = 158.54):
    """BASELINE config 2: 48 x Abeta16-22 (KLVFFAE) as two identical species of 24 chains, L = 158.54 A
    (qfile/script.sh:7, parameters/identity.inp, code/inputinfo.f:78)."""
    h = n_chains // 2
    return generate_box(["KLVFFAE", "KLVFFAE"], [h, n_chains - h], boxl, tstar, tables, seed)
