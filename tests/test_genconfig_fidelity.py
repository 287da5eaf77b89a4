"""genconfig restatement against the reference generator's own output.

The shipped snapshot genconfig/results/run0000.config (tests/golden/systemA_run0000.config) holds 8 chains of
GVAYVGSKTKEGVVHGVATVAE that gen_config_random-SQZ.f90 built from its 31-residue template and placed WITHOUT rotation.
Its side-chain beads come from a brute-force search on a 0.005 A grid (gen_config_random-SQZ.f90:250-309);
genconfig.build_chain places them in closed form (trilateration + the same feasibility rule).  After a rigid fit of
the backbone, every bead of our chain must sit within 2.5 grid steps (0.0125 A) of the generator's (its L1 objective is
flat near the optimum, so the grid optimum may sit a cell or two from the exact solution; measured: <= 0.0104 A), and
every bonded side-chain distance within one grid step (0.005 A).
The generator read its OWN copy of rcarnrco.data, whose His row has R-Ca = 3.160 A instead of the 3.150 A of
parameters/rcarnrco.data: the comparison uses that value."""
import copy
import os

import numpy as np

from conftest import GOLDEN, SEQ_A
from parallel_dmd_for_biomolecules_b200 import fileio, genconfig

L_BOX, NB, NRES = 110.0, 84, len(SEQ_A)
HIS_ROW = 6  # rows of rcarnrco.data: G R N D Q E H K P S T A C I L M F W Y V


def _unwrap_chain(x):
    """box units, beads wrapped one by one -> a connected chain (each bead next to a bonded neighbour)"""
    y = x.copy()
    for k in range(1, NRES):
        d = x[k] - x[k - 1]
        y[k] = y[k - 1] + d - np.round(d)
    side = 0
    for k in range(NRES):
        for off in (NRES, 2 * NRES):
            d = x[off + k] - x[k]
            y[off + k] = y[k] + d - np.round(d)
        if SEQ_A[k] != "G":
            j = 3 * NRES + side
            side += 1
            d = x[j] - x[k]
            y[j] = y[k] + d - np.round(d)
    return y


def _fit(ours, ref, nfit):
    """proper rigid motion (Kabsch) that takes ours[:nfit] onto ref[:nfit], applied to all of ours"""
    p0, q0 = ours[:nfit].mean(0), ref[:nfit].mean(0)
    u, _, vt = np.linalg.svd((ours[:nfit] - p0).T @ (ref[:nfit] - q0))
    rot = u @ np.diag([1.0, 1.0, np.sign(np.linalg.det(u @ vt))]) @ vt
    return (ours - p0) @ rot + q0


def test_build_chain_reproduces_the_shipped_chains(tab):
    t = copy.deepcopy(tab)
    rc = np.asarray(t.rcarnrco, dtype=np.float64).reshape(20, 6).copy()
    assert rc[HIS_ROW, 0] == 3.150
    rc[HIS_ROW, 0] = 3.160  # genconfig/parameters/rcarnrco.data
    for k, v in enumerate(rc.reshape(-1)):
        t.rcarnrco[k] = v
    ours = genconfig.build_chain(SEQ_A, t)
    assert ours.shape == (NB, 3)
    sv = fileio.sv_from_files(os.path.join(GOLDEN, "systemA_run0000.config"), os.path.join(GOLDEN, "systemA_run0000.lastvel"))
    bb = 3 * NRES
    for c in range(8):
        ref = _unwrap_chain(sv[c * NB:(c + 1) * NB, :3]) * L_BOX
        fit = _fit(ours, ref, bb)
        disp = np.linalg.norm(fit - ref, axis=1)
        assert disp[:bb].max() < 2e-3          # backbone = the template (the file keeps ~1e-3 A)
        assert disp[bb:].max() < 0.0125, (c, disp[bb:].max())
        # bonded side-chain distances R-Ca, R-N, R-C: within the grid step of the generator's
        side = 0
        for k in range(NRES):
            if SEQ_A[k] == "G":
                continue
            j = bb + side
            side += 1
            for other in (k, NRES + k, 2 * NRES + k):
                a, b = np.linalg.norm(ref[j] - ref[other]), np.linalg.norm(ours[j] - ours[other])
                assert abs(a - b) < 0.005, (c, k, a, b)
    # all intra-chain distances of chain 1
    ref = _unwrap_chain(sv[:NB, :3]) * L_BOX
    d_ref = np.linalg.norm(ref[:, None] - ref[None], axis=2)
    d_our = np.linalg.norm(ours[:, None] - ours[None], axis=2)
    assert np.abs(d_ref - d_our).max() < 2 * 0.0125
