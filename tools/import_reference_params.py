#!/usr/bin/env python
"""One-off importer (run in the build container, where /root/reference exists): turns the reference's
PRIME20 parameter DATA files into the JSON the package ships, and the shipped system-A snapshot
(genconfig/results/run0000.*) + genconfig/checks known answers into tests/golden fixtures.
Only data tables are imported -- no reference source code."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from parallel_dmd_for_biomolecules_b200 import fileio, tables  # noqa: E402

REF = "/root/reference/parallel-dmd-PRIME20"
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def main():
    t = tables.read_parameters_dir(REF)
    topo_b = tables.read_topology_dir(REF, [24, 24], [7, 7], [28, 28])
    out = {
        "source": "parallel-dmd-PRIME20/parameters/*.data, parametersep/ep19p_ha55a_weakhp.data",
        "tables": tables.tables_to_dict(t),
        "system_B": {"species": [dict(n_chains=s.n_chains, identity=s.identity, hp=s.hp, firstside=s.firstside)
                                 for s in topo_b.species], "boxl": 158.54},
    }
    os.makedirs(os.path.join(ROOT, "parallel_dmd_for_biomolecules_b200/data"), exist_ok=True)
    with open(os.path.join(ROOT, "parallel_dmd_for_biomolecules_b200/data/prime20_ha55a.json"), "w") as f:
        json.dump(out, f)
    # ---- system A fixtures
    g = os.path.join(ROOT, "tests/golden")
    os.makedirs(g, exist_ok=True)
    for name in ("run0000.config", "run0000.lastvel"):
        with open(os.path.join(REF, "genconfig/results", name), "rb") as src, open(os.path.join(g, "systemA_" + name), "wb") as dst:
            dst.write(src.read())
    ident = np.loadtxt(os.path.join(REF, "genconfig/checks/identity.out"), dtype=int)[:, 1]
    masses = np.loadtxt(os.path.join(REF, "genconfig/checks/masses.out"))
    sumvel = np.loadtxt(os.path.join(REF, "genconfig/checks/sumvelcheck.out"))
    np.savez_compressed(os.path.join(g, "systemA_checks.npz"), identity_with_gly=ident, masses=masses, sumvelcheck=sumvel)
    # genconfig inputs needed by the box generator (chain template of the 31-residue extended peptide)
    tpl = []
    for ax in "xyz":
        vals = tables._floats(os.path.join(REF, f"genconfig/inputs/peptide{ax}.inp"))
        tpl.append(vals[3:3 + 124])
    with open(os.path.join(ROOT, "parallel_dmd_for_biomolecules_b200/data/chain_template31.json"), "w") as f:
        json.dump({"source": "genconfig/inputs/peptide{x,y,z}.inp (124 beads: Ca,N,C,R x 31, Angstrom)", "xyz": tpl}, f)
    print("ok")


if __name__ == "__main__":
    main()
