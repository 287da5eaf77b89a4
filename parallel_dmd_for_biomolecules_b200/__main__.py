"""``python -m parallel_dmd_for_biomolecules_b200 [options] < temp_018`` -- the reference's ``./dmd < temp_0xx``.

Reads T* and the number of collisions from stdin (main.F90:126-128), the parameter files from ``<root>/parameters``
and ``<root>/parametersep``, restarts from the last ``<root>/results/runNNNN.*`` and writes the next run's files
(driver.py).  What the reference fixes at compile time with ``-Dnop1 -Dnop2 -Dchnln1 -Dchnln2 -Dnumbeads1
-Dnumbeads2`` (qfile/script.sh:7) and hard-codes in inputinfo.f:78 are options here."""
from __future__ import annotations

import argparse
import json
import sys

from . import driver, tables


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m parallel_dmd_for_biomolecules_b200", description=__doc__)
    ap.add_argument("--root", default=".", help="directory holding parameters/, parametersep/ and results/")
    ap.add_argument("--nop1", type=int, default=672, help="beads of species 1 (-Dnop1)")
    ap.add_argument("--nop2", type=int, default=672, help="beads of species 2 (-Dnop2); 0 = one species")
    ap.add_argument("--chnln1", type=int, default=7)
    ap.add_argument("--chnln2", type=int, default=7)
    ap.add_argument("--numbeads1", type=int, default=28)
    ap.add_argument("--numbeads2", type=int, default=28)
    ap.add_argument("--boxl", type=float, default=158.540, help="box length in Angstrom (inputinfo.f:78)")
    ap.add_argument("--nve", action="store_true", help="build without -Dcanon (no Andersen thermostat)")
    ap.add_argument("--engine", type=int, default=0, choices=[0, 1, 2, 3])
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    vals = []
    for line in sys.stdin:
        tok = line.split("#")[0].split()
        if tok:
            vals.append(tok[0].replace("D", "E").replace("d", "e"))
    if len(vals) < 2:
        print("expected T* and the number of collisions on stdin (a temp_0xx file)", file=sys.stderr)
        return 2
    tstar, ncoll = float(vals[0]), int(float(vals[1]))
    n_chains = [args.nop1 // args.numbeads1] + ([args.nop2 // args.numbeads2] if args.nop2 else [])
    chnln = [args.chnln1] + ([args.chnln2] if args.nop2 else [])
    numbeads = [args.numbeads1] + ([args.numbeads2] if args.nop2 else [])
    tab = tables.read_parameters_dir(args.root)
    topo = tables.read_topology_dir(args.root, n_chains, chnln, numbeads)
    print("noptotal", topo.n_beads)
    print("number of collisions requested", ncoll)
    res = driver.run_temperature(args.root, topo, tab, tstar, ncoll, boxl=args.boxl, canon=not args.nve,
                                 engine=args.engine, device=args.device)
    print(json.dumps(res))
    return 0


if __name__ == "__main__":
    sys.exit(main())
