#!/usr/bin/env python
"""SASS listing of the event-loop kernel annotated with executed warp instructions PER EVENT (ncu source page of one
captured launch joined with nvdisasm -g line info), plus a per-function instruction histogram (static footprint and
executed count).  usage: sass_hist.py report.ncu-rep lib.so kernel_regex events out.sass [min_exec_per_event]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def funcs_of(path):
    out = []
    for i, t in enumerate(open(path), 1):
        m = re.match(r"^(?:template <.*>\s*)?(?:DMD_(?:DEV|COLD)|__device__ __noinline__|__global__)\s+[\w:<>\s\*&()]*?\s+(\w+)\(", t)
        if m:
            out.append((i, m.group(1)))
    return out


def main():
    rep, lib, kern, events, outp = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), sys.argv[5]
    thr = float(sys.argv[6]) if len(sys.argv) > 6 else 0.02
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    ix = {h: i for i, h in enumerate(rows[1])}
    data = [r for r in rows[2:] if len(r) > ix["Instructions Executed"]]
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
    insts, in_k, cur = [], False, ("?", 0, ())
    for line in dis:
        if line.startswith("//---") and ".text." in line:
            in_k = bool(re.search(kern, line))
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
        if m:
            inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            cur = (os.path.basename(m.group(1)), int(m.group(2)), tuple((os.path.basename(f), int(l)) for f, l in inl))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            insts.append(cur)
    n = min(len(data), len(insts))
    csrc = os.path.join(ROOT, "parallel_dmd_for_biomolecules_b200", "csrc")
    ftab = {f: funcs_of(os.path.join(csrc, f)) for f in ("dmd_engine.h", "dmd_lockstep.h", "dmd_physics.h", "dmd_cuda.cu", "dmd_topology.h", "dmd_warp.h", "dmd_math.h")}

    def fn(f, l):
        name = None
        for ln, nm in ftab.get(f, []):
            if ln <= l:
                name = nm
        return name

    def owner(c):
        frames = [(c[0], c[1])] + list(c[2])
        for f, l in frames:  # innermost frame that belongs to one of our sources
            nm = fn(f, l)
            if nm:
                return "%s:%s" % (f.replace("dmd_", "").split(".")[0], nm)
        return frames[0][0]

    stat, ex, hot = Counter(), Counter(), Counter()
    lines = []
    tot = 0.0
    for k in range(n):
        e = int(data[k][ix["Instructions Executed"]] or 0) / events
        o = owner(insts[k])
        stat[o] += 1
        ex[o] += e
        tot += e
        if e >= 0.25:
            hot[o] += 1
        if e >= thr:
            lines.append("%6d  %8.3f  %-62s %s:%d  [%s]" % (k, e, data[k][ix["Source"]].strip()[:62], insts[k][0], insts[k][1], o))
    with open(outp, "w") as f:
        f.write("# dmd_event_loop_kernel, sm_100a: SASS annotated with executed warp instructions per event (%s, %d events)\n" % (os.path.basename(rep), events))
        f.write("# total %.1f warp instructions per event over the whole grid (event-loop CTAs + list-rebuild service CTAs incl. idle polling)\n" % tot)
        f.write("#\n# per-function histogram: static instructions | of which executed >= 0.25 x per event | executed per event | share\n")
        for o, v in ex.most_common(60):
            f.write("#   %-40s %6d %6d %10.1f %6.2f %%\n" % (o, stat[o], hot[o], v, 100 * v / tot))
        f.write("#\n# instructions executed >= %.2f x per event:  index | per event | SASS | source line | [function]\n" % thr)
        f.write("\n".join(lines) + "\n")
    print(open(outp).read()[:5000])


if __name__ == "__main__":
    main()
