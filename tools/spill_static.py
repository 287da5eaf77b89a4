#!/usr/bin/env python
"""Static count of local-memory (spill) instructions of dmd_event_loop_kernel by source region of dmd_engine.h
(nvdisasm -g line info).  A cheap proxy to compare builds before spending GPU time.  usage: spill_static.py lib.so"""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

REGIONS = [("calendar/pop", "group_min_update", "pack_type"), ("segmented_pass", "seg_argmin", "repuls_del_b(Rep& r, int n, int cb);"),
           ("partial_events", "partial_events_t", "redo_lane"), ("pair_event(hot)", "void pair_event(", "cell_coords"),
           ("step/run", "void process_one", "void retemp")]


def main():
    lib = os.path.abspath(sys.argv[1])
    kern = sys.argv[2] if len(sys.argv) > 2 else "dmd_event_loop_kernel"
    eng = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "parallel_dmd_for_biomolecules_b200", "csrc", "dmd_engine.h")
    src = open(eng).read().splitlines()

    def find(tok, start=0):
        for k in range(start, len(src)):
            if tok in src[k]:
                return k + 1
        return len(src)
    spans = []
    for name, a, b in REGIONS:
        la = find(a)
        spans.append((name, la, find(b, la)))
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
    in_k, cur = False, None
    tot, spl = Counter(), Counter()
    for line in dis:
        if line.startswith("//---") and ".text." in line:
            in_k = kern in line
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
        if m:
            inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            cur = [(os.path.basename(m.group(1)), int(m.group(2)))] + [(os.path.basename(f), int(l)) for f, l in inl]
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            key = None
            for f, l in cur or []:
                if f == "dmd_engine.h":
                    key = l
                    break
            reg = "other"
            for name, a, b in spans:
                if key is not None and a <= key < b:
                    reg = name
            tot[reg] += 1
            if "LDL" in line or "STL" in line:
                spl[reg] += 1
    for name in [s[0] for s in spans] + ["other"]:
        print("%-18s insts %6d  LDL/STL %5d" % (name, tot[name], spl[name]))


if __name__ == "__main__":
    main()
