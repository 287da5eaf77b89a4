#!/usr/bin/env python
"""Join an ncu SASS source page (per-instruction executed counts / stall samples) with nvdisasm -g line info of
the same cubin, and aggregate by CUDA source line / inlined function.  usage: ncu_lines.py rep.ncu-rep lib.so kernel_regex"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter


def main():
    rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) > ix["Instructions Executed"]]
    execd = [int(r[ix["Instructions Executed"]] or 0) for r in data]
    samples = [int(r[ix["# Samples"]] or 0) for r in data]
    sass = [r[ix["Source"]] for r in data]
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
    # walk the kernel's section; remember the current (file, line, inline chain) before each instruction
    insts = []
    in_k = False
    cur = ("?", 0)
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
    print("sass rows in report:", len(data), " instructions in disassembly:", len(insts))
    n = min(len(data), len(insts))
    tot = sum(execd[:n])
    tots = sum(samples[:n])
    by_line, by_line_s = Counter(), Counter()
    for k in range(n):
        key = insts[k][:2]
        by_line[key] += execd[k]
        by_line_s[key] += samples[k]
    print("total executed %d, samples %d" % (tot, tots))
    # ---- by engine component: which function of dmd_engine.h (by line range) the instruction was inlined through
    eng = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", "dmd_engine.h")
    funcs = []
    if os.path.exists(eng):
        for ln, text in enumerate(open(eng), 1):
            m2 = re.match(r"^DMD_(?:DEV|COLD)\s+[\w:<>\s\*&]+?\s+(\w+)\(", text)
            if m2:
                funcs.append((ln, m2.group(1)))

    def func_of(line):
        name = "?"
        for ln, nm in funcs:
            if ln <= line:
                name = nm
        return name

    by_fn, by_fn_s = Counter(), Counter()
    for k in range(n):
        f, l, chain = insts[k] if len(insts[k]) == 3 else (insts[k][0], insts[k][1], ())
        frames = [(f, l)] + list(chain)
        comp = "other"
        for ff, ll in frames:  # innermost engine frame
            if ff == "dmd_engine.h":
                comp = func_of(ll)
                break
        else:
            comp = frames[-1][0]
        by_fn[comp] += execd[k]
        by_fn_s[comp] += samples[k]
    print("---- by engine function (innermost dmd_engine.h frame)")
    for fn, v in by_fn.most_common(30):
        print("%6.2f%% exec  %6.2f%% samples  %s" % (100 * v / tot, 100 * by_fn_s[fn] / max(tots, 1), fn))
    print("---- top source lines by executed instructions")
    for (f, l), v in by_line.most_common(45):
        print("%6.2f%% exec  %6.2f%% samples  %s:%d" % (100 * v / tot, 100 * by_line_s[(f, l)] / max(tots, 1), f, l))
    print("---- top source lines by stall samples")
    for (f, l), v in by_line_s.most_common(30):
        print("%6.2f%% samples  %6.2f%% exec  %s:%d" % (100 * v / max(tots, 1), 100 * by_line[(f, l)] / tot, f, l))


if __name__ == "__main__":
    main()
