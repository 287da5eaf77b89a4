#!/usr/bin/env python
"""Turn the scratch ncu outputs in gpurun_out/ into the small tracked summaries under profiles/.
usage: summarize_profiles.py TAG [launches.csv] [report.ncu-rep kernel_regex]"""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum ",
        "smsp__inst_executed.sum,", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sass__inst_executed_register_spilling",
        "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_per_inst_issued",
        "sm__throughput.avg.pct", "sm__cycles_elapsed.max ", "lts__t_sectors_srcunit_tex_op_read.sum ",
        # instruction supply: SM instruction cache (32 KB) hit rate and the GPC-level cache behind it
        "sm__icc_request_hit_rate", "sm__icc_requests.sum", "gcc__cache_requests_type_instruction",
        "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate", "l1tex__t_requests_pipe_lsu_mem_local_op"]


def launches(tag, path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0] != "ID"]
    agg = {}
    for r in rows:
        k = r[4].split("(")[0]
        a = agg.setdefault(k, [0, 0.0, r[8], r[7]])
        a[0] += 1
        a[1] += float(r[-1]) / 1e6
    tot = sum(a[1] for a in agg.values())
    out = os.path.join(ROOT, "profiles", tag + "_launches.csv")
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold-cache): compare SHARES\n")
        f.write("kernel,launches,total_ms,share,grid,block\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%s,%d,%.3f,%.4f,\"%s\",\"%s\"\n" % (k, a[0], a[1], a[1] / tot, a[2], a[3]))
    print(open(out).read())


def raw(tag, rep, kern):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = os.path.join(ROOT, "profiles", tag + "_ncu_raw.txt")
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on; selected raw metrics per captured launch\n")
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write("== %s  grid %s block %s\n" % (name, vals[hdr.index("Grid Size")], vals[hdr.index("Block Size")]))
            for h, u, v in zip(hdr, units, vals):
                if any(h.startswith(k.strip(" ,")) for k in KEYS):
                    f.write("%-90s %-12s %s\n" % (h, u, v))
    print(open(out).read()[:6000])
    lib = os.environ.get("DMDB_LIB") or os.path.join(ROOT, "parallel_dmd_for_biomolecules_b200", "libdmdb200.so")  # the profiled build
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, lib, kern],
                           capture_output=True, text=True).stdout
    with open(os.path.join(ROOT, "profiles", tag + "_ncu_lines.txt"), "w") as f:
        f.write("# per-source-line / per-function executed warp instructions and stall samples (tools/ncu_lines.py)\n")
        f.write(lines)


if __name__ == "__main__":
    tag = sys.argv[1]
    if len(sys.argv) > 2 and sys.argv[2] != "-":
        launches(tag, sys.argv[2])
    if len(sys.argv) > 4:
        raw(tag, sys.argv[3], sys.argv[4])
