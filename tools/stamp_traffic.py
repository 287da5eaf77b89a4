#!/usr/bin/env python
"""Write profiles/event_loop_traffic.json from an `ncu --set full` capture of dmd_event_loop_kernel: DRAM bytes per
event of the captured launch, stamped with the hash of the kernel sources so that bench.py only uses it for the code it
was measured on.  usage: stamp_traffic.py report.ncu-rep events_in_captured_launch "description of the capture" """
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rep, events, desc = sys.argv[1], float(sys.argv[2]), sys.argv[3]
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (float(v.replace(",", "")), u) for h, u, v in zip(hdr, units, vals) if h in ("dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum")}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
rd = m["dram__bytes_read.sum"][0] * scale[m["dram__bytes_read.sum"][1]]
wr = m["dram__bytes_write.sum"][0] * scale[m["dram__bytes_write.sum"][1]]
out = {"kernel": "dmd_event_loop_kernel", "source": desc, "source_sha16": bench.kernel_source_hash(), "dram_bytes_read": rd,
       "dram_bytes_write": wr, "events": events, "dram_bytes_per_event": (rd + wr) / events,
       "warp_instructions_per_event": m["smsp__inst_executed.sum"][0] / events,
       "note": "smsp__inst_executed over the whole grid (event-loop CTAs, list-rebuild service CTAs and their idle polling) / events"}
with open(os.path.join(ROOT, "profiles", "event_loop_traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
