#!/usr/bin/env python3
"""Summarise an `ncu --set full` capture of the bench workload into profiles/traffic.json: per kernel
the DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum), the warp instructions executed and the
launch shape.  bench.py prints `roofline.traffic` / `issue_frac` from it only when the shapes match
what it launches.   usage: tools/make_traffic.py gpurun_out/x_prof.ncu-rep profiles/x_ncu_full.txt"""
import csv
import json
import os
import subprocess
import sys

rep, source = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
kern = {}
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = r[col["Kernel Name"]].split("(")[0]
    name = name.replace("void ", "").split("<")[0].strip()      # template instances under the kernel's plain name

    def num(k):
        return float(r[col[k]].replace(",", ""))

    def unit_bytes(k):
        u = rows[1][col[k]].lower()
        return num(k) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    grid = int(num("launch__grid_size"))
    block = int(num("launch__block_size"))
    rec = {"traffic_bytes": unit_bytes("dram__bytes_read.sum") + unit_bytes("dram__bytes_write.sum"),
           "dram_read_bytes": unit_bytes("dram__bytes_read.sum"), "dram_write_bytes": unit_bytes("dram__bytes_write.sum"),
           "warp_insts": num("smsp__inst_executed.sum"), "grid": grid, "block": block,
           "duration_ms_under_ncu": num("gpu__time_duration.sum") / 1e6 if rows[1][col["gpu__time_duration.sum"]] in ("ns", "nsecond") else num("gpu__time_duration.sum"),
           "threads_per_inst": num("smsp__thread_inst_executed_per_inst_executed.ratio")}
    kern[name] = rec          # the last capture of a kernel wins
json.dump({"source": source, "kernels": kern}, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(kern, indent=1))
