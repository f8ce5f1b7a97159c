#!/usr/bin/env python3
"""Group a kernel's SASS (ncu source page) into runs of equal execution count and print each
run's share of executed warp-instructions and of stall samples.
usage: tools/ncu_regions.py rep kernel_regex [top_n]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        if data:
            break
        continue
    if len(r) == len(hdr):
        data.append(r)
S, X, T = ci["# Samples"], ci["Instructions Executed"], ci["Avg. Threads Executed"]
totx = sum(int(r[X]) for r in data)
tots = sum(int(r[S]) for r in data)
regions, cur = [], None
for k, r in enumerate(data):
    x = int(r[X])
    if cur and abs(x - cur["x"]) <= 0.02 * max(x, cur["x"], 1):
        cur["n"] += 1; cur["sx"] += x; cur["ss"] += int(r[S]); cur["end"] = k
    else:
        cur = {"start": k, "end": k, "x": x, "n": 1, "sx": x, "ss": int(r[S]), "thr": r[T]}
        regions.append(cur)
print("kernel %s: %d SASS instructions, %d warp-instructions executed (profiled pass), %d samples" % (rows[0][1][:60], len(data), totx, tots))
for g in sorted(regions, key=lambda g: -g["sx"])[:topn]:
    print("instr %4d-%4d n=%3d exec/instr=%10d thr=%5s  warp-instr %5.1f%%  samples %5.1f%%   %s" % (
        g["start"], g["end"], g["n"], g["x"], g["thr"], 100 * g["sx"] / totx, 100 * g["ss"] / tots, data[g["start"]][ci["Source"]].strip()[:44]))
