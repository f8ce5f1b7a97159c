#!/usr/bin/env python3
"""Hot SASS instructions of one kernel from an .ncu-rep (source page): samples, executions,
average active threads.  usage: tools/ncu_hot.py rep kernel_regex [top_n] [context]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
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
tot = sum(int(r[S] or 0) for r in data)
totx = sum(int(r[X] or 0) for r in data)
print("kernel", rows[0][1][:80], "instructions", len(data), "samples", tot, "warp-instr executed", totx)
top = sorted(range(len(data)), key=lambda k: -int(data[k][S] or 0))[:topn]
acc = 0
for k in sorted(top):
    r = data[k]
    acc += int(r[S] or 0)
    print("%5d %-78s smp %6s (%4.1f%%) exec %10s thr %5s" % (k, r[ci["Source"]].strip()[:78], r[S], 100.0 * int(r[S] or 0) / max(tot, 1), r[X], r[T]))
print("shown: %.1f%% of samples" % (100.0 * acc / max(tot, 1)))
if len(sys.argv) > 4:
    lo, hi2 = [int(v) for v in sys.argv[4].split(":")]
    print("---- range")
    for k in range(lo, hi2):
        r = data[k]
        print("%5d %-78s smp %6s exec %10s thr %5s" % (k, r[ci["Source"]].strip()[:78], r[S], r[X], r[T]))
