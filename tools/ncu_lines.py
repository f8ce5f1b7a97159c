#!/usr/bin/env python3
"""Per-source-line shares of executed warp-instructions and stall samples for one kernel of an .ncu-rep
(needs --import-source on and -lineinfo).  usage: tools/ncu_lines.py rep kernel_regex [top_n]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
cur, out, h = None, [], None
for r in csv.reader(txt.splitlines()):
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        h = r
    elif h and len(r) > 8 and r[0].isdigit() and r[2] == "-":
        out.append((int(r[h.index("Instructions Executed")] or 0), int(r[h.index("# Samples")] or 0), cur, r[0], r[1].strip()[:120]))
tot, ts = sum(o[0] for o in out), sum(o[1] for o in out)
print("warp-instructions", tot, "samples", ts)
byfile = {}
for o in out:
    a = byfile.setdefault(o[2], [0, 0]); a[0] += o[0]; a[1] += o[1]
for f, a in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print("  %-22s instr %5.1f%%  samples %5.1f%%" % (f, 100.0 * a[0] / max(tot, 1), 100.0 * a[1] / max(ts, 1)))
for o in sorted(out, reverse=True)[:topn]:
    print("%5.1f%% smp %5.1f%% %s:%s  %s" % (100.0 * o[0] / max(tot, 1), 100.0 * o[1] / max(ts, 1), o[2], o[3], o[4]))
