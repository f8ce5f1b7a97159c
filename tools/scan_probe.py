#!/usr/bin/env python3
"""GPU tuning aid: per-stream cycle / step counts of the frame-boundary scan, from a
-DDCSB_SCAN_DEBUG build of the library (tools/build_debug.sh -> /tmp/libdcsb200_dbg.so)."""
import ctypes as C
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["DCSB200_LIB"] = os.path.join(ROOT, "gpurun_dbg", "libdcsb200_dbg.so")
import bench
import torch
import dcsexplorer_b200 as dx

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
streams, n_unique, src = bench.build_corpus(4096, 10.0, 0)
streams = streams[:n]
ctx = dx.Context(0)
ctx.set_overlap(False)      # (the decode kernel's own probes share the debug buffer)
L = ctx._L
L.dcsb_batch_scan_debug.argtypes = [C.c_void_p, C.c_void_p]
for lanes in sys.argv[2:] or ["2"]:
    os.environ["DCSB_SCAN_LANES"] = lanes
    batch = ctx.batch(streams, os_version=dx.OS94, master_volume=255, mixing_level=0x64, tail_frames=2)
    for i in range(3):
        batch.decode()
        torch.cuda.synchronize()
    ms = batch.kernel_ms(0)
    raw = np.zeros((max(2048, n), 8), dtype=np.uint32)          # the library copies max(2048, n) * 32 bytes
    assert L.dcsb_batch_scan_debug(batch._h, raw.ctypes.data) == 0
    dbg, laps = raw[:n, :4], raw[:n, 4:]
    if os.environ.get("LAPS"):
        nf = np.array([(s[0] << 8) | s[1] for s in streams], dtype=np.float64)
        print("n=%d lanes=%s scan %.3f ms" % (n, lanes, ms))
        for ty in range(3):
            for rate in (0, 3, 5):
                idx = np.array([i for i in range(n) if i % 3 == ty and (i // 3) % 6 == rate])
                print("  type %s rate %6d: per frame cycles: topup %.0f header %.0f huffman %.0f rest %.0f" % (
                    (bench.TYPES[ty], bench.RATES[rate]) + tuple((laps[idx, k] / nf[idx]).mean() for k in range(4))), flush=True)
        batch.close()
        continue
    cyc = dbg[:, 0].astype(np.float64) + dbg[:, 1].astype(np.float64) * 2 ** 32
    nf = np.array([(s[0] << 8) | s[1] for s in streams], dtype=np.float64)
    nbits = np.array([len(s) * 8 for s in streams], dtype=np.float64)
    print("n=%d lanes=%s scan %.3f ms; cycles/frame: mean %.0f p50 %.0f max %.0f; steps/frame mean %.1f max %.1f; hdr steps/frame %.1f; bits/frame %.0f"
          % (n, lanes, ms, (cyc / nf).mean(), np.median(cyc / nf), (cyc / nf).max(), (dbg[:, 2] / nf).mean(), (dbg[:, 2] / nf).max(),
             (dbg[:, 3] / nf).mean(), (nbits / nf).mean()), flush=True)
    for ty in range(3):
        for rate in range(6):
            idx = [i for i in range(n) if i % 3 == ty and (i // 3) % 6 == rate]
            if not idx:
                continue
            idx = np.array(idx)
            print("  type %s rate %6d: cyc/frame %.0f  steps/frame %.1f hdr %.1f bits/frame %.0f  cyc/step %.0f" % (
                bench.TYPES[ty], bench.RATES[rate], (cyc[idx] / nf[idx]).mean(), (dbg[idx, 2] / nf[idx]).mean(),
                (dbg[idx, 3] / nf[idx]).mean(), (nbits[idx] / nf[idx]).mean(),
                (cyc[idx] / np.maximum(1, dbg[idx, 2] + dbg[idx, 3])).mean()), flush=True)
    batch.close()
