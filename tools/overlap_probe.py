#!/usr/bin/env python3
"""GPU tuning aid (-DDCSB_SCAN_DEBUG build): when do the persistent decode warps work relative to
the scan running beside them?"""
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

streams, n_unique, src = bench.build_corpus(4096, 10.0, 0)
ctx = dx.Context(0)
L = ctx._L
L.dcsb_batch_scan_debug.argtypes = [C.c_void_p, C.c_void_p]
batch = ctx.batch(streams, os_version=dx.OS94, master_volume=255, mixing_level=0x64, tail_frames=2)
for i in range(3):
    batch.decode()
    torch.cuda.synchronize()
print("scan %.3f decode span %.3f step %.3f ms" % (batch.kernel_ms(0), batch.kernel_ms(1), batch.kernel_ms(2)))
dbg = np.zeros((max(2048, len(streams)), 4), dtype=np.uint64)
assert L.dcsb_batch_scan_debug(batch._h, dbg.ctypes.data) == 0
d = dbg[:444 * 4]
d = d[d[:, 3] > 0]
t0 = d[:, 0].min()
print("decode warps that worked: %d; first item ready at +0, last done at +%.3f ms" % (len(d), (d[:, 1].max() - t0) / 1e6))
print("items per warp: mean %.1f min %d max %d; waiting per warp: mean %.3f ms max %.3f ms" % (
    d[:, 3].mean(), d[:, 3].min(), d[:, 3].max(), d[:, 2].mean() / 1e6, d[:, 2].max() / 1e6))
first = np.sort((d[:, 0] - t0) / 1e6)
print("first-item times (ms): p0 %.3f p50 %.3f p90 %.3f p100 %.3f" % (first[0], first[len(first) // 2], first[int(len(first) * .9)], first[-1]))
