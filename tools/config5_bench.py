#!/usr/bin/env python3
"""GPU measurement aid: BASELINE config 5's per-GPU share -- very many short streams (default
131,072 x 1 s: a pool of unique streams from the reference encoder, replicated so that every stream
has its own bytes in HBM).  Resident decode, scan and decode kernels beside each other and alone."""
import os
import sys
import time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import torch
import dcsexplorer_b200 as dx

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
pool_n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
pool, n_unique, src = bench.build_corpus(pool_n, 1.0, 1000, budget_s=120.0)
streams = [pool[i % n_unique] for i in range(n)]
ctx = dx.Context(0)
t0 = time.time()
batch = ctx.batch(streams, os_version=dx.OS94, master_volume=255, mixing_level=0x64, tail_frames=2)
print("batch: %d streams (%d unique), %d frames, %.2f GB in, %.2f GB PCM out, created in %.1f s" % (
    n, n_unique, batch.total_frames, batch.compressed_bytes / 1e9, batch.total_samples * 2 / 1e9, time.time() - t0), flush=True)
d_pcm = torch.empty(batch.total_samples, dtype=torch.int16, device="cuda")
st = torch.cuda.current_stream()
for overlap in (1, 0):
    ctx.set_overlap(overlap)
    ks, kd, kt = [], [], []
    for i in range(5):
        batch.decode(d_pcm.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        if i >= 2:
            ks.append(batch.kernel_ms(0)); kd.append(batch.kernel_ms(1)); kt.append(batch.kernel_ms(2))
    res = batch.results(st.cuda_stream)
    bad = sum(1 for r in res if r["status"] != 0)
    x = 0
    for r in res[:n_unique]:
        x ^= r["checksum"]
    print("overlap=%d: scan %.2f ms decode %.2f ms step %.2f ms -> %.1f Gsamples/s, %.0f GB/s algorithmic; errors %d, xor(first pool) %016x" % (
        overlap, np.mean(ks), np.mean(kd), np.mean(kt), batch.total_samples / np.mean(kt) / 1e6,
        (batch.compressed_bytes + batch.total_samples * 2) / np.mean(kt) / 1e6, bad, x), flush=True)
