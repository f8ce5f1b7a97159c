#!/usr/bin/env python3
"""GPU measurement aid: BASELINE config 3 -- a batch of 1993-era streams (0x9302 types 0 / 1 and
0x9301 type 0 from the reference encoder, plus fuzzer-made OS93a type 1 streams no encoder
produces), resident decode, kernels beside each other and alone, spot-checked against the oracle."""
import os
import sys
import time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import dcsfuzz
import torch
import dcsexplorer_b200 as dx
from oracle import ref, orc

KINDS = [(0x9302, 0), (0x9302, 1), (0x9301, 0)]


def _enc(args):
    seed, seconds = args
    fmt, ty = KINDS[seed % 3]
    data, nf = ref.encode(bench.synth_source(seed, seconds), fmt=fmt, stype=ty, subtype=0,
                          bit_rate=bench.RATES[(seed // 3) % 6], power_cut=bench.CUTS[(seed // 18) % 3])
    return data, fmt


n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pool_n = int(sys.argv[2]) if len(sys.argv) > 2 else 192
seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 10.0
import multiprocessing as mp
with mp.get_context("fork").Pool(len(os.sched_getaffinity(0))) as p:
    pool = p.map(_enc, [(5000 + i, seconds) for i in range(pool_n)], chunksize=2)
rng = np.random.default_rng(9)
nf = int(seconds * 31250 / 240)
pool += [(dcsfuzz.fuzz93a1(rng, nf), 0x9301) for _ in range(pool_n // 3)]
streams = [(pool[i % len(pool)][0], pool[i % len(pool)][1], 255, 0x64, 2) for i in range(n)]
ctx = dx.Context(0)
batch = ctx.batch(streams)
print("batch: %d streams (%d unique: 0x9302 type 0 / type 1, 0x9301 type 0, OS93a type 1), %d frames, %.2f GB in, %.2f GB PCM out" % (
    n, len(pool), batch.total_frames, batch.compressed_bytes / 1e9, batch.total_samples * 2 / 1e9), flush=True)
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
    print("overlap=%d: scan %.2f ms decode %.2f ms step %.2f ms -> %.1f Gsamples/s, %.0f GB/s algorithmic; streams with errors %d" % (
        overlap, np.mean(ks), np.mean(kd), np.mean(kt), batch.total_samples / np.mean(kt) / 1e6,
        (batch.compressed_bytes + batch.total_samples * 2) / np.mean(kt) / 1e6, bad), flush=True)
h = d_pcm.cpu().numpy()
for i in (0, 1, 2, len(pool) - 1):
    d, os_, vol, lvl, tail = streams[i]
    want, _ = orc.decode(d, os_, vol, lvl, ((d[0] << 8) | d[1]) + tail)
    o = batch.pcm_offset(i)
    assert np.array_equal(h[o:o + want.size], want), i
print("spot check vs oracle: bit-exact")
