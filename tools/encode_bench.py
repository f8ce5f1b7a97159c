#!/usr/bin/env python3
"""GPU measurement aid: the forward path (dcsb_encode_streams) on clips of the bench corpus' shape -- stream types
{0.0, 1.0, 1.3} x bit rates 32k..256k x power cut {.90, .97, 1} -- against the reference DCSEncoder fed the same framing
on the host cores (oracle/_ref: dcsref_encode_framed): bytes compared stream by stream, throughput of both.
  encode_bench.py [n_clips=256] [seconds=10]"""
import multiprocessing as mp
import os
import sys
import time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import dcsexplorer_b200 as dx
from oracle import ref

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0


def params(seed):
    ty, sub = bench.TYPES[seed % 3]
    return (ty, sub, bench.RATES[(seed // 3) % 6], bench.CUTS[(seed // 18) % 3])


def _ref_one(seed):
    p = params(seed)
    d, nf = ref.encode_framed(bench.synth_source(seed, seconds), p[0], p[1], p[2], p[3])
    return d


ncpu = max(1, len(os.sched_getaffinity(0)))
with mp.get_context("fork").Pool(ncpu) as pool:
    t0 = time.perf_counter()
    clips = [bench.synth_source(9000 + i, seconds) for i in range(n)]
    t_src = time.perf_counter() - t0
    t0 = time.perf_counter()
    want = pool.map(_ref_one, [9000 + i for i in range(n)], chunksize=2)
    t_ref = time.perf_counter() - t0
# (the reference timing includes making the clip inside each worker; measured separately: t_src single-threaded)
ctx = dx.Context(0)
pl = [params(9000 + i) for i in range(n)]
ts = []
for it in range(3):
    t0 = time.perf_counter()
    got = ctx.encode_streams(clips, pl)
    ts.append(time.perf_counter() - t0)
bad = sum(1 for a, b in zip(got, want) if a != b)
samples = sum(c.size for c in clips)
print("encode: %d clips x %.1f s (%d samples, %d bytes of streams); GPU %.1f ms per call (first %.1f) = %.1f Msamples/s incl. PCM upload and "
      "stream download; reference DCSEncoder (same framing) on %d host threads: %.1f Msamples/s (incl. ~%.0f%% making the clips); "
      "streams byte-identical: %d of %d" % (n, seconds, samples, sum(len(g) for g in got), min(ts[1:]) * 1e3, ts[0] * 1e3,
                                            samples / min(ts[1:]) / 1e6, ncpu, samples / t_ref / 1e6, 100.0 * t_src / ncpu / max(t_ref, 1e-9), n - bad, n))
sys.exit(1 if bad else 0)
