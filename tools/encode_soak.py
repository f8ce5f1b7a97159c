#!/usr/bin/env python3
"""GPU soak of the forward path (1994 layout: the four stream types and the wildcard; $9302: types 0, 1 and the wildcard; $9301: type 0): seeded random clips (noise, tones, sweeps, bursts, near-silence, clipped; 100..60000
samples) with random parameters through dcsb_encode_streams, every stream's bytes against the reference DCSEncoder fed
the same framing (oracle/_ref) on the host cores.   usage: tools/encode_soak.py [n_clips=4000] [seed=1]"""
import multiprocessing as mp
import os
import sys
import time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dcsexplorer_b200 as dx
from oracle import ref


from encode_soak_cases import make


def want(args):
    x, p = make(args)
    return ref.encode_framed(x, p[0], p[1], p[2], p[3], p[4], p[5], fmt=p[6])[0]


n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = dx.Context(0)
t0 = time.time()
bad = total = 0
with mp.get_context("fork").Pool(max(1, len(os.sched_getaffinity(0)))) as pool:
    for b0 in range(0, n, 500):
        idx = [(seed, i) for i in range(b0, min(n, b0 + 500))]
        w = pool.map(want, idx, chunksize=4)
        cp = pool.map(make, idx, chunksize=4)
        got = ctx.encode_streams([c for c, _ in cp], [p for _, p in cp])
        d = [i for i in range(len(idx)) if got[i] != w[i]]
        total += len(idx)
        bad += len(d)
        print("clips %d..%d: mismatches %d%s, %.0f s" % (b0, b0 + len(idx) - 1, len(d), (" first %d %s" % (b0 + d[0], cp[d[0]][1])) if d else "", time.time() - t0), flush=True)
print("encode soak: %d clips, mismatches %d, %.0f s" % (total, bad, time.time() - t0))
sys.exit(1 if bad else 0)
