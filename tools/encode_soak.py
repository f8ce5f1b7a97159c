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


def make(args):
    seed, i = args
    rng = np.random.default_rng([seed, i])
    n = int(rng.integers(100, 60000))
    t = np.arange(n)
    kind = i % 6
    if kind == 0:
        x = rng.standard_normal(n) * rng.uniform(0.001, 0.5)
    elif kind == 1:
        x = rng.uniform(0.05, 1.0) * np.sin(t * rng.uniform(0.001, 3.0))
    elif kind == 2:
        x = 0.5 * np.sin(t * t * rng.uniform(1e-6, 1e-4))
    elif kind == 3:
        x = (rng.standard_normal(n) * 0.4) * (np.sin(t * 0.002) > 0.7)
    elif kind == 4:
        x = rng.standard_normal(n) * 1e-4
    else:
        x = np.clip(rng.standard_normal(n) * 1.5, -1.0, 1.0)
    p = (int(rng.integers(0, 2)), int(rng.choice([0, 3])), int(rng.choice([8000, 32000, 64000, 96000, 128000, 192000, 256000, 512000])),
         float(rng.choice([0.5, 0.9, 0.97, 1.0])), float(rng.choice([1.0, 10.0, 100.0])) / 32768.0, float(rng.choice([0.0, 10.0, 200.0])) / 32768.0)
    fmt = int(rng.choice([0x9400, 0x9400, 0x9302, 0x9301]))
    if fmt == 0x9400:
        if i % 11 == 0:
            p = (-1, -1) + p[2:]                # the wildcard: every format tried, the first of the smallest kept
    elif fmt == 0x9302:
        p = (int(rng.choice([0, 1, -1])), 0) + p[2:]
    else:
        p = (0, 0) + p[2:]                      # $9301: stream type 0 (nobody encodes OS93a type 1)
    return x.astype(np.float32), p + (fmt,)


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
