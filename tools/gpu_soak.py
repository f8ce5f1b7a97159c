#!/usr/bin/env python3
"""GPU soak beyond the test-suite's: batches of fuzzer-made streams of every layout (7 of 8 damaged) with seeds the
tests do not use, decoded through dcsb_decode_streams (pinned, time-sliced) and through a resident batch (scan beside
decode), every stream's checksum against the oracle's PCM.   usage: tools/gpu_soak.py [first_seed=2000] [n_seeds=1840]"""
import multiprocessing as mp
import os
import sys
import time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dcsexplorer_b200 as dx
import test_gpu_parity as T

s0 = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 1840
ctx = dx.Context(0)
t0 = time.time()
total = bad_total = 0
with mp.get_context("fork").Pool(max(1, len(os.sched_getaffinity(0)))) as pool:
    for b0 in range(s0, s0 + ns, 460):
        seeds = range(b0, min(b0 + 460, s0 + ns))
        streams = [s for part in pool.map(T._soak_make, seeds, chunksize=4) for s in part]
        streams = [s for s in streams if len(s[0]) >= 3 and ((s[0][0] << 8) | s[0][1]) > 0]
        chunks = [streams[i:i + 500] for i in range(0, len(streams), 500)]
        want = [c for part in pool.map(T._soak_expect, chunks) for c in part]
        pcm, offs, res = ctx.decode_streams_pinned(streams)
        bad = [i for i in range(len(streams)) if res[i]["checksum"] != want[i]]
        b = ctx.batch(streams)
        b.decode()
        res2 = b.results()
        bad2 = [i for i in range(len(streams)) if res2[i]["checksum"] != want[i]]
        b.close()
        total += len(streams)
        bad_total += len(bad) + len(bad2)
        print("seeds %d..%d: %d streams, mismatches %d (host-buffer call) / %d (resident batch), %.0f s" % (
            seeds[0], seeds[-1], len(streams), len(bad), len(bad2), time.time() - t0), flush=True)
        if bad or bad2:
            i = (bad or bad2)[0]
            print("  first: stream %d os %#x status %d" % (i, streams[i][1], res[i]["status"]))
print("gpu soak: %d streams x 2 paths, mismatches %d, %.0f s" % (total, bad_total, time.time() - t0))
sys.exit(1 if bad_total else 0)
