#!/usr/bin/env python3
"""GPU measurement aid: BASELINE config 1 -- ONE 10 s clip (0.5 sin 440 Hz + 0.2 sin 2.5 kHz +
N(0, 0.05)) encoded by the reference's DCSEncoder as a 1994+ type 1.3 stream at 128 kbit/s
(1303 frames), decoded through dcsb_decode_streams (host in, host out): single-stream latency,
bit-exact against the unmodified reference decoder run on the same bytes."""
import os
import sys
import time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dcsexplorer_b200 as dx
from oracle import ref, orc

rng = np.random.Generator(np.random.MT19937(12345))
t = np.arange(312500) / 31250.0
x = (0.5 * np.sin(2 * np.pi * 440 * t) + 0.2 * np.sin(2 * np.pi * 2500 * t) + rng.normal(0, 0.05, t.size)).astype(np.float32)
data, nf = ref.encode(x, fmt=0x9400, stype=1, subtype=3, bit_rate=128000, power_cut=0.97)
t0 = time.perf_counter()
want = ref.decode(data, 0x9400, 255, 0x64, nf + 2)
tref = time.perf_counter() - t0
ctx = dx.Context(0)
ts = []
for i in range(8):
    t0 = time.perf_counter()
    pcm, offs, res = ctx.decode_streams([(data, 0x9400, 255, 0x64, 2)])
    ts.append(time.perf_counter() - t0)
assert res[0]["status"] == 0 and np.array_equal(pcm[:want.size], want)
print("config 1: %d frames, %d bytes; dcsb_decode_streams (host in/out) %.2f ms (first call %.2f ms) = %.1f Msamples/s; "
      "reference DCSDecoderNative, 1 thread: %.2f ms = %.1f Msamples/s; bit-exact, fnv %016x" % (
          nf, len(data), min(ts[1:]) * 1e3, ts[0] * 1e3, want.size / min(ts[1:]) / 1e6, tref * 1e3, want.size / tref / 1e6,
          orc.fnv1a(pcm[:want.size])))
ctx.close()
