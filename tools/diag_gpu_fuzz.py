"""Diagnostic (GPU box): decode the fuzz corpus of tests/test_gpu_parity.py through the C-ABI
and print where the CUDA path and the oracle disagree."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dcsexplorer_b200 as dx
import dcsfuzz
from oracle import orc
rng = np.random.default_rng(4242)
streams, labels = [], []
for seed in range(10):
    for os_, d, label in dcsfuzz.corpus(seed + 300, n_each=3, nframes=int(rng.integers(1, 100))):
        streams.append((d, os_, int(rng.integers(0, 256)), int(rng.integers(0, 256)), int(rng.integers(0, 5))))
        labels.append(label)
for i in range(8):
    streams.append((dcsfuzz.fuzz94(rng, 50, type1=i & 1, max_code=6, error_frame=int(rng.integers(0, 50)), escape_p=0.2), 0x9400, 255, 0x64, 2))
    labels.append("err%d" % i)
ctx = dx.Context(0)
for rep in range(2):
    pcm, offs, res = ctx.decode_streams(streams)
    nbad = 0
    for i, (d, os_, vol, lvl, tail) in enumerate(streams):
        nf = (d[0] << 8) | d[1]
        want, rc = orc.decode(d, os_, vol, lvl, nf + tail)
        got = pcm[offs[i]:offs[i] + want.size]
        if not np.array_equal(got, want):
            bad = np.nonzero(got != want)[0]
            nbad += 1
            print("rep", rep, "stream", i, labels[i], hex(os_), "nf", nf, "tail", tail, "vol", vol, "lvl", lvl,
                  "first bad sample", bad[0], "frame", bad[0] // 240, "+", bad[0] % 240, "nbad", len(bad),
                  "last bad frame", bad[-1] // 240, "hdr", d[2:18].hex(), res[i])
            fr = sorted(set((bad // 240).tolist()))
            print("   bad frames:", fr[:40])
            k = bad[0]
            print("   got ", got[k:k + 8], "want", want[k:k + 8])
    print("rep", rep, "mismatching streams:", nbad, "of", len(streams))
