#!/usr/bin/env python3
"""GPU tuning aid: end-to-end time of dcsb_decode_streams (pinned host in/out) on the bench
workload under different pipeline shapes (chunks x frames per time slice)."""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
streams, n_unique, src = bench.build_corpus(n, 10.0, 0)
if os.environ.get("MIXED"):
    # streams of different lengths: cut every stream after a random number of frames (the frame count
    # in the preamble is patched; the bytes behind the last frame are simply never read)
    rng = np.random.default_rng(1)
    cut = []
    for s in streams:
        nf = int(rng.integers(200, ((s[0] << 8) | s[1]) + 1))
        cut.append(bytes([nf >> 8, nf & 255]) + s[2:])
    streams = cut
ctx = dx.Context(0)
descs, keep = dx.make_descs(streams, os_version=dx.OS94, master_volume=255, mixing_level=0x64, tail_frames=2)
blob = torch.empty(sum(len(s) for s in streams), dtype=torch.uint8).pin_memory()
off, total = 0, 0
bnp = blob.numpy()
for i, s in enumerate(streams):
    bnp[off:off + len(s)] = np.frombuffer(s, dtype=np.uint8)
    descs[i].data = blob.data_ptr() + off
    off += len(s)
    total += (((s[0] << 8) | s[1]) + 2) * 240
h_pcm = torch.empty(total, dtype=torch.int16).pin_memory()
res = (dx.Result * n)()
L = ctx._L
ref = None
shapes = [(0, -1), (0, 0), (8, 126), (8, 63), (4, 0), (4, 63), (2, 63), (1, 63), (8, 315)]
if len(sys.argv) > 2:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[2:]]
for chunks, sl in shapes:
    ctx.set_pipeline(chunks, sl)
    ts = []
    for i in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = L.dcsb_decode_streams(ctx._h, descs, n, h_pcm.data_ptr(), None, res)
        ts.append((time.perf_counter() - t0) * 1e3)
        assert rc == 0, rc
    x = 0
    for i in range(n):
        x ^= res[i].checksum
    ref = x if ref is None else ref
    print("chunks %d slice %4d: %.2f ms (min %.2f)  %.1f Msamples/s  xor %016x %s" % (
        chunks, sl, np.mean(ts[1:]), min(ts[1:]), total / np.mean(ts[1:]) / 1e3, x, "OK" if x == ref else "MISMATCH"), flush=True)
# the copy engine alone: the same bytes, one 1-D copy each way
d = torch.empty(total, dtype=torch.int16, device="cuda")
dblob = torch.empty(blob.numel(), dtype=torch.uint8, device="cuda")
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dblob.copy_(blob, non_blocking=True); h_pcm.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) * 1e3
print("copies alone (H2D %d MB then D2H %d MB, one stream): %.2f ms" % (blob.numel() >> 20, total * 2 >> 20, t))
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h_pcm.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) * 1e3
print("D2H alone: %.2f ms" % t)
