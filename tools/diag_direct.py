#!/usr/bin/env python3
"""GPU diagnostic: the mixed fuzz batch of test_gpu_scan_direct_variant_vs_oracle through the public calls,
every stream against the oracle; prints which streams differ and how."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dcsexplorer_b200 as dx
from oracle import orc
import test_gpu_parity as T
ctx = dx.Context(0)
streams = []
for seed in range(12):
    streams += [s for s in T._soak_make(900 + seed) if len(s[0]) >= 3 and ((s[0][0] << 8) | s[0][1]) > 0]
print("streams", len(streams), "env", {k: v for k, v in os.environ.items() if k.startswith("DCSB_")})
for name, fn in (("decode_streams", ctx.decode_streams), ("decode_streams_pinned", ctx.decode_streams_pinned)):
    pcm, offs, res = fn(streams)
    bad = []
    for i, (d, os_, vol, lvl, tail) in enumerate(streams):
        want, rc = T._expect(d, os_, vol, lvl, tail)
        got = pcm[offs[i]:offs[i] + want.size]
        if not np.array_equal(got, want):
            k = int(np.nonzero(got != want)[0][0])
            bad.append((i, hex(os_), res[i]["status"], rc, k // 240, want.size // 240))
    print(name, "bad", len(bad), bad[:12])
for ov in (1, 0):
    ctx.set_overlap(ov)
    b = ctx.batch(streams)
    import torch
    d_pcm = torch.full((b.total_samples,), 0x5A5A, dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    b.decode(d_pcm.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    h = d_pcm.cpu().numpy()
    res = b.results(torch.cuda.current_stream().cuda_stream)
    bad = []
    for i, (d, os_, vol, lvl, tail) in enumerate(streams):
        want, rc = T._expect(d, os_, vol, lvl, tail)
        o = b.pcm_offset(i)
        got = h[o:o + want.size]
        if not np.array_equal(got, want):
            k = int(np.nonzero(got != want)[0][0])
            fr = sorted(set((np.nonzero(got != want)[0] // 240).tolist()))
            bad.append((i, hex(os_), res[i]["status"], rc, fr, want.size // 240, (d[0] << 8) | d[1], res[i], got[k:k + 6].tolist(), want[k:k + 6].tolist(), k % 240))
    print("batch overlap", ov, "bad", len(bad), bad[:12])
