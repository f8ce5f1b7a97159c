#!/usr/bin/env python3
"""Generator of tests/golden/encode_golden.json: SHA-256 of the streams the UNMODIFIED reference DCSEncoder (oracle/_ref,
fed the direct framing: dcsref_encode_framed) makes of seeded clips, for every stream format it can write.  Run in the
build container (needs /root/reference compiled into oracle/_ref); the GPU tests compare dcsb_encode_streams with these
hashes, so the parity check does not depend on the reference library travelling to the GPU box."""
import hashlib
import json
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref


def clip(seed):
    """seeded clip, numpy only (the GPU test rebuilds it from the seed)"""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(200, 30000))
    t = np.arange(n)
    kind = seed % 5
    if kind == 0:
        x = rng.standard_normal(n) * rng.uniform(0.01, 0.5)
    elif kind == 1:
        x = rng.uniform(0.1, 0.9) * np.sin(t * rng.uniform(0.005, 2.0)) + 0.05 * rng.standard_normal(n)
    elif kind == 2:
        x = 0.5 * np.sin(t * t * rng.uniform(1e-6, 5e-5))
    elif kind == 3:
        x = (rng.standard_normal(n) * 0.4) * (np.sin(t * 0.003) > 0.5)
    else:
        x = np.clip(rng.standard_normal(n) * 1.2, -1.0, 1.0)
    return x.astype(np.float32)


CASES = []
for i in range(48):
    fmt = (0x9400, 0x9400, 0x9400, 0x9302, 0x9302, 0x9301)[i % 6]
    if fmt == 0x9400:
        ty, sub = ((0, 0), (0, 3), (1, 0), (1, 3), (-1, -1), (1, -1))[(i // 6) % 6]
    elif fmt == 0x9302:
        ty, sub = ((0, 0), (1, 0), (-1, 0))[(i // 6) % 3]
    else:
        ty, sub = 0, 0
    CASES.append(dict(seed=500 + i, fmt=fmt, type=ty, subtype=sub, bit_rate=(32000, 64000, 128000, 256000)[i % 4],
                      power_cut=(0.9, 0.97, 1.0)[i % 3], max_err=(10.0, 1.0, 100.0)[(i // 2) % 3] / 32768.0, min_range=10.0 / 32768.0))

if __name__ == "__main__":
    out = []
    for c in CASES:
        d, nf = ref.encode_framed(clip(c["seed"]), c["type"], c["subtype"], c["bit_rate"], c["power_cut"], c["max_err"], c["min_range"], fmt=c["fmt"])
        out.append(dict(c, n_frames=nf, n_bytes=len(d), sha256=hashlib.sha256(d).hexdigest()))
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "encode_golden.json"), "w"), indent=0)
    print("wrote %d cases" % len(out))
