#!/usr/bin/env python3
"""Generate tests/golden/golden.npz from the UNMODIFIED reference decoder/encoder compiled
into oracle/_ref (oracle/Makefile).  Runs only in the build container (the reference tree is
not available on the GPU box); the resulting fixture file is committed.

Contents per fixture i: stream bytes, os_version, master volume, mixing level, frames pulled,
and the reference's PCM (full for short streams; FNV-1a-64 + head/tail for long ones), plus
frame bit offsets from the reference's own bit pointer for scan parity."""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle import ref          # noqa: E402
import dcsfuzz                  # noqa: E402


def fnv1a(pcm):
    h = 1469598103934665603
    for b in np.ascontiguousarray(pcm, dtype="<i2").tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def main():
    rng = np.random.default_rng(20261017)
    fx = []
    # (1) encoder-made clips: config-1 style source (sine + sine + noise), 1 s, every layout the
    # reference encoder supports (DCSEncoder.cpp:785-842, :2053-2470)
    t = np.arange(31250) / 31250.0
    x = 0.5 * np.sin(2 * np.pi * 440 * t) + 0.2 * np.sin(2 * np.pi * 2500 * t) + rng.normal(0, 0.05, t.size)
    for fmt, ty, sub, br in [(0x9400, 1, 3, 128000), (0x9400, 0, 0, 128000), (0x9400, 1, 0, 64000),
                             (0x9400, 0, 3, 256000), (0x9302, 0, 0, 128000), (0x9302, 1, 0, 128000),
                             (0x9301, 0, 0, 96000)]:
        d, nf = ref.encode(x, fmt=fmt, stype=ty, subtype=sub, bit_rate=br)
        fx.append((d, fmt, 255, 0x64, nf + 2, "enc-%04x-%d.%d" % (fmt, ty, sub)))
    # loud clip: exercises saturation in the 1994 transform
    xl = np.clip(1.6 * np.sin(2 * np.pi * 997 * t[:8000]) + rng.normal(0, 0.4, 8000), -1, 1)
    d, nf = ref.encode(xl, fmt=0x9400, stype=1, subtype=3, bit_rate=256000, power_cut=1.0)
    fx.append((d, 0x9400, 255, 0x7F, nf + 2, "enc-loud"))
    # (2) bit-level fuzz: branches no encoder produces (incl. OS93a type 1 and the error path)
    for os_, d, label in dcsfuzz.corpus(seed=11, n_each=4, nframes=10):
        fx.append((d, os_, int(rng.integers(1, 256)), int(rng.integers(0, 128)), ((d[0] << 8) | d[1]) + 2, "fuzz-" + label))
    for i in range(6):
        d = dcsfuzz.fuzz94(rng, 12, type1=i & 1, max_code=6, error_frame=int(rng.integers(0, 12)), escape_p=0.25)
        fx.append((d, 0x9400, 255, 0x64, 14, "fuzz-94-error-%d" % i))
    # volume edge cases on one stream
    d0 = fx[0][0]
    for vol, lvl in [(0, 0x64), (1, 0), (128, 0x7F), (255, 0xFF), (67, 0x40)]:
        fx.append((d0, 0x9400, vol, lvl, 20, "vol-%d-%d" % (vol, lvl)))

    out = {}
    meta = []
    for i, (d, os_, vol, lvl, nfr, label) in enumerate(fx):
        pcm = ref.decode(d, os_, vol, lvl, nfr)
        bp, bt, _, stop = ref.probe_frames(d, os_)
        out["s%d" % i] = np.frombuffer(d, dtype=np.uint8)
        out["p%d" % i] = pcm if pcm.size <= 240 * 40 else np.concatenate([pcm[:2400], pcm[-2400:]])
        out["b%d" % i] = bp
        meta.append((os_, vol, lvl, nfr, fnv1a(pcm), int(stop.any()), label))
    out["meta"] = np.array([(m[0], m[1], m[2], m[3], m[5]) for m in meta], dtype=np.int64)
    out["fnv"] = np.array([m[4] for m in meta], dtype=np.uint64)
    out["labels"] = np.array([m[6] for m in meta])
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    print("wrote %d fixtures" % len(fx))


if __name__ == "__main__":
    main()
