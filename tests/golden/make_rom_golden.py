#!/usr/bin/env python3
"""Generate tests/golden/rom_golden.npz: the UNMODIFIED reference decoder (oracle/_ref) playing
the seeded ROM scenarios of tests/romscen.py (BASELINE config 4).  Runs only in the build
container; the fixture is committed so the GPU box needs no reference tree.

Per scenario: SHA-256 of the ROM images (guards against generator drift), FNV-1a-64 of the
reference PCM, a 32-bit sum per frame (to localise a mismatch), the first and last 20 frames of
PCM, and the bytes the decoder sent back to the host."""
import hashlib
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle import ref, orc     # noqa: E402
import romscen                  # noqa: E402


def images_digest(images):
    h = hashlib.sha256()
    for chip in sorted(images):
        h.update(bytes([chip]))
        h.update(images[chip])
    return h.hexdigest()


def frame_sums(pcm):
    return pcm.reshape(-1, 240).astype(np.int64).sum(axis=1).astype(np.uint32)


def main():
    out = {"names": np.array([n for n, _ in romscen.SCENARIOS])}
    for name, kw in romscen.SCENARIOS:
        sc = romscen.make_scenario(**kw)
        rp = ref.RomPlayer(sc["images"], sc["master_volume"])
        info = rp.info()
        assert info["check"] == 1, (name, info)
        pcm = rp.render_timeline(sc["writes"], sc["n_frames"])
        hb = rp.host_bytes()
        out[name + "/digest"] = np.array(images_digest(sc["images"]))
        out[name + "/fnv"] = np.array(orc.fnv1a(pcm), dtype=np.uint64)
        out[name + "/sums"] = frame_sums(pcm)
        out[name + "/head"] = pcm[:240 * 20]
        out[name + "/tail"] = pcm[-240 * 20:]
        out[name + "/host"] = np.frombuffer(hb, dtype=np.uint8)
        out[name + "/streams"] = np.array(rp.list_streams(), dtype=np.uint32)
        out[name + "/info"] = np.array([info["os"], info["hw"], info["max_track"], info["channels"]], dtype=np.int32)
        print("%-12s frames %4d  nonzero frames %4d  host bytes %s" % (
            name, sc["n_frames"], int((pcm.reshape(-1, 240) != 0).any(axis=1).sum()), hb.hex()))
    np.savez_compressed(os.path.join(HERE, "rom_golden.npz"), **out)


if __name__ == "__main__":
    main()
