"""CPU side of the forward path: the encoder's kernel bodies (dcsexplorer_b200/csrc/dcsb_encode.cuh) compiled by g++
into tests/hostsim, a loop playing each grid, against (1) the committed golden hashes made from the unmodified
reference encoder and (2), when oracle/_ref is there, the reference encoder itself -- frames bit for bit, streams byte
for byte.  Test infrastructure: the product has no CPU path."""
import hashlib
import json
import os
import sys
import numpy as np
import pytest
from oracle import ref
import simutil

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, HERE)
import make_encode_golden as mg


def test_hostsim_encode_matches_the_committed_reference_hashes(built):
    cases = json.load(open(os.path.join(HERE, "encode_golden.json")))
    assert len(cases) == len(mg.CASES)
    # the wildcard is host logic of the product's entry point: here, the explicit cases
    cases = [c for c in cases if c["type"] >= 0 and c["subtype"] >= 0]
    assert len(cases) >= 30
    streams = simutil.encode_streams([mg.clip(c["seed"]) for c in cases],
                                     [(c["type"], c["subtype"], c["bit_rate"], c["power_cut"], c["max_err"], c["min_range"], c["fmt"]) for c in cases])
    bad = [(c["seed"], hex(c["fmt"]), c["type"]) for c, s in zip(cases, streams) if hashlib.sha256(s).hexdigest() != c["sha256"]]
    assert not bad, bad


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_hostsim_encode_frames_and_bytes_equal_the_reference(built):
    rng = np.random.default_rng(8)
    clips = [(rng.standard_normal(int(rng.integers(100, 6000))) * rng.uniform(0.001, 0.7)).astype(np.float32) for _ in range(24)]
    clips += [np.zeros(700, dtype=np.float32), (0.9 * np.sin(np.arange(3000) * 0.05)).astype(np.float32)]
    params = []
    for i in range(len(clips)):
        fmt = (0x9400, 0x9400, 0x9302, 0x9301)[i % 4]
        ty = (i // 4) % 2 if fmt != 0x9301 else 0
        sub = 3 * ((i // 8) % 2) if fmt == 0x9400 else 0
        params.append((ty, sub, int(rng.choice([32000, 128000, 256000])), float(rng.choice([0.9, 0.97, 1.0])), 10.0 / 32768.0, 10.0 / 32768.0, fmt))
    streams, frames = simutil.encode_streams(clips, params, want_frames=True)
    for c, p, s, fr in zip(clips, params, streams, frames):
        want, nf, wfr = ref.encode_framed(c, p[0], p[1], p[2], p[3], p[4], p[5], want_frames=True, fmt=p[6])
        assert np.array_equal(fr.view(np.uint32), wfr.view(np.uint32)), p
        assert s == want, p
