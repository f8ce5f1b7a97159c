"""The oracle (oracle/dcs_oracle.c) against the committed golden fixtures generated from the
reference, and -- in the build container, where oracle/_ref exists -- against the reference
itself on fresh random inputs."""
import numpy as np
import pytest
from conftest import check_against_golden
from oracle import orc, ref
import dcsfuzz

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (reference tree absent)")


def test_oracle_matches_golden_pcm(built, golden):
    for it in golden.items:
        pcm, rc = orc.decode(it["stream"], it["os"], it["vol"], it["lvl"], it["nframes_out"])
        check_against_golden(it, pcm)
        assert (rc == 0) or it["stop"] or rc < 0


def test_oracle_scan_matches_golden_bitpos(built, golden):
    for it in golden.items:
        if it["stop"]:
            continue
        rc, bp, bt, stop = orc.scan(it["stream"], it["os"])
        assert rc == len(it["bitpos"]) - 1
        assert np.array_equal(bp, it["bitpos"]), it["label"]


def test_fnv_implementation():
    from conftest import fnv1a
    x = np.arange(-50, 50, dtype=np.int16)
    assert orc.fnv1a(x) == fnv1a(x)


def test_oracle_rejects_empty_and_short(built):
    pcm, rc = orc.decode(bytes([0, 0] + [0x7F] * 16), 0x9400, 255, 100, 3)
    assert rc == -1 and not pcm.any()
    pcm, rc = orc.decode(bytes([0, 1]), 0x9400, 255, 100, 3)
    assert rc == -4 and not pcm.any()


def test_header_with_no_bands_consumes_no_bits(built):
    d = bytes([0, 5] + [0x7F] * 16)
    rc, bp, bt, stop = orc.scan(d, 0x9400)
    assert rc == 5 and not bp.any()


@needs_ref
@pytest.mark.ref
def test_oracle_vs_reference_fuzz(built):
    for seed in range(6):
        for os_, d, label in dcsfuzz.corpus(seed + 1000, n_each=2, nframes=9):
            want = ref.decode(d, os_, 200, 0x50)
            got, _ = orc.decode(d, os_, 200, 0x50)
            assert np.array_equal(got, want), label
            bp, bt, bins, stop = ref.probe_frames(d, os_)
            rc, obp, obt, ostop = orc.scan(d, os_)
            assert np.array_equal(bp, obp), label


@needs_ref
@pytest.mark.ref
def test_oracle_transform_vs_reference_extremes(built):
    rng = np.random.default_rng(3)
    for it in range(400):
        os_ = [0x9400, 0x9302][it & 1]
        bins = rng.choice([-32768, 32767, 0, -1, 1, 12345, -12345], 256) if it % 3 == 0 else rng.integers(-32768, 32768, 256)
        ovl = rng.integers(-32768, 32768, 16)
        vs = int(rng.integers(0, 9))
        p1, o1, _ = ref.transform(bins, ovl, os_, vs)
        p2, o2 = orc.transform(bins, ovl, os_, vs)
        assert np.array_equal(p1, p2) and np.array_equal(o1, o2)
