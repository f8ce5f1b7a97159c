"""CPU simulation of the CUDA kernel bodies (tests/hostsim) against the oracle and the golden
fixtures: the same dcsb_scan_stream / dcsb_decode_tile code the GPU runs, compiled as C++."""
import numpy as np
from conftest import check_against_golden
from oracle import orc
import dcsfuzz
import simutil


def test_sim_matches_golden(built, golden):
    streams = [(it["stream"], it["os"], it["vol"], it["lvl"], it["nframes_out"] - ((it["stream"][0] << 8) | it["stream"][1]))
               for it in golden.items]
    pcm, offs, res, bp, bt = simutil.decode_streams(streams)
    for i, it in enumerate(golden.items):
        check_against_golden(it, pcm[offs[i]:offs[i] + it["nframes_out"] * 240])
        assert res[i]["status"] == (-5 if it["stop"] else 0)


def test_sim_fuzz_vs_oracle(built):
    rng = np.random.default_rng(77)
    streams = []
    for seed in range(4):
        for os_, d, label in dcsfuzz.corpus(seed + 50, n_each=2, nframes=int(rng.integers(1, 80))):
            streams.append((d, os_, int(rng.integers(0, 256)), int(rng.integers(0, 256)), int(rng.integers(0, 5))))
    pcm, offs, res, bp, bt = simutil.decode_streams(streams)
    fbase = 0
    for i, (d, os_, vol, lvl, tail) in enumerate(streams):
        nf = (d[0] << 8) | d[1]
        want, rc = orc.decode(d, os_, vol, lvl, nf + tail)
        assert np.array_equal(pcm[offs[i]:offs[i] + want.size], want), (i, hex(os_))
        orc_rc, obp, obt, ostop = orc.scan(d, os_)
        assert np.array_equal(bp[fbase:fbase + nf], obp[:nf])
        assert np.array_equal(bt[fbase:fbase + nf], obt[:nf])
        # checksum definition: sum (uint16)s[i] * (2i+1) mod 2^64
        s = want.astype(np.uint16).astype(np.uint64)
        w = (2 * np.arange(s.size, dtype=np.uint64) + 1)
        assert res[i]["checksum"] == int((s * w).sum(dtype=np.uint64))
        assert res[i]["stream_bytes"] == 2 + (1 if (os_ == 0x9301 and d[2] & 0x80) else 16) + (int(obp[nf]) + 7) // 8
        fbase += nf


def test_sim_edge_cases(built):
    empty = bytes([0, 0] + [0x10] * 16)
    short = bytes([0, 3, 0x10])
    nobands = bytes([0, 4] + [0x7F] * 16)
    ok = dcsfuzz.fuzz94(np.random.default_rng(1), 31 * 3, type1=1)       # exactly three tiles
    one = dcsfuzz.fuzz94(np.random.default_rng(2), 1, type1=0)
    trunc = ok[: len(ok) // 2]
    streams = [(empty, 0x9400, 255, 100, 2), (short, 0x9400, 255, 100, 2), (nobands, 0x9400, 255, 100, 1),
               (ok, 0x9400, 255, 100, 0), (one, 0x9400, 255, 100, 3), (trunc, 0x9400, 255, 100, 2)]
    pcm, offs, res, bp, bt = simutil.decode_streams(streams)
    assert [r["status"] for r in res[:5]] == [-1, -4, 0, 0, 0]
    assert res[5]["status"] in (-2, -3, -5)
    for i, (d, os_, vol, lvl, tail) in enumerate(streams):
        nf = ((d[0] << 8) | d[1]) if i != 1 else 3
        if i == 1:
            assert not pcm[offs[i]:offs[i] + (nf + tail) * 240].any()
            continue
        want, rc = orc.decode(d, os_, vol, lvl, nf + tail)
        assert np.array_equal(pcm[offs[i]:offs[i] + want.size], want), i


def test_sim_time_slices(built):
    """Resumed scans (frames [k*slice, (k+1)*slice) per launch) + per-slice work items give the
    same PCM, checkpoints, statuses and checksums as the one-pass form."""
    rng = np.random.default_rng(5)
    streams = []
    for seed in range(3):
        for os_, d, label in dcsfuzz.corpus(seed + 90, n_each=2, nframes=int(rng.integers(40, 150))):
            streams.append((d, os_, int(rng.integers(0, 256)), int(rng.integers(0, 256)), int(rng.integers(0, 4))))
    ok = dcsfuzz.fuzz94(np.random.default_rng(1), 100, type1=1)
    streams.append((ok[: len(ok) // 2], 0x9400, 255, 100, 2))           # truncated: fails in some slice
    streams.append((bytes([0, 0] + [0x10] * 16), 0x9400, 255, 100, 2))  # empty
    want = simutil.decode_streams(streams)
    for sl in (1, 7, 31, 64, 1000):
        got = simutil.decode_streams(streams, slice_frames=sl)
        assert np.array_equal(got[0], want[0]), sl
        assert got[2] == want[2], sl
        assert np.array_equal(got[3], want[3]) and np.array_equal(got[4], want[4]), sl


def test_hostsim_fuzz_soak_vs_oracle():
    """The damaged-stream soak of tests/test_gpu_parity.py (cut short, flipped bits in data and header) on
    the CPU build of the kernel bodies: 46 seeds = about 5,000 streams of every layout against the oracle.
    (Found on this route: a band switch that topped the slot budget up instead of replacing it let a lane
    walk off the end of its band list on frames with several overrunning bands.)"""
    import multiprocessing as mp
    import os
    import test_gpu_parity as t
    with mp.get_context("fork").Pool(max(1, len(os.sched_getaffinity(0)))) as pool:
        streams = [s for part in pool.map(t._soak_make, range(46), chunksize=2) for s in part]
        streams = [s for s in streams if len(s[0]) >= 3 and ((s[0][0] << 8) | s[0][1]) > 0]
        chunks = [streams[i:i + 200] for i in range(0, len(streams), 200)]
        want = [c for part in pool.map(t._soak_expect, chunks) for c in part]
    pcm, offs, res, _, _ = simutil.decode_streams(streams)
    bad = [i for i in range(len(streams)) if res[i]["checksum"] != want[i]]
    assert len(streams) > 4500 and not bad, (len(bad), bad[:5])
