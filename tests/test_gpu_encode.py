"""GPU forward path (dcsb_encode_streams) against the reference's DCSEncoder fed the same framing
(oracle/ref_shim.cpp: dcsref_encode_framed -- the reference minus its resampler): the transformed frames and the
stream BYTES must be the reference's; the streams must decode, and decode to the same PCM through our decoder and the
reference's."""
import numpy as np
import pytest
from oracle import ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    import dcsexplorer_b200 as dx
    c = dx.Context(0)
    yield c
    c.close()


def test_gpu_encode_matches_the_committed_reference_hashes(ctx):
    """tests/golden/encode_golden.json (made by make_encode_golden.py from the unmodified reference encoder in the build
    container): SHA-256 of the stream of every case -- every format the reference writes, wildcards included.  Needs
    nothing but the fixture on the GPU box."""
    import hashlib
    import json
    import os
    import sys
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, here)
    import make_encode_golden as mg
    cases = json.load(open(os.path.join(here, "encode_golden.json")))
    assert len(cases) == len(mg.CASES) >= 48
    clips = [mg.clip(c["seed"]) for c in cases]
    params = [(c["type"], c["subtype"], c["bit_rate"], c["power_cut"], c["max_err"], c["min_range"], c["fmt"]) for c in cases]
    streams = ctx.encode_streams(clips, params)
    bad = [(c["seed"], hex(c["fmt"]), c["type"]) for c, s_ in zip(cases, streams)
           if len(s_) != c["n_bytes"] or hashlib.sha256(s_).hexdigest() != c["sha256"]]
    assert not bad, bad


def _clips():
    import bench
    rng = np.random.default_rng(11)
    clips = [bench.synth_source(40 + i, 1.5 + 0.37 * i) for i in range(6)]
    clips.append(np.zeros(1000, dtype=np.float32))                                  # silence
    clips.append((0.9 * np.sin(np.arange(5000) * 0.05)).astype(np.float32))         # loud tone
    clips.append((rng.standard_normal(240 * 7) * 0.3).astype(np.float32))           # a whole number of frames
    clips.append((rng.standard_normal(17) * 0.01).astype(np.float32))               # less than one frame
    return clips


PARAMS = [(0, 0, 128000, 0.97), (0, 3, 64000, 0.90), (1, 0, 128000, 0.97), (1, 3, 256000, 1.0), (1, 3, 32000, 0.97), (0, 0, 96000, 1.0)]


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref did not travel with the snapshot")
def test_gpu_encode_frames_and_bytes_equal_the_reference(ctx):
    clips = _clips()
    jobs = [(c, p) for c in clips for p in PARAMS]
    streams, frames = ctx.encode_streams([j[0] for j in jobs], [j[1] for j in jobs], want_frames=True)
    nbad_frames = nbad_bytes = 0
    for (clip, p), got, fr in zip(jobs, streams, frames):
        want, nf, wfr = ref.encode_framed(clip, p[0], p[1], p[2], p[3], want_frames=True)
        assert fr.shape[0] == nf
        if not np.array_equal(fr.view(np.uint32), wfr.view(np.uint32)):
            nbad_frames += 1
        if got != want:
            nbad_bytes += 1
    assert nbad_frames == 0, "%d of %d clips: transformed frames differ from the reference's (bit pattern)" % (nbad_frames, len(jobs))
    assert nbad_bytes == 0, "%d of %d streams differ from the reference encoder's bytes" % (nbad_bytes, len(jobs))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref did not travel with the snapshot")
def test_gpu_encoded_streams_decode_like_the_reference(ctx):
    clips = _clips()[:6]
    params = [PARAMS[i % len(PARAMS)] for i in range(len(clips))]
    streams = ctx.encode_streams(clips, params)
    pcm, offs, res = ctx.decode_streams([(s, 0x9400, 255, 0x64, 2) for s in streams])
    for i, s in enumerate(streams):
        assert res[i]["status"] == 0
        want = ref.decode(s)
        assert np.array_equal(pcm[offs[i]:offs[i] + want.size], want)
        # and it is the clip (the codec delays it by its 16-sample overlap; lossy, so only loosely): best correlation
        # with the source over small lags
        n = min(len(clips[i]), want.size) - 64
        a = clips[i][:n].astype(np.float64)
        best = max(np.corrcoef(a, want[lag:lag + n].astype(np.float64))[0, 1] for lag in range(0, 48))
        assert best > 0.5, (i, best)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref did not travel with the snapshot")
def test_gpu_encode_random_clips_and_parameters_equal_the_reference(ctx):
    """seeded clips of several kinds (noise, tones, sweeps, bursts, near-silence, clipped) with random stream types, bit rates
    8k..512k, power cuts, quantisation-error and dynamic-range thresholds: every stream byte-identical to the reference's"""
    rng = np.random.default_rng(2024)
    clips, params = [], []
    for i in range(160):
        n = int(rng.integers(100, 40000))
        t = np.arange(n)
        kind = i % 6
        if kind == 0:
            x = rng.standard_normal(n) * rng.uniform(0.001, 0.5)
        elif kind == 1:
            x = rng.uniform(0.05, 1.0) * np.sin(t * rng.uniform(0.001, 3.0))
        elif kind == 2:
            x = 0.5 * np.sin(t * t * rng.uniform(1e-6, 1e-4))
        elif kind == 3:
            x = (rng.standard_normal(n) * 0.4) * (np.sin(t * 0.002) > 0.7)
        elif kind == 4:
            x = rng.standard_normal(n) * 1e-4
        else:
            x = np.clip(rng.standard_normal(n) * 1.5, -1.0, 1.0)
        clips.append(x.astype(np.float32))
        ty = int(rng.integers(0, 2))
        params.append((ty, int(rng.choice([0, 3])), int(rng.choice([8000, 32000, 64000, 128000, 256000, 512000])),
                       float(rng.choice([0.5, 0.9, 0.97, 1.0])), float(rng.choice([1.0, 10.0, 100.0])) / 32768.0,
                       float(rng.choice([0.0, 10.0, 200.0])) / 32768.0))
    streams = ctx.encode_streams(clips, params)
    bad = []
    for i, (c, p) in enumerate(zip(clips, params)):
        want, nf = ref.encode_framed(c, p[0], p[1], p[2], p[3], p[4], p[5])
        if streams[i] != want:
            bad.append((i, p, len(streams[i]), len(want)))
    assert not bad, "%d of %d streams differ from the reference encoder's: %s" % (len(bad), len(clips), bad[:5])


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref did not travel with the snapshot")
def test_gpu_encode_wildcard_stream_type_picks_like_the_reference(ctx):
    """-1 as stream type / subtype: every matching format is tried and the first of the smallest streams kept
    (CloseStream, DCSEncoder.cpp:779-836)"""
    clips = _clips()[:8]
    jobs = [(c, p) for c in clips for p in ((-1, -1, 128000, 0.97), (-1, 3, 64000, 0.9), (1, -1, 96000, 1.0), (-1, 0, 256000, 0.97))]
    streams = ctx.encode_streams([j[0] for j in jobs], [j[1] for j in jobs])
    for (clip, p), got in zip(jobs, streams):
        want, nf = ref.encode_framed(clip, p[0], p[1], p[2], p[3])
        assert got == want, p


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref did not travel with the snapshot")
def test_gpu_encode_1993_layouts_equal_the_reference(ctx):
    """format versions $9302 / $9301 (CompressFrame93b), stream type 0 and -- $9302 -- type 1 and the wildcard: bytes
    identical to the reference's, and the streams decode through our 1993 path like the oracle decodes them"""
    from oracle import orc
    rng = np.random.default_rng(93)
    clips = _clips() + [(rng.standard_normal(int(rng.integers(300, 20000))) * rng.uniform(0.001, 0.6)).astype(np.float32) for _ in range(40)]
    jobs = []
    for i, c in enumerate(clips):
        fmt = 0x9302 if i % 2 else 0x9301
        ty = (0, 1, -1)[(i // 2) % 3] if fmt == 0x9302 else 0
        jobs.append((c, (ty, 0, int(rng.choice([32000, 96000, 128000, 256000])), float(rng.choice([0.9, 0.97, 1.0])),
                         float(rng.choice([1.0, 10.0, 100.0])) / 32768.0, 10.0 / 32768.0, fmt)))
    streams = ctx.encode_streams([j[0] for j in jobs], [j[1] for j in jobs])
    bad = []
    for i, ((clip, p), got) in enumerate(zip(jobs, streams)):
        want, nf = ref.encode_framed(clip, p[0], p[1], p[2], p[3], p[4], p[5], fmt=p[6])
        if got != want:
            bad.append((i, hex(p[6]), len(got), len(want)))
    assert not bad, "%d of %d streams differ from the reference encoder's: %s" % (len(bad), len(jobs), bad[:5])
    pcm, offs, res = ctx.decode_streams([(s, j[1][6], 255, 0x64, 2) for s, j in zip(streams, jobs)])
    for i, s_ in enumerate(streams):
        want, rc = orc.decode(s_, jobs[i][1][6], 255, 0x64, ((s_[0] << 8) | s_[1]) + 2)
        assert res[i]["status"] == 0 and np.array_equal(pcm[offs[i]:offs[i] + want.size], want), i


def test_gpu_encode_rejects_bad_arguments(ctx):
    import dcsexplorer_b200 as dx
    clip = np.zeros(480, dtype=np.float32)
    with pytest.raises(dx.DcsbError):
        ctx.encode_streams([clip], [(2, 0, 128000, 0.97)])
    with pytest.raises(dx.DcsbError):
        ctx.encode_streams([clip], [(0, 1, 128000, 0.97)])
    with pytest.raises(dx.DcsbError):
        ctx.encode_streams([clip], [(-2, 0, 128000, 0.97)])
    with pytest.raises(dx.DcsbError):
        ctx.encode_streams([clip], [(1, 0, 128000, 0.97, 10 / 32768, 10 / 32768, 0x9301)])      # no encoder for OS93a type 1
    with pytest.raises(dx.DcsbError):
        ctx.encode_streams([clip], [(0, 0, 128000, 0.97, 10 / 32768, 10 / 32768, 0x9500)])
    with pytest.raises(dx.DcsbError):
        ctx.encode_streams([np.zeros(0, dtype=np.float32)], [(0, 0, 128000, 0.97)])
    with pytest.raises(dx.DcsbError):
        ctx.encode_streams([np.zeros(240 * 65536 + 1, dtype=np.float32)], [(0, 0, 128000, 0.97)])
