"""GPU parity tests proper: the CUDA path through the C-ABI vs the oracle and the golden
fixtures (bit-exact; the format is integer throughout)."""
import numpy as np
import pytest
from conftest import check_against_golden
from oracle import orc
import dcsfuzz

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    import dcsexplorer_b200 as dx
    c = dx.Context(0)
    yield c
    c.close()


def _expect(d, os_, vol, lvl, tail):
    nf = (d[0] << 8) | d[1]
    return orc.decode(d, os_, vol, lvl, nf + tail)


def test_gpu_matches_golden(ctx, golden):
    streams = [(it["stream"], it["os"], it["vol"], it["lvl"], it["nframes_out"] - ((it["stream"][0] << 8) | it["stream"][1]))
               for it in golden.items]
    pcm, offs, res = ctx.decode_streams(streams)
    for i, it in enumerate(golden.items):
        check_against_golden(it, pcm[offs[i]:offs[i] + it["nframes_out"] * 240])
        assert res[i]["status"] == (-5 if it["stop"] else 0)


def test_gpu_fuzz_all_layouts_vs_oracle(ctx):
    rng = np.random.default_rng(4242)
    streams = []
    for seed in range(10):
        for os_, d, label in dcsfuzz.corpus(seed + 300, n_each=3, nframes=int(rng.integers(1, 100))):
            streams.append((d, os_, int(rng.integers(0, 256)), int(rng.integers(0, 256)), int(rng.integers(0, 5))))
    for i in range(8):
        streams.append((dcsfuzz.fuzz94(rng, 50, type1=i & 1, max_code=6, error_frame=int(rng.integers(0, 50)), escape_p=0.2),
                        0x9400, 255, 0x64, 2))
    pcm, offs, res = ctx.decode_streams(streams)
    for i, (d, os_, vol, lvl, tail) in enumerate(streams):
        want, rc = _expect(d, os_, vol, lvl, tail)
        assert np.array_equal(pcm[offs[i]:offs[i] + want.size], want), (i, hex(os_))
        s = want.astype(np.uint16).astype(np.uint64)
        assert res[i]["checksum"] == int((s * (2 * np.arange(s.size, dtype=np.uint64) + 1)).sum(dtype=np.uint64))


def test_gpu_scan_checkpoints_vs_oracle(ctx):
    rng = np.random.default_rng(9)
    streams = [(d, os_, 255, 100, 2) for os_, d, _ in dcsfuzz.corpus(901, n_each=2, nframes=45)]
    b = ctx.batch(streams)
    b.decode()
    res = b.results()
    for i, (d, os_, *_r) in enumerate(streams):
        nf = (d[0] << 8) | d[1]
        bp, bt = b.read_scan(i, nf)
        rc, obp, obt, ostop = orc.scan(d, os_)
        assert np.array_equal(bp, obp[:nf]) and np.array_equal(bt, obt[:nf])
        assert res[i]["frames_decoded"] == nf
    b.close()


def test_gpu_edge_cases(ctx):
    empty = bytes([0, 0] + [0x10] * 16)
    short = bytes([0, 3, 0x10])
    nobands = bytes([0, 4] + [0x7F] * 16)
    ok = dcsfuzz.fuzz94(np.random.default_rng(1), 31 * 3, type1=1)
    one = dcsfuzz.fuzz94(np.random.default_rng(2), 1, type1=0)
    trunc = ok[: len(ok) // 2]
    streams = [(empty, 0x9400, 255, 100, 2), (short, 0x9400, 255, 100, 2), (nobands, 0x9400, 255, 100, 1),
               (ok, 0x9400, 255, 100, 0), (one, 0x9400, 255, 100, 3), (trunc, 0x9400, 255, 100, 2)]
    pcm, offs, res = ctx.decode_streams(streams)
    assert [r["status"] for r in res[:5]] == [-1, -4, 0, 0, 0]
    for i, (d, os_, vol, lvl, tail) in enumerate(streams):
        if i == 1:
            continue
        want, rc = _expect(d, os_, vol, lvl, tail)
        assert np.array_equal(pcm[offs[i]:offs[i] + want.size], want), i
    pcm, offs, res = ctx.decode_streams([])
    assert pcm.size == 0 and res == []


def test_gpu_resident_batch_is_idempotent(ctx):
    """Decoding the same resident batch twice gives identical PCM and checksums."""
    streams = [(d, os_, 255, 100, 2) for os_, d, _ in dcsfuzz.corpus(55, n_each=2, nframes=64)]
    b = ctx.batch(streams)
    b.decode()
    r1 = b.results()
    p1 = [b.read_pcm(i, r1[i]["frames"] * 240) for i in range(len(streams))]
    b.decode()
    r2 = b.results()
    p2 = [b.read_pcm(i, r2[i]["frames"] * 240) for i in range(len(streams))]
    assert r1 == r2
    assert all(np.array_equal(a, c) for a, c in zip(p1, p2))
    b.close()


def test_gpu_many_streams_checksum_of_checksums(ctx):
    """Larger batch: per-stream device checksums must equal the checksum of the oracle's PCM
    for a sample of streams, and replicas of a stream must agree with each other."""
    rng = np.random.default_rng(31)
    base = [dcsfuzz.fuzz94(rng, 120, type1=1) for _ in range(8)]
    streams = [(base[i % 8], 0x9400, 255, 100, 2) for i in range(2048)]
    b = ctx.batch(streams)
    b.decode()
    res = b.results()
    for i in range(8):
        want, _ = _expect(base[i], 0x9400, 255, 100, 2)
        s = want.astype(np.uint16).astype(np.uint64)
        cs = int((s * (2 * np.arange(s.size, dtype=np.uint64) + 1)).sum(dtype=np.uint64))
        assert all(res[j]["checksum"] == cs for j in range(i, 2048, 8))
    b.close()
