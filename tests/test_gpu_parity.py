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


def _soak_make(seed):
    """one seed's worth of soak streams: fuzzer-made streams of every layout plus damaged copies
    (cut short, one or two flipped bits in the frame data, a flipped bit in the stream header)"""
    rng = np.random.default_rng(70000 + seed)
    out = []
    for os_, d, _ in dcsfuzz.corpus(70000 + seed, n_each=2, nframes=int(rng.integers(6, 40))):
        vol, lvl, tail = int(rng.integers(0, 256)), int(rng.integers(0, 256)), int(rng.integers(0, 4))
        out.append((d, os_, vol, lvl, tail))
        hl = 3 if len(d) < 18 else 18
        for k in range(7):
            b = bytearray(d)
            if k < 2 and len(b) > hl + 2:
                b = b[:int(rng.integers(hl + 1, len(b)))]                       # cut short
            elif k < 5 and len(b) > hl + 1:
                for _ in range(1 + (k == 4)):
                    i = int(rng.integers(hl, len(b)))
                    b[i] ^= 1 << int(rng.integers(0, 8))                        # flipped bit(s) in the frame data
            else:
                i = int(rng.integers(2, min(hl, len(b))))
                b[i] ^= 1 << int(rng.integers(0, 7))                            # flipped bit in the stream header
            out.append((bytes(b), os_, vol, lvl, tail))
    return out


def _soak_expect(chunk):
    res = []
    for d, os_, vol, lvl, tail in chunk:
        nf = ((d[0] << 8) | d[1]) if len(d) >= 2 else 0
        pcm, rc = orc.decode(d, os_, vol, lvl, nf + tail)
        u = pcm.view(np.uint16).astype(np.uint64)
        res.append(int((u * (2 * np.arange(u.size, dtype=np.uint64) + 1)).sum(dtype=np.uint64)))
    return res


def test_gpu_fuzz_soak_50k_streams_vs_oracle(ctx):
    """One batch of more than 50,000 fuzzer-made streams of every layout, most of them damaged, decoded on
    the GPU; every stream's checksum against the checksum of the oracle's PCM.  This is the net under
    code-generation bugs of the device build (the CPU-side soak cannot see them)."""
    import multiprocessing as mp
    import os
    ncpu = max(1, len(os.sched_getaffinity(0)))
    with mp.get_context("fork").Pool(ncpu) as pool:
        streams = [s for part in pool.map(_soak_make, range(460), chunksize=4) for s in part]
        # streams the host rejects up front (no frames / too short) are not part of this test's point
        streams = [s for s in streams if len(s[0]) >= 3 and ((s[0][0] << 8) | s[0][1]) > 0]
        assert len(streams) >= 50000
        chunks = [streams[i:i + 500] for i in range(0, len(streams), 500)]
        want = [c for part in pool.map(_soak_expect, chunks) for c in part]
    pcm, offs, res = ctx.decode_streams_pinned(streams)
    bad = [i for i in range(len(streams)) if res[i]["checksum"] != want[i]]
    assert not bad, "%d of %d streams differ from the oracle; first: stream %d os %#x status %d" % (
        len(bad), len(streams), bad[0], streams[bad[0]][1], res[bad[0]]["status"])
    # and the PCM itself on a sample (the checksum is computed by the kernel from what it stores)
    rng = np.random.default_rng(5)
    for i in rng.integers(0, len(streams), 200):
        d, os_, vol, lvl, tail = streams[int(i)]
        w, _ = _expect(d, os_, vol, lvl, tail)
        assert np.array_equal(pcm[offs[i]:offs[i] + w.size], w), int(i)


def test_gpu_scan_direct_variant_vs_oracle(ctx, monkeypatch):
    """The scan variant multi-wave batches use (stream words read from global memory through L1 instead
    of shared-memory rings, up to eight warps per CTA), forced onto a small batch: same checkpoints, same
    PCM as the oracle -- valid streams of every 1994 flavour, damaged ones, and the other layouts riding
    along in the same warps."""
    monkeypatch.setenv("DCSB_SCAN_DIRECT", "1")
    monkeypatch.setenv("DCSB_SCAN_WARPS", "8")
    streams = []
    for seed in range(12):
        streams += [s for s in _soak_make(900 + seed) if len(s[0]) >= 3 and ((s[0][0] << 8) | s[0][1]) > 0]
    pcm, offs, res = ctx.decode_streams(streams)
    for i, (d, os_, vol, lvl, tail) in enumerate(streams):
        want, rc = _expect(d, os_, vol, lvl, tail)
        assert np.array_equal(pcm[offs[i]:offs[i] + want.size], want), (i, hex(os_))
    b = ctx.batch(streams[:300])
    b.decode()
    bres = b.results()
    for i, (d, os_, *_r) in enumerate(streams[:300]):
        rc, obp, obt, ostop = orc.scan(d, os_)
        nf = (d[0] << 8) | d[1]
        if rc == nf and bres[i]["status"] == 0:         # every frame decodable: compare the checkpoints
            bp, bt = b.read_scan(i, nf)
            assert np.array_equal(bp, obp[:nf]) and np.array_equal(bt, obt[:nf]), i
    b.close()


def test_gpu_decode_streams_refuses_wrap_empty_flag(ctx):
    """DCSB_STREAM_WRAP_EMPTY would make a zero-count stream render 65,536 frames into a buffer the caller
    sized from the count as written: dcsb_decode_streams refuses the flag instead of overrunning pcm_out."""
    import dcsexplorer_b200 as dx
    ok = dcsfuzz.fuzz94(np.random.default_rng(3), 5, type1=1)
    empty = bytes([0, 0] + [0x10] * 16) + bytes(64)
    descs, keep = dx.make_descs([(ok, 0x9400, 255, 100, 2), (empty, 0x9400, 255, 100, 2)])
    descs[1].reserved = 1
    pcm = np.full(2 * 7 * 240 + 64, 0x1234, dtype=np.int16)
    res = (dx.Result * 2)()
    rc = ctx._L.dcsb_decode_streams(ctx._h, descs, 2, pcm.ctypes.data, None, res)
    assert rc == dx.E_ARG and (pcm == 0x1234).all()
    descs[1].reserved = 0
    assert ctx._L.dcsb_decode_streams(ctx._h, descs, 2, pcm.ctypes.data, None, res) == 0
    assert res[1].status == dx.E_EMPTY and (pcm[9 * 240:] == 0x1234).all()


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


def _many_streams(seed, n, nframes):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        for os_, d, _ in dcsfuzz.corpus(int(rng.integers(0, 1 << 30)), n_each=1, nframes=nframes):
            out.append((d, os_, 255, 0x64, 2))
    return out[:n]


def test_gpu_decode_streams_pipeline_lanes_pinned_in_place(ctx):
    """Enough PCM that dcsb_decode_streams splits the batch over several pipeline lanes; inputs in
    one pinned blob (uploaded in place, arbitrary byte alignment) and a pinned output buffer
    (PCM lands by DMA), against the same call on pageable memory and against the oracle."""
    import ctypes as C
    import torch
    import dcsexplorer_b200 as dx
    streams = _many_streams(11, 520, 1000)           # ~125 M samples -> 2+ lanes
    descs, keep = dx.make_descs(streams)
    n = len(streams)
    blob = torch.empty(sum(len(s[0]) for s in streams) + n, dtype=torch.uint8).pin_memory()
    bnp = blob.numpy()
    off = 0
    for i, s in enumerate(streams):                  # odd gaps: streams start at arbitrary byte offsets
        bnp[off:off + len(s[0])] = np.frombuffer(s[0], dtype=np.uint8)
        descs[i].data = blob.data_ptr() + off
        off += len(s[0]) + (i & 1)
    total = sum((((s[0][0] << 8) | s[0][1]) + 2) * 240 for s in streams)
    h_pcm = torch.zeros(total, dtype=torch.int16).pin_memory()
    res = (dx.Result * n)()
    for _ in range(2):                               # second call reuses the lanes' buffers
        rc = ctx._L.dcsb_decode_streams(ctx._h, descs, n, h_pcm.data_ptr(), None, res)
        assert rc == 0, ctx._L.dcsb_last_error(ctx._h)
    pcm2, offs, res2 = ctx.decode_streams(streams)   # pageable in / out, packed by the library
    assert np.array_equal(h_pcm.numpy(), pcm2)
    for i in range(n):
        assert res[i].checksum == res2[i]["checksum"] and res[i].status == res2[i]["status"]
    for i in list(range(0, n, 37)) + [n - 1]:
        d, os_, vol, lvl, tail = streams[i]
        want, rc = _expect(d, os_, vol, lvl, tail)
        assert np.array_equal(pcm2[offs[i]:offs[i] + want.size], want), (i, hex(os_))


def test_gpu_decode_streams_custom_offsets(ctx):
    import ctypes as C
    import dcsexplorer_b200 as dx
    streams = _many_streams(5, 9, 20)
    descs, keep = dx.make_descs(streams)
    n = len(streams)
    sizes = [(((s[0][0] << 8) | s[0][1]) + 2) * 240 for s in streams]
    offs = np.zeros(n, dtype=np.uint64)
    pos = 7
    for i in reversed(range(n)):                     # reversed order with gaps
        offs[i] = pos
        pos += sizes[i] + 13
    out = np.full(pos, 0x5A5A, dtype=np.int16)
    res = (dx.Result * n)()
    rc = ctx._L.dcsb_decode_streams(ctx._h, descs, n, out.ctypes.data, offs.ctypes.data, res)
    assert rc == 0
    for i, (d, os_, vol, lvl, tail) in enumerate(streams):
        want, _ = _expect(d, os_, vol, lvl, tail)
        assert np.array_equal(out[int(offs[i]):int(offs[i]) + want.size], want), i
    assert out[0] == 0x5A5A and out[int(offs[0]) + sizes[0]] == 0x5A5A


def test_gpu_config5_shape_many_short_streams(ctx):
    """BASELINE config 5 in miniature: tens of thousands of ~1 s streams (a pool replicated so that
    every replica has its own bytes in HBM).  More streams than one scan wave holds, so the scan
    CTAs walk stream groups grid-stride while the persistent decode kernel drains the ready queue.
    Size-independent checks: replicas agree, the checksum of checksums matches the oracle's."""
    rng = np.random.default_rng(2026)
    pool = [dcsfuzz.fuzz94(rng, int(rng.integers(100, 140)), type1=i % 3 != 0, max_code=15 if i % 3 == 0 else 9) for i in range(48)]
    pool += [dcsfuzz.fuzz94(rng, 60, type1=1, max_code=6, error_frame=31, escape_p=0.2), bytes([0, 0] + [0x10] * 16)]
    n = 24000
    streams = [(pool[i % len(pool)], 0x9400, 255, 100, 2) for i in range(n)]
    want = []
    for d in pool:
        pcm, rc = _expect(d, 0x9400, 255, 100, 2)
        s = pcm.astype(np.uint16).astype(np.uint64)
        want.append(int((s * (2 * np.arange(s.size, dtype=np.uint64) + 1)).sum(dtype=np.uint64)))
    for overlap in (True, False):
        ctx.set_overlap(overlap)
        b = ctx.batch(streams)
        b.decode()
        res = b.results()
        for i in range(n):
            assert res[i]["checksum"] == want[i % len(pool)], (overlap, i)
        assert res[len(pool) - 2]["status"] == -5 and res[len(pool) - 1]["status"] == -1
        k = 17 * len(pool) + 5
        pcm, _ = _expect(pool[5], 0x9400, 255, 100, 2)
        assert np.array_equal(b.read_pcm(k, pcm.size), pcm)
        b.close()
    ctx.set_overlap(True)


def test_gpu_decode_streams_time_slices(ctx):
    """Uniform chunks (every stream renders the same number of frames into a packed pinned buffer)
    are cut in time: resumed scans + per-slice work items + strided PCM copies.  Same PCM, statuses
    and checksums as the one-pass form, for slice lengths that do and do not divide the streams,
    for all layouts, and with streams that fail in the middle of a slice."""
    rng = np.random.default_rng(99)
    nf = 150
    streams = []
    for k in range(12):
        streams.append((dcsfuzz.fuzz94(rng, nf, type1=k & 1, max_code=15 if k % 3 == 0 else 9), 0x9400, 255, 100, 2))
    streams.append((dcsfuzz.fuzz94(rng, nf, type1=1, max_code=6, error_frame=77, escape_p=0.2), 0x9400, 200, 64, 2))
    ok = dcsfuzz.fuzz94(rng, nf, type1=1)
    streams.append((ok[: len(ok) // 2], 0x9400, 255, 100, 2))                # truncated half way
    for os_, d, label in dcsfuzz.corpus(seed=3, n_each=1, nframes=nf):
        streams.append((d, os_, 220, 0x64, 2))
    ctx.set_pipeline(0, -1)
    want, offs, wres = ctx.decode_streams_pinned(streams)
    for i in (0, 5, 12, 13, len(streams) - 1):
        d, os_, vol, lvl, tail = streams[i]
        exp, _ = _expect(d, os_, vol, lvl, tail)
        assert np.array_equal(want[offs[i]:offs[i] + exp.size], exp), i
    try:
        for chunks, sl in ((1, 31), (1, 40), (2, 63), (1, 151), (3, 7)):
            ctx.set_pipeline(chunks, sl)
            got, offs2, res = ctx.decode_streams_pinned(streams)
            assert offs2 == offs
            assert np.array_equal(got, want), (chunks, sl)
            assert res == wres, (chunks, sl, [(i, a, b) for i, (a, b) in enumerate(zip(res, wres)) if a != b][:3])
    finally:
        ctx.set_pipeline(0, 0)


def test_gpu_decode_streams_time_slices_mixed_lengths(ctx):
    """Chunks whose streams differ in length are cut in time as well: one batched copy per slice with a
    piece per stream that reaches into it (short streams drop out of the later slices)."""
    rng = np.random.default_rng(123)
    streams = []
    for k in range(40):
        nf = int(rng.integers(1, 400))
        streams.append((dcsfuzz.fuzz94(rng, nf, type1=k & 1, max_code=15 if k % 3 == 0 else 9), 0x9400, 255, 100, int(rng.integers(0, 4))))
    for os_, d, label in dcsfuzz.corpus(seed=11, n_each=1, nframes=90):
        streams.append((d, os_, 220, 0x64, 2))
    streams.append((bytes([0, 0] + [0x10] * 16), 0x9400, 255, 100, 2))
    ctx.set_pipeline(0, -1)
    want, offs, wres = ctx.decode_streams_pinned(streams)
    for i in range(0, len(streams), 7):
        d, os_, vol, lvl, tail = streams[i]
        exp, _ = _expect(d, os_, vol, lvl, tail)
        assert np.array_equal(want[offs[i]:offs[i] + exp.size], exp), i
    try:
        for chunks, sl in ((1, 31), (2, 50), (3, 64), (1, 333)):
            ctx.set_pipeline(chunks, sl)
            got, offs2, res = ctx.decode_streams_pinned(streams)
            assert np.array_equal(got, want), (chunks, sl)
            assert res == wres, (chunks, sl, [(i, a, b) for i, (a, b) in enumerate(zip(res, wres)) if a != b][:3])
        # the caller's own offsets (reversed order, gaps) in a pinned buffer: same pieces, other destinations
        import torch
        import dcsexplorer_b200 as dx
        descs, keep = dx.make_descs(streams)
        n = len(streams)
        sizes = [(offs[i + 1] if i + 1 < n else want.size) - offs[i] for i in range(n)]
        coffs = np.zeros(n, dtype=np.uint64)
        pos = 5
        for i in reversed(range(n)):
            coffs[i] = pos
            pos += sizes[i] + 11
        out = torch.full((pos,), 0x5A5A, dtype=torch.int16).pin_memory()
        res2 = (dx.Result * n)()
        ctx.set_pipeline(2, 40)
        assert ctx._L.dcsb_decode_streams(ctx._h, descs, n, out.data_ptr(), coffs.ctypes.data, res2) == 0
        o = out.numpy()
        for i in range(n):
            assert np.array_equal(o[int(coffs[i]):int(coffs[i]) + sizes[i]], want[offs[i]:offs[i] + sizes[i]]), i
        assert o[0] == 0x5A5A and o[int(coffs[0]) + sizes[0]] == 0x5A5A
    finally:
        ctx.set_pipeline(0, 0)
