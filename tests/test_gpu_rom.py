"""GPU parity for ROM / track playback (BASELINE config 4): the product's sequencer + K1 scan +
K4 mix kernels through the C-ABI against the golden fixtures frozen from the reference, the
CPU-side simulator and (where oracle/_ref travelled with the snapshot) the reference itself."""
import os
import numpy as np
import pytest
from oracle import orc, ref
import rombuild as rb
import romscen
import simutil
from test_rom import check_rom_golden, images_digest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ctx(built):
    import dcsexplorer_b200 as dx
    c = dx.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def rom_golden():
    return np.load(os.path.join(HERE, "golden", "rom_golden.npz"))


@pytest.mark.parametrize("name,kw", romscen.SCENARIOS)
def test_gpu_timeline_matches_golden(ctx, rom_golden, name, kw):
    import dcsexplorer_b200 as dx
    sc = romscen.make_scenario(**kw)
    assert images_digest(sc["images"]) == str(rom_golden[name + "/digest"])
    rom = dx.Rom(sc["images"])
    pcm, res = ctx.render_timelines(rom, [(sc["writes"], sc["n_frames"], sc["master_volume"])])
    assert res[0]["status"] == 0 and res[0]["frames"] == sc["n_frames"]
    s = pcm[0].astype(np.uint16).astype(np.uint64)
    assert res[0]["checksum"] == int((s * (2 * np.arange(s.size, dtype=np.uint64) + 1)).sum(dtype=np.uint64))
    # host bytes come from the player interface (same sequencer)
    p = dx.Player(ctx, rom)
    p.set_master_volume(sc["master_volume"])
    out = np.zeros(sc["n_frames"] * 240, dtype=np.int16)
    frames = sorted(set([f for f, _ in sc["writes"]] + [0, sc["n_frames"]]))
    w = 0
    for a, b in zip(frames[:-1], frames[1:]):          # chunk boundaries where the host writes the port
        while w < len(sc["writes"]) and sc["writes"][w][0] <= a:
            p.write_data_port(sc["writes"][w][1])
            w += 1
        out[a * 240:b * 240] = p.render(b - a)
    assert np.array_equal(out, pcm[0]), "chunked player render differs from the one-shot timeline render"
    check_rom_golden(rom_golden, name, pcm[0], p.host_bytes())
    p.close()
    rom.close()


@pytest.mark.parametrize("name,kw", romscen.SCENARIOS)
def test_gpu_device_sequencer_equals_host_sequencer(ctx, monkeypatch, name, kw):
    """dcsb_render_timelines runs the track interpreter on the GPU (one thread per timeline, dcsb_seq_kernel);
    DCSB_SEQ_HOST=1 runs the same core on host threads.  Same PCM, checksums, status and host-byte counts, for
    many timelines with shifted command times, different volumes and lengths (so that the 32 instances of a warp
    are not in step)."""
    import dcsexplorer_b200 as dx
    sc = romscen.make_scenario(**kw)
    rom = dx.Rom(sc["images"])
    tls = [([(f + (i * 7) % 23, b) for f, b in sc["writes"]], sc["n_frames"] - (i % 5) * 17, 255 - (i * 3) % 120) for i in range(70)]
    tls.append(([], 40, 255))                                   # nothing ever happens
    tls.append(([(0, 0x7F), (0, 0x7F)] * 3, 30, 200))           # commands for tracks that do not exist
    pcm_d, res_d = ctx.render_timelines(rom, tls)
    monkeypatch.setenv("DCSB_SEQ_HOST", "1")
    pcm_h, res_h = ctx.render_timelines(rom, tls)
    monkeypatch.delenv("DCSB_SEQ_HOST")
    for i in range(len(tls)):
        assert np.array_equal(pcm_d[i], pcm_h[i]), (name, i)
        assert res_d[i] == res_h[i], (name, i, res_d[i], res_h[i])
    assert any(r["n_host_bytes"] for r in res_d) or name == "os93b"
    rom.close()


def _reference_timeline(images, tl):
    """one timeline (writes, n_frames, master_volume) rendered by the unmodified reference decoder"""
    rp = ref.RomPlayer(images, tl[2])
    want = rp.render_timeline(tl[0], tl[1])
    rp.close()
    return want


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref did not travel with the snapshot")
def test_gpu_many_timelines_vs_reference(ctx):
    """64 timelines on one ROM in one launch (different volumes / command times); a sample is
    compared with the reference decoder fed the same data-port bytes, every timeline's checksum
    with its own PCM, and two timelines that differ only in master volume with each other."""
    import dcsexplorer_b200 as dx
    sc = romscen.make_scenario(os_version=rb.OS95, seed=77, n_frames=300, version=0x0105)
    rom = dx.Rom(sc["images"])
    tls = []
    for i in range(64):
        shift = i % 5
        tls.append(([(f + shift, b) for f, b in sc["writes"]], 260 + (i % 7) * 11, 255 - 3 * (i % 32)))
    pcm, res = ctx.render_timelines(rom, tls)
    for i in (0, 1, 6, 33, 63):
        want = _reference_timeline(sc["images"], tls[i])
        bad = np.nonzero(pcm[i] != want)[0]
        assert bad.size == 0, "timeline %d: first differing frame %d" % (i, bad[0] // 240)
    for i in range(64):
        assert res[i]["status"] == 0 and res[i]["frames"] == tls[i][1]
        s = pcm[i].astype(np.uint16).astype(np.uint64)
        assert res[i]["checksum"] == int((s * (2 * np.arange(s.size, dtype=np.uint64) + 1)).sum(dtype=np.uint64)), i
    # timelines 0 and 35 share command times (shift 0) and length class but not the master volume
    n = min(tls[0][1], tls[35][1]) * 240
    assert tls[0][0] == tls[35][0] and tls[0][2] != tls[35][2] and not np.array_equal(pcm[0][:n], pcm[35][:n])
    rom.close()


@pytest.mark.parametrize("name", ["os94", "os95-v105", "os93a"])
def test_gpu_player_lookahead_is_frame_accurate(ctx, rom_golden, name):
    """dcsb_player_set_lookahead(16): frames are rendered 16 at a time, yet data-port bytes written between any
    two frames act on the frame they arrive at -- PCM, host bytes and IsStreamPlaying polls are those of
    frame-at-a-time rendering, and the PCM is the golden PCM of the reference."""
    import dcsexplorer_b200 as dx
    sc = romscen.make_scenario(**dict(romscen.SCENARIOS)[name])
    rom = dx.Rom(sc["images"])
    runs = []
    for look in (0, 16, 5):
        p = dx.Player(ctx, rom)
        assert ctx._L.dcsb_player_set_lookahead(p._h, look) == 0
        p.set_master_volume(sc["master_volume"])
        out = np.zeros(sc["n_frames"] * 240, dtype=np.int16)
        hb, playing, w = b"", [], 0
        for f in range(sc["n_frames"]):
            while w < len(sc["writes"]) and sc["writes"][w][0] <= f:
                p.write_data_port(sc["writes"][w][1])
                w += 1
            out[f * 240:(f + 1) * 240] = p.render(1)
            hb += p.host_bytes()
            playing.append(tuple(p.is_stream_playing(ch) for ch in range(6)))
        runs.append((out, hb, playing))
        p.close()
    for out, hb, playing in runs[1:]:
        assert np.array_equal(out, runs[0][0]) and hb == runs[0][1] and playing == runs[0][2]
    check_rom_golden(rom_golden, name, runs[1][0], runs[1][1])
    rom.close()


def test_gpu_player_load_audio_stream_equals_batch_decode(ctx):
    """LoadAudioStream on a player (DCSExplorer's stream extraction protocol) and the batch
    decode of the same stream bytes are two routes to the same PCM."""
    import dcsexplorer_b200 as dx
    for osv, seed in ((rb.OS94, 31), (rb.OS93B, 32), (rb.OS93A, 33)):
        sc = romscen.make_scenario(os_version=osv, seed=seed, n_frames=100)
        rom = dx.Rom(sc["images"])
        rom.check()
        for addr in rom.list_streams()[:5]:
            p = dx.Player(ctx, rom)
            p.set_master_volume(255)
            p.load_audio_stream(0, addr, 0x64)
            assert p.is_stream_playing(0)
            hdr = rom.stream_bytes(addr, 2)
            nf = (hdr[0] << 8) | hdr[1]
            got = p.render(nf + 2)
            assert not p.is_stream_playing(0)
            data = rom.stream_bytes(addr, 1 << 20)
            pcm, offs, res = ctx.decode_streams([(data, osv, 255, 0x64, 2)])
            assert np.array_equal(got, pcm[:got.size]), (hex(osv), hex(addr))
            want, _ = orc.decode(data[:res[0]["stream_bytes"] + 8], osv, 255, 0x64, nf + 2)
            assert np.array_equal(got, want)
            p.close()
        rom.close()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref did not travel with the snapshot")
@pytest.mark.parametrize("os_version,seed", [(rb.OS94, 301), (rb.OS95, 302), (rb.OS93B, 303), (rb.OS93A, 304)])
def test_gpu_timeline_vs_reference_fresh(ctx, os_version, seed):
    import dcsexplorer_b200 as dx
    sc = romscen.make_scenario(os_version=os_version, seed=seed, n_frames=600)
    rp = ref.RomPlayer(sc["images"], sc["master_volume"])
    want = rp.render_timeline(sc["writes"], sc["n_frames"])
    rom = dx.Rom(sc["images"])
    pcm, res = ctx.render_timelines(rom, [(sc["writes"], sc["n_frames"], sc["master_volume"])])
    bad = np.nonzero(pcm[0] != want)[0]
    assert bad.size == 0, "first differing frame %d" % (bad[0] // 240)
    rom.close()


def test_gpu_render_timelines_output_placements(ctx):
    """dcsb_render_timelines writes the same PCM whichever way the caller wants it: packed into a
    pageable buffer (blocking copy), packed into a page-locked buffer (asynchronous chunk-by-chunk
    download beside the sequencer threads), or at the caller's own offsets; several chunks
    (> 65 536 frames per chunk boundary) so that the pipeline really has more than one stage."""
    import ctypes as C
    import torch
    import dcsexplorer_b200 as dx
    sc = romscen.make_scenario(os_version=rb.OS94, seed=88, n_frames=400)
    rom = dx.Rom(sc["images"])
    tls = [([(f + i % 7, b) for f, b in sc["writes"]], 380 + (i % 5) * 9, 255 - (i % 50)) for i in range(520)]
    pcm, res = ctx.render_timelines(rom, tls)                     # pageable, packed
    total = sum(t[1] for t in tls)
    assert total > 3 * 65536 and all(r["status"] == 0 for r in res)
    tl_arr, keep = dx.make_timelines(tls)
    resarr = (dx.TimelineResult * len(tls))()
    pinned = torch.zeros(total * 240, dtype=torch.int16).pin_memory()
    assert ctx._L.dcsb_render_timelines(ctx._h, rom._h, tl_arr, len(tls), pinned.data_ptr(), None, resarr) == 0
    assert np.array_equal(pinned.numpy(), np.concatenate(pcm))
    assert [resarr[i].checksum for i in range(len(tls))] == [r["checksum"] for r in res]
    # caller-chosen offsets: timelines in reverse order with a gap of 100 samples between them
    offs = np.zeros(len(tls), dtype=np.uint64)
    o = 0
    for i in reversed(range(len(tls))):
        offs[i] = o
        o += tls[i][1] * 240 + 100
    out = np.full(o, 0x5A5A, dtype=np.int16)
    assert ctx._L.dcsb_render_timelines(ctx._h, rom._h, tl_arr, len(tls), out.ctypes.data, offs.ctypes.data, resarr) == 0
    for i in (0, 1, 257, len(tls) - 1):
        assert np.array_equal(out[int(offs[i]):int(offs[i]) + tls[i][1] * 240], pcm[i]), i
        assert (out[int(offs[i]) + tls[i][1] * 240:int(offs[i]) + tls[i][1] * 240 + 100] == 0x5A5A).all()
    # a sample against the reference decoder (or, without oracle/_ref, the golden-pinned CPU simulator)
    for i in (3, 519):
        want = _reference_timeline(sc["images"], tls[i]) if ref.available() else simutil.rom_render(sc["images"], [tls[i]])[0][0]
        assert np.array_equal(pcm[i], want), i
    rom.close()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref did not travel with the snapshot")
def test_gpu_device_sequencer_soak_vs_reference(ctx):
    """The GPU track interpreter against the unmodified reference decoder on freshly seeded ROM scenarios (all four
    OS versions, every third with malformed programs / damaged streams, some with the 1.05 opcode set): for each
    scenario a few timelines with shifted command times and different volumes in ONE dcsb_render_timelines call;
    PCM and host-byte counts must be the reference's.  (The CPU twin of this test, tools/soak_cpu.py rom, runs the
    same core on the host; this one is the net under code-generation differences of the device build.)"""
    import time
    import dcsexplorer_b200 as dx
    import simutil
    t0, n, checked = time.time(), 0, 0
    for k in range(160):
        if time.time() - t0 > 60 and n >= 24:      # (bounded in time, but never fewer than 24 scenarios)
            break
        osv = (rb.OS94, rb.OS95, rb.OS93B, rb.OS93A)[k % 4]
        sc = romscen.make_scenario(os_version=osv, seed=7000 + k, n_frames=240, with_errors=(k % 3 == 0),
                                   version=(0x0105 if (osv == rb.OS95 and k % 8 == 1) else None))
        rom = dx.Rom(sc["images"])
        tls = [([(f + sh, b) for f, b in sc["writes"]], sc["n_frames"] + sh, vol) for sh, vol in ((0, sc["master_volume"]), (3, 255), (11, 140))]
        pcm, res = ctx.render_timelines(rom, tls)
        for i, tl in enumerate(tls):
            rp = ref.RomPlayer(sc["images"], tl[2])
            want = rp.render_timeline(tl[0], tl[1])
            nhost = len(rp.host_bytes()) if hasattr(rp, "host_bytes") else None
            rp.close()
            assert np.array_equal(pcm[i], want), (k, hex(osv), i)
            if nhost is not None:
                assert res[i]["n_host_bytes"] == nhost, (k, i, res[i], nhost)
            checked += 1
        rom.close()
        n += 1
    assert n >= 24, n
