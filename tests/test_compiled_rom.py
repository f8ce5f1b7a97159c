"""BASELINE config 4 as BASELINE.json words it: "full synthetic ROM image built by DCSCompiler, all
tracks rendered with multi-channel mix and volume commands".  The ROM sets in
tests/golden/compiled_rom.npz were built by the reference's own compiler (DCSCompiler +
DCSEncoder, unmodified) from a generated 72-track script; the expected PCM is the unmodified
reference decoder's.  CPU: the product's ROM model through the C-ABI and the sequencer + kernel
bodies in the simulator.  GPU: dcsb_render_timelines, the main timeline and all 72 tracks in one call."""
import numpy as np
import pytest
import compiledrom
import simutil


@pytest.mark.parametrize("name", compiledrom.NAMES)
def test_compiled_rom_model_through_the_c_abi(built, name):
    import dcsexplorer_b200 as dx
    c = compiledrom.load(name)
    rom = dx.Rom(c["images"])
    assert rom.check() == 1
    info = rom.info()
    assert info["os"] == c["os"] and info["n_tracks"] == c["n_tracks"] == 72 and info["channels"] == 6
    assert "built by DCSCompiler" in info["signature"]
    assert rom.list_streams() == [int(a) for a in c["g"][name + "/streams"]]
    # GetTrackInfo of every track equals the reference's (valid, address, channel, type, deferCode, time, looping)
    for t, want in enumerate(c["g"][name + "/tracks"]):
        ti = rom.track_info(t)
        if not want[0]:
            assert ti is None, t
            continue
        assert [ti["address"], ti["channel"], ti["type"], ti["defer_code"], ti["time"], int(ti["looping"])] == [int(v) for v in want[1:]], hex(t)
    assert rom.track_info(c["n_tracks"]) is None
    rom.close()


@pytest.mark.parametrize("name", compiledrom.NAMES)
def test_compiled_rom_decompile_matches_reference(built, name):
    """DecompileTrackProgram (DCSDecoder.h:481): every step of all 72 track programs -- offsets, loop
    nesting and parents, delay counts, opcodes, operand bytes, the mnemonic and the hex text -- equals
    what the reference's decompiler produced (records frozen in the fixture, byte for byte)."""
    import dcsexplorer_b200 as dx
    c = compiledrom.load(name)
    rom = dx.Rom(c["images"])
    assert rom.check() == 1
    counts = c["g"][name + "/decompile_counts"]
    blob = c["g"][name + "/decompile"].tobytes()
    o = 0
    kinds = set()
    for t in range(c["n_tracks"]):
        raw, n = rom.decompile_track(t, raw=True)
        assert n == int(counts[t]), hex(t)
        want = blob[o:o + 128 * n]
        o += 128 * n
        if raw != want:
            for k in range(n):
                assert raw[128 * k:128 * k + 128] == want[128 * k:128 * k + 128], "track $%04X step %d: %r" % (t, k, want[128 * k + 24:128 * k + 88])
        for st in rom.decompile_track(t):
            kinds.add(st["opcode"])
    assert o == len(blob)
    assert {0x00, 0x01, 0x02, 0x03, 0x04, 0x05, 0x06, 0x07, 0x08, 0x09, 0x0A, 0x0B, 0x0C, 0x0E, 0x0F} <= kinds
    assert rom.decompile_track(0x35) == [] and rom.decompile_track(c["n_tracks"]) == []     # deferred track / no such track
    st = rom.decompile_track(0x2F)[1]
    assert st["desc"].startswith("Play(channel 4,stream $") and st["desc"].endswith(", repeat 2);") and st["opcode"] == 1
    rom.close()


@pytest.mark.parametrize("name", compiledrom.NAMES)
def test_compiled_rom_sim_matches_reference(built, name):
    c = compiledrom.load(name)
    pcm, res, info, hb = simutil.rom_render(c["images"], [(c["writes"], c["n_frames"], c["master_volume"])])
    assert res[0]["status"] == 0 and info["post"] == 1
    compiledrom.check_main(c, pcm[0], hb)
    pcms, res, _, hb = simutil.rom_render(c["images"], c["track_timelines"])
    compiledrom.check_tracks(c, pcms)
    assert hb == b"".join(c["track_host"])
    assert [r["n_host_bytes"] for r in res] == [len(h) for h in c["track_host"]]


@pytest.fixture(scope="module")
def ctx(built):
    import dcsexplorer_b200 as dx
    c = dx.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", compiledrom.NAMES)
def test_gpu_compiled_rom_all_tracks_and_timeline(ctx, name):
    import dcsexplorer_b200 as dx
    c = compiledrom.load(name)
    rom = dx.Rom(c["images"])
    tls = [(c["writes"], c["n_frames"], c["master_volume"])] + c["track_timelines"]
    pcm, res = ctx.render_timelines(rom, tls)               # 73 decoder instances, one call
    assert all(r["status"] == 0 for r in res)
    compiledrom.check_main(c, pcm[0], None)
    compiledrom.check_tracks(c, pcm[1:])
    assert [r["n_host_bytes"] for r in res[1:]] == [len(h) for h in c["track_host"]]
    # the single-instance player interface on the same timeline, host bytes included
    p = dx.Player(ctx, rom)
    p.set_master_volume(c["master_volume"])
    out = np.zeros(c["n_frames"] * 240, dtype=np.int16)
    frames = sorted(set([f for f, _ in c["writes"]] + [0, c["n_frames"]]))
    w = 0
    for a, b in zip(frames[:-1], frames[1:]):
        while w < len(c["writes"]) and c["writes"][w][0] <= a:
            p.write_data_port(c["writes"][w][1])
            w += 1
        out[a * 240:b * 240] = p.render(b - a)
    compiledrom.check_main(c, out, p.host_bytes())
    p.close()
    rom.close()
