"""The DCSDecoder-compatible C++ front (include/DCSDecoderB200.h) used from a C++ client the way
the reference's clients use DCSDecoderNative: CPU = it builds, links and refuses to run without a
GPU; GPU = ROM zip in, data-port timeline in, one GetNextSample() per sample out, against the
golden PCM frozen from the reference."""
import os
import subprocess
import zipfile
import numpy as np
import pytest
import rombuild as rb
import romscen
from test_rom import check_rom_golden

HERE = os.path.dirname(os.path.abspath(__file__))
DEMO = os.path.join(HERE, "cpp", "decoder_b200_demo")


@pytest.fixture(scope="module")
def demo(built):
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "cpp")])
    return DEMO


def _write_inputs(tmp_path, sc):
    z = tmp_path / "rom_2.zip"
    with zipfile.ZipFile(z, "w", zipfile.ZIP_DEFLATED) as f:
        for chip, img in sc["images"].items():
            f.writestr("snd_u%d.rom" % chip, img)
    tl = tmp_path / "timeline.txt"
    tl.write_text("".join("%d %d\n" % w for w in sc["writes"]))
    return z, tl


def test_cpp_front_builds_and_has_no_cpu_fallback(demo, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    sc = romscen.make_scenario(os_version=rb.OS94, seed=101)
    z, tl = _write_inputs(tmp_path, sc)
    r = subprocess.run([demo, str(z), str(tl), "10", "255", "1", str(tmp_path / "o.pcm")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name,chunk", [("os94", 1), ("os95-v105", 1), ("os93a", 1)])
def test_cpp_front_matches_golden(demo, tmp_path, name, chunk):
    g = np.load(os.path.join(HERE, "golden", "rom_golden.npz"))
    sc = romscen.make_scenario(**dict(romscen.SCENARIOS)[name])
    z, tl = _write_inputs(tmp_path, sc)
    out = tmp_path / "o.pcm"
    r = subprocess.run([demo, str(z), str(tl), str(sc["n_frames"]), str(sc["master_volume"]), str(chunk), str(out)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    pcm = np.fromfile(out, dtype=np.int16)
    line = r.stdout.splitlines()[0]
    hb = bytes(int(x, 16) for x in line.split("host bytes")[1].split())
    check_rom_golden(g, name, pcm, hb)
    assert "6 channels" in line and ("DCS-95" in line) == (sc["os"] == rb.OS95)
    assert "track 0: type 1 channel 0" in r.stdout and "stream $" in r.stdout
    # ExplainTrackProgram / DecompileTrackProgram: track 0 = SetMixingLevel(level 100); Play(...); wait forever
    assert "  SetMixingLevel(level 100);" in r.stdout and "Play(stream $" in r.stdout and "Wait(Forever) " in r.stdout
    assert "track 0: 3 steps" in r.stdout


@pytest.mark.gpu
def test_cpp_front_chunked_latency(demo, tmp_path):
    """chunkFrames = 16 (rendered ahead on the GPU): the same audio as frame-at-a-time rendering wherever the
    data-port bytes arrive -- an input takes effect at its own frame, as in the reference."""
    sc = romscen.make_scenario(os_version=rb.OS94, seed=55, n_frames=320)
    z, tl = _write_inputs(tmp_path, sc)
    outs = []
    for chunk in (1, 16):
        out = tmp_path / ("o%d.pcm" % chunk)
        r = subprocess.run([demo, str(z), str(tl), "320", "200", str(chunk), str(out)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append(np.fromfile(out, dtype=np.int16))
    assert outs[0].any() and np.array_equal(outs[0], outs[1])


@pytest.mark.gpu
def test_cpp_front_standalone_stream_protocol(demo, tmp_path, golden):
    """InitStandalone -> SoftBoot -> SetMasterVolume -> LoadAudioStream(0, ROMPointer(0, data, n), level) -> GetNextSample:
    the reference's stream-extraction protocol through the C++ front, against the golden PCM of every layout."""
    from oracle import ref, orc
    seen = set()
    for k, it in enumerate(golden.items):
        if it["os"] in seen or it["stop"] or it["nframes_out"] * 240 != it["pcm"].size:
            continue
        seen.add(it["os"])
        src = tmp_path / ("s%d.bin" % k)
        src.write_bytes(it["stream"] + bytes(64))
        out = tmp_path / ("s%d.pcm" % k)
        nfr = it["nframes_out"] + 3
        r = subprocess.run([demo, "--standalone", "%x" % it["os"], str(src), str(it["vol"]), str(it["lvl"]), str(nfr), str(out)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        pcm = np.fromfile(out, dtype=np.int16)
        want = it["pcm"]
        n = min(pcm.size, want.size)
        assert np.array_equal(pcm[:n], want[:n]), it["label"]
        assert not pcm[want.size:].any()                        # silence behind the stream and its overlap tail
        nf = (it["stream"][0] << 8) | it["stream"][1]
        assert ("standalone: %d frames" % nf) in r.stdout and ("playing for %d samples" % (nf * 240)) in r.stdout
    assert len(seen) >= 3


@pytest.mark.gpu
def test_cpp_front_standalone_unsized_pointer(demo, tmp_path, golden):
    """LoadAudioStream(0, ROMPointer(0, data), level) WITHOUT a size -- what the reference's own client passes
    (DCSEncoder.cpp:553) -- and GetStreamInfo on it: the front finds the extent by walking the frames on the
    host (dcsb_stream_extent), then decodes on the GPU; same PCM, and the reference's stream size."""
    import dcsexplorer_b200 as dx
    from dcsexplorer_b200 import _capi
    import ctypes as C
    seen = set()
    for k, it in enumerate(golden.items):
        if it["os"] in seen or it["stop"] or it["nframes_out"] * 240 != it["pcm"].size:
            continue
        seen.add(it["os"])
        src = tmp_path / ("u%d.bin" % k)
        src.write_bytes(it["stream"])
        out = tmp_path / ("u%d.pcm" % k)
        r = subprocess.run([demo, "--standalone", "%x" % it["os"], str(src), str(it["vol"]), str(it["lvl"]), str(it["nframes_out"]), str(out), "unsized"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        pcm = np.fromfile(out, dtype=np.int16)
        n = min(pcm.size, it["pcm"].size)
        assert np.array_equal(pcm[:n], it["pcm"][:n]), it["label"]
        # the extent the host walk reports = the bytes the batch decoder reports for the same stream
        buf = np.frombuffer(it["stream"] + bytes(32), dtype=np.uint8)
        ext = _capi.lib().dcsb_stream_extent(buf.ctypes.data, it["os"])
        nf = (it["stream"][0] << 8) | it["stream"][1]
        assert ("standalone: %d frames, %d bytes" % (nf, ext)) in r.stdout, (r.stdout, ext)
        assert 0 < ext <= len(it["stream"])
    assert len(seen) >= 3


@pytest.mark.gpu
def test_cpp_front_hard_boot_sequence(demo, tmp_path):
    """HardBoot(): 7812 samples of silence, the POST code to the host, then the startup bong (POST code 1 = one
    bong of 23437 samples: a square wave flipping every 81 samples under a decay of 0x7f80/0x8000 every 32
    samples, DCSDecoder.cpp:1584-1619, :1697-1728), then the decoder runs.  Fast-boot mode skips the bong; a
    data-port byte during the 250 ms wait soft-boots at once and is not queued."""
    sc = romscen.make_scenario(os_version=rb.OS94, seed=101)
    z, tl = _write_inputs(tmp_path, sc)
    out = tmp_path / "boot.pcm"
    n = 7812 + 23437 + 480
    r = subprocess.run([demo, "--hardboot", str(z), "0", "-1", str(n), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    pcm = np.fromfile(out, dtype=np.int16)
    assert not pcm[:7812].any()
    # the bong, restated: sample k (from 0) of the bong has |level| after k // 32 decays, sign flipping every 81 samples from -1
    level, want = 0x0FFF, []
    env = sgn_n = 0
    sign = -1
    for k in range(23437):
        if env >= 31:
            level = ((level * 0x7F80) << 1 >> 16) & 0xFFFF
            env = 0
        else:
            env += 1
        if sgn_n >= 80:
            sign, sgn_n = -sign, 0
        else:
            sgn_n += 1
        want.append(sign * level)
    assert np.array_equal(pcm[7812:7812 + 23437], np.array(want, dtype=np.int16))
    assert not pcm[7812 + 23437:].any()                         # running, nothing playing
    assert "host bytes 79 01 | running 1" in r.stdout
    assert "snd_u2.rom=2(" in r.stdout and "snd_u3.rom=3(" in r.stdout     # the reference's zip file list: names and chip numbers
    # fast boot: no bong
    r = subprocess.run([demo, "--hardboot", str(z), "1", "-1", "9000", str(out)], capture_output=True, text=True)
    assert r.returncode == 0 and not np.fromfile(out, dtype=np.int16).any() and "host bytes 79 01 | running 1" in r.stdout
    # a data-port byte during the boot wait: soft boot at once, no POST code, no bong
    r = subprocess.run([demo, "--hardboot", str(z), "0", "100", "9000", str(out)], capture_output=True, text=True)
    assert r.returncode == 0 and not np.fromfile(out, dtype=np.int16).any() and "host bytes | running 1" in r.stdout
