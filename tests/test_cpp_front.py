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
    """chunkFrames = 16: same audio as chunk 1 when the data-port bytes arrive on chunk boundaries."""
    sc = romscen.make_scenario(os_version=rb.OS94, seed=55, n_frames=320)
    sc["writes"] = [((f + 15) // 16 * 16, b) for f, b in sc["writes"]]
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
