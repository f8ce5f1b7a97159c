"""The reference-side plugin (dcsexplorer_b200/plugin/DCSDecoderB200Plugin.cpp): a subclass of the
REFERENCE's abstract DCSDecoder, registered in the reference's own decoder registry as "b200"
(DCSDecoder.h:1115-1128) and linked with the reference's unmodified DCSDecoder.o.  The client
(tests/cpp/ref_plugin_host.cpp) picks its decoder by registry name like DCSExplorer does
(DCSExplorer.cpp:459-537) and talks to it through the base class only (AddROM, CheckROMs,
SoftBoot, SetMasterVolume, WriteDataPort, GetNextSample).
CPU: both implementations are registered, "native" plays, "b200" refuses without a GPU.
GPU: the same binary run with "native" and with "b200" gives bit-identical PCM and host bytes."""
import os
import subprocess
import numpy as np
import pytest
import rombuild as rb
import romscen
from test_rom import check_rom_golden

HERE = os.path.dirname(os.path.abspath(__file__))
HOST = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "ref_plugin_host")
needs_host = pytest.mark.skipif(not os.path.exists(HOST), reason="oracle/_ref/ref_plugin_host not built (reference tree absent)")


def _run(decoder, sc, tmp_path, n_frames=None):
    tl = tmp_path / "timeline.txt"
    tl.write_text("".join("%d %d\n" % w for w in sc["writes"]))
    chips = []
    for chip, img in sc["images"].items():
        p = tmp_path / ("u%d.bin" % chip)
        p.write_bytes(img)
        chips.append("%d=%s" % (chip, p))
    out = tmp_path / ("%s.pcm" % decoder)
    r = subprocess.run([HOST, decoder, str(tl), str(n_frames or sc["n_frames"]), str(sc["master_volume"]), str(out)] + chips,
                       capture_output=True, text=True)
    if r.returncode != 0:
        return r, None, None
    line = r.stdout.splitlines()[0]
    hb = bytes(int(x, 16) for x in line.split("host bytes")[1].split())
    return r, np.fromfile(out, dtype=np.int16), hb


@needs_host
def test_b200_is_registered_beside_native(built):
    r = subprocess.run([HOST, "--list"], capture_output=True, text=True)
    names = [l.split("\t")[0] for l in r.stdout.splitlines()]
    assert r.returncode == 0 and "native" in names and "b200" in names


@needs_host
def test_plugin_host_native_matches_golden_and_b200_has_no_cpu_fallback(built, tmp_path):
    import torch
    g = np.load(os.path.join(HERE, "golden", "rom_golden.npz"))
    sc = romscen.make_scenario(**dict(romscen.SCENARIOS)["os94"])
    r, pcm, hb = _run("native", sc, tmp_path)
    assert r.returncode == 0, r.stderr
    check_rom_golden(g, "os94", pcm, hb)
    if not torch.cuda.is_available():
        r, _, _ = _run("b200", sc, tmp_path)
        assert r.returncode == 3 and "b200" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@needs_host
@pytest.mark.parametrize("name", ["os94", "os95-v105", "os93b", "os93a", "os94-errors"])
def test_b200_plugin_equals_native_behind_the_reference_base_class(built, tmp_path, name):
    sc = romscen.make_scenario(**dict(romscen.SCENARIOS)[name])
    rn, pcm_n, hb_n = _run("native", sc, tmp_path)
    rb_, pcm_b, hb_b = _run("b200", sc, tmp_path)
    assert rn.returncode == 0, rn.stderr
    assert rb_.returncode == 0, rb_.stderr
    assert pcm_n.any() and np.array_equal(pcm_n, pcm_b)
    assert hb_n == hb_b
    # the base class's own lookups (version, tracks, channels, streams) see the same ROM set
    assert rn.stdout.split("|", 1)[1] == rb_.stdout.split("|", 1)[1]
    assert rb_.stdout.startswith("b200 |") and rn.stdout.startswith("Universal native decoder |")      # Name() of each
