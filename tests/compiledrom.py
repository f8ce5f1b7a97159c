"""TEST INFRASTRUCTURE: access to tests/golden/compiled_rom.npz -- BASELINE config 4's ROM images
built by the reference's own DCSCompiler (tests/golden/make_compiled_rom_golden.py) with the
reference decoder's output on them."""
import os
import numpy as np
import rombuild as rb

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["c94", "c95"]
TRACK_FRAMES = 160


def load(name):
    g = np.load(os.path.join(HERE, "golden", "compiled_rom.npz"))
    images = {int(c): g["%s/u%d" % (name, c)].tobytes() for c in g[name + "/chips"]}
    n_frames, vol, ntracks, osv = [int(v) for v in g[name + "/params"]]
    th, host = g[name + "/track_host"].tobytes(), []
    p = 0
    for _ in range(ntracks):
        host.append(th[p + 1:p + 1 + th[p]])
        p += 1 + th[p]
    return dict(g=g, name=name, images=images, writes=[(int(f), int(b)) for f, b in g[name + "/writes"]], n_frames=n_frames,
                master_volume=vol, n_tracks=ntracks, os=osv, track_host=host,
                track_timelines=[([(1, b) for b in rb.command_bytes(t)], TRACK_FRAMES, 255) for t in range(ntracks)])


def frame_sums(pcm):
    return pcm.reshape(-1, 240).astype(np.int64).sum(axis=1).astype(np.uint32)


def check_main(c, pcm, host_bytes):
    from oracle import orc
    g, name = c["g"], c["name"]
    sums = frame_sums(pcm)
    bad = np.nonzero(sums != g[name + "/sums"])[0]
    assert bad.size == 0, "%s: first differing frame %d" % (name, bad[0])
    assert np.array_equal(pcm[:240 * 20], g[name + "/head"]) and np.array_equal(pcm[-240 * 20:], g[name + "/tail"])
    assert orc.fnv1a(pcm) == int(g[name + "/fnv"])
    if host_bytes is not None:
        assert host_bytes == g[name + "/host"].tobytes()


def check_tracks(c, pcms):
    want = c["g"][c["name"] + "/track_sums"]
    for t, pcm in enumerate(pcms):
        bad = np.nonzero(frame_sums(pcm) != want[t])[0]
        assert bad.size == 0, "%s track $%04X: first differing frame %d" % (c["name"], t, bad[0])
