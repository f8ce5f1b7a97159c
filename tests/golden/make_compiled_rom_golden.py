#!/usr/bin/env python3
"""Generate tests/golden/compiled_rom.npz: BASELINE config 4 -- a ROM image built by the REFERENCE's
own ROM compiler (DCSCompiler, unmodified, behind oracle/_ref/ref_dcscompile) from a forged
prototype ROM and a generated script with 72 tracks, then played by the UNMODIFIED reference
decoder (oracle/_ref).  Runs only in the build container; the fixture (ROM images + expected
output) is committed so that neither the CPU suite nor the GPU box needs the reference tree.

Sources: synthetic 16-bit WAV clips (sine mixtures + noise, 31 250 Hz) encoded by the reference's
DCSEncoder with a mix of stream types / bit rates, plus two fuzzer-made raw .dcs streams taken over
without transcoding (DCSEncoder::EncodeDCSFile).  The script uses every program statement of the
compiler's language: Play (own / other channel, repeat), Wait (frames / stream / forever),
SetMixingLevel (level / increase / decrease, with and without steps, own and other channels),
Loop (nested), Queue, Stop (own / other / all), WriteDataPort, SetVariable, StartDeferred,
Defer and Defer Indirect tracks.

Fixture contents: the ROM images; a main timeline (track commands overlapping on >= 4 channels,
master-volume and channel-volume sequences) with FNV / per-frame sums / head / tail / host bytes of
the reference's PCM; one timeline PER TRACK ("all tracks rendered") with per-frame sums; the
reference's GetTrackInfo of every track and its ListStreams."""
import os
import subprocess
import sys
import tempfile
import wave
import zipfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle import ref, orc     # noqa: E402
import rombuild as rb           # noqa: E402
from rombuild import Track      # noqa: E402
import dcsfuzz                  # noqa: E402
import dcsexplorer_b200 as dx   # noqa: E402  (only its host-side .dcs container writer)

COMPILER = os.path.join(ROOT, "oracle", "_ref", "ref_dcscompile")
N_CLIPS = 22
TRACK_FRAMES = 160


def write_wav(path, x, rate=31250):
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(rate)
        w.writeframes((np.clip(x, -1, 1) * 32767).astype("<i2").tobytes())


def make_clip(rng, seconds):
    n = int(seconds * 31250)
    t = np.arange(n) / 31250.0
    x = np.zeros(n)
    for _ in range(int(rng.integers(1, 5))):
        f = float(np.exp(rng.uniform(np.log(50), np.log(12000))))
        x += rng.uniform(0.05, 0.4) * np.sin(2 * np.pi * f * t + rng.uniform(0, 6.28))
    x += 10 ** (rng.uniform(-40, -14) / 20) * rng.standard_normal(n)
    if rng.random() < 0.3:                      # a silent gap
        a = int(rng.integers(0, n // 2))
        x[a:a + n // 10] = 0
    env = np.minimum(1.0, np.minimum(t, t[-1] - t) * 50)
    return x * env


def make_script(rng, clips, raw):
    L = []
    A = L.append
    A('Signature "dcsb200 config-4 ROM built by DCSCompiler";')
    A("Default encoding parameters (Type=*, Subtype=*, BitRate=128000, PowerCut=97);")
    types = ["Type=0, Subtype=0", "Type=1, Subtype=0", "Type=1, Subtype=3", "Type=*, Subtype=*"]
    rates = [48000, 64000, 96000, 128000, 192000, 256000]
    for i, c in enumerate(clips):
        A('Stream C%d "%s" (%s, BitRate=%d, PowerCut=%d);' % (i, c, types[i % 4], rates[i % 6], (90, 97, 100)[i % 3]))
    for i, r in enumerate(raw):
        A('Stream R%d "%s";' % (i, r))
    A("Var VA;")
    A("Var VB;")
    A("Deferred Indirect table TabA ($0030, $0031, $0032);")
    A("Deferred Indirect table TabB ($0033, $0034);")
    S = lambda k: "C%d" % (k % len(clips))
    ntracks = 0

    def track(n, ch, body):
        nonlocal ntracks
        A("Track $%04X channel %d {\n%s\n};" % (n, ch, "\n".join("   " + b for b in body)))
        ntracks = max(ntracks, n + 1)

    # $00-$17: one clip each, channel k % 6, levels spread, some repeated, some waiting on the stream
    for k in range(24):
        ch = k % 6
        body = ["SetMixingLevel($%02X);" % (0x30 + 3 * k)]
        if k % 4 == 1:
            body += ["Play(Stream %s, Repeat %d);" % (S(k), 2 + k % 3), "Wait(stream);"]
        elif k % 4 == 2:
            body += ["Play(%s);" % S(k), "Wait(stream - 10);", "SetMixingLevel(decrease $20, steps 8);", "Wait(12);"]
        elif k % 4 == 3:
            body += ["Play(%s);" % ("R%d" % (k % len(raw)) if raw else S(k)), "Wait(forever);"]
        else:
            body += ["Play(%s);" % S(k), "Wait(stream);"]
        track(k, ch, body)
    # $18-$23: level moves on other channels while a clip plays (ducking), with and without steps
    for k in range(24, 36):
        ch = k % 6
        other = (ch + 1 + k % 4) % 6
        kind = ("Level $%02X" % (0x20 + k), "Increase $%02X" % (4 + k % 9), "Decrease $%02X" % (6 + k % 11))[k % 3]
        steps = ("", ", Steps %d" % (3 + k % 17), ", Steps 0.2 sec")[k % 3 if k % 2 else 0]
        track(k, ch, ["SetMixingLevel($64);", "SetMixingLevel(Channel %d, %s%s);" % (other, kind, steps), "Play(%s);" % S(k + 3),
                      "Wait(stream - %d);" % (5 + k % 9), "SetMixingLevel(Channel %d, Level $50, Steps %d);" % (other, 2 + k % 6),
                      "Wait(%d);" % (3 + k % 5)])
    # $24-$2B: loops (nested), host bytes
    for k in range(36, 44):
        ch = k % 6
        track(k, ch, ["SetMixingLevel($%02X);" % (0x50 + k), "Loop (%d) {" % (2 + k % 3), "   Play(%s);" % S(k + 5),
                      "   Wait(%d) WriteDataPort($%02X);" % (4 + k % 7, 0x40 + k), "   Loop (2) {",
                      "      Wait(%d) SetMixingLevel(Increase $%02X);" % (2 + k % 3, 3 + k % 5), "   }", "   Wait(stream);", "}",
                      "WriteDataPort($%02X);" % (0x80 + k)])
    # $2C-$2F: queue / stop
    track(0x2C, 0, ["SetMixingLevel($60);", "Play(%s);" % S(7), "Wait(10);", "Queue($0002);", "Wait(12);", "Queue(Track $0009);", "Wait(stream);"])
    track(0x2D, 1, ["Stop(Channel 0);", "Wait(4) Stop(2);", "Stop(1);"])
    track(0x2E, 5, ["Wait(1) Stop(*);", "Stop(5);"])
    track(0x2F, 3, ["SetMixingLevel($70);", "Play(Channel 4, Stream %s, Repeat 2);" % S(11), "Play(Channel 3, Stream %s);" % S(12),
                    "Wait(stream);", "Stop(Channel 4);"])
    # $30-$34: targets of the deferred-indirect tables; $35-$3B deferred machinery
    for k in range(0x30, 0x35):
        track(k, k % 6, ["SetMixingLevel($%02X);" % (0x58 + k % 8), "Play(%s);" % S(k), "Wait(stream);"])
    A("Track $0035 channel 2 Defer($0004);")
    A("Track $0036 channel 3 Defer Indirect(TabA[VA]);")
    A("Track $0037 channel 4 Defer Indirect(TabB[VB]);")
    ntracks = max(ntracks, 0x38)
    track(0x38, 0, ["SetMixingLevel($40);", "Play(%s);" % S(3), "Wait(15);", "StartDeferred(Channel 2);", "Wait(stream);"])
    track(0x39, 1, ["SetVariable(Var VA, Value 2);", "SetVariable(Var VB, Value 1);", "Wait(5);", "StartDeferred(Channel 3);",
                    "Wait(9) StartDeferred(4);"])
    track(0x3A, 1, ["SetVariable(Var VA, Value 0);", "SetVariable(Var VB, Value 0);"])
    track(0x3B, 5, ["SetMixingLevel($7F);", "Loop {", "   Play(%s);" % S(9), "   Wait(stream);", "   Wait(6);", "}"])
    # $3C-$47: cross-channel plays, several channels from one program
    for k in range(0x3C, 0x48):
        ch = k % 6
        body = ["SetMixingLevel($5A);"]
        for j in range(3):
            oc = (ch + j) % 6
            body.append("SetMixingLevel(Channel %d, Level $%02X);" % (oc, 0x40 + 5 * j + k % 16))
            body.append("Wait(%d) Play(Channel %d, Stream %s%s);" % (j * (2 + k % 4), oc, S(k + j), ", Repeat 2" if (k + j) % 5 == 0 else ""))
        body += ["Wait(stream);", "Wait(10);"]
        track(k, ch, body)
    return "\n".join(L) + "\n", ntracks


def frame_sums(pcm):
    return pcm.reshape(-1, 240).astype(np.int64).sum(axis=1).astype(np.uint32)


def build(os_version, seed, tmp, seconds):
    rng = np.random.default_rng(seed)
    proto, _ = rb.build_rom(os_version, [Track(0).mix(0, 0, 100).play("s0").wait_forever()],
                            {"s0": dcsfuzz.fuzz94(rng, 8, type1=False)}, n_chips=1)
    pz = os.path.join(tmp, "proto.zip")
    with zipfile.ZipFile(pz, "w", zipfile.ZIP_DEFLATED) as z:
        for c, img in proto.items():
            z.writestr("snd_u%d.rom" % c, img)
    clips = []
    for i in range(N_CLIPS):
        p = os.path.join(tmp, "clip%02d.wav" % i)
        write_wav(p, make_clip(rng, float(rng.uniform(*seconds))))
        clips.append(p)
    raw = []
    for i in range(2):
        p = os.path.join(tmp, "raw%d.dcs" % i)
        dx.write_dcs_file(p, 0x9400, dcsfuzz.fuzz94(rng, 30 + 25 * i, type1=bool(i), max_code=15 if i == 0 else 9))
        raw.append(p)
    script, ntracks = make_script(rng, clips, raw)
    sp = os.path.join(tmp, "rom.txt")
    open(sp, "w").write(script)
    oz = os.path.join(tmp, "out.zip")
    r = subprocess.run([COMPILER, pz, sp, oz, str(512 * 1024)], capture_output=True, text=True, cwd=tmp)
    if r.returncode != 0:
        raise RuntimeError("ref_dcscompile failed:\n" + r.stdout + r.stderr)
    print(r.stdout.strip())
    images = {}
    with zipfile.ZipFile(oz) as z:
        for n in z.namelist():
            images[int(n.rsplit(".", 1)[0][-1])] = z.read(n)      # snd_u2.rom (DCS) / snd_s2.rom (DCS-95)
    return images, ntracks, script


def make_timeline(rng, ntracks):
    writes = []
    f = 2
    order = [0x3B, 0x01, 0x02, 0x1A, 0x2F, 0x24, 0x38, 0x35, 0x39, 0x36, 0x37, 0x3C, 0x05, 0x2C, 0x1D, 0x27, 0x3A, 0x36, 0x39,
             0x40, 0x0B, 0x2D, 0x13, 0x45, 0x20, 0x2E, 0x07, 0x3E, 0x29, 0x16, 0x47, 0x2D, 0x03, 0x22]
    for k, t in enumerate(order):
        for b in rb.command_bytes(t):
            writes.append((f, b))
        if k % 5 == 2:
            for b in rb.volume_bytes(int(rng.integers(70, 256))):
                writes.append((f + 1, b))
        if k % 7 == 3:
            for b in rb.channel_volume_bytes(int(rng.integers(0, 6)), int(rng.integers(100, 256))):
                writes.append((f + 2, b))
        f += int(rng.integers(6, 40))
    writes.sort(key=lambda w: w[0])
    return writes, f + 200


def main():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, osv, seed, seconds in (("c94", rb.OS94, 9401, (0.8, 3.2)), ("c95", rb.OS95, 9501, (0.3, 1.0))):
            d = os.path.join(tmp, name)
            os.makedirs(d)
            images, ntracks, script = build(osv, seed, d, seconds)
            rng = np.random.default_rng(seed + 1)
            writes, n_frames = make_timeline(rng, ntracks)
            vol = int(rng.integers(150, 256))
            rp = ref.RomPlayer(images, vol)
            info = rp.info()
            assert info["check"] == 1 and info["max_track"] == ntracks - 1, info
            pcm = rp.render_timeline(writes, n_frames)
            hb = rp.host_bytes()
            out[name + "/chips"] = np.array(sorted(images), dtype=np.int32)
            for c in images:
                out[name + "/u%d" % c] = np.frombuffer(images[c], dtype=np.uint8)
            out[name + "/script"] = np.array(script)
            out[name + "/writes"] = np.array(writes, dtype=np.int32)
            out[name + "/params"] = np.array([n_frames, vol, ntracks, osv], dtype=np.int32)
            out[name + "/fnv"] = np.array(orc.fnv1a(pcm), dtype=np.uint64)
            out[name + "/sums"] = frame_sums(pcm)
            out[name + "/head"] = pcm[:240 * 20]
            out[name + "/tail"] = pcm[-240 * 20:]
            out[name + "/host"] = np.frombuffer(hb, dtype=np.uint8)
            out[name + "/streams"] = np.array(rp.list_streams(), dtype=np.uint32)
            out[name + "/tracks"] = np.array([rp.track_info(t) for t in range(ntracks)], dtype=np.uint32)
            dec = [rp.decompile(t) for t in range(ntracks)]        # DecompileTrackProgram: dcsb_opcode-shaped records
            out[name + "/decompile_counts"] = np.array([n for _, n in dec], dtype=np.int32)
            out[name + "/decompile"] = np.frombuffer(b"".join(b for b, _ in dec), dtype=np.uint8)
            chans = (pcm.reshape(-1, 240) != 0).any(axis=1).sum()
            rp.close()
            # every track on its own: fresh decoder, command at frame 1
            tsums = np.zeros((ntracks, TRACK_FRAMES), dtype=np.uint32)
            thost = []
            for t in range(ntracks):
                rp = ref.RomPlayer(images, 255)
                p = rp.render_timeline([(1, b) for b in rb.command_bytes(t)], TRACK_FRAMES)
                tsums[t] = frame_sums(p)
                thost.append(rp.host_bytes())
                rp.close()
            out[name + "/track_sums"] = tsums
            out[name + "/track_host"] = np.frombuffer(b"".join(bytes([len(h)]) + h for h in thost), dtype=np.uint8)
            print("%s: %d tracks, %d streams, timeline %d frames (%d audible), host bytes %d, silent solo tracks %d" % (
                name, ntracks, len(out[name + "/streams"]), n_frames, chans, len(hb), int((tsums.sum(axis=1) == 0).sum())))
    np.savez_compressed(os.path.join(HERE, "compiled_rom.npz"), **out)
    print("fixture bytes:", os.path.getsize(os.path.join(HERE, "compiled_rom.npz")))


if __name__ == "__main__":
    main()
