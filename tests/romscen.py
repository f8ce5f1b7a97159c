"""TEST INFRASTRUCTURE: seeded ROM-playback scenarios (BASELINE config 4): a synthetic ROM set
with tracks that exercise every track opcode (play on own / other channels, stream repeat, stop,
queue, deferred + indirect deferred links, variables, mixing levels set / increase / decrease
with and without fades, loops, host bytes and the OS93a host timer), plus a timeline of data
port traffic (track commands, master volume and channel volume sequences, junk bytes).
Streams come from the bit-level fuzzer so that the scenario can be rebuilt anywhere from its
seed (the GPU box has no reference tree); the expected PCM is frozen in tests/golden/."""
import numpy as np
import dcsfuzz
import rombuild as rb
from rombuild import Track


def make_streams(os_version, rng, n, frames=(6, 90)):
    out = {}
    for i in range(n):
        nf = int(rng.integers(frames[0], frames[1]))
        if os_version in (rb.OS94, rb.OS95):
            k = i % 3
            d = dcsfuzz.fuzz94(rng, nf, type1=(k != 0), max_code=15 if k == 0 else 9)
        elif os_version == rb.OS93B:
            d = dcsfuzz.fuzz93(rng, nf, type1=i & 1, small=bool(i & 2))
        else:
            d = dcsfuzz.fuzz93(rng, nf, type1=0, small=bool(i & 2)) if i % 3 else dcsfuzz.fuzz93a1(rng, nf)
        out["s%d" % i] = d
    return out


def make_scenario(os_version, seed, n_streams=14, n_frames=700, version=None, with_errors=False):
    """Returns dict(images, writes [(frame, byte)], n_frames, master_volume, os)."""
    rng = np.random.default_rng(seed)
    streams = make_streams(os_version, rng, n_streams)
    if with_errors and os_version in (rb.OS94, rb.OS95):
        # a stream whose decoder error path fires mid-way (channel stop), and one with a zero frame
        # count: the reference's 16-bit frame counter wraps, so it plays on (65,536 frames).  It is
        # long enough that the timeline replaces it before the decoder runs off its end into
        # whatever follows in the ROM (band types leave 0..15 there: undefined in the reference).
        streams["bad"] = dcsfuzz.fuzz94(rng, 40, type1=1, max_code=6, error_frame=17, escape_p=0.2)
        good = dcsfuzz.fuzz94(rng, 420, type1=0)
        streams["zero"] = bytes([0, 0]) + good[2:]
    keys = list(streams.keys())
    nch = 6
    S = lambda i: keys[i % len(keys)]
    is93a = os_version == rb.OS93A
    is93 = os_version in (rb.OS93A, rb.OS93B)
    hb = (lambda t, b, wait=0: t.host_timer93a(b, 0, wait)) if is93a else (lambda t, b, wait=0: t.host_byte(b, wait))
    tracks = []
    # 0: plain play on channel 0, level 100
    tracks.append(Track(0).mix(0, 0, 100).play(S(0)).wait_forever())
    # 1: music bed on channel 1: endless stream repeat, level with fade in
    tracks.append(Track(1).mix(0, 1, 20).mix(1, 1, 80, steps=40).play(S(1), repeat=0).wait_forever())
    # 2: effect on channel 2 that ducks channel 1 while it plays, then restores it
    tracks.append(hb(Track(2).mix(0, 2, 110).mix(2, 1, 40, steps=10).play(S(2)), 0x42)
                  .mix(1, 1, 40, steps=25, wait=30).stop(wait=60))
    # 3: loop of two short plays on channel 3, 3 times, then stop
    tracks.append(Track(3).mix(0, 3, 90).loop(3).play(S(3)).play(S(4), wait=25).end_loop(wait=20).stop(wait=5))
    # 4: track on channel 4 that starts streams on channels 4 and 5 and queues track 0
    tracks.append(Track(4).mix(0, 4, 100).mix(0, 5, 70).play(S(5)).play(S(6), channel=5, repeat=2).queue(0, wait=12).wait_forever())
    # 5: stops channel 1 (music) and itself
    tracks.append(Track(5).stop_channel(1).stop_channel(5))
    # 6: deferred link (type 2) on channel 2 -> track 2; 7: triggers it from channel 0 after a wait
    tracks.append(Track(2, ttype=2, link=2))
    tracks.append(Track(0).mix(0, 0, 60).play(S(7)).start_deferred(2, wait=15).wait_forever())
    # 8: indirect deferred (type 3): variable 3 selects from table 0; 9 sets the variable and triggers
    tracks.append(Track(3, ttype=3, link=(3 << 8) | 0))
    t9 = Track(1).mix(0, 1, 100)
    if is93:
        t9.nop93_06()
    else:
        t9.set_var(3, 1)
    tracks.append(t9.play(S(8)).start_deferred(3, wait=8).wait_forever())
    # 10: nested loops + level steps on another channel + nops
    tracks.append(Track(0).mix(0, 0, 127).loop(2).loop(2).play(S(9)).mix(2, 0, 15, wait=10).nop(wait=3).end_loop(wait=6)
                  .mix(1, 0, 30, steps=5).end_loop(wait=4).stop(wait=30))
    # 11: endless program loop replaying a short stream (stopped from outside)
    tracks.append(Track(5).mix(0, 5, 85).loop(0).play(S(10)).end_loop(wait=18))
    # 12: unpopulated slot; 13: negative levels and an over-range fade
    tracks.append(None)
    tracks.append(Track(2).mix(0, 2, -20).mix(1, 2, 127, steps=3).mix(1, 2, 127).play(S(11), repeat=3).mix(2, 2, 100, steps=200, wait=5).stop(wait=120))
    if is93a:
        # 14: OS93a host event timer (byte every 9 frames) while a stream plays
        tracks.append(Track(4).mix(0, 4, 100).host_timer93a(0x33, 9).play(S(12)).host_timer93a(0, 0, wait=50).stop(wait=5))
    else:
        # 14: 1994+ mystery opcodes are skipped correctly
        t = Track(4).mix(0, 4, 100)
        if not is93:
            t.op10(1, 5).op11(2, 3, 7).op11(2, 3, 7, dec=True)
        tracks.append(t.play(S(12)).host_byte(0x69).host_byte(0x6A, wait=20).stop(wait=40))
    if with_errors and "bad" in streams:
        tracks.append(Track(3).mix(0, 3, 100).play("bad").wait_forever())         # 15
        tracks.append(Track(5).mix(0, 5, 100).play("zero").wait_forever())        # 16
    images, addr = rb.build_rom(os_version, tracks, streams, n_chips=3, version=version,
                                indirect_tables=[[0, 3, 10]])
    # timeline
    ntr = len(tracks)
    writes = []
    f = 2
    order = [1, 0, 2, 4, 3, 7, 6, 9, 8, 10, 13, 11, 14, 2, 5, 0, 1, 12, 3, 11, 2, 10, 5]
    if with_errors and "bad" in streams:
        order = order[:6] + [15, 16] + order[6:]
    for k, t in enumerate(order):
        for b in rb.command_bytes(t):
            writes.append((f, b))
        if k % 5 == 2:
            for b in rb.volume_bytes(int(rng.integers(60, 256))):
                writes.append((f + 1, b))
        if k % 7 == 3:
            for b in rb.channel_volume_bytes(int(rng.integers(0, nch)), int(rng.integers(90, 256))):
                writes.append((f + 2, b))
        if k % 6 == 4:
            writes.append((f + 3, 0x55))            # half a sequence that times out
            writes.append((f + 3 + 15, 0x7F))       # first byte of a new (invalid: no second byte soon) command ...
            writes.append((f + 3 + 15, 0xF0))       # ... completed: command $7FF0 = out of range, ignored
            f += 20                                 # (let the data port settle before the next command)
        if k % 9 == 5:
            for b in (0x55, 0xC2, 0x55, 0xC3):      # version queries
                writes.append((f + 1, b))
        f += int(rng.integers(8, 45))
    writes.sort(key=lambda w: w[0])
    n_frames = max(n_frames, f + 150)
    return dict(images=images, writes=writes, n_frames=int(n_frames), master_volume=int(rng.integers(120, 256)),
                os=os_version, n_tracks=ntr, stream_addr=addr)


SCENARIOS = [
    ("os94", dict(os_version=rb.OS94, seed=101)),
    ("os95", dict(os_version=rb.OS95, seed=102)),
    ("os95-v105", dict(os_version=rb.OS95, seed=103, version=0x0105)),
    ("os93b", dict(os_version=rb.OS93B, seed=104)),
    ("os93a", dict(os_version=rb.OS93A, seed=105)),
    ("os94-errors", dict(os_version=rb.OS94, seed=106, with_errors=True)),
]
