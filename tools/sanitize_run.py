#!/usr/bin/env python3
"""A small run of every kernel path, meant to be executed under compute-sanitizer (memcheck / racecheck /
synccheck) on the GPU box:  compute-sanitizer --tool memcheck python tools/sanitize_run.py
Paths: resident batch with the scan beside the persistent decode kernel (ready queue, gate kernel), the two
kernels one after the other, dcsb_decode_streams with pinned buffers (time-sliced, resumed scans), the same with
a pageable buffer, the scan variant without rings, 1993 layouts, damaged streams, ROM timeline rendering, the encoder."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dcsfuzz
import romscen
import rombuild as rb
import dcsexplorer_b200 as dx
import test_gpu_parity as t

ctx = dx.Context(0)
# racecheck serialises kernels: the scan cannot run beside the persistent decode kernel, whose warps then give up
# waiting (and the call reports DCSB_E_CUDA, as designed) -- SANITIZE_SERIAL=1 runs the kernels one after the other
SERIAL = os.environ.get("SANITIZE_SERIAL") == "1"
rng = np.random.default_rng(1)
streams = [(dcsfuzz.fuzz94(rng, 140, type1=i & 1, max_code=15 if i % 3 == 0 else 9), 0x9400, 255, 0x64, 2) for i in range(96)]
streams += [s for seed in range(3) for s in t._soak_make(seed) if len(s[0]) >= 3 and ((s[0][0] << 8) | s[0][1]) > 0]
want = None
for overlap in ((0,) if SERIAL else (1, 0)):
    ctx.set_overlap(overlap)
    b = ctx.batch(streams)
    b.decode()
    res = b.results()
    cs = [r["checksum"] for r in res]
    assert want is None or cs == want
    want = cs
    b.close()
ctx.set_overlap(0 if SERIAL else 1)
for pinned in (True, False):
    ctx.set_pipeline(3, 40)
    pcm, offs, res = (ctx.decode_streams_pinned if pinned else ctx.decode_streams)(streams)
    assert [r["checksum"] for r in res] == want
ctx.set_pipeline(0, 0)
os.environ["DCSB_SCAN_DIRECT"] = "1"
os.environ["DCSB_SCAN_WARPS"] = "4"
pcm, offs, res = ctx.decode_streams(streams)
assert [r["checksum"] for r in res] == want
del os.environ["DCSB_SCAN_DIRECT"], os.environ["DCSB_SCAN_WARPS"]
sc = romscen.make_scenario(os_version=rb.OS95, seed=77, n_frames=120, version=0x0105)
rom = dx.Rom(sc["images"])
tls = [([(f + i % 3, b) for f, b in sc["writes"]], 100 + i, 255 - i) for i in range(6)]
ctx.render_timelines(rom, tls)
p = dx.Player(ctx, rom)
ctx._L.dcsb_player_set_lookahead(p._h, 8)
for f in range(40):
    if f == 13:
        p.write_data_port(0)
        p.write_data_port(1)
    p.render(1)
p.close()
rom.close()
# the forward path: a few clips of every stream type (incl. one shorter than a frame), decoded again
erng = np.random.default_rng(3)
clips = [(erng.standard_normal(n) * 0.2).astype(np.float32) for n in (17, 240, 1000, 5000, 12345)]
enc = ctx.encode_streams(clips * 4, [(t, u, 96000, 0.97) for t, u in ((0, 0), (0, 3), (1, 0), (1, 3)) for _ in clips])
pcm, offs, res = ctx.decode_streams([(e, 0x9400, 255, 0x64, 2) for e in enc])
assert all(r["status"] == 0 for r in res)
enc = ctx.encode_streams(clips * 3 + clips[:2], [(t, 0, 96000, 0.97, 10 / 32768, 10 / 32768, v) for t, v in ((0, 0x9301), (0, 0x9302), (1, 0x9302)) for _ in clips]
                         + [(-1, -1, 64000, 0.9), (-1, 0, 64000, 0.9, 10 / 32768, 10 / 32768, 0x9302)])
pcm, offs, res = ctx.decode_streams([(e, v, 255, 0x64, 2) for e, v in zip(enc, [0x9301] * 5 + [0x9302] * 10 + [0x9400, 0x9302])])
assert all(r["status"] == 0 for r in res)
ctx.close()
print("sanitize_run: ok, %d streams" % len(streams))
