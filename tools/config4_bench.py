#!/usr/bin/env python3
"""GPU measurement aid: BASELINE config 4 -- track playback on the ROM sets built by the
reference's DCSCompiler (tests/golden/compiled_rom.npz): N decoder instances (timelines with
shifted command times and different master volumes, plus every track on its own) rendered by ONE
dcsb_render_timelines call (K5 track-program interpreter kernel -> mix schedule -> K4 mix kernel -> PCM to
host; the ROM's streams are scanned once per ROM; DCSB_SEQ_HOST=1 runs the sequencers on host threads instead),
against the unmodified reference decoder on the host cores.
  config4_bench.py [n_timelines=2048]"""
import os
import sys
import time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import compiledrom
import torch
import dcsexplorer_b200 as dx
from oracle import ref

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ctx = dx.Context(0)
for name in compiledrom.NAMES:
    c = compiledrom.load(name)
    rom = dx.Rom(c["images"])
    tls = []
    for i in range(n):
        sh = i % 13
        tls.append(([(f + sh, b) for f, b in c["writes"]], c["n_frames"] + sh, 255 - (i % 100)))
    tls += c["track_timelines"]
    frames = sum(t[1] for t in tls)
    # the C call alone: descriptors and the (page-locked) output buffer are made once, outside the clock
    tl_arr, keep = dx.make_timelines(tls)
    out = torch.empty(frames * 240, dtype=torch.int16).pin_memory()
    out.zero_()
    resarr = (dx.TimelineResult * len(tls))()
    ts = []
    for it in range(4):
        t0 = time.perf_counter()
        rc = ctx._L.dcsb_render_timelines(ctx._h, rom._h, tl_arr, len(tls), out.data_ptr(), None, resarr)
        ts.append(time.perf_counter() - t0)
        assert rc == 0
    h = out.numpy()
    pcm, o = [], 0
    for t in tls:
        pcm.append(h[o:o + t[1] * 240])
        o += t[1] * 240
    res = [dict(status=resarr[i].status) for i in range(len(tls))]
    assert all(r["status"] == 0 for r in res)
    compiledrom.check_tracks(c, pcm[n:])
    # timeline i against the reference (fresh decoder each)
    t0 = time.perf_counter()
    nref = 0
    for i in (0, 1, n // 2, n - 1):
        rp = ref.RomPlayer(c["images"], tls[i][2])
        want = rp.render_timeline(tls[i][0], tls[i][1])
        rp.close()
        nref += tls[i][1]
        assert np.array_equal(pcm[i], want), i
    tref = time.perf_counter() - t0
    best = min(ts[1:])
    print("%s: %d timelines + %d solo tracks, %d output frames (%.1f h of audio), one call: %.1f ms (first %.1f) = %.2f Gsamples/s "
          "incl. sequencer and PCM download; reference decoder, 1 thread: %.2f Msamples/s; spot checks bit-exact" % (
              name, n, len(c["track_timelines"]), frames, frames * 240 / 31250 / 3600, best * 1e3, ts[0] * 1e3,
              frames * 240 / best / 1e9, nref * 240 / tref / 1e6), flush=True)
    rom.close()
ctx.close()
