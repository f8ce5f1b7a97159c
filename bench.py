#!/usr/bin/env python3
"""bench.py -- decoded PCM Msamples/s of the DCS batch decoder (BASELINE.json metric).

A "step" is one pass of the hot path (frame-boundary scan + decode/transform/writeback) over
one batch of synthetic DCS streams.  At N=1 the workload is BASELINE.json configs[1]: 4,096
synthetic 1994+ format streams (mixed types / bit rates), 10 s each, produced by the reference's
own DCSEncoder (oracle/_ref) from generated sources.  With N>1 every rank decodes its own
4,096 streams (streams are independent: sharded by stream, no data-path collective; weak
scaling); the only collective is the gather of the per-rank time / checksum.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`value`  : whole-job Msamples/s with the compressed streams already resident in HBM.
`e2e`    : same metric through the host-buffer C-ABI call dcsb_decode_streams (pinned host
           input, H2D, kernels, D2H of all PCM inside the timed region).
`roofline`: the dominant kernel against the measured HBM copy bandwidth, plus its share of the
           SMs' issue rate; DRAM traffic / instruction counts come from the committed ncu capture
           named in `traffic_source` and are only printed when that capture's launch shapes are
           the ones this run launched.
`cpu_baseline`: the reference DCSDecoderNative (oracle/_ref, kind "reference") -- or the
           oracle port when the reference did not compile -- on a bounded sample, all host cores;
           its per-stream checksums are compared with the GPU's (`parity`).
`configs`: (N = 1) BASELINE.json configs 1, 3 and 4, bounded to a few seconds each.
`config5`: (N > 1) BASELINE.json config 5: 1,048,576 one-second streams sharded over the ranks.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SAMPLE_RATE = 31250
FRAME = 240
# DRAM traffic and warp-instruction counts per launch come from the committed `ncu --set full` capture of
# this workload, summarised by tools/make_traffic.py into profiles/traffic.json together with the launch
# shapes of the captured kernels
TRAFFIC_JSON = os.path.join(ROOT, "profiles", "traffic.json")


# ------------------------------------------------------------------------------------------
# synthetic corpus
def synth_source(seed, seconds):
    rng = np.random.default_rng(seed)
    n = int(seconds * SAMPLE_RATE)
    t = np.arange(n) / SAMPLE_RATE
    x = np.zeros(n)
    for _ in range(int(rng.integers(1, 5))):
        f = float(np.exp(rng.uniform(np.log(50.0), np.log(12000.0))))
        x += rng.uniform(0.05, 0.4) * np.sin(2 * np.pi * f * t + rng.uniform(0, 6.28))
    level = 10 ** (rng.uniform(-40, -10) / 20)
    noise = rng.normal(0, 1, n)
    if rng.random() < 0.5:                       # pink noise (Kellet's 3-pole approximation)
        from scipy.signal import lfilter
        noise = lfilter([0.049922035, -0.095993537, 0.050612699, -0.004408786],
                        [1, -2.494956002, 2.017265875, -0.522189400], noise)
        noise /= max(1e-9, noise.std())
    x += level * noise
    if rng.random() < 0.05:                      # silent gap
        g0 = int(rng.integers(0, max(1, n - SAMPLE_RATE // 2)))
        x[g0:g0 + SAMPLE_RATE // 2] = 0
    return np.clip(x, -1, 1).astype(np.float32)


TYPES = [(0, 0), (1, 0), (1, 3)]
RATES = [32000, 64000, 96000, 128000, 192000, 256000]
CUTS = [0.90, 0.97, 1.0]


def _encode_one(args):
    seed, seconds = args
    from oracle import ref
    ty, sub = TYPES[seed % 3]
    rate = RATES[(seed // 3) % 6]
    cut = CUTS[(seed // 18) % 3]
    data, nf = ref.encode(synth_source(seed, seconds), fmt=0x9400, stype=ty, subtype=sub, bit_rate=rate, power_cut=cut)
    return data


def build_corpus(n_streams, seconds, seed0, budget_s=75.0, log=None):
    """Returns (list of stream bytes, n_unique, source description).  Unique streams come from
    the reference encoder (cached under corpus_cache/); when fewer than n_streams can be made
    inside the time budget the pool is replicated (every replica gets its own bytes in HBM)."""
    from oracle import ref
    cache_dir = os.path.join(ROOT, "corpus_cache")
    os.makedirs(cache_dir, exist_ok=True)
    cache = os.path.join(cache_dir, "c2_%d_%gs_seed%d.npz" % (n_streams, seconds, seed0))
    pool = []
    if os.path.exists(cache):
        z = np.load(cache)
        blob, offs = z["blob"], z["offs"]
        pool = [blob[offs[i]:offs[i + 1]].tobytes() for i in range(len(offs) - 1)]
    src = "reference DCSEncoder (oracle/_ref)"
    if len(pool) < n_streams:
        if ref.available():
            import multiprocessing as mp
            ncpu = len(os.sched_getaffinity(0))
            t0 = time.time()
            with mp.get_context("fork").Pool(ncpu) as pool_mp:
                it = pool_mp.imap(_encode_one, [(seed0 + i, seconds) for i in range(len(pool), n_streams)], chunksize=2)
                for d in it:
                    pool.append(d)
                    if time.time() - t0 > budget_s and len(pool) >= 32:
                        pool_mp.terminate()
                        break
            offs = np.cumsum([0] + [len(p) for p in pool]).astype(np.int64)
            np.savez(cache, blob=np.frombuffer(b"".join(pool), dtype=np.uint8), offs=offs)
        else:
            import dcsfuzz
            src = "bit-level fuzzer (reference encoder unavailable)"
            rng = np.random.default_rng(seed0)
            nf = int(seconds * SAMPLE_RATE / FRAME)
            pool = [dcsfuzz.fuzz94(rng, nf, type1=i & 1, max_code=9) for i in range(min(n_streams, 32))]
    n_unique = min(len(pool), n_streams)
    streams = [pool[i % n_unique] for i in range(n_streams)]
    return streams, n_unique, src


def load_corpus_cache(n_streams, seconds, seed0):
    """The pool build_corpus left in corpus_cache/ (ranks other than 0 read what rank 0 made, so that
    every rank sees the same pool even when the encoder's time budget cut it short)."""
    cache = os.path.join(ROOT, "corpus_cache", "c2_%d_%gs_seed%d.npz" % (n_streams, seconds, seed0))
    z = np.load(cache)
    blob, offs = z["blob"], z["offs"]
    pool = [blob[offs[i]:offs[i + 1]].tobytes() for i in range(len(offs) - 1)]
    n_unique = min(len(pool), n_streams)
    return [pool[i % n_unique] for i in range(n_streams)], n_unique, "reference DCSEncoder (oracle/_ref)"


# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_run(streams, vol, lvl, tail, threads, budget_streams, os_version=0x9400):
    """Times the reference CPU decoder (or the oracle port) on streams[:budget_streams].
    Returns (Msamples/s, kind, threads, sample description, seconds, per-stream checksums or None)."""
    from oracle import ref, orc
    sample = streams[:budget_streams]
    nsamp = sum((((s[0] << 8) | s[1]) + tail) * FRAME for s in sample)
    cs_each = None
    if ref.available():
        L = ref.lib()
        n = len(sample)
        bufs = [np.frombuffer(s, dtype=np.uint8) for s in sample]
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        nbytes = np.array([b.size for b in bufs], dtype=np.uint32)
        nfr = np.array([((s[0] << 8) | s[1]) + tail for s in sample], dtype=np.uint32)
        cs = C.c_uint64(0)
        cs_each = np.zeros(n, dtype=np.uint64)
        secs = L.dcsref_decode_batch_timed(ptrs, nbytes.ctypes.data, nfr.ctypes.data, n, os_version, vol, lvl,
                                           threads, None, C.byref(cs), cs_each.ctypes.data)
        kind = "reference"
    else:
        t0 = time.time()
        cs_each = np.zeros(len(sample), dtype=np.uint64)
        for i, s in enumerate(sample):
            pcm, _ = orc.decode(s, os_version, vol, lvl, ((s[0] << 8) | s[1]) + tail)
            u = pcm.view(np.uint16).astype(np.uint64)
            cs_each[i] = np.sum(u * (2 * np.arange(u.size, dtype=np.uint64) + 1), dtype=np.uint64)
        secs = time.time() - t0
        kind, threads = "port", 1
    return nsamp / secs / 1e6, kind, threads, "%d of the workload's streams (%.1f s of audio each), %.2f s wall" % (
        len(sample), (((sample[0][0] << 8) | sample[0][1]) * FRAME) / SAMPLE_RATE, secs), secs, cs_each


def load_traffic(shapes):
    """profiles/traffic.json -> {kernel: record}, keeping only the kernels whose captured launch shape is
    the one this run launches (a capture of other kernels / other grids says nothing about this run)."""
    try:
        t = json.load(open(TRAFFIC_JSON))
    except Exception:
        return {}, None
    out = {}
    for k, rec in t.get("kernels", {}).items():
        if k in shapes and list(shapes[k]) == [rec.get("grid"), rec.get("block")]:
            out[k] = rec
    return out, t.get("source")


# ------------------------------------------------------------------------------------------
# BASELINE.json configs 1, 3, 4 (N = 1) and 5 (N > 1): bounded side measurements
def config1(ctx, dx):
    """one 10 s clip -> one 1994+ type 1.3 stream at 128 kbit/s, host-to-host latency, checked against the reference"""
    from oracle import ref, orc
    if not ref.available():
        return {"skipped": "reference encoder unavailable"}
    rng = np.random.Generator(np.random.MT19937(12345))
    t = np.arange(312500) / 31250.0
    x = (0.5 * np.sin(2 * np.pi * 440 * t) + 0.2 * np.sin(2 * np.pi * 2500 * t) + rng.normal(0, 0.05, t.size)).astype(np.float32)
    data, nf = ref.encode(x, fmt=0x9400, stype=1, subtype=3, bit_rate=128000, power_cut=0.97)
    t0 = time.perf_counter()
    want = ref.decode(data, 0x9400, 255, 0x64, nf + 2)
    tref = time.perf_counter() - t0
    ts = []
    for _ in range(6):
        t0 = time.perf_counter()
        pcm, offs, res = ctx.decode_streams([(data, 0x9400, 255, 0x64, 2)])
        ts.append(time.perf_counter() - t0)
    ok = res[0]["status"] == 0 and bool(np.array_equal(pcm[:want.size], want))
    return {"workload": "one 10 s sine+noise clip, 1994+ type 1.3 at 128 kbit/s (%d frames, %d bytes)" % (nf, len(data)),
            "latency_ms": min(ts[1:]) * 1e3, "first_call_ms": ts[0] * 1e3, "value": want.size / min(ts[1:]) / 1e6, "unit": "Msamples/s",
            "api": "dcsb_decode_streams (host in, host out)", "reference_1_thread_ms": tref * 1e3, "bit_exact_vs_reference": ok}


_KINDS3 = [(0x9302, 0), (0x9302, 1), (0x9301, 0)]


def _encode3(args):
    seed, seconds = args
    from oracle import ref
    fmt, ty = _KINDS3[seed % 3]
    data, nf = ref.encode(synth_source(seed, seconds), fmt=fmt, stype=ty, subtype=0,
                          bit_rate=RATES[(seed // 3) % 6], power_cut=CUTS[(seed // 18) % 3])
    return data, fmt


def config3(ctx, dx, torch, n=4096, pool_n=96, seconds=10.0):
    """1993-era layouts: 0x9302 types 0 / 1 and 0x9301 type 0 from the reference encoder + fuzzer-made OS93a type 1"""
    from oracle import ref, orc
    import dcsfuzz
    if not ref.available():
        return {"skipped": "reference encoder unavailable"}
    import multiprocessing as mp
    with mp.get_context("fork").Pool(len(os.sched_getaffinity(0))) as p:
        pool = p.map(_encode3, [(5000 + i, seconds) for i in range(pool_n)], chunksize=2)
    rng = np.random.default_rng(9)
    nf = int(seconds * SAMPLE_RATE / FRAME)
    pool += [(dcsfuzz.fuzz93a1(rng, nf), 0x9301) for _ in range(8)]
    streams = [(pool[i % len(pool)][0], pool[i % len(pool)][1], 255, 0x64, 2) for i in range(n)]
    batch = ctx.batch(streams)
    d_pcm = torch.empty(batch.total_samples, dtype=torch.int16, device="cuda")
    st = torch.cuda.current_stream()
    ms = []
    for i in range(6):
        batch.decode(d_pcm.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        if i >= 2:
            ms.append(batch.kernel_ms(2))
    res = batch.results(st.cuda_stream)
    # every unique stream once against the oracle (the pool is small)
    h = d_pcm.cpu().numpy()
    bad = 0
    for i in range(len(pool)):
        d, os_, vol, lvl, tail = streams[i]
        want, _ = orc.decode(d, os_, vol, lvl, ((d[0] << 8) | d[1]) + tail)
        o = batch.pcm_offset(i)
        bad += 0 if np.array_equal(h[o:o + want.size], want) else 1
    out = {"workload": "%d streams x %g s (%d unique: 0x9302 type 0 / type 1, 0x9301 type 0 from the reference encoder, 8 fuzzer-made OS93a type 1)" % (n, seconds, len(pool)),
           "ms_per_step": float(np.mean(ms)), "value": batch.total_samples / float(np.mean(ms)) / 1e3, "unit": "Msamples/s",
           "frames": int(batch.total_frames), "streams_with_errors": sum(1 for r in res if r["status"] != 0),
           "parity_checked_streams": len(pool), "mismatches": bad, "checked_against": "oracle/dcs_oracle.c (pinned to the reference)"}
    batch.close()
    del d_pcm
    return out


def _ref_encode_framed(args):
    from oracle import ref
    seed, seconds = args
    ty, sub = TYPES[seed % 3]
    d, nf = ref.encode_framed(synth_source(seed, seconds), ty, sub, RATES[(seed // 3) % 6], CUTS[(seed // 18) % 3])
    return d


def encoder_record(ctx, dx, n=128, seconds=10.0):
    """SURVEY 8(f)4, the forward path: n clips of the corpus' parameter grid through dcsb_encode_streams, every stream's
    bytes against the reference DCSEncoder fed the same framing (a bounded sample on the host cores)"""
    from oracle import ref
    seeds = [9000 + i for i in range(n)]
    clips = [synth_source(sd, seconds) for sd in seeds]
    pl = [(TYPES[sd % 3][0], TYPES[sd % 3][1], RATES[(sd // 3) % 6], CUTS[(sd // 18) % 3]) for sd in seeds]
    ts = []
    for it in range(3):
        t0 = time.perf_counter()
        got = ctx.encode_streams(clips, pl)
        ts.append(time.perf_counter() - t0)
    samples = sum(c.size for c in clips)
    res = {"workload": "%d clips x %.0f s, stream types {0.0, 1.0, 1.3} x bit rates 32k..256k x power cut {.90, .97, 1}" % (n, seconds),
           "ms_per_call": min(ts[1:]) * 1e3, "value": samples / min(ts[1:]) / 1e6, "unit": "Msamples/s encoded",
           "api": "dcsb_encode_streams (host PCM in, host stream bytes out)", "stream_bytes": int(sum(len(g) for g in got))}
    # round trip: every encoded stream back through the decode path; signal-to-noise ratio against the source (the codec
    # is lossy and delays the clip by its 16-sample overlap: best lag of 0..31) on a sample of the clips
    pcm, offs, dres = ctx.decode_streams_pinned([(g, 0x9400, 255, 0x64, 2) for g in got])
    snr = []
    for i in range(0, n, max(1, n // 16)):
        a = clips[i].astype(np.float64)
        best = -1e9
        for lag in range(32):
            b = pcm[offs[i] + lag:offs[i] + lag + a.size].astype(np.float64) / 32768.0
            m = min(a.size, b.size)
            g_ = np.dot(a[:m], b[:m]) / max(np.dot(b[:m], b[:m]), 1e-30)          # (the decoder's output gain is not unity)
            e = a[:m] - g_ * b[:m]
            best = max(best, 10.0 * np.log10(max(np.dot(a[:m], a[:m]), 1e-30) / max(np.dot(e, e), 1e-30)))
        snr.append(best)
    res["round_trip"] = {"decoded_streams": n, "streams_with_errors": sum(1 for r in dres if r["status"] != 0),
                         "median_snr_db_of_sample": float(np.median(snr)), "min_snr_db_of_sample": float(np.min(snr)), "sampled_clips": len(snr)}
    if ref.available():
        import multiprocessing as mp
        ncpu = max(1, len(os.sched_getaffinity(0)))
        k = min(n, 4 * ncpu)
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(ncpu) as pool:
            want = pool.map(_ref_encode_framed, [(sd, seconds) for sd in seeds[:k]], chunksize=1)
        tref = time.perf_counter() - t0
        res.update({"parity_checked_streams": k, "mismatches": sum(1 for a, b in zip(got[:k], want) if a != b),
                    "checked_against": "stream bytes of the reference DCSEncoder fed the same framing (oracle/_ref, dcsref_encode_framed)",
                    "reference_host_threads": ncpu, "reference_value": sum(c.size for c in clips[:k]) / tref / 1e6})
    return res


def config4(ctx, dx, torch, n=2048):
    """track playback on the ROM built by the reference's DCSCompiler: n timelines + every track solo, one call"""
    import compiledrom
    from oracle import ref
    try:
        c = compiledrom.load(compiledrom.NAMES[0])
    except Exception as e:
        return {"skipped": "compiled ROM fixture unavailable: %s" % e}
    rom = dx.Rom(c["images"])
    tls = []
    for i in range(n):
        sh = i % 13
        tls.append(([(f + sh, b) for f, b in c["writes"]], c["n_frames"] + sh, 255 - (i % 100)))
    tls += c["track_timelines"]
    frames = sum(t[1] for t in tls)
    tl_arr, keep = dx.make_timelines(tls)
    out = torch.empty(frames * 240, dtype=torch.int16).pin_memory()
    resarr = (dx.TimelineResult * len(tls))()
    ts = []
    for it in range(4):
        t0 = time.perf_counter()
        rc = ctx._L.dcsb_render_timelines(ctx._h, rom._h, tl_arr, len(tls), out.data_ptr(), None, resarr)
        ts.append(time.perf_counter() - t0)
        if rc != 0:
            return {"error": "dcsb_render_timelines: %d" % rc}
    h = out.numpy()
    checked, bad = 0, 0
    if ref.available():
        offs = np.concatenate([[0], np.cumsum([t[1] * 240 for t in tls])])
        for i in (0, 1, n // 2, n - 1, n, len(tls) - 1):
            rp = ref.RomPlayer(c["images"], tls[i][2])
            want = rp.render_timeline(tls[i][0], tls[i][1])
            rp.close()
            checked += 1
            bad += 0 if np.array_equal(h[offs[i]:offs[i + 1]], want) else 1
    best = min(ts[1:])
    res = {"workload": "%d timelines (shifted command times, master volumes 156..255) + %d solo tracks of the ROM set compiled by the reference's DCSCompiler (%s)" % (
               n, len(c["track_timelines"]), compiledrom.NAMES[0]),
           "output_frames": int(frames), "ms_per_call": best * 1e3, "value": frames * 240 / best / 1e6, "unit": "Msamples/s",
           "api": "dcsb_render_timelines (host in, pinned host out; track-program interpreter kernel + mix kernel + PCM download inside)",
           "timelines_with_errors": sum(1 for i in range(len(tls)) if resarr[i].status != 0),
           "parity_checked_timelines": checked, "mismatches": bad, "checked_against": "reference DCSDecoderNative (oracle/_ref)"}
    rom.close()
    return res


def config5(ctx, dx, torch, dist, rank, world, local_rank, total_streams, steps=3):
    """BASELINE.json config 5: total_streams one-second streams (a pool of unique streams from the reference
    encoder, replicated so that every stream has its own bytes in HBM), LPT-partitioned over the ranks by frame
    count, resident decode, per-rank checksum gathered over NCCL."""
    POOL = 16384
    if rank == 0:
        build_corpus(POOL, 1.0, 100000, budget_s=60.0)
    dist.barrier()
    pool, n_unique, src = load_corpus_cache(POOL, 1.0, 100000)                  # every rank: what rank 0 left in the cache
    pool = pool[:n_unique]
    blob = np.frombuffer(b"".join(pool), dtype=np.uint8)
    offs = np.concatenate([[0], np.cumsum([len(p) for p in pool])]).astype(np.int64)
    pframes = np.array([(p[0] << 8) | p[1] for p in pool], dtype=np.uint32)
    gidx = (np.arange(total_streams, dtype=np.int64) * 7919) % n_unique          # every rank computes the same job
    part, load = dx.partition_streams(pframes[gidx], world)
    mine = gidx[np.asarray(part) == rank]
    descs, keep = dx.make_descs_pool(blob, offs, mine, os_version=dx.OS94, master_volume=255, mixing_level=0x64, tail_frames=2)
    t0 = time.time()
    batch = dx.Batch(ctx, None, descs=descs)
    t_create = time.time() - t0
    d_pcm = torch.empty(batch.total_samples, dtype=torch.int16, device="cuda")
    st = torch.cuda.current_stream()
    for _ in range(2):
        batch.decode(d_pcm.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(st)
    for _ in range(steps):
        batch.decode(d_pcm.data_ptr(), st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ctx.set_overlap(False)
    batch.decode(d_pcm.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    scan_ms, dec_ms = batch.kernel_ms(0), batch.kernel_ms(1)
    ctx.set_overlap(True)
    res = batch.results_np(st.cuda_stream)
    # (a sum, not an xor: every pool stream is replicated an even number of times per rank)
    xor = int(np.sum(res["checksum"], dtype=np.uint64)) if res.size else 0
    # the pool's streams must decode to the same PCM wherever they sit: per-pool-stream checksums of this rank
    bad = int(np.count_nonzero(res["status"]))
    first = {}
    for k in range(min(res.size, 200000)):
        g = int(mine[k])
        if g in first:
            bad += 0 if first[g] == int(res["checksum"][k]) else 1
        else:
            first[g] = int(res["checksum"][k])
    t = torch.tensor([ms, float(batch.total_samples), float(batch.compressed_bytes), scan_ms, dec_ms], dtype=torch.float64, device="cuda")
    g = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    cs = torch.tensor([xor & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device="cuda")
    gc = [torch.zeros_like(cs) for _ in range(world)]
    dist.all_gather(gc, cs)                         # the checksum gather: 8 bytes per rank over NCCL
    per_rank_ms = [float(x[0].item()) for x in g]
    samples = sum(float(x[1].item()) for x in g)
    cbytes = sum(float(x[2].item()) for x in g)
    peak, _ = measured_peak_gbs()
    worst = max(per_rank_ms)
    out = {"workload": "%d streams x 1 s (pool of %d unique 1994+ streams from the reference encoder, every stream with its own bytes in HBM), "
                       "LPT-partitioned by frame count over %d ranks, resident decode" % (total_streams, n_unique, world),
           "value": samples / (worst * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": worst, "per_rank_ms": per_rank_ms,
           "per_rank_scan_ms": [float(x[3].item()) for x in g], "per_rank_decode_ms": [float(x[4].item()) for x in g],
           "streams_per_rank": int(mine.size), "frames_per_rank": int(batch.total_frames), "audio_hours": samples / SAMPLE_RATE / 3600,
           "compressed_bytes": cbytes, "pcm_bytes": samples * 2,
           "frac": (cbytes + samples * 2) / (worst * 1e-3) / 1e9 / (peak * world),
           "frac_note": "algorithmic bytes (compressed in + PCM out) / slowest rank's step time, against %d x the measured HBM copy bandwidth" % world,
           "errors_or_replica_mismatches_this_rank": bad, "checksum_sum_per_rank": ["%016x" % int(x.item()) for x in gc],
           "batch_create_s": t_create}
    batch.close()
    del d_pcm
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------
def main():
    # the contract is ONE JSON line on stdout: libraries that write to fd 1 themselves (NCCL prints its
    # version there) are sent to stderr, the line goes to the real stdout
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    try:
        _main(real_stdout)
    finally:
        real_stdout.flush()


def _main(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=4096)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs 1/3/4 (N = 1) and config 5 (N > 1) side measurements")
    ap.add_argument("--config5-streams", type=int, default=1048576)
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    vol, lvl, tail = 255, 0x64, 2
    workload = "%d synthetic 1994+ streams x %g s, types {0.0,1.0,1.3} x bit rates 32k..256k x power cut {.90,.97,1}" % (a.streams, a.seconds)
    ncores = len(os.sched_getaffinity(0))

    # ---------------- reference arm: the reference's own CPU decoder on the host cores
    if a.impl == "reference":
        if rank != 0:
            return
        streams, n_unique, src = build_corpus(a.streams, a.seconds, 0)
        # a step = a bounded sample of the workload, large enough for >= 1 s of wall time on all cores
        per_step = min(max(ncores * 128, 512), len(streams))
        vals = []
        for s in range(a.warmup + a.steps):
            lo = (s * per_step) % max(1, len(streams) - per_step + 1)
            v, kind, thr, sample, secs, _ = cpu_reference_run(streams[lo:lo + per_step], vol, lvl, tail, ncores, per_step)
            if s >= a.warmup:
                vals.append((v, secs))
        value = float(np.mean([v for v, _ in vals]))
        line = {"impl": "reference", "metric": "decoded PCM Msamples/s (bit-exact)", "value": value, "unit": "Msamples/s",
                "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": float(np.mean([s for _, s in vals]) * 1e3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int16/int32 fixed point", "data": "synthetic: " + src,
                "config": {"workload": workload, "step": "%d streams per step (bounded sample of the workload, a rate)" % per_step,
                           "unique_streams": n_unique},
                "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": thr, "kind": kind, "sample": sample},
                "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), file=out)
        return

    # ---------------- our arm
    import torch
    import dcsexplorer_b200 as dx
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the decoder has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if rank == 0:
        streams, n_unique, src = build_corpus(a.streams, a.seconds, 0)
    if world > 1:
        dist.barrier()
        if rank != 0:
            streams, n_unique, src = load_corpus_cache(a.streams, a.seconds, 0)   # cache written by rank 0
        # the job is world x 4,096 streams, sharded by stream with the library's own partition
        # function (LPT on frame counts; every rank computes the same assignment, no collective)
        pool = streams
        glob = [pool[(g * 997) % len(pool)] for g in range(world * a.streams)]
        part, load = dx.partition_streams([dx.stream_frames(s) for s in glob], world)
        streams = [glob[g] for g in range(len(glob)) if part[g] == rank]

    ctx = dx.Context(local_rank)
    batch = ctx.batch(streams, os_version=dx.OS94, master_volume=vol, mixing_level=lvl, tail_frames=tail)
    total_samples = batch.total_samples
    comp_bytes = int(batch.compressed_bytes)
    alg_bytes = comp_bytes + total_samples * 2
    d_pcm = torch.empty(total_samples, dtype=torch.int16, device="cuda")
    st = torch.cuda.current_stream()

    def step():
        batch.decode(d_pcm.data_ptr(), st.cuda_stream)

    for _ in range(max(3, a.warmup)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    kms = [[], []]
    torch.cuda.synchronize()
    ev[0].record(st)
    for i in range(a.steps):
        step()
        ev[i + 1].record(st)
    torch.cuda.synchronize()
    total_ms = ev[0].elapsed_time(ev[a.steps])
    my_ms = total_ms / a.steps
    # per-kernel durations: the timed steps above run the scan BESIDE the decode kernel (the
    # library's default), so a kernel's own duration is read from K more steps with the two
    # kernels one after the other, from the library's CUDA-event pairs on the launching stream
    spans = [[], []]
    for i in range(min(a.steps, 5)):
        step()
        spans[0].append(batch.kernel_ms(0))
        spans[1].append(batch.kernel_ms(1))
    ctx.set_overlap(False)
    serial_ms = []
    for i in range(a.steps):
        step()
        kms[0].append(batch.kernel_ms(0))
        kms[1].append(batch.kernel_ms(1))
        serial_ms.append(batch.kernel_ms(2))
    ctx.set_overlap(True)
    clocks = sampler.stop()
    res = batch.results(st.cuda_stream)
    bad = [r["status"] for r in res if r["status"] != 0]
    xor, csum = 0, 0
    for r in res:
        xor ^= r["checksum"]
        csum = (csum + r["checksum"]) & 0xFFFFFFFFFFFFFFFF     # (N > 1: a rank may hold both copies of a stream, which cancel in the xor)

    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    cs = torch.tensor([csum & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device="cuda")
    per_rank_ms = [my_ms]
    if world > 1:
        mine_t = torch.tensor([my_ms], dtype=torch.float64, device="cuda")
        gt = [torch.zeros_like(mine_t) for _ in range(world)]
        dist.all_gather(gt, mine_t)
        per_rank_ms = [float(x.item()) for x in gt]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(gathered, cs)               # the checksum gather: 8 bytes per rank over NCCL
        rank_sums = [int(x.item()) for x in gathered]
    else:
        rank_sums = [csum & 0x7FFFFFFFFFFFFFFF]
    total_ms = float(t.item())
    value = world * total_samples * a.steps / (total_ms * 1e-3) / 1e6

    # ---------------- end-to-end through the host-buffer C-ABI call (rank-local, then max over ranks)
    descs, keep = dx.make_descs(streams, os_version=dx.OS94, master_volume=vol, mixing_level=lvl, tail_frames=tail)
    # inputs in pinned host memory: one pinned blob, descriptors point into it
    blob = torch.empty(sum(len(s) for s in streams), dtype=torch.uint8).pin_memory()
    off = 0
    bnp = blob.numpy()
    for i, s in enumerate(streams):
        bnp[off:off + len(s)] = np.frombuffer(s, dtype=np.uint8)
        descs[i].data = blob.data_ptr() + off
        off += len(s)
    h_pcm = torch.empty(total_samples, dtype=torch.int16).pin_memory()
    resarr = (dx.Result * len(streams))()
    L = ctx._L
    e2e_ms = []
    for i in range((1 + a.e2e_steps) if a.e2e_steps > 0 else 0):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        rc = L.dcsb_decode_streams(ctx._h, descs, len(streams), h_pcm.data_ptr(), None, resarr)
        t1 = time.perf_counter()
        if rc != 0:
            raise SystemExit("dcsb_decode_streams failed: %d" % rc)
        if i > 0:
            e2e_ms.append((t1 - t0) * 1e3)
    e2e_t = torch.tensor([float(np.mean(e2e_ms)) if e2e_ms else float("inf")], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * total_samples / (float(e2e_t.item()) * 1e-3) / 1e6
    # the whole e2e output against the resident-path output, and its per-stream checksums
    same = bool(torch.equal(h_pcm, d_pcm.cpu())) if e2e_ms else None
    e2e_cs_same = all(resarr[i].checksum == res[i]["checksum"] for i in range(len(streams))) if e2e_ms else None
    # what the PCIe link alone takes for the PCM bytes (one device-to-host copy of the same size)
    d2h_ms = None
    if e2e_ms:
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            h_pcm.copy_(d_pcm, non_blocking=True)
            torch.cuda.synchronize()
            d2h_ms = (time.perf_counter() - t0) * 1e3
    shapes = {"dcsb_scan_kernel": batch.launch_shape(0), "dcsb_decode94_kernel": batch.launch_shape(1),
              "dcsb_decode94_queue_kernel": batch.launch_shape(2)}
    n_launch = batch.launches()
    del h_pcm, blob

    # ---------------- side measurements
    c5 = None
    if world > 1 and not a.no_configs:
        batch.close()
        del d_pcm
        torch.cuda.empty_cache()
        c5 = config5(ctx, dx, torch, dist, rank, world, local_rank, a.config5_streams)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        dec_ms = float(np.mean(kms[1]))
        scan_ms = float(np.mean(kms[0]))
        traffic, traffic_src = load_traffic(shapes)
        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        clk_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        # algorithmic bytes per launch: the scan reads every compressed byte once;
        # the decode kernel reads them once more and writes every PCM byte once
        kern = [
            {"kernel": "dcsb_scan_kernel", "ms": scan_ms, "algorithmic_bytes_per_launch": comp_bytes},
            {"kernel": "dcsb_decode94_kernel", "ms": dec_ms, "algorithmic_bytes_per_launch": int(alg_bytes)},
        ]
        for k in kern:
            k["achieved"] = k["algorithmic_bytes_per_launch"] / (k["ms"] * 1e-3) / 1e9
            k["frac"] = k["achieved"] / peak
            k["grid"], k["block"] = shapes[k["kernel"]]
            tr = traffic.get(k["kernel"])
            k["traffic"] = tr["traffic_bytes"] if tr else None
            # share of the SMs' issue rate: warp instructions of the capture / (duration x SMs x 4 schedulers x clock)
            k["issue_frac"] = tr["warp_insts"] / (k["ms"] * 1e-3 * sms * 4 * clk_hz) if tr else None
        dom = max(kern, key=lambda k: k["ms"])
        line = {
            "metric": "decoded PCM Msamples/s (bit-exact)", "value": value, "unit": "Msamples/s", "n_gpus": world,
            "steps": a.steps, "warmup": max(3, a.warmup), "ms_per_step": total_ms / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16/int32 fixed point", "data": "synthetic: " + src,
            "config": {"workload": workload, "per_gpu_streams": a.streams, "unique_streams": n_unique,
                       "frames_per_gpu": int(total_samples // FRAME), "compressed_bytes_per_gpu": comp_bytes,
                       "pcm_bytes_per_gpu": int(total_samples * 2), "parallelism": "streams sharded by rank, no data-path collective",
                       "l2": "inputs+outputs (%.2f GB) exceed the 126 MB L2; no flush needed" % (alg_bytes / 1e9),
                       "master_volume": vol, "mixing_level": lvl, "tail_frames": tail},
            "per_rank_ms": per_rank_ms,
            "kernels_ms": {"scan": scan_ms, "decode_transform": dec_ms, "serial_step": float(np.mean(serial_ms)),
                           "overlapped_scan_span": float(np.mean(spans[0])), "overlapped_decode_span": float(np.mean(spans[1])),
                           "note": "value/ms_per_step: scan and decode kernels resident together (default); "
                                   "scan / decode_transform: each kernel alone (dcsb_set_overlap(ctx, 0))"},
            "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": peak, "unit": "GB/s",
                         "frac": dom["frac"], "traffic": dom["traffic"], "traffic_source": traffic_src if dom["traffic"] is not None else None,
                         "issue_frac": dom["issue_frac"], "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                         "note": "dominant kernel by its own duration; it is bound by the per-stream dependent chain (latency) and "
                                 "by integer issue, not by HBM -- see DESIGN.md section 4.  traffic / issue_frac are null when the "
                                 "committed capture (profiles/traffic.json) is not of the launch shapes this run used",
                         "kernels": kern,
                         "whole_step_frac": alg_bytes / (total_ms / a.steps * 1e-3) / 1e9 / peak},
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": comp_bytes,
                    "d2h_bytes_per_step": int(total_samples * 2), "ms_per_step": float(e2e_t.item()),
                    "api": "dcsb_decode_streams (pinned host in/out)", "matches_resident_path": same,
                    "checksums_match_resident_path": e2e_cs_same, "d2h_copy_alone_ms": d2h_ms,
                    "note": "bound by the PCIe link: the PCM is 4.7x the compressed bytes; d2h_copy_alone_ms = one "
                            "cudaMemcpy of the same PCM bytes on this box"},
            "gpu_launches": n_launch * a.steps,
            "clocks": clocks,
            "status": {"streams_with_errors": len(bad), "checksum_xor": "%016x" % xor,
                       "checksum_sum_per_rank": ["%016x" % v for v in rank_sums],
                       "note": "checksum_xor: xor of rank 0's per-stream checksums (at N > 1 a rank may hold both copies of a "
                               "stream, which cancel); checksum_sum_per_rank: their sum mod 2^63, gathered over NCCL"},
        }
        if world == 1 and not a.no_cpu_baseline:
            # the reference decoder on a bounded sample of the SAME streams (~20 s of CPU work), all host cores;
            # its per-stream checksums (computed from its PCM inside oracle/ref_shim.cpp) against the GPU's
            nsample = min(2048, len(streams))
            v, kind, thr, sample, secs, cs_ref = cpu_reference_run(streams, vol, lvl, tail, ncores, nsample)
            line["cpu_baseline"] = {"value": v, "unit": "Msamples/s", "cores": thr, "kind": kind, "sample": sample}
            gpu_cs = np.array([res[i]["checksum"] for i in range(nsample)], dtype=np.uint64)
            e2e_cs = np.array([resarr[i].checksum for i in range(nsample)], dtype=np.uint64) if e2e_ms else gpu_cs
            line["parity"] = {"parity_checked_streams": int(nsample), "mismatches": int(np.count_nonzero(gpu_cs != cs_ref)),
                              "mismatches_e2e_path": int(np.count_nonzero(e2e_cs != cs_ref)),
                              "checked_against": "per-stream checksums of the PCM of %s" % ("the reference DCSDecoderNative (oracle/_ref)" if kind == "reference" else "the oracle port"),
                              "checksum": "sum over samples i of (uint16)s[i] * (2 i + 1) mod 2^64 (include/dcsb200.h)"}
        elif world > 1:
            line["cpu_baseline_note"] = "measured at N = 1 only (the other ranks would share the host cores with it)"
        if world == 1 and not a.no_configs:
            batch.close()
            del d_pcm
            torch.cuda.empty_cache()
            cfg = {}
            for name, fn in (("config1", lambda: config1(ctx, dx)), ("config3", lambda: config3(ctx, dx, torch)),
                             ("config4", lambda: config4(ctx, dx, torch)), ("encoder", lambda: encoder_record(ctx, dx))):
                t0 = time.time()
                try:
                    cfg[name] = fn()
                except Exception as e:          # a side measurement never takes the line down
                    cfg[name] = {"error": "%s: %s" % (type(e).__name__, e)}
                cfg[name]["wall_s"] = time.time() - t0
            line["configs"] = cfg
        if c5 is not None:
            line["config5"] = c5
        print(json.dumps(line), file=out)
        out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
