"""dcsexplorer_b200 -- B200-native batch decoder for DCS compressed audio streams.

Python is only the test/bench harness language here: the product is libdcsb200.so
(hand-written sm_100a CUDA kernels behind the C-ABI of include/dcsb200.h) plus the C++
DCSDecoder-compatible front end.  This module mirrors the C-ABI one to one.
"""
import ctypes as C
import numpy as np

from . import _capi
from ._capi import (OS93A, OS93B, OS94, OS95, OK, E_EMPTY, E_TRUNCATED, E_BANDTYPE, E_SHORT,
                    E_STOPPED, E_ARG, E_CUDA, E_NOMEM, StreamDesc, Result)

from ._capi import RomInfo, TrackInfo, PortWrite, Timeline, TimelineResult

__all__ = ["Context", "Batch", "Rom", "Player", "make_descs", "make_timelines", "OS93A", "OS93B", "OS94", "OS95"]


class DcsbError(RuntimeError):
    pass


def make_descs(streams, os_version=OS94, master_volume=255, mixing_level=0x64, tail_frames=2):
    """streams: list of bytes-like or (bytes, os_version[, master_volume, mixing_level, tail]) tuples.
    Returns (ctypes array of StreamDesc, keepalive list)."""
    n = len(streams)
    arr = (StreamDesc * max(1, n))()
    keep = []
    for i, s in enumerate(streams):
        osv, vol, lvl, tail = os_version, master_volume, mixing_level, tail_frames
        if isinstance(s, tuple):
            data = s[0]
            if len(s) > 1 and s[1] is not None: osv = s[1]
            if len(s) > 2 and s[2] is not None: vol = s[2]
            if len(s) > 3 and s[3] is not None: lvl = s[3]
            if len(s) > 4 and s[4] is not None: tail = s[4]
        else:
            data = s
        buf = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        keep.append(buf)
        arr[i].data = buf.ctypes.data if buf.size else None
        arr[i].nbytes = buf.size
        arr[i].os_version = osv
        arr[i].master_volume = vol
        arr[i].mixing_level = lvl
        arr[i].tail_frames = tail
    return arr, keep


DESC_DTYPE = np.dtype({"names": ["data", "nbytes", "os_version", "master_volume", "mixing_level", "tail_frames", "reserved"],
                       "formats": ["<u8", "<u4", "<u2", "u1", "u1", "<u2", "<u2"],
                       "offsets": [0, 8, 12, 14, 15, 16, 18], "itemsize": C.sizeof(StreamDesc)})
RESULT_DTYPE = np.dtype({"names": ["status", "frames", "frames_decoded", "stream_bytes", "checksum"],
                         "formats": ["<i4", "<u4", "<u4", "<u4", "<u8"],
                         "offsets": [0, 4, 8, 12, 16], "itemsize": C.sizeof(Result)})


def make_descs_pool(blob, offs, idx, os_version=OS94, master_volume=255, mixing_level=0x64, tail_frames=2):
    """Descriptors for very many streams without a Python loop: stream i is bytes
    blob[offs[idx[i]] : offs[idx[i] + 1]] of a (kept alive) uint8 numpy blob.  Several streams may
    share host bytes (a pool replicated into a large batch): dcsb_batch_create copies every stream
    to its own place in HBM.  Returns (numpy array with the dcsb_stream_desc layout, keepalive)."""
    idx = np.asarray(idx, dtype=np.int64)
    offs = np.asarray(offs, dtype=np.int64)
    d = np.zeros(max(1, idx.size), dtype=DESC_DTYPE)
    d["data"][:idx.size] = blob.ctypes.data + offs[idx]
    d["nbytes"][:idx.size] = offs[idx + 1] - offs[idx]
    d["os_version"] = os_version
    d["master_volume"] = master_volume
    d["mixing_level"] = mixing_level
    d["tail_frames"] = tail_frames
    return d, blob


def _results_to_list(res, n):
    return [dict(status=res[i].status, frames=res[i].frames, frames_decoded=res[i].frames_decoded,
                 stream_bytes=res[i].stream_bytes, checksum=res[i].checksum) for i in range(n)]


class Context:
    """dcsb_create / dcsb_destroy."""

    def __init__(self, device=0):
        self._L = _capi.lib()
        h = C.c_void_p()
        rc = self._L.dcsb_create(device, C.byref(h))
        if rc != OK:
            raise DcsbError("dcsb_create(device=%d) failed with %d: no usable CUDA device "
                            "(the decoder has no CPU fallback)" % (device, rc))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.dcsb_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc, what):
        if rc != OK:
            raise DcsbError("%s failed (%d): %s" % (what, rc, self._L.dcsb_last_error(self._h).decode()))

    def set_overlap(self, on):
        self._L.dcsb_set_overlap(self._h, 1 if on else 0)

    def decode_streams(self, streams, pcm_out=None, **kw):
        """dcsb_decode_streams with host buffers.  Returns (pcm int16 array, offsets, results)."""
        descs, keep = make_descs(streams, **kw)
        n = len(streams)
        offs, total = [], 0
        for i in range(n):
            nf = 0
            if descs[i].nbytes >= 2:
                nf = (int(keep[i][0]) << 8) | int(keep[i][1])
            offs.append(total)
            total += (nf + descs[i].tail_frames) * 240
        if pcm_out is None:
            pcm_out = np.zeros(max(total, 1), dtype=np.int16)
        res = (Result * max(1, n))()
        rc = self._L.dcsb_decode_streams(self._h, descs, n, pcm_out.ctypes.data, None, res)
        self._check(rc, "dcsb_decode_streams")
        return pcm_out[:total], offs, _results_to_list(res, n)

    def set_pipeline(self, max_chunks=0, slice_frames=0):
        """dcsb_set_pipeline: chunks / frames per time slice of decode_streams (0 = choose)."""
        self._check(self._L.dcsb_set_pipeline(self._h, max_chunks, slice_frames), "dcsb_set_pipeline")

    def decode_streams_pinned(self, streams, **kw):
        """decode_streams with the streams packed into one pinned host blob and a pinned PCM buffer
        (what a throughput-minded caller does): uploads happen in place, PCM is copied straight
        into the output, uniform chunks are time-sliced.  Returns (pcm, offsets, results)."""
        import torch
        descs, keep = make_descs(streams, **kw)
        n = len(streams)
        blob = torch.empty(max(1, sum(int(k.size) for k in keep)), dtype=torch.uint8).pin_memory()
        bnp = blob.numpy()
        offs, total, o = [], 0, 0
        for i, s in enumerate(keep):
            bnp[o:o + s.size] = s
            descs[i].data = blob.data_ptr() + o
            o += int(s.size)
            nf = ((int(s[0]) << 8) | int(s[1])) if s.size >= 2 else 0
            offs.append(total)
            total += (nf + descs[i].tail_frames) * 240
        h_pcm = torch.zeros(max(total, 1), dtype=torch.int16).pin_memory()
        res = (Result * max(1, n))()
        rc = self._L.dcsb_decode_streams(self._h, descs, n, h_pcm.data_ptr(), None, res)
        self._check(rc, "dcsb_decode_streams")
        return h_pcm.numpy()[:total].copy(), offs, _results_to_list(res, n)

    def encode_streams(self, clips, params, want_frames=False):
        """dcsb_encode_streams: clips = list of float32 arrays (mono, 31,250 Hz), params = list of
        (stream_type, stream_subtype, bit_rate, power_cut[, max_quantization_error, min_dynamic_range[, format_version]]).
        Returns the list of stream bytes (and the transformed frames, float32 [n_frames, 256] per clip)."""
        import numpy as np
        n = len(clips)
        keep = [np.ascontiguousarray(c, dtype=np.float32) for c in clips]
        ptrs = (C.c_void_p * max(1, n))(*[k.ctypes.data for k in keep])
        ns = (C.c_uint64 * max(1, n))(*[k.size for k in keep])
        pa = np.zeros(max(1, n), dtype=np.dtype([("t", "<i4"), ("s", "<i4"), ("r", "<i4"), ("c", "<f4"), ("q", "<f4"), ("d", "<f4"), ("v", "<i4")]))
        for i, p in enumerate(params):
            pa[i] = (p[0], p[1], p[2], p[3], p[4] if len(p) > 4 else 10.0 / 32768.0, p[5] if len(p) > 5 else 10.0 / 32768.0,
                     p[6] if len(p) > 6 else 0)
        cap = sum(int(self._L.dcsb_encode_bound(int(k.size))) for k in keep) + 64
        out = np.zeros(cap, dtype=np.uint8)
        offs = (C.c_uint64 * (n + 1))()
        nfr = [(int(k.size) + 239) // 240 for k in keep]
        frames = np.zeros((max(1, sum(nfr)), 256), dtype=np.float32) if want_frames else None
        rc = self._L.dcsb_encode_streams(self._h, ptrs, ns, n, pa.ctypes.data, out.ctypes.data, cap, offs,
                                         frames.ctypes.data if want_frames else None)
        self._check(rc, "dcsb_encode_streams")
        streams = [out[offs[i]:offs[i + 1]].tobytes() for i in range(n)]
        if not want_frames:
            return streams
        fl, o = [], 0
        for k in nfr:
            fl.append(frames[o:o + k])
            o += k
        return streams, fl

    def batch(self, streams, **kw):
        return Batch(self, streams, **kw)


class Batch:
    """dcsb_batch_*: streams resident in HBM."""

    def __init__(self, ctx, streams, descs=None, **kw):
        """streams: list of streams (see make_descs), or None with descs = a DESC_DTYPE numpy array
        (make_descs_pool) for batches too large for a Python loop."""
        self.ctx = ctx
        self._L = ctx._L
        if descs is not None:
            self.n = int(descs.size)
            dptr = descs.ctypes.data_as(C.POINTER(StreamDesc))
        else:
            self.n = len(streams)
            dptr, keep = make_descs(streams, **kw)
        h = C.c_void_p()
        ctx._check(self._L.dcsb_batch_create(ctx._h, dptr, self.n, C.byref(h)), "dcsb_batch_create")
        self._h = h
        self.total_samples = self._L.dcsb_batch_total_samples(h)
        self.total_frames = self._L.dcsb_batch_total_frames(h)
        self.compressed_bytes = self._L.dcsb_batch_compressed_bytes(h)

    def close(self):
        if getattr(self, "_h", None):
            self._L.dcsb_batch_destroy(self._h)
            self._h = None

    __del__ = close

    def decode(self, d_pcm=None, stream=None):
        self.ctx._check(self._L.dcsb_batch_decode(self._h, d_pcm, stream), "dcsb_batch_decode")

    def launches(self):
        return self._L.dcsb_batch_launches(self._h)

    def results(self, stream=None):
        res = (Result * max(1, self.n))()
        self.ctx._check(self._L.dcsb_batch_results(self._h, stream, res), "dcsb_batch_results")
        return _results_to_list(res, self.n)

    def results_np(self, stream=None):
        """the same as a numpy structured array (RESULT_DTYPE): for batches of 10^5 .. 10^6 streams"""
        res = (Result * max(1, self.n))()
        self.ctx._check(self._L.dcsb_batch_results(self._h, stream, res), "dcsb_batch_results")
        return np.frombuffer(res, dtype=RESULT_DTYPE, count=self.n).copy()

    def launch_shape(self, which):
        g, b = C.c_int(0), C.c_int(0)
        self.ctx._check(self._L.dcsb_batch_launch_shape(self._h, which, C.byref(g), C.byref(b)), "dcsb_batch_launch_shape")
        return g.value, b.value

    def pcm_offset(self, i):
        return self._L.dcsb_batch_pcm_offset(self._h, i)

    def read_pcm(self, i, n_samples):
        out = np.zeros(n_samples, dtype=np.int16)
        rc = self._L.dcsb_batch_read_pcm(self._h, i, out.ctypes.data, n_samples)
        if rc < 0:
            self.ctx._check(rc, "dcsb_batch_read_pcm")
        return out[:rc]

    def read_scan(self, i, n_frames):
        bp = np.zeros(n_frames, dtype=np.uint32)
        bt = np.zeros((n_frames, 16), dtype=np.uint8)
        rc = self._L.dcsb_batch_read_scan(self._h, i, bp.ctypes.data, bt.ctypes.data, n_frames)
        if rc < 0:
            self.ctx._check(rc, "dcsb_batch_read_scan")
        return bp[:rc], bt[:rc]

    def kernel_ms(self, which):
        return self._L.dcsb_batch_last_kernel_ms(self._h, which)


def make_timelines(timelines):
    """timelines: list of (writes [(frame, byte)], n_frames, master_volume) -> (ctypes array, keepalive)"""
    arr = (Timeline * max(1, len(timelines)))()
    keep = []
    for i, (writes, n_frames, vol) in enumerate(timelines):
        w = (PortWrite * max(1, len(writes)))()
        for k, (f, b) in enumerate(writes):
            w[k].frame, w[k].byte = f, b
        keep.append(w)
        arr[i].writes = w
        arr[i].n_writes = len(writes)
        arr[i].n_frames = n_frames
        arr[i].master_volume = vol
    return arr, keep


class Rom:
    """dcsb_rom_*: a ROM set (host side: AddROM / LoadROMFromZipFile / CheckROMs / track and stream lookup)."""

    def __init__(self, images=None, zip_path=None):
        self._L = _capi.lib()
        h = C.c_void_p()
        if self._L.dcsb_rom_create(C.byref(h)) != OK:
            raise DcsbError("dcsb_rom_create failed")
        self._h = h
        if images:
            for chip, data in images.items():
                buf = np.frombuffer(bytes(data), dtype=np.uint8)
                rc = self._L.dcsb_rom_add(h, chip, buf.ctypes.data, buf.size)
                if rc != OK:
                    raise DcsbError("dcsb_rom_add(U%d) failed: %d" % (chip, rc))
        if zip_path:
            rc = self._L.dcsb_rom_load_zip(h, str(zip_path).encode(), None)
            if rc != 0:
                raise DcsbError("dcsb_rom_load_zip: %d %s" % (rc, self._L.dcsb_rom_last_error(h).decode()))

    def close(self):
        if getattr(self, "_h", None):
            self._L.dcsb_rom_destroy(self._h)
            self._h = None

    __del__ = close

    def check(self):
        return self._L.dcsb_rom_check(self._h)

    def info(self):
        i = RomInfo()
        self._L.dcsb_rom_get_info(self._h, C.byref(i))
        return dict(os=i.os_version, hw=i.hw_version, channels=i.n_channels, version=i.version_number,
                    n_tracks=i.n_tracks, catalog=i.catalog_offset, post=i.post_code, signature=i.signature.decode())

    def track_info(self, track):
        t = TrackInfo()
        if not self._L.dcsb_rom_track_info(self._h, track, C.byref(t)):
            return None
        return dict(address=t.address, channel=t.channel, type=t.type, defer_code=t.defer_code,
                    looping=bool(t.looping), time=t.time)

    def decompile_track(self, track, raw=False):
        """DecompileTrackProgram: list of step dicts (raw=True: the dcsb_opcode records as bytes, and their count)"""
        n = self._L.dcsb_rom_decompile_track(self._h, track, None, 0)
        arr = (_capi.Opcode * max(1, n))()
        self._L.dcsb_rom_decompile_track(self._h, track, arr, n)
        if raw:
            return bytes(arr)[:n * C.sizeof(_capi.Opcode)], n
        return [dict(offset=o.offset, nesting_level=o.nesting_level, loop_parent=o.loop_parent, delay_count=o.delay_count,
                     opcode=o.opcode, operands=bytes(o.operand_bytes[:o.n_operand_bytes]), desc=o.desc.decode(),
                     hex_desc=o.hex_desc.decode()) for o in arr[:n]]

    def list_streams(self):
        n = self._L.dcsb_rom_list_streams(self._h, None, 0)
        out = np.zeros(max(1, n), dtype=np.uint32)
        self._L.dcsb_rom_list_streams(self._h, out.ctypes.data, n)
        return [int(x) for x in out[:n]]

    def stream_bytes(self, address, nbytes):
        left = C.c_uint32(0)
        p = self._L.dcsb_rom_pointer(self._h, address, C.byref(left))
        if not p:
            return b""
        return C.string_at(p, min(nbytes, left.value))


class Player:
    """dcsb_player_*: one decoder instance on a ROM set."""

    def __init__(self, ctx, rom):
        self.ctx, self.rom, self._L = ctx, rom, ctx._L
        h = C.c_void_p()
        ctx._check(self._L.dcsb_player_create(ctx._h, rom._h, C.byref(h)), "dcsb_player_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.dcsb_player_destroy(self._h)
            self._h = None

    __del__ = close

    def set_master_volume(self, v): self._L.dcsb_player_set_master_volume(self._h, v)
    def write_data_port(self, b): self._L.dcsb_player_write_data_port(self._h, b)
    def add_track_command(self, t): self._L.dcsb_player_add_track_command(self._h, t)
    def clear_tracks(self): self._L.dcsb_player_clear_tracks(self._h)
    def is_stream_playing(self, ch): return bool(self._L.dcsb_player_is_stream_playing(self._h, ch))

    def load_audio_stream(self, ch, address, level):
        self.ctx._check(self._L.dcsb_player_load_audio_stream(self._h, ch, address, level), "dcsb_player_load_audio_stream")

    def render(self, n_frames):
        pcm = np.zeros(max(1, n_frames * 240), dtype=np.int16)
        self.ctx._check(self._L.dcsb_player_render(self._h, n_frames, pcm.ctypes.data), "dcsb_player_render")
        return pcm[:n_frames * 240]

    def stream_info(self, address):
        si = _capi.StreamInfo()
        self.ctx._check(self._L.dcsb_player_stream_info(self._h, address, C.byref(si)), "dcsb_player_stream_info")
        return dict(nFrames=si.n_frames, nBytes=si.n_bytes, type=si.stream_type, subtype=si.stream_subtype,
                    status=si.status, header=bytes(si.header))

    def host_bytes(self):
        out = np.zeros(65536, dtype=np.uint8)
        n = self._L.dcsb_player_host_bytes(self._h, out.ctypes.data, out.size)
        return out[:n].tobytes()


def _render_timelines(self, rom, timelines):
    """dcsb_render_timelines: list of (writes, n_frames, master_volume) -> (list of pcm arrays, results)"""
    tl, keep = make_timelines(timelines)
    total = sum(t[1] for t in timelines)
    pcm = np.zeros(max(1, total * 240), dtype=np.int16)
    res = (TimelineResult * max(1, len(timelines)))()
    self._check(self._L.dcsb_render_timelines(self._h, rom._h, tl, len(timelines), pcm.ctypes.data, None, res),
                "dcsb_render_timelines")
    out, o = [], 0
    for t in timelines:
        out.append(pcm[o:o + t[1] * 240])
        o += t[1] * 240
    return out, [dict(status=res[i].status, frames=res[i].frames, checksum=res[i].checksum,
                      n_host_bytes=res[i].n_host_bytes) for i in range(len(timelines))]


Context.render_timelines = _render_timelines


def partition_streams(frames, n_parts):
    """dcsb_partition_streams: returns (part index per stream, load per part).  Host side, no GPU."""
    f = np.ascontiguousarray(frames, dtype=np.uint32)
    part = np.zeros(max(1, f.size), dtype=np.uint32)
    load = np.zeros(n_parts, dtype=np.uint64)
    rc = _capi.lib().dcsb_partition_streams(f.ctypes.data, f.size, n_parts, part.ctypes.data, load.ctypes.data)
    if rc != OK:
        raise DcsbError("dcsb_partition_streams failed: %d" % rc)
    return part[:f.size], load


def stream_frames(data):
    """frame count in a stream's preamble"""
    return ((data[0] << 8) | data[1]) if len(data) >= 2 else 0


def write_wav(path, pcm):
    """dcsb_write_wav: mono 16-bit 31,250 Hz"""
    a = np.ascontiguousarray(pcm, dtype=np.int16)
    if _capi.lib().dcsb_write_wav(str(path).encode(), a.ctypes.data, a.size) != OK:
        raise DcsbError("dcsb_write_wav(%s) failed" % path)


def write_dcs_file(path, os_version, stream):
    b = np.frombuffer(bytes(stream), dtype=np.uint8)
    if _capi.lib().dcsb_write_dcs_file(str(path).encode(), os_version, b.ctypes.data, b.size) != OK:
        raise DcsbError("dcsb_write_dcs_file(%s) failed" % path)


def read_dcs_file(path):
    """dcsb_read_dcs_file -> (format version, stream bytes)"""
    L = _capi.lib()
    osv = C.c_uint16(0)
    n = L.dcsb_read_dcs_file(str(path).encode(), C.byref(osv), None, 0)
    if n < 0:
        raise DcsbError("dcsb_read_dcs_file(%s): not a DCSa file (%d)" % (path, n))
    out = np.zeros(max(1, n), dtype=np.uint8)
    if L.dcsb_read_dcs_file(str(path).encode(), C.byref(osv), out.ctypes.data, n) != n:
        raise DcsbError("dcsb_read_dcs_file(%s): short read" % path)
    return osv.value, out[:n].tobytes()
