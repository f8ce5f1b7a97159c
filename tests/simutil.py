"""TEST INFRASTRUCTURE: ctypes access to tests/hostsim/libhostsim.so (CPU simulation of the
kernel bodies).  Never imported by the product."""
import ctypes as C
import os
import subprocess
import numpy as np
from dcsexplorer_b200 import make_descs, Result

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        d = os.path.join(_HERE, "hostsim")
        subprocess.check_call(["make", "-s", "-C", d])
        L = C.CDLL(os.environ.get("HOSTSIM_LIB", os.path.join(d, "libhostsim.so")))
        L.hostsim_decode_streams.restype = C.c_int
        L.hostsim_decode_streams.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def decode_streams(streams, slice_frames=0, **kw):
    """Same contract as dcsexplorer_b200.Context.decode_streams, executed by the simulator.
    slice_frames > 0: time-sliced (resumed scans), as dcsb_decode_streams does for uniform chunks."""
    descs, keep = make_descs(streams, **kw)
    n = len(streams)
    offs, total, nframes = [], 0, 0
    for i in range(n):
        nf = (int(keep[i][0]) << 8) | int(keep[i][1]) if descs[i].nbytes >= 2 else 0
        offs.append(total)
        total += (nf + descs[i].tail_frames) * 240
        nframes += nf
    pcm = np.zeros(max(total, 2), dtype=np.int16)
    res = (Result * max(1, n))()
    bitpos = np.zeros(max(nframes, 1), dtype=np.uint32)
    bt = np.zeros((max(nframes, 1), 16), dtype=np.uint8)
    if slice_frames:
        L = lib()
        L.hostsim_decode_streams_sliced.restype = C.c_int
        L.hostsim_decode_streams_sliced.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        rc = L.hostsim_decode_streams_sliced(descs, n, pcm.ctypes.data, res, bitpos.ctypes.data, bt.ctypes.data, slice_frames)
    else:
        rc = lib().hostsim_decode_streams(descs, n, pcm.ctypes.data, res, bitpos.ctypes.data, bt.ctypes.data)
    assert rc == 0, rc
    results = [dict(status=res[i].status, frames=res[i].frames, frames_decoded=res[i].frames_decoded,
                    stream_bytes=res[i].stream_bytes, checksum=res[i].checksum) for i in range(n)]
    return pcm[:total], offs, results, bitpos, bt


class PortWrite(C.Structure):
    _fields_ = [("frame", C.c_uint32), ("byte", C.c_uint8), ("pad", C.c_uint8 * 3)]


class Timeline(C.Structure):
    _fields_ = [("writes", C.POINTER(PortWrite)), ("n_writes", C.c_uint32), ("n_frames", C.c_uint32),
                ("master_volume", C.c_uint8), ("pad", C.c_uint8 * 3)]


class TimelineResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("frames", C.c_uint32), ("checksum", C.c_uint64),
                ("n_host_bytes", C.c_uint32), ("reserved", C.c_uint32)]


class RomInfo(C.Structure):
    _fields_ = [("os_version", C.c_uint16), ("hw_version", C.c_uint8), ("n_channels", C.c_uint8),
                ("version_number", C.c_uint16), ("n_tracks", C.c_uint16), ("catalog_offset", C.c_uint32),
                ("post_code", C.c_int32), ("signature", C.c_char * 128)]


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


def rom_render(images, timelines):
    """The product's host sequencer + the K1/K4 kernel bodies, executed on the CPU.
    Returns (list of pcm arrays, results, rom info dict, host bytes)."""
    L = lib()
    n = len(images)
    bufs = [np.frombuffer(bytes(v), dtype=np.uint8) for v in images.values()]
    ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
    sizes = (C.c_size_t * n)(*[b.size for b in bufs])
    chips = (C.c_int * n)(*list(images.keys()))
    tl, keep = make_timelines(timelines)
    total = sum(t[1] for t in timelines)
    pcm = np.zeros(max(1, total) * 240, dtype=np.int16)
    res = (TimelineResult * max(1, len(timelines)))()
    info = RomInfo()
    hb = np.zeros(1 << 16, dtype=np.uint8)
    L.hostsim_rom_render.restype = C.c_int
    L.hostsim_rom_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    rc = L.hostsim_rom_render(ptrs, sizes, chips, n, tl, len(timelines), pcm.ctypes.data, res, C.byref(info), hb.ctypes.data, hb.size)
    assert rc == 0, rc
    out, o = [], 0
    for t in timelines:
        out.append(pcm[o:o + t[1] * 240])
        o += t[1] * 240
    nh = sum(res[i].n_host_bytes for i in range(len(timelines)))
    results = [dict(status=res[i].status, frames=res[i].frames, checksum=res[i].checksum, n_host_bytes=res[i].n_host_bytes)
               for i in range(len(timelines))]
    inf = dict(os=info.os_version, hw=info.hw_version, channels=info.n_channels, n_tracks=info.n_tracks,
               catalog=info.catalog_offset, post=info.post_code, version=info.version_number)
    return out, results, inf, hb[:nh].tobytes()


def encode_streams(clips, params, want_frames=False):
    """hostsim_encode_streams: the encoder's kernel bodies on the CPU (explicit stream types).  clips / params as for
    dcsexplorer_b200.Context.encode_streams; returns the list of stream bytes (and the frames per clip)."""
    L = lib()
    n = len(clips)
    keep = [np.ascontiguousarray(c, dtype=np.float32) for c in clips]
    ptrs = (C.c_void_p * max(1, n))(*[k.ctypes.data for k in keep])
    ns = (C.c_uint64 * max(1, n))(*[k.size for k in keep])
    pa = np.zeros(max(1, n), dtype=np.dtype([("t", "<i4"), ("s", "<i4"), ("r", "<i4"), ("c", "<f4"), ("q", "<f4"), ("d", "<f4"), ("v", "<i4")]))
    for i, p in enumerate(params):
        pa[i] = (p[0], p[1], p[2], p[3], p[4] if len(p) > 4 else 10.0 / 32768.0, p[5] if len(p) > 5 else 10.0 / 32768.0, p[6] if len(p) > 6 else 0)
    nfr = [(int(k.size) + 239) // 240 for k in keep]
    cap = sum(18 + 527 * f + 8 for f in nfr) + 64
    out = np.zeros(cap, dtype=np.uint8)
    offs = (C.c_uint64 * (n + 1))()
    frames = np.zeros((max(1, sum(nfr)), 256), dtype=np.float32) if want_frames else None
    L.hostsim_encode_streams.restype = C.c_int
    L.hostsim_encode_streams.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    rc = L.hostsim_encode_streams(ptrs, ns, n, pa.ctypes.data, out.ctypes.data, cap, offs, frames.ctypes.data if want_frames else None)
    if rc != 0:
        raise RuntimeError("hostsim_encode_streams: %d" % rc)
    streams = [out[offs[i]:offs[i + 1]].tobytes() for i in range(n)]
    if not want_frames:
        return streams
    fl, o = [], 0
    for k in nfr:
        fl.append(frames[o:o + k])
        o += k
    return streams, fl
