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

__all__ = ["Context", "Batch", "make_descs", "OS93A", "OS93B", "OS94", "OS95"]


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

    def batch(self, streams, **kw):
        return Batch(self, streams, **kw)


class Batch:
    """dcsb_batch_*: streams resident in HBM."""

    def __init__(self, ctx, streams, **kw):
        self.ctx = ctx
        self._L = ctx._L
        self.n = len(streams)
        descs, keep = make_descs(streams, **kw)
        h = C.c_void_p()
        ctx._check(self._L.dcsb_batch_create(ctx._h, descs, self.n, C.byref(h)), "dcsb_batch_create")
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
