"""TEST INFRASTRUCTURE ONLY: ctypes access to oracle/liboracle.so (the plain-C restatement,
oracle/dcs_oracle.c).  Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Frame(C.Structure):
    _fields_ = [("bitpos", C.c_uint32), ("bt", C.c_uint8 * 16)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.dcso_scan.restype = C.c_int
        L.dcso_scan.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        L.dcso_decode_frame.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint16, C.c_void_p, C.POINTER(C.c_int)]
        L.dcso_transform.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.dcso_master_multiplier.restype = C.c_uint16
        L.dcso_master_multiplier.argtypes = [C.c_int]
        L.dcso_level_multiplier.restype = C.c_uint16
        L.dcso_level_multiplier.argtypes = [C.c_int] * 4
        L.dcso_decode_stream.restype = C.c_int
        L.dcso_decode_stream.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.dcso_fnv1a.restype = C.c_uint64
        L.dcso_fnv1a.argtypes = [C.c_void_p, C.c_size_t]
        _LIB = L
    return _LIB


def _padded(data):
    a = np.zeros(len(data) + 64, dtype=np.uint8)
    a[:len(data)] = np.frombuffer(data, dtype=np.uint8)
    return a


def n_frames(data):
    return (data[0] << 8) | data[1]


def scan(data, os_version=0x9400):
    a = _padded(data)
    nf = n_frames(data)
    fr = (Frame * (nf + 1))()
    stop = C.c_int(-1)
    rc = lib().dcso_scan(a.ctypes.data, len(data), os_version, fr, C.byref(stop))
    bitpos = np.array([f.bitpos for f in fr], dtype=np.uint32)
    bt = np.array([list(f.bt) for f in fr], dtype=np.uint8)
    return rc, bitpos, bt, stop.value


def decode_frames(data, os_version=0x9400, mix_mult=0x7FFF):
    """bins each frame adds into a zeroed frame buffer (for comparison with ref.probe_frames)"""
    a = _padded(data)
    nf = n_frames(data)
    fr = (Frame * (nf + 1))()
    stop = C.c_int(-1)
    rc = lib().dcso_scan(a.ctypes.data, len(data), os_version, fr, C.byref(stop))
    assert rc == nf, rc
    bins = np.zeros((nf, 512), dtype=np.uint16)
    for f in range(nf):
        lib().dcso_decode_frame(a.ctypes.data, os_version, C.byref(fr[f]), mix_mult, bins[f].ctypes.data, None)
    return bins


def transform(bins, overlap, os_version=0x9400, vol_shift=0):
    b = np.zeros(512, dtype=np.uint16)
    b[:len(bins)] = np.asarray(bins).astype(np.uint16)
    o = np.ascontiguousarray(np.asarray(overlap).astype(np.uint16))
    pcm = np.zeros(240, dtype=np.int16)
    lib().dcso_transform(os_version, b.ctypes.data, o.ctypes.data, vol_shift, pcm.ctypes.data)
    return pcm, o.astype(np.int16)


def decode(data, os_version=0x9400, master_volume=255, mixing_level=0x64, n_frames_out=None):
    a = _padded(data)
    if n_frames_out is None:
        n_frames_out = n_frames(data) + 2
    pcm = np.zeros(n_frames_out * 240, dtype=np.int16)
    rc = lib().dcso_decode_stream(a.ctypes.data, len(data), os_version, master_volume, mixing_level,
                                  n_frames_out, pcm.ctypes.data)
    return pcm, rc


def fnv1a(pcm):
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    return lib().dcso_fnv1a(pcm.ctypes.data, pcm.size)
