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


def decode_streams(streams, **kw):
    """Same contract as dcsexplorer_b200.Context.decode_streams, executed by the simulator."""
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
    rc = lib().hostsim_decode_streams(descs, n, pcm.ctypes.data, res, bitpos.ctypes.data, bt.ctypes.data)
    assert rc == 0, rc
    results = [dict(status=res[i].status, frames=res[i].frames, frames_decoded=res[i].frames_decoded,
                    stream_bytes=res[i].stream_bytes, checksum=res[i].checksum) for i in range(n)]
    return pcm[:total], offs, results, bitpos, bt
