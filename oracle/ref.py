"""TEST INFRASTRUCTURE ONLY: ctypes access to oracle/_ref/libdcsref.so (the UNMODIFIED
reference decoder/encoder compiled by oracle/Makefile).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libdcsref.so"))


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(os.path.join(_HERE, "_ref", "libdcsref.so"))
        L.dcsref_decode_stream.restype = C.c_int
        L.dcsref_decode_stream.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.dcsref_stream_info.restype = C.c_int
        L.dcsref_stream_info.argtypes = [C.c_void_p, C.c_size_t, C.c_int] + [C.POINTER(C.c_int)] * 4
        L.dcsref_probe_frames.restype = C.c_int
        L.dcsref_probe_frames.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint16, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.dcsref_transform.restype = C.c_int
        L.dcsref_transform.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.dcsref_encode.restype = C.c_size_t
        L.dcsref_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_float, C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]
        L.dcsref_read_dcs_file.restype = C.c_long
        L.dcsref_read_dcs_file.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]
        L.dcsref_decode_batch_timed.restype = C.c_double
        L.dcsref_decode_batch_timed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                                                C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p]
        L.dcsref_rom_open_zip.restype = C.c_void_p
        L.dcsref_rom_open_zip.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
        L.dcsref_rom_close.argtypes = [C.c_void_p]
        L.dcsref_rom_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 5
        L.dcsref_rom_write_port.argtypes = [C.c_void_p, C.c_uint8]
        L.dcsref_rom_set_master_volume.argtypes = [C.c_void_p, C.c_int]
        L.dcsref_rom_render.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.dcsref_rom_host_bytes.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.dcsref_rom_list_streams.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.dcsref_rom_open_images.restype = C.c_void_p
        L.dcsref_rom_open_images.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.dcsref_rom_track_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.dcsref_rom_decompile.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _LIB = L
    return _LIB


def encode(pcm, sample_rate=31250, fmt=0x9400, stype=1, subtype=3, bit_rate=128000, power_cut=0.97):
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    cap = 64 + pcm.size * 2 + 65536
    out = np.zeros(cap, dtype=np.uint8)
    nf = C.c_int(0)
    n = lib().dcsref_encode(pcm.ctypes.data, pcm.size, sample_rate, fmt, stype, subtype, bit_rate,
                            power_cut, out.ctypes.data, cap, C.byref(nf))
    if n == 0:
        raise RuntimeError("reference encoder refused the stream")
    return out[:n].tobytes(), nf.value


def encode_framed(pcm, stype=1, subtype=3, bit_rate=128000, power_cut=0.97, max_quant_error=-1.0, min_dynamic_range=-1.0,
                  want_frames=False, fmt=0x9400):
    """The reference encoder without its resampler (samples framed directly, then the reference's own TransformFrame /
    CloseStream): the oracle for the GPU encoder.  Returns (stream bytes, n_frames[, frames float32 [n_frames, 256]])."""
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    cap = 64 + pcm.size * 3 + 65536
    out = np.zeros(cap, dtype=np.uint8)
    nf = C.c_int(0)
    nfr = (pcm.size + 239) // 240 + 1
    frames = np.zeros((nfr, 256), dtype=np.float32) if want_frames else None
    L = lib()
    L.dcsref_encode_framed.restype = C.c_size_t
    L.dcsref_encode_framed.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                       C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.c_void_p, C.c_size_t, C.c_int]
    n = L.dcsref_encode_framed(pcm.ctypes.data, pcm.size, stype, subtype, bit_rate, power_cut, max_quant_error, min_dynamic_range,
                               out.ctypes.data, cap, C.byref(nf), frames.ctypes.data if want_frames else None, nfr, fmt)
    if n == 0:
        raise RuntimeError("reference encoder refused the stream")
    if want_frames:
        return out[:n].tobytes(), nf.value, frames[:nf.value]
    return out[:n].tobytes(), nf.value


def decode(data, os_version=0x9400, master_volume=255, mixing_level=0x64, n_frames=None):
    a = np.frombuffer(data, dtype=np.uint8)
    if n_frames is None:
        n_frames = ((int(a[0]) << 8) | int(a[1])) + 2
    pcm = np.zeros(n_frames * 240, dtype=np.int16)
    lib().dcsref_decode_stream(a.ctypes.data, a.size, os_version, master_volume, mixing_level, n_frames, pcm.ctypes.data)
    return pcm


def stream_info(data, os_version=0x9400):
    a = np.frombuffer(data, dtype=np.uint8)
    v = [C.c_int(0) for _ in range(4)]
    lib().dcsref_stream_info(a.ctypes.data, a.size, os_version, *[C.byref(x) for x in v])
    return dict(nFrames=v[0].value, nBytes=v[1].value, type=v[2].value, subtype=v[3].value)


def probe_frames(data, os_version=0x9400, mix_mult=0x7FFF, max_frames=65535):
    a = np.frombuffer(data, dtype=np.uint8)
    nf = min((int(a[0]) << 8) | int(a[1]), max_frames)
    bitpos = np.zeros(nf + 1, dtype=np.uint32)
    bt = np.zeros((nf, 16), dtype=np.uint16)
    bins = np.zeros((nf, 512), dtype=np.uint16)
    stop = np.zeros(nf, dtype=np.uint8)
    lib().dcsref_probe_frames(a.ctypes.data, a.size, os_version, mix_mult, nf, bitpos.ctypes.data,
                              bt.ctypes.data, bins.ctypes.data, stop.ctypes.data)
    return bitpos, bt, bins, stop


def transform(bins, overlap, os_version=0x9400, vol_shift=0):
    b = np.zeros(512, dtype=np.uint16)
    b[:len(bins)] = np.asarray(bins).astype(np.uint16)
    o = np.ascontiguousarray(np.asarray(overlap).astype(np.uint16))
    pcm = np.zeros(240, dtype=np.int16)
    lib().dcsref_transform(os_version, vol_shift, b.ctypes.data, o.ctypes.data, pcm.ctypes.data)
    return pcm, o.astype(np.int16), b


class RomPlayer:
    """The reference decoder on a ROM set: images = {chip number: bytes}.  SoftBoot +
    SetMasterVolume(master_volume), then write_port / render as the host would."""

    def __init__(self, images, master_volume=255):
        self._keep = [np.frombuffer(bytes(v), dtype=np.uint8) for v in images.values()]
        n = len(self._keep)
        ptrs = (C.c_void_p * n)(*[k.ctypes.data for k in self._keep])
        sizes = (C.c_size_t * n)(*[k.size for k in self._keep])
        chips = (C.c_int * n)(*list(images.keys()))
        self._h = lib().dcsref_rom_open_images(ptrs, sizes, chips, n, master_volume)

    def info(self):
        v = [C.c_int(0) for _ in range(5)]
        lib().dcsref_rom_info(self._h, *[C.byref(x) for x in v])
        return dict(os=v[0].value, hw=v[1].value, max_track=v[2].value, channels=v[3].value, check=v[4].value)

    def write_port(self, b):
        lib().dcsref_rom_write_port(self._h, b)

    def render(self, n_frames):
        pcm = np.zeros(n_frames * 240, dtype=np.int16)
        lib().dcsref_rom_render(self._h, n_frames, pcm.ctypes.data)
        return pcm

    def host_bytes(self):
        out = np.zeros(65536, dtype=np.uint8)
        n = lib().dcsref_rom_host_bytes(self._h, out.ctypes.data, out.size)
        return out[:n].tobytes()

    def list_streams(self):
        out = np.zeros(4096, dtype=np.uint32)
        n = lib().dcsref_rom_list_streams(self._h, out.ctypes.data, out.size)
        return [int(x) for x in out[:n]]

    def track_info(self, track):
        out = np.zeros(7, dtype=np.uint32)
        lib().dcsref_rom_track_info(self._h, track, out.ctypes.data)
        return [int(x) for x in out]

    def decompile(self, track, cap=512):
        """DecompileTrackProgram as raw records (the byte layout of dcsb_opcode, 128 bytes each)"""
        buf = np.zeros(cap * 128, dtype=np.uint8)
        n = lib().dcsref_rom_decompile(self._h, track, buf.ctypes.data, cap)
        return buf[:min(n, cap) * 128].tobytes(), n

    def render_timeline(self, writes, n_frames):
        """writes: list of (frame, byte), sorted by frame"""
        out = np.zeros(n_frames * 240, dtype=np.int16)
        w = 0
        for f in range(n_frames):
            while w < len(writes) and writes[w][0] <= f:
                self.write_port(writes[w][1])
                w += 1
            out[f * 240:(f + 1) * 240] = self.render(1)
        return out

    def close(self):
        if self._h:
            lib().dcsref_rom_close(self._h)
            self._h = None

    __del__ = close


def read_dcs_file(path):
    """DCSEncoder::IsDCSFile + EncodeDCSFile (pass-through branch): (format version, stream bytes, frames),
    or None when the reference does not take the file for a DCS stream file"""
    fmt, nf = C.c_int(0), C.c_int(0)
    out = np.zeros(1 << 22, dtype=np.uint8)
    n = lib().dcsref_read_dcs_file(str(path).encode(), C.byref(fmt), out.ctypes.data, out.size, C.byref(nf))
    if n < 0:
        return None
    return fmt.value, out[:n].tobytes(), nf.value
