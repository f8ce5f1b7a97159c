"""TEST INFRASTRUCTURE: builds synthetic DCS ROM sets (U2..U9 images) from track programs and
compressed streams, in the layout the reference decoder parses (DCSDecoder.cpp:26-76 AddROM /
MakeROMPointer, :207-234 FindCatalog, :236-504 CheckROMs version probes, :622-651 channel-count
probe).  No commercial data: the "code" areas only hold the opcode patterns the version
detection searches for.  Never imported by the product."""
import struct
import numpy as np

OS93A, OS93B, OS94, OS95 = 0x9301, 0x9302, 0x9400, 0x9500


class Track:
    """Assembler for a track: type 1 = byte-code program, 2/3 = deferred link."""

    def __init__(self, channel, ttype=1, link=0):
        self.channel, self.type, self.link = channel, ttype, link
        self.code = bytearray()
        self.fixups = []            # (offset of a U24 operand, stream key)

    def _op(self, wait, op, operands=b""):
        self.code += struct.pack(">HB", wait, op) + bytes(operands)
        return self

    def stop(self, wait=0): return self._op(wait, 0x00)

    def play(self, stream_key, channel=None, repeat=1, wait=0):
        ch = self.channel if channel is None else channel
        self._op(wait, 0x01, bytes([ch, 0, 0, 0, repeat]))
        self.fixups.append((len(self.code) - 4, stream_key))
        return self

    def stop_channel(self, ch, wait=0): return self._op(wait, 0x02, bytes([ch]))
    def queue(self, track, wait=0): return self._op(wait, 0x03, struct.pack(">H", track))
    def host_byte(self, b, wait=0): return self._op(wait, 0x04, bytes([b]))
    def host_timer93a(self, b, counter, wait=0): return self._op(wait, 0x04, bytes([b]) + struct.pack(">H", counter))
    def start_deferred(self, ch, wait=0): return self._op(wait, 0x05, bytes([ch]))
    def set_var(self, idx, val, wait=0): return self._op(wait, 0x06, bytes([idx, val]))
    def nop93_06(self, wait=0): return self._op(wait, 0x06)             # 1993 software: opcode 6 has no operands

    def mix(self, mode, ch, level, steps=None, wait=0):
        """mode 0 = set, 1 = increase, 2 = decrease; level is a signed byte"""
        if steps is None:
            return self._op(wait, 0x07 + mode, bytes([ch, level & 0xFF]))
        return self._op(wait, 0x0A + mode, bytes([ch, level & 0xFF]) + struct.pack(">H", steps))

    def nop(self, wait=0): return self._op(wait, 0x0D)
    def loop(self, count, wait=0): return self._op(wait, 0x0E, bytes([count]))
    def end_loop(self, wait=0): return self._op(wait, 0x0F)
    def op10(self, ch, val, wait=0): return self._op(wait, 0x10, bytes([ch, val]))
    def op11(self, ch, delta, steps, dec=False, wait=0): return self._op(wait, 0x12 if dec else 0x11, bytes([ch, delta]) + struct.pack(">H", steps))
    def wait_forever(self): self.code += b"\xFF\xFF\x0D"; return self
    def raw(self, b): self.code += bytes(b); return self

    def image(self):
        if self.type == 1:
            return bytes([1, self.channel]) + bytes(self.code)
        return bytes([self.type, self.channel]) + struct.pack(">H", self.link)


def _put_opcodes(u2, offset, hexops):
    for i, h in enumerate(hexops.split()):
        v = int(h, 16)
        u2[offset + 4 * i: offset + 4 * i + 3] = bytes([(v >> 16) & 0xFF, (v >> 8) & 0xFF, v & 0xFF])


def _checksum(img):
    a = np.frombuffer(bytes(img), dtype=np.uint8)
    return ((int(a[0::2].sum()) & 0xFF) << 8) | (int(a[1::2].sum()) & 0xFF)


def build_rom(os_version, tracks, streams, n_chips=2, chip_size=1 << 19, channels=6, version=None,
              signature=b"DCSB200 synthetic test ROM (C) 2026", indirect_tables=None, empty_tracks=()):
    """tracks: list of Track (index = track number; None = unpopulated slot); streams: dict
    key -> bytes.  Returns {chip number: bytes}.  Streams are spread round-robin over the chips
    (U2 first) to exercise the bank addressing of both board generations."""
    dcs95 = os_version == OS95
    shift = 21 if dcs95 else 20
    cat = 0x6000 if dcs95 else 0x4000
    chips = [bytearray(b"\xFF" * chip_size) for _ in range(n_chips)]
    u2 = chips[0]
    u2[0:4] = b"\x18\x00\x0F\xFF"                      # JUMP in the reset vector
    u2[4:4 + len(signature) + 1] = signature + b"\0"
    for c in range(1, n_chips):                         # sound ROM labels, as the zip loader matches them
        lab = ("S%d synthetic  01/02/26" % (c + 2)).encode() + b"\0"
        chips[c][0:len(lab)] = lab
    # channel-count probe pattern (GetNumChannels)
    _put_opcodes(u2, 0x0100, "22200F 4000%X4 26E20F 221800 90000A 80000A 400%02X4 26E20F 180001" % (channels, (1 << channels) - 1))
    if os_version in (OS93A, OS93B):
        _put_opcodes(u2, 0x1400 + 0x40, "380026 3C1005 0C00C0")
    if os_version == OS93A:
        _put_opcodes(u2, 0x2800 + 0x40, "47FFF2 47C946")
    if dcs95 and version:
        v = "%04X" % version
        _put_opcodes(u2, 0x2C00 + 0x40, "4%sE 0F16F8 93300E 18000F 4%sE 0F1608 0F16F8 93300E 18000F" % (v, v))
    # data area: tracks and tables in U2 behind the catalog, streams everywhere
    free = [cat + 0x1000] + [0x100] * (n_chips - 1)

    def alloc(chip, n, align=1):
        o = (free[chip] + align - 1) // align * align
        if o + n > chip_size:
            raise ValueError("chip %d full" % chip)
        free[chip] = o + n
        return o

    stream_addr = {}
    for i, (key, data) in enumerate(streams.items()):
        chip = i % n_chips
        o = alloc(chip, len(data) + 8)
        chips[chip][o:o + len(data)] = data
        stream_addr[key] = (chip << shift) | o
    ntr = len(tracks)
    index = alloc(0, 3 * ntr + 3)
    for t, tr in enumerate(tracks):
        if tr is None or t in empty_tracks:
            u2[index + 3 * t: index + 3 * t + 3] = b"\xFF\xFF\xFF"
            continue
        img = bytearray(tr.image())
        for ofs, key in tr.fixups:
            a = stream_addr[key]
            img[2 + ofs: 2 + ofs + 3] = bytes([(a >> 16) & 0xFF, (a >> 8) & 0xFF, a & 0xFF])
        o = alloc(0, len(img) + 4)
        u2[o:o + len(img)] = img
        u2[index + 3 * t: index + 3 * t + 3] = bytes([(o >> 16) & 0xFF, (o >> 8) & 0xFF, o & 0xFF])      # chip 0: linear address = offset
    di = alloc(0, 64)
    if indirect_tables:
        for k, table in enumerate(indirect_tables):
            o = alloc(0, 2 * len(table) + 2)
            for j, trk in enumerate(table):
                u2[o + 2 * j: o + 2 * j + 2] = struct.pack(">H", trk)
            u2[di + 3 * k: di + 3 * k + 3] = bytes([(o >> 16) & 0xFF, (o >> 8) & 0xFF, o & 0xFF])
    # catalog
    u2[cat:cat + 0x48] = b"\0" * 0x48
    u2[cat + 0x40: cat + 0x43] = bytes([(index >> 16) & 0xFF, (index >> 8) & 0xFF, index & 0xFF])
    u2[cat + 0x43: cat + 0x46] = bytes([(di >> 16) & 0xFF, (di >> 8) & 0xFF, di & 0xFF])
    u2[cat + 0x46: cat + 0x48] = struct.pack(">H", ntr)
    for c in range(n_chips):
        sel = (c << 9) if dcs95 else (c << 8)
        ck = _checksum(chips[c]) if c else 0
        u2[cat + 6 * c: cat + 6 * c + 6] = struct.pack(">HHH", chip_size // 4096, sel, ck)
    # balance bytes: U2's own even / odd byte sums must come out as 0
    u2[cat + 0x32] = u2[cat + 0x33] = 0
    ck = _checksum(u2)
    u2[cat + 0x32] = (-(ck >> 8)) & 0xFF
    u2[cat + 0x33] = (-(ck & 0xFF)) & 0xFF
    assert _checksum(u2) == 0
    return {c + 2: bytes(chips[c]) for c in range(n_chips)}, stream_addr


def command_bytes(track):
    return [(track >> 8) & 0xFF, track & 0xFF]


def volume_bytes(vol):
    return [0x55, 0xAA, vol & 0xFF, (~vol) & 0xFF]


def channel_volume_bytes(ch, lvl):
    return [0x55, 0xAB + ch, lvl & 0xFF, (~lvl) & 0xFF]
