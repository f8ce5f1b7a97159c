"""Bit-stream fuzzer: emits random VALID DCS streams directly at the bit level for every
frame layout the decoder knows (1994 type 0/1 all subtypes, 1993 types 0/1, OS93a type 1),
reaching branches the reference encoder never produces (fixed-width band types 7..15,
half-density bands, 'two zeros' escapes at band end, zero-type repeat/ramp bands, the
OS93a vector-quantised format).  The code tables are read from the generated
oracle/dcs_tables.h, so the fuzzer shares no decode logic with either implementation.
Layouts follow DCSDecoder/DCSDecoderNative.cpp:1679-2261, :2293-2684, :2831-3032."""
import os
import re
import numpy as np

_T = None


def _tables():
    global _T
    if _T is not None:
        return _T
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "dcs_tables.h")
    s = open(path).read()
    t = {}
    for m in re.finditer(r"dcs_code_t (\w+)\[\d+\] = \{(.*?)\};", s, re.S):
        t[m.group(1)] = [(int(c, 16), int(l), int(v, 16))
                         for c, l, v in re.findall(r"\{0x([0-9a-f]+),(\d+),0x([0-9a-f]+)\}", m.group(2))]
    for m in re.finditer(r"uint8_t (dcs94_xlat_\w+)\[16\]\[2\] = \{(.*?)\};", s, re.S):
        t[m.group(1)] = [(int(a), int(b)) for a, b in re.findall(r"\{(\d+),(\d+)\}", m.group(2))]
    m = re.search(r"dcs93a_inputs_per_band\[18\] = \{(.*?)\};", s)
    t["ipb"] = [int(x) for x in m.group(1).split(",")]
    _T = t
    return t


class BitWriter:
    def __init__(self):
        self.bits = []

    def put(self, value, n):
        for i in range(n - 1, -1, -1):
            self.bits.append((int(value) >> i) & 1)

    def code(self, entry):
        self.put(entry[0], entry[1])

    def tobytes(self):
        b = self.bits + [0] * ((-len(self.bits)) % 8)
        return np.packbits(np.array(b, dtype=np.uint8)).tobytes() if b else b""


BAND_LEN94 = [7, 8] + [16] * 13 + [32]


def fuzz94(rng, nframes=8, type1=None, subtype=None, nbands=None, error_frame=None,
           max_code=15, escape_p=0.15, hold_p=0.5):
    """Returns stream bytes (1994+ layout)."""
    T = _tables()
    type1 = int(rng.integers(0, 2)) if type1 is None else type1
    subtype = int(rng.integers(0, 4)) if subtype is None else subtype
    nbands = int(rng.integers(1, 17)) if nbands is None else nbands
    hdr = bytearray([0x7F] * 16)
    for b in range(nbands):
        hdr[b] = int(rng.integers(0, 0x3F)) | (0x40 if rng.random() < 0.2 else 0)
        if (hdr[b] & 0x7F) == 0x7F:
            hdr[b] = 0x3E
    if type1:
        hdr[0] |= 0x80
    if subtype & 2:
        hdr[1] |= 0x80
    if subtype & 1:
        hdr[2] |= 0x80
    hdr_codes = {v - 0x2E: (c, l, v) for c, l, v in T["dcs94_hdr"]}
    cbs = [None] + [T["dcs94_cb%d" % k] for k in range(1, 7)]
    w = BitWriter()
    bt = [0] * 16
    for f in range(nframes):
        for b in range(nbands):
            if f > 0 and rng.random() < hold_p:
                new = bt[b]
            else:
                new = int(rng.integers(0, max_code + 1))
            if (new - bt[b]) not in hdr_codes:
                new = bt[b]
            w.code(hdr_codes[new - bt[b]])
            bt[b] = new
        for b in range(nbands):
            count = BAND_LEN94[b]
            if hdr[b] & 0x40:
                count //= 2
            code = bt[b]
            if code == 0:
                continue
            if type1:
                x = T["dcs94_xlat_lo" if b < 3 else "dcs94_xlat_mid" if b < 6 else "dcs94_xlat_hi"]
                code = x[code][0]
            if code <= 6:
                cb = cbs[code]
                esc = [e for e in cb if e[2] & 0x80][0]
                vals = [e for e in cb if not (e[2] & 0x80)]
                remaining = count
                while remaining > 0:
                    if remaining >= 2 and rng.random() < escape_p:
                        w.code(esc)
                        remaining -= 2
                    elif remaining == 1 and error_frame == f:
                        w.code(esc)      # the malformed case the reference flags (:2213-2218)
                        remaining -= 1
                    else:
                        w.code(vals[int(rng.integers(0, len(vals)))])
                        remaining -= 1
            else:
                lim = 1 << code
                for _ in range(count):
                    w.put(int(rng.integers(0, lim)), code)
    body = w.tobytes()
    return bytes([nframes >> 8, nframes & 0xFF]) + bytes(hdr) + body


def fuzz93(rng, nframes=8, type1=None, nbands=None, small=False):
    """1993 layout (OS93b types 0/1; OS93a type 0 is the same as OS93b type 0)."""
    T = _tables()
    type1 = int(rng.integers(0, 2)) if type1 is None else type1
    nbands = int(rng.integers(1, 17)) if nbands is None else nbands
    hdr = bytearray([0x7F] * 16)
    span = 1
    for b in range(nbands):
        stride2 = rng.random() < 0.2
        need = (32 if not type1 else 16) if stride2 else 16
        if span + need > 500:
            nbands = b
            break
        span += need
        hdr[b] = int(rng.integers(0x10 if small else 0, 0x30 if small else 0x3F)) | (0x40 if stride2 else 0)
        if (hdr[b] & 0x7F) == 0x7F:
            hdr[b] = 0x3E
    if type1:
        hdr[0] |= 0x80
    codes93 = {}
    for c, l, v in T["dcs93_hdr"]:
        codes93[v] = (c, l, v)
    w = BitWriter()
    bt = [0] * 16
    for f in range(nframes):
        subtype = 0 if type1 else 2
        reuse = False
        code = 0
        first = True
        for b in range(nbands):
            stride2 = bool(hdr[b] & 0x40)
            if not type1:
                n = 16
            else:
                n = 8 if stride2 else (15 if first else 16)
            if reuse:
                reuse = rng.random() < 0.5
                w.put(1 if reuse else 0, 1)
            if not reuse:
                if not type1:
                    if rng.random() < 0.4:
                        w.put(1, 1)
                        up = int(rng.integers(0, 2))
                        w.put(up, 1)
                        subtype = [1, 2, 0][subtype] if up else [2, 0, 1][subtype]
                    else:
                        w.put(0, 1)
                    code = int(rng.integers(0, 16)) if rng.random() < 0.8 else 0
                    w.put(code, 4)
                else:
                    # pick a new band type reachable with an existing delta code
                    for _ in range(50):
                        new = int(rng.integers(0, 16)) if rng.random() < 0.7 else bt[b]
                        if rng.random() < 0.25:
                            new = 0
                        d = new - bt[b]
                        flip = rng.random() < 0.3
                        v = d + 0x2E if flip else d + 0x0F
                        if flip and not (0x1E <= v <= 0x3D):
                            continue
                        if (not flip) and not (0 <= v < 0x1E):
                            continue
                        if v in codes93:
                            break
                    else:
                        new, flip, v = bt[b], False, 0x0F
                    w.code(codes93[v])
                    if flip:
                        subtype = 0 if subtype else 1
                    bt[b] = new
                    code = new
            if code == 0:
                reuse = True
            else:
                width = code + (0 if type1 else 1)
                lim = 1 << width
                for _ in range(n):
                    if small and width > 4:
                        v = int(rng.integers(-8, 8)) & (lim - 1)
                    else:
                        v = int(rng.integers(0, lim))
                    w.put(v, width)
            first = False
    return bytes([nframes >> 8, nframes & 0xFF]) + bytes(hdr) + w.tobytes()


def fuzz93a1(rng, nframes=8, sel=None, nbands=None):
    """OS93a type-1 layout (one header byte: 1 pp bbbbb)."""
    T = _tables()
    sel = int(rng.integers(0, 4)) if sel is None else sel
    nbands = int(rng.integers(1, 19)) if nbands is None else nbands
    hb = 0x80 | (sel << 5) | nbands
    bbt = T["dcs93a_bandbits%d" % sel]
    ends = [e for e in bbt if e[2] == 0xFF]
    normal = [e for e in bbt if e[2] != 0xFF]
    scale = T["dcs93a_scale"]
    w = BitWriter()
    for f in range(nframes):
        for b in range(nbands):
            if rng.random() < 0.04:
                w.code(ends[0])
                break
            e = normal[int(rng.integers(0, len(normal)))]
            w.code(e)
            bits = e[2]
            if bits == 0:
                continue
            if rng.random() < 0.7:
                se = [x for x in scale if x[2] <= 3][int(rng.integers(0, 4))]
            else:
                se = scale[int(rng.integers(0, len(scale)))]
            w.code(se)
            for _ in range(T["ipb"][b]):
                w.put(int(rng.integers(0, 1 << bits)), bits)
    return bytes([nframes >> 8, nframes & 0xFF, hb]) + w.tobytes()


def corpus(seed=0, n_each=6, nframes=12):
    """A mixed bag: list of (os_version, stream_bytes, label)."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_each):
        out.append((0x9400, fuzz94(rng, nframes, type1=0, max_code=15), "94-t0-%d" % i))
        out.append((0x9400, fuzz94(rng, nframes, type1=1), "94-t1-%d" % i))
        out.append((0x9400, fuzz94(rng, nframes, type1=i & 1, max_code=6, escape_p=0.4), "94-huff-%d" % i))
        out.append((0x9302, fuzz93(rng, nframes, type1=0, small=bool(i & 1)), "93b-t0-%d" % i))
        out.append((0x9302, fuzz93(rng, nframes, type1=1, small=bool(i & 1)), "93b-t1-%d" % i))
        out.append((0x9301, fuzz93(rng, nframes, type1=0, small=bool(i & 1)), "93a-t0-%d" % i))
        out.append((0x9301, fuzz93a1(rng, nframes), "93a-t1-%d" % i))
    return out
