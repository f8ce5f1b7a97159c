/* TEST INFRASTRUCTURE ONLY -- see dcs_oracle.h.  Plain-C restatement of the reference
 * decode path (DCSDecoder/DCSDecoderNative.cpp).  Structured the way the CUDA path is:
 * scan (frame checkpoints) -> per-frame decode from a checkpoint -> transform/overlap,
 * so the decomposition itself is validated against the sequential reference. */
#include <string.h>
#include "dcs_oracle.h"
#include "dcs_tables.h"

#define NCODES(t) ((int)(sizeof(t) / sizeof((t)[0])))

/* ---- bit reader: MSB-first, position-addressed (ROMBitPointer, DCSDecoderNative.h:229-289).
 * The reference's look-ahead never changes a returned value, so a pure function of
 * (bytes, bit position) is equivalent. */
static uint32_t peek_bits(const uint8_t *d, uint32_t pos, int n)
{
    uint64_t w = 0;
    const uint8_t *p = d + (pos >> 3);
    for (int i = 0; i < 5; ++i) w = (w << 8) | p[i];
    return (uint32_t)((w >> (40 - (pos & 7) - n)) & ((n >= 32) ? 0xFFFFFFFFu : ((1u << n) - 1)));
}
static uint32_t get_bits(const uint8_t *d, uint32_t *pos, int n)
{
    uint32_t v = peek_bits(d, *pos, n);
    *pos += n;
    return v;
}
static int32_t get_signed(const uint8_t *d, uint32_t *pos, int n)
{
    uint32_t v = get_bits(d, pos, n);
    if (v & (1u << (n - 1))) v |= ~0u << n;
    return (int32_t)v;
}
/* prefix-code decode by exhaustive match (codes are at most 30 bits) */
static int get_code(const dcs_code_t *t, int n, const uint8_t *d, uint32_t *pos)
{
    uint32_t code = 0;
    for (int len = 1; len <= 30; ++len) {
        code = (code << 1) | get_bits(d, pos, 1);
        for (int i = 0; i < n; ++i)
            if (t[i].len == len && t[i].code == code) return t[i].val;
    }
    return -1;
}

static inline int16_t s16(uint32_t v) { return (int16_t)(uint16_t)v; }
static inline int32_t sat16(int32_t v) { return v < -32768 ? -32768 : v > 32767 ? 32767 : v; }

/* ---- ADSP-2105 MAC rounding, in 32-bit wrap-around form (DCSDecoderNative.cpp:3503-3554).
 * Only bits 16..31 of MR are ever consumed, so arithmetic mod 2^32 is exact.
 * r = MR1( 2ab -/+ 2cd, rounded; "if low word of the second product == 0x8000 clear bit 16") */
static inline uint16_t mac_round(int32_t a, int32_t b, int32_t c, int32_t d, int sub)
{
    uint32_t p1 = (uint32_t)(a * b) << 1;
    uint32_t p2 = (uint32_t)(c * d) << 1;
    uint32_t r = (sub ? p1 - p2 : p1 + p2) + 0x8000u;
    if ((p2 & 0xFFFFu) == 0x8000u) r &= ~0x10000u;
    return (uint16_t)(r >> 16);
}
/* MultiplyAndRound(a,b) (:3526-3538) */
static inline uint16_t mul_round(int32_t a, int32_t b) { return mac_round(0, 0, a, b, 0); }

int dcso_calc_exp32(uint32_t x)       /* :3447-3459 */
{
    int res = 0;
    if (x & 0x80000000u) { for (; x & 0x40000000u; --res, x <<= 1) ; }
    else { for (; res > -31 && !(x & 0x40000000u); --res, x <<= 1) ; }
    return res;
}

/* ---- stream preamble (InitChannelStream :1433-1463) */
int dcso_header_len(const uint8_t *s, int os)
{
    return (os == DCSO_OS93A && (s[2] & 0x80)) ? 1 : 16;
}

static const uint16_t kMant[4] = { 0x8000, 0x9838, 0xb505, 0xd745 };   /* :1978, :2342 */
static const uint8_t kBandLen94[16] = { 7, 8, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 32 }; /* :1848 */
static const uint8_t kMaxW94[7] = { 0, 2, 3, 5, 7, 8, 9 };             /* :2005 */
static const uint8_t kPreAdj0[16] = { 0,0,0,0,1,1,1,1,1,1,1,1,1,1,1,1 };   /* :1744 */
static const uint8_t kPreAdj3[16] = { 0,0,0,0,1,2,3,4,4,4,4,4,4,4,4,4 };   /* :1747 */

static const dcs_code_t *cb94(int k, int *n)
{
    switch (k) {
    case 1: *n = NCODES(dcs94_cb1); return dcs94_cb1;
    case 2: *n = NCODES(dcs94_cb2); return dcs94_cb2;
    case 3: *n = NCODES(dcs94_cb3); return dcs94_cb3;
    case 4: *n = NCODES(dcs94_cb4); return dcs94_cb4;
    case 5: *n = NCODES(dcs94_cb5); return dcs94_cb5;
    default: *n = NCODES(dcs94_cb6); return dcs94_cb6;
    }
}

/* add one dequantised sample into a bin: the 16x16 scale product is truncated to its
 * low word, then (low word as unsigned) + (low word as signed)*mult is added at 16.16
 * (:2244-2250, :2434-2443) */
static inline void add_bin(uint16_t *fb, int idx, int32_t sample, uint16_t scale, uint16_t mult)
{
    if (idx < 0 || idx >= 512) return;
    uint16_t ss = (uint16_t)((uint32_t)sample * scale);
    int32_t c = ((int32_t)ss + (int32_t)s16(ss) * (int32_t)mult) >> 16;
    fb[idx] = (uint16_t)(fb[idx] + c);
}
/* "first sample moves from bin 1 to bin 0" (:2254-2257, :2608-2611); the caller passes the
 * value bin 1 held before this channel was added */
static void fix_bin01(uint16_t *fb, uint16_t old1)
{
    int32_t delta = sat16((int32_t)s16(fb[1]) - (int32_t)s16(old1));
    fb[0] = (uint16_t)sat16(delta + (int32_t)s16(fb[0]));
    fb[1] = old1;
}

/* ================= 1994+ frame (DecoderImpl94x::DecompressFrame :1679-2261) ==========
 * One walker serves both the scan (fb == NULL: lengths only) and the decode.
 * Returns 0, or DCSO_E_BANDTYPE.  bt[] is updated to the post-header state. */
static int walk94(const uint8_t *s, uint32_t *ppos, uint8_t bt[16], uint16_t mult, uint16_t *fb, int *stop)
{
    const uint8_t *hdr = s + 2, *d = s + 18;
    uint32_t pos = *ppos;
    int type1 = hdr[0] >> 7;                                         /* :1707 */
    int sub = ((hdr[1] & 0x80) >> 6) | ((hdr[2] & 0x80) >> 7);       /* :1712 */
    const uint8_t *pre = sub == 0 ? kPreAdj0 : kPreAdj3;             /* :1750 */
    int preadj[3];
    for (int i = 0; i < 3; ++i) preadj[i] = pre[bt[i]];              /* from the PRIOR frame, :1771-1773 */

    /* frame header: one delta per populated band (:1780-1834) */
    for (int i = 0; i < 16 && (hdr[i] & 0x7F) != 0x7F; ++i) {
        int v = get_code(dcs94_hdr, NCODES(dcs94_hdr), d, &pos);
        int nbt = bt[i] + (v - 0x2E);
        if (v < 0 || nbt < 0 || nbt > 15) { *ppos = pos; return DCSO_E_BANDTYPE; }
        bt[i] = (uint8_t)nbt;
    }

    uint16_t old1 = fb ? fb[1] : 0;
    int idx = 1, valid = 1;
    for (int b = 0; b < 16; ++b) {
        int hb = hdr[b] & 0x7F;
        if (hb == 0x7F) break;
        int count = kBandLen94[b], inc = 1;
        if (hb & 0x40) { inc = 2; count /= 2; }                      /* :1858-1862 */
        int code = bt[b];
        if (code == 0) { idx += count; continue; }                   /* :1878-1887 (sic: not count*inc) */
        int sc = hb;
        if (type1) {                                                 /* :1907-1961 */
            const uint8_t (*x)[2] = b < 3 ? dcs94_xlat_lo : b < 6 ? dcs94_xlat_mid : dcs94_xlat_hi;
            if (b < 3) hb += preadj[b];
            sc = hb + x[code][1];
            code = x[code][0];
        }
        uint16_t scale = (uint16_t)(kMant[sc & 3] >> (15 - ((sc >> 2) & 15)));   /* :1978-1979 */
        int32_t smp[32];
        int n = 0;
        if (code <= 6) {                                             /* :1992-2226 */
            int nc; const dcs_code_t *cb = cb94(code, &nc);
            int ref = 1 << (code - 1);
            for (int i = count; i != 0; --i) {
                int v = get_code(cb, nc, d, &pos);
                if (v & 0x80) {
                    if (i >= 2) { smp[n++] = 0; smp[n++] = 0; --i; }
                    else { valid = 0; if (stop) *stop = 1; i = 1; }
                } else smp[n++] = v - ref;
            }
        } else {                                                     /* :2227-2234 */
            for (int i = 0; i < count; ++i) smp[n++] = (int16_t)get_signed(d, &pos, code);
        }
        (void)kMaxW94;
        if (fb) {
            for (int i = 0; i < count; ++i, idx += inc)
                add_bin(fb, idx, (valid && i < n) ? smp[i] : 0, scale, mult);   /* :2238-2250 */
        } else idx += count * inc;
    }
    if (fb) fix_bin01(fb, old1);
    *ppos = pos;
    return 0;
}

/* ================= 1993 frame (DecoderImpl93::DecompressFrame :2293-2615) =========== */
static int walk93(const uint8_t *s, uint32_t *ppos, uint8_t bt[16], uint16_t mult, uint16_t *fb)
{
    const uint8_t *hdr = s + 2, *d = s + 18;
    uint32_t pos = *ppos;
    int type1 = hdr[0] >> 7;
    int subtype = type1 ? 0 : 2;                                     /* :2309 */
    uint16_t prv = 0, prvd = 0;
    int reuse = 0, code = 0, first = 1;
    uint16_t dummy[512];
    if (!fb) { memset(dummy, 0, sizeof(dummy)); fb = dummy; mult = 0; }
    uint16_t old1 = fb[1];
    int idx = 1;
    for (int band = 0; band < 16; ++band) {
        int hb = hdr[band] & 0x7F;
        if (hb == 0x7F) break;
        uint16_t scale = (uint16_t)(kMant[hb & 3] >> (15 - ((hb >> 2) & 15)));   /* :2337-2343 */
        int stridecode = hb >> 6;
        int n, inc, fixup, stride;
        if (!type1) {                                                /* :2351-2368 */
            if (!stridecode) { n = 16; inc = 1; fixup = 0; stride = 16; }
            else { ++idx; n = 16; inc = 2; fixup = -1; stride = 31; }
        } else {                                                     /* :2369-2383 */
            if (!stridecode) { inc = 1; fixup = 0; n = stride = first ? 15 : 16; }
            else { inc = 2; fixup = 0; n = stride = 8; }
        }
        if (reuse) reuse = (int)get_bits(d, &pos, 1);                /* :2388-2389 */
        if (!reuse) {
            if (!type1) {                                            /* :2396-2419 */
                if (get_bits(d, &pos, 1)) {
                    static const uint8_t dec[3] = { 2, 0, 1 }, incr[3] = { 1, 2, 0 };
                    subtype = get_bits(d, &pos, 1) ? incr[subtype] : dec[subtype];
                }
                code = (int)get_bits(d, &pos, 4);
            } else {                                                 /* :2420-2430, ReadHuff93 :2618-2684 */
                int v = get_code(dcs93_hdr, NCODES(dcs93_hdr), d, &pos);
                int delta;
                if (v < 0x1E) delta = v - 0x0F;
                else { delta = v - 0x2E; subtype = subtype ? 0 : 1; }
                int nbt = bt[band] + delta;
                if (v < 0 || nbt < 0 || nbt > 15) { *ppos = pos; return DCSO_E_BANDTYPE; }
                bt[band] = (uint8_t)nbt;
                code = nbt;
            }
        }
        if (code == 0) {                                             /* :2446-2547 */
            reuse = 1;
            if (subtype == 0) { idx += stride; prv = 0; prvd = 0; }
            else if (subtype == 1) {
                /* the low word of the running product carries from sample to sample (:2513-2535) */
                int64_t prod = (int64_t)s16(prv) * scale;
                int16_t plow = (int16_t)(prod & 0xFFFF);
                for (int i = 0; i < n; ++i) {
                    if (idx >= 0 && idx < 512) {
                        prod = (prod & 0xFFFF) | ((int64_t)s16(fb[idx]) << 16);
                        prod += (int64_t)plow * mult;
                        fb[idx] = (uint16_t)((prod >> 16) & 0xFFFF);
                    }
                    idx += inc;
                }
                prvd = 0;
                idx += fixup;
            } else {
                for (int i = 0; i < n; ++i) { prv = (uint16_t)(prv + prvd); add_bin(fb, idx, s16(prv), scale, mult); idx += inc; }
                idx += fixup;
            }
        } else {                                                     /* :2548-2603 */
            int width = code + (type1 ? 0 : 1);
            uint16_t in[16];
            for (int i = 0; i < n; ++i) in[i] = (uint16_t)get_signed(d, &pos, width);
            if (subtype == 0) {
                for (int i = 0; i < n; ++i) { add_bin(fb, idx, s16(in[i]), scale, mult); idx += inc; }
                prv = in[n - 1];
                prvd = (uint16_t)(prv - in[n - 2]);
            } else if (subtype == 1) {
                for (int i = 0; i < n; ++i) { prvd = in[i]; prv = (uint16_t)(prv + prvd); add_bin(fb, idx, s16(prv), scale, mult); idx += inc; }
            } else {
                for (int i = 0; i < n; ++i) { prvd = (uint16_t)(prvd + in[i]); prv = (uint16_t)(prv + prvd); add_bin(fb, idx, s16(prv), scale, mult); idx += inc; }
            }
            idx += fixup;
        }
        first = 0;
    }
    fix_bin01(fb, old1);
    *ppos = pos;
    return 0;
}

/* ================= OS93a type-1 frame (DecoderImpl93a::DecompressFrame :2831-3032) == */
static int walk93a1(const uint8_t *s, uint32_t *ppos, uint16_t mult, uint16_t *fb)
{
    const uint8_t *d = s + 3;
    uint32_t pos = *ppos;
    int hb = s[2];
    int sel = (hb & 0x60) >> 5, nbands = hb & 0x1F;
    static const dcs_code_t *const bbt[4] = { dcs93a_bandbits0, dcs93a_bandbits1, dcs93a_bandbits2, dcs93a_bandbits3 };
    static const int bbn[4] = { NCODES(dcs93a_bandbits0), NCODES(dcs93a_bandbits1), NCODES(dcs93a_bandbits2), NCODES(dcs93a_bandbits3) };
    int prvscale = 0x1A, idx = 0;
    for (int b = 0; b < nbands && b < 18; ++b) {
        int ninputs = dcs93a_inputs_per_band[b];
        int bits = get_code(bbt[sel], bbn[sel], d, &pos);
        if (bits == 0xFF) break;                                     /* :2922 */
        if (bits == 0) { idx += ninputs * 2; continue; }
        int v = get_code(dcs93a_scale, NCODES(dcs93a_scale), d, &pos);   /* :2938-2970 */
        int sc = prvscale + v - 1 + bits * 2;                        /* :2975-2981 */
        if (sc > 0x39) sc -= 0x36;
        prvscale = sc - bits * 2;
        uint32_t sf = 0x8000;
        for (int i = 0; i < (sc & 3); ++i) sf = (sf * 0x9838u) >> 15;    /* :2986-2991 */
        sf <<= ((sc >> 2) & 31);   /* x86 shift-count masking = the reference as compiled (count >= 32 only on malformed streams) */
        sf = ((sf >> 16) * mult) >> 15;                              /* :2995 */
        const uint16_t *base = &dcs93a_pairs[2 << bits];
        for (int i = 0; i < ninputs; ++i) {
            uint32_t smp = get_bits(d, &pos, bits);
            for (int k = 0; k < 2; ++k, ++idx) {
                if (!fb || idx >= 512) continue;
                /* MultiplyRoundAdd with MR = bin<<16 (:3010-3015, :3540-3546) */
                uint32_t p = (uint32_t)((int32_t)s16(base[smp * 2 + k]) * (int32_t)s16(sf)) << 1;
                uint32_t r = ((uint32_t)fb[idx] << 16) + p + 0x8000u;
                if ((p & 0xFFFFu) == 0x8000u) r &= ~0x10000u;
                fb[idx] = (uint16_t)(r >> 16);
            }
        }
    }
    *ppos = pos;
    return 0;
}

static int is93a1(const uint8_t *s, int os) { return os == DCSO_OS93A && (s[2] & 0x80); }

static int walk(const uint8_t *s, int os, uint32_t *pos, uint8_t bt[16], uint16_t mult, uint16_t *fb, int *stop)
{
    if (os == DCSO_OS93A || os == DCSO_OS93B)
        return is93a1(s, os) ? walk93a1(s, pos, mult, fb) : walk93(s, pos, bt, mult, fb);
    return walk94(s, pos, bt, mult, fb, stop);
}

int dcso_scan(const uint8_t *s, size_t nbytes, int os, dcso_frame_t *fr, int *stop_frame)
{
    if (stop_frame) *stop_frame = -1;
    if (nbytes < 3) return DCSO_E_SHORT;
    int nf = (s[0] << 8) | s[1];
    if (nf == 0) return DCSO_E_EMPTY;
    size_t hl = (size_t)dcso_header_len(s, os);
    if (nbytes < 2 + hl) return DCSO_E_SHORT;
    uint64_t nbits = (uint64_t)(nbytes - 2 - hl) * 8;
    uint32_t pos = 0;
    uint8_t bt[16];
    memset(bt, 0, sizeof(bt));                                       /* InitStreamPlayback :1640 */
    for (int f = 0; f < nf; ++f) {
        fr[f].bitpos = pos;
        memcpy(fr[f].bt, bt, 16);
        int stop = 0;
        int rc = walk(s, os, &pos, bt, 0, NULL, &stop);
        if (rc == 0 && pos > nbits) rc = DCSO_E_TRUNCATED;
        if (rc) { if (stop_frame) *stop_frame = f; return rc; }
        if (stop && stop_frame && *stop_frame < 0) { *stop_frame = f; }
        if (stop) {   /* the reference clears the channel on the next main loop (:95-116) */
            for (int g = f + 1; g <= nf; ++g) { fr[g].bitpos = pos; memcpy(fr[g].bt, bt, 16); }
            return nf;
        }
    }
    fr[nf].bitpos = pos;
    memcpy(fr[nf].bt, bt, 16);
    return nf;
}

void dcso_decode_frame(const uint8_t *s, int os, const dcso_frame_t *f, uint16_t mult, uint16_t *fb, int *stop)
{
    uint32_t pos = f->bitpos;
    uint8_t bt[16];
    memcpy(bt, f->bt, 16);
    int st = 0;
    walk(s, os, &pos, bt, mult, fb, &st);
    if (stop) *stop = st;
}

/* ================= transforms ======================================================= */
static inline int rev7(int x) { int r = 0; for (int i = 0; i < 7; ++i) r |= ((x >> i) & 1) << (6 - i); return r; }
static inline int rev9(int x) { int r = 0; for (int i = 0; i < 9; ++i) r |= ((x >> i) & 1) << (8 - i); return r; }

/* overlap-add of the first 16 samples (:538-555, :787-802): MulSU both terms, round */
static inline uint16_t overlap_mix(uint16_t cur, uint16_t prev, int i)
{
    uint32_t a = (uint32_t)((int32_t)s16(cur) * (int32_t)dcs_overlap_win[i]) << 1;
    uint32_t b = (uint32_t)((int32_t)s16(prev) * (int32_t)dcs_overlap_win[15 - i]) << 1;
    return (uint16_t)((a + b + 0x8000u) >> 16);
}

static void transform94(uint16_t *fb, uint16_t *ovl, int vs, int16_t *pcm)
{
    const uint16_t *sn = dcs_twiddle, *cs = dcs_twiddle + 128;
    /* bins -> complex pairs; MulSS(x,0x8000) == wrap(-x) (:403-418) */
    fb[0x100] = fb[0x101] = 0;   /* the frame buffer is 512 long and zero above 255 (:92) */
    fb[0x80] = (uint16_t)(-(int32_t)s16(fb[0x80]));
    fb[0x81] = (uint16_t)(-(int32_t)s16((uint16_t)(-(int32_t)s16(fb[0x81]))));
    for (int i = 0; i < 64; ++i) {
        uint16_t *p0 = fb + 2 * i, *p1 = fb + 0x100 - 2 * i;
        int32_t x0 = s16(p0[0]), y0 = s16(p1[0]), x1 = s16(p0[1]), y1 = s16(p1[1]);
        p0[0] = (uint16_t)(-sat16(x0 + y0));
        p1[0] = (uint16_t)(-sat16(x0 - y0));
        p0[1] = (uint16_t)(-sat16(x1 - y1));
        p1[1] = (uint16_t)(-sat16(x1 + y1));
    }
    /* twiddle pass (:420-456): coefficient index bitRev9[2+4i] / bitRev9[4i] */
    for (int i = 0; i < 64; ++i) {
        uint16_t *p4 = fb + 2 * i, *p5 = fb + 0x100 - 2 * i;
        int32_t c0 = s16(dcs_twiddle[rev9(2 + 4 * i)]), c1 = s16(dcs_twiddle[rev9(4 * i)]);
        int32_t x0 = s16(p4[0]), x1 = s16(p4[1]), xn0 = s16(p5[0]), xn1 = s16(p5[1]);
        int32_t prod0 = s16(mac_round(xn1, c1, xn0, c0, 1));
        int32_t prod1 = s16(mac_round(xn1, c0, xn0, c1, 0));
        p4[0] = (uint16_t)sat16(prod1 + x0);
        p4[1] = (uint16_t)sat16(prod0 + x1);
        p5[0] = (uint16_t)sat16(x0 - prod1);
        p5[1] = (uint16_t)sat16(prod0 - x1);
    }
    /* half fold (:458-471) */
    for (int i = 0; i < 128; ++i) {
        int32_t x = s16(fb[i]), y = s16(fb[i + 0x80]);
        fb[i] = (uint16_t)sat16(x + y);
        fb[i + 0x80] = (uint16_t)sat16(x - y);
    }
    /* 6 radix-2 stages over two 64-point halves, saturating (:480-524) */
    for (int st = 0, np = 2, ps = 0x40; st < 6; ++st, np *= 2, ps /= 2) {
        for (int p = 0; p < np; ++p) {
            int32_t sv = s16(sn[p]), cv = s16(cs[p]);
            uint16_t *p0 = fb + p * 2 * ps, *p1 = p0 + ps;
            for (int j = 0; j < ps / 2; ++j, p0 += 2, p1 += 2) {
                int32_t ar = s16(p1[0]), ai = s16(p1[1]);
                int32_t tr = s16(mac_round(ar, cv, ai, sv, 1));
                int32_t ti = s16(mac_round(ai, cv, ar, sv, 0));
                int32_t ur = s16(p0[0]), ui = s16(p0[1]);
                p0[0] = (uint16_t)sat16(ur - tr); p0[1] = (uint16_t)sat16(ui - ti);
                p1[0] = (uint16_t)sat16(ur + tr); p1[1] = (uint16_t)sat16(ui + ti);
            }
        }
    }
    /* volume normalisation, overlap, bit-reversed gather (:532-575):
     * time samples (2m, 2m+1) are the (re, im) of complex element bitrev7(m) */
    for (int i = 0; i < 256; ++i) fb[i] = (uint16_t)((int32_t)s16(fb[i]) >> vs);
    for (int n = 0; n < 256; ++n) {
        uint16_t v = fb[2 * rev7(n >> 1) + (n & 1)];
        if (n < 16) v = overlap_mix(v, ovl[n], n);
        if (n < 240) pcm[n] = (int16_t)v;
    }
    for (int n = 240; n < 256; ++n) ovl[n - 240] = fb[2 * rev7(n >> 1) + (n & 1)];
}

static void transform93(uint16_t *fb, uint16_t *ovl, int vs, int16_t *pcm)
{
    const uint16_t *sn = dcs_twiddle, *cs = dcs_twiddle + 128;
    /* |bin0 + i*bin1| by the 1.15 Taylor series (:633-710); 40-bit MR kept in int64 */
    uint16_t AR = fb[0];
    int neg = s16(AR) < 0;
    if (neg) AR = (uint16_t)(-(int32_t)s16(AR));
    int64_t MR = (((int64_t)s16(fb[1]) * s16(fb[1])) << 1) + (((int64_t)s16(AR) * s16(AR)) << 1);
    uint32_t SR = (uint32_t)(MR & 0xFFFFFFFF);
    int exponent = dcso_calc_exp32(SR);
    if (exponent < 0) SR <<= -exponent;
    AR = (uint16_t)(SR >> 16);
    if (AR != 0) {
        static const int32_t k[5] = { 0x5D1D, -22035, 0x46D6, -8790, 0x072D };
        uint64_t mr = 0x0D490000u;
        uint16_t mf = AR;
        for (int t = 0; t < 5; ++t) {
            mr += (uint64_t)(((int64_t)k[t] * (int64_t)s16(mf)) << 1);
            if (t < 4) mf = mul_round(s16(AR), s16(mf));
        }
        if (exponent & 1) {
            /* MultiplyAndRound(MR, MR1(MR), 0x5A82) replaces MR by the rounded product (:3526-3531) */
            int32_t prod = (int32_t)((uint32_t)((int32_t)s16((uint16_t)(mr >> 16)) * 0x5A82) << 1);
            int64_t r = (int64_t)prod + 0x8000;
            if ((prod & 0xFFFF) == 0x8000) r &= ~0x10000LL;
            mr = (uint64_t)r;
            exponent += 1;
        }
        exponent = exponent / 2 + 1;
        int32_t v = (int32_t)(uint32_t)(mr & 0xFFFFFFFF);
        uint32_t sr;                                                 /* BitShiftSigned32 (:3486-3501) */
        if (exponent >= 0) sr = (uint32_t)v << exponent;
        else if (v >= 0) sr = (uint32_t)v >> -exponent;
        else sr = ((uint32_t)v >> -exponent) | (~0u << (32 + exponent));
        AR = (uint16_t)(sr >> 16);
        if (neg) AR = (uint16_t)(-(int32_t)s16(AR));
    }
    fb[0] = fb[0x100] = AR;
    fb[1] = fb[0x101] = 0;
    /* 256 -> 512 expansion, wrapping (:714-732) */
    for (int i = 0; i < 64; ++i) {
        uint16_t *i0 = fb + 2 + 2 * i, *i1 = fb + 0xFE - 2 * i, *i2 = fb + 0x102 + 2 * i, *i3 = fb + 0x1FE - 2 * i;
        int32_t xr = s16(i0[0]), xi = s16(i0[1]), yr = s16(i1[0]), yi = s16(i1[1]);
        i0[0] = i1[0] = (uint16_t)(xr + yr);
        i2[0] = (uint16_t)(xr - yr);
        i3[0] = (uint16_t)(yr - xr);
        i2[1] = i3[1] = (uint16_t)(xi + yi);
        i0[1] = (uint16_t)(xi - yi);
        i1[1] = (uint16_t)(yi - xi);
    }
    /* 7 radix-2 stages over 256 complex points, wrapping adds (:742-778) */
    for (int st = 0, np = 2, ps = 0x80; st < 7; ++st, np *= 2, ps /= 2) {
        for (int p = 0; p < np; ++p) {
            int32_t sv = s16(sn[p]), cv = s16(cs[p]);
            uint16_t *p0 = fb + p * 2 * ps, *p1 = p0 + ps;
            for (int j = 0; j < ps / 2; ++j, p0 += 2, p1 += 2) {
                int32_t a0 = s16(p1[0]), a1 = s16(p1[1]), y0 = s16(p0[0]), y1 = s16(p0[1]);
                int32_t x0 = s16(mac_round(a0, cv, a1, sv, 1));
                int32_t x1 = s16(mac_round(a1, cv, a0, sv, 0));
                p0[0] = (uint16_t)(y0 - x0); p0[1] = (uint16_t)(y1 - x1);
                p1[0] = (uint16_t)(x0 + y0); p1[1] = (uint16_t)(x1 + y1);
            }
        }
    }
    /* sample n is the REAL part at bit-reversed index (:782-812) */
    for (int n = 0; n < 256; ++n) {
        uint16_t v = (uint16_t)((int32_t)s16(fb[rev9(n)]) >> vs);
        if (n < 16) pcm[n] = (int16_t)overlap_mix(v, ovl[n], n);
        else if (n < 240) pcm[n] = (int16_t)v;
        else ovl[n - 240] = v;
    }
}

void dcso_transform(int os, uint16_t *fb, uint16_t *ovl, int vs, int16_t *pcm)
{
    if (os == DCSO_OS93A || os == DCSO_OS93B) transform93(fb, ovl, vs, pcm);
    else transform94(fb, ovl, vs, pcm);
}

/* ================= gain ============================================================= */
uint16_t dcso_master_multiplier(int vol)      /* SetMasterVolume :3250-3282 */
{
    if (vol > 255) vol = 255;
    if (vol <= 0) return 0;   /* note: the reference tests the unclamped value against 0 */
    uint32_t s = (uint32_t)vol, x = 0x3fff, y = 0x7d98;
    for (int i = 0; i < 8; ++i) {
        if (!(s & 1)) x = ((x * y) >> 15) & 0xFFFF;
        y = ((y * y) >> 15) & 0xFFFF;
        s >>= 1;
    }
    return (uint16_t)(x << 1);
}

uint16_t dcso_level_multiplier(int sum, int os, int chan_vol, int max_override)   /* :3071-3121 */
{
    if (sum > 8191) sum = 8191; else if (sum < -8191) sum = -8191;
    uint16_t e = (uint16_t)(((sum >> 6) & 0x3FF) + 0x80);
    uint16_t m = (os == DCSO_OS93A) ? 0x7FFF : (uint16_t)(chan_vol << 7);
    if (max_override) m = 0xFF << 7;
    uint16_t prod = 0x7C94;
    for (int j = 0, bit = 1; j < 8; ++j, bit <<= 1) {
        if (!(e & bit)) m = (uint16_t)(((uint32_t)m * prod) >> 15);
        prod = (uint16_t)(((uint32_t)prod * prod) >> 15);
    }
    return (uint16_t)(m << 1);
}

int dcso_gain_stage(const uint16_t mix[8], unsigned active, unsigned maxovr, uint16_t vol, uint16_t eff[8])   /* :227-269 */
{
    uint64_t sum = 0;
    for (int i = 0; i < 8; ++i) {
        if (maxovr & (1u << i)) sum += (uint64_t)mix[i] * 0x7FFE;
        else if (active & (1u << i)) sum += (uint64_t)mix[i] * vol;
    }
    sum >>= 2;
    int vs = -(dcso_calc_exp32((uint32_t)sum) + 3);
    vs = vs < 0 ? 0 : vs > 8 ? 8 : vs;
    for (int i = 0; i < 8; ++i) {
        uint16_t v = (maxovr & (1u << i)) ? 0x7FFE : vol;
        uint64_t m = ((uint64_t)mix[i] * v) << 1;
        eff[i] = (uint16_t)((m << vs) >> 16);
    }
    return vs;
}

/* ================= whole-stream protocol ===================== */
int dcso_decode_stream(const uint8_t *s, size_t nbytes, int os, int vol, int level, int nout, int16_t *pcm)
{
    static dcso_frame_t fr[65537];
    int stopf = -1;
    int nf = dcso_scan(s, nbytes, os, fr, &stopf);
    int status = nf < 0 ? nf : DCSO_OK;
    int nplay;            /* frames actually decoded before the channel goes silent */
    if (nf == DCSO_E_EMPTY || nf == DCSO_E_SHORT) nplay = 0;
    else if (nf < 0) nplay = stopf;              /* frames before the undecodable one */
    else nplay = stopf >= 0 ? stopf + 1 : nf;    /* the stop frame itself is still output */
    uint16_t volm = dcso_master_multiplier(vol);
    uint16_t mix[8], eff[8], ovl[16];
    for (int i = 0; i < 8; ++i) mix[i] = 0x7FFF;                     /* DCSDecoderNative.h:514 */
    memset(ovl, 0, sizeof(ovl));
    uint16_t lvl = dcso_level_multiplier(level << 6, os, 0xFF, 0);
    uint16_t idle = dcso_level_multiplier(0, os, 0xFF, 0);
    for (int f = 0; f < nout; ++f) {
        uint16_t fb[512];
        memset(fb, 0, sizeof(fb));
        int active = f < nplay;
        int vs = dcso_gain_stage(mix, active ? 1u : 0u, 0, volm, eff);
        if (active) dcso_decode_frame(s, os, &fr[f], eff[0], fb, NULL);
        dcso_transform(os, fb, ovl, vs, pcm + (size_t)f * 240);
        /* UpdateMixingLevels recomputes every channel from its level sum (:3071-3121) */
        mix[0] = lvl;
        for (int i = 1; i < 8; ++i) mix[i] = idle;
    }
    return status;
}

uint64_t dcso_fnv1a(const int16_t *pcm, size_t n)
{
    uint64_t h = 1469598103934665603ULL;
    const uint8_t *p = (const uint8_t *)pcm;
    for (size_t i = 0; i < n * 2; ++i) { h ^= p[i]; h *= 1099511628211ULL; }
    return h;
}
