// dcsb200 fast path for the 1994+ frame layout (the format BASELINE.json's metric is quoted on).
//
//   dcsb_lane_decode94     K2 phase A: one LANE decodes one frame from its checkpoint into
//                          16-bit frequency bins in its own shared-memory row.
//   dcsb_lane_transform94  K2 phase B: the same lane runs the exact fixed-point inverse
//                          transform on its row (registers + lane-private shared memory,
//                          no warp synchronisation).
//   dcsb_tile_output94     K2 phase C: the warp applies the 16-sample overlap-add and
//                          writes PCM with coalesced stores (+ checksum).
//
// Row layout (per lane, 129 words so that equal offsets in different rows hit different
// banks): two planes of 128 int16 -- real parts at halfwords [0,128), imaginary parts at
// [128,256) -- with complex element e of the reference's frameBuffer (bins 2e, 2e+1) stored at
// position bitrev7(e).  Bin idx therefore lives at halfword  __brev(idx) >> 24.  The
// transform's last pass then leaves sample pair m (samples 2m, 2m+1) at position m, i.e. the
// bit-reversed gather of DCSDecoderNative.cpp:559-565 costs nothing.
//
// Compiled by nvcc for sm_100a (the product) and by g++ for the CPU-side kernel simulator
// (tests/hostsim, test infrastructure only).
#pragma once
#include "dcsb_core.cuh"

// ---------------------------------------------------------------------------------------
// Register bit window: MSB-first reader over big-endian 32-bit words (ROMBitPointer,
// DCSDecoderNative.h:229-289).  w0 holds the current bit at offset s, w1 the next word,
// nx is prefetched one word ahead and kept in memory byte order: it is only byte-swapped
// when it moves into w1, so a refill never waits for the load it has just issued.
struct DcsbWin {
    const uint32_t *base;   // 4-byte aligned word containing bit 0 of the stream data
    const uint32_t *p;      // next word to prefetch
    uint32_t w0, w1, nx;
    uint32_t s;             // 0..31
    uint32_t bias;          // bit offset of the stream's first data bit inside base[0]

    DCSB_HD void seek(uint32_t pos)
    {
        const uint32_t a = pos + bias;
        const uint32_t *q = base + (a >> 5);
        s = a & 31;
        w0 = DcsbBits::be(DCSB_LDG(q));
        w1 = DcsbBits::be(DCSB_LDG(q + 1));
        nx = DCSB_LDG(q + 2);
        p = q + 3;
    }
    DCSB_HD uint32_t pos() const { return (uint32_t)(p - base - 3) * 32u + s - bias; }
    // next 32 bits, left aligned
    DCSB_HD uint32_t peek32() const
    {
#if DCSB_DEVICE_PASS
        return __funnelshift_l(w1, w0, s);
#else
        return s ? ((w0 << s) | (w1 >> (32 - s))) : w0;
#endif
    }
    // consume n <= 32 bits
    DCSB_HD void skip(uint32_t n)
    {
        s += n;
        refill();
    }
    // Two short reads per refill: peek_wide() stays valid while s + bits <= 64, so a loop may
    // advance() twice (each by <= 15 bits) and then drop at most one word.
    DCSB_HD uint32_t peek_wide() const { return (uint32_t)((((((uint64_t)w0) << 32) | w1) << s) >> 32); }
    DCSB_HD void advance(uint32_t n) { s += n; }
    DCSB_HD void refill()
    {
        if (s >= 32) {
            s -= 32;
            w0 = w1;
            w1 = DcsbBits::be(nx);
            nx = DCSB_LDG(p);
            ++p;
        }
    }
};

DCSB_HD DcsbWin dcsb_make_window(const uint8_t *slab, const DcsbStreamRec &s, uint32_t pos)
{
    const uint64_t start = s.data_off + 2 + s.hdr_len;
    DcsbWin w;
    w.base = reinterpret_cast<const uint32_t *>(slab + (start & ~3ull));
    w.bias = (uint32_t)(start & 3) * 8;
    w.seek(pos);
    return w;
}

DCSB_HD int dcsb_nib(uint64_t bt, int i) { return (int)((bt >> (4 * i)) & 15); }
DCSB_HD int dcsb_clz(uint32_t v)
{
#if DCSB_DEVICE_PASS
    return __clz((int)v);
#else
    int n = 0;
    while (n < 32 && !(v & (0x80000000u >> n))) ++n;
    return n;
#endif
}
DCSB_HD uint32_t dcsb_brev(uint32_t v)
{
#if DCSB_DEVICE_PASS
    return __brev(v);
#else
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
    return r;
#endif
}

// 1994 band geometry (DCSDecoderNative.cpp:1848-1862)
DCSB_HD int dcsb_band_count94(int b) { return b == 0 ? 7 : (b == 1 ? 8 : (b == 15 ? 32 : 16)); }
// sample codebook k: widest code / LUT offset inside the DCSB_LUT_CB block (:2005-2175)
DCSB_HD int dcsb_cb_maxw(int k) { return k <= 2 ? k + 1 : (k == 3 ? 5 : k + 3); }
DCSB_HD int dcsb_cb_ofs(int k) { return k == 1 ? 0 : k == 2 ? 4 : k == 3 ? 12 : k == 4 ? 44 : k == 5 ? 172 : 428; }

// (K1, the frame-boundary scan of the 1994 layout, lives in dcsb_scan94.cuh)

// =======================================================================================
// K2 phase A: decode the bands of one 1994 frame (DCSDecoderNative.cpp:1836-2257).
// `win` stands at the first band; btp / btc are the band types of the previous / this frame.
// ACCUM=false: the row is zero on entry and every bin is written at most once, so the
// contribution is stored; ACCUM=true adds modulo 2^16 (more channels mixed into one frame).
// dequantise one sample into its bin (:2244-2250); bin idx lives at halfword brev(idx) >> 24
template <bool ACCUM>
DCSB_HD void dcsb_store_bin94(int16_t *r16, int idx, int val, uint32_t scale, uint32_t mult)
{
    const uint32_t ss = ((uint32_t)val * scale) & 0xFFFFu;
    const int c = ((int)ss + dcsb_s16(ss) * (int)mult) >> 16;
    const uint32_t a = dcsb_brev((uint32_t)idx) >> 24;
    if (ACCUM) r16[a] = (int16_t)(r16[a] + c);
    else r16[a] = (int16_t)c;
}

template <bool ACCUM>
DCSB_HD void dcsb_lane_decode94(const uint8_t *hdr, const uint16_t *lut, DcsbWin &win, uint64_t btp, uint64_t btc,
                                uint32_t mult, int zero_from, int16_t *r16)
{
    const int type1 = hdr[0] >> 7;
    const int sub = ((hdr[1] & 0x80) >> 6) | ((hdr[2] & 0x80) >> 7);
    const int old1 = ACCUM ? (int)r16[128] : 0;      // bin 1 = imaginary part of element 0
    int idx = 1;
    for (int b = 0; b < 16; ++b) {
        int hb = hdr[b] & 0x7F;
        if (hb == 0x7F) break;
        int count = dcsb_band_count94(b);
        int inc = 1;
        if (hb & 0x40) { inc = 2; count >>= 1; }                         // :1858-1862
        int code = dcsb_nib(btc, b);
        if (code == 0) { idx += count; continue; }                       // :1878-1887
        int sc = hb;
        if (type1) {                                                     // :1907-1961
            const uint32_t x = lut[DCSB_LUT_XLAT + (b < 3 ? 0 : (b < 6 ? 16 : 32)) + code];
            if (b < 3) {
                const int t = dcsb_nib(btp, b);                          // :1744-1773
                hb += t < 4 ? 0 : (sub == 0 ? 1 : (t > 7 ? 4 : t - 3));
            }
            sc = hb + (int)(x & 0xFF);
            code = (int)(x >> 8);
        }
        const uint32_t scale = dcsb_scale_factor(sc);
        const bool add = b < zero_from;
        if (code <= 6) {                                                 // :1992-2226
            // sample codebook entry: bits 12..15 = code length, bit 11 = 'two zeros', low byte = signed value
            const uint16_t *cb = lut + DCSB_LUT_CB + dcsb_cb_ofs(code);
            const int sh = 32 - dcsb_cb_maxw(code);
            int rem = count;
            while (rem > 0) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (rem > 0) {
                        const uint32_t e = cb[win.peek_wide() >> sh];
                        win.advance(e >> 12);
                        const bool dz = (e & 0x800u) != 0;
                        if (!dz && add) dcsb_store_bin94<ACCUM>(r16, idx, (int)(int8_t)(e & 0xFF), scale, mult);
                        const int st = dz ? (rem < 2 ? rem : 2) : 1;     // :2199-2219
                        idx += st * inc;
                        rem -= st;
                    }
                }
                win.refill();
            }
        } else {                                                         // :2227-2234
            const int sh = 32 - code;
            for (int i = 0; i < count; i += 2) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (i + u < count) {
                        const int val = (int)win.peek_wide() >> sh;
                        win.advance((uint32_t)code);
                        if (add) dcsb_store_bin94<ACCUM>(r16, idx, val, scale, mult);
                        idx += inc;
                    }
                }
                win.refill();
            }
        }
    }
    // move this frame's bin-1 contribution to bin 0 with saturation (:2255-2257)
    const int delta = dcsb_sat16((int)r16[128] - old1);
    r16[0] = (int16_t)dcsb_sat16(delta + (int)r16[0]);
    r16[128] = (int16_t)old1;
}

// =======================================================================================
// K2 phase B: TransformFrame for the 1994 layout (DCSDecoderNative.cpp:397-576), one lane
// per frame.  Twiddles come pre-doubled (2*cos, 2*sin) so that a multiply-accumulate
// produces the ADSP's left-shifted MR directly.
//
// Between the pre-pass and the last store the elements are held BIASED (v + 0x8000, 0..0xFFFF): the saturating
// add of a biased and a plain value is max(min(x + t, 0xFFFF), 0), ONE instruction (VIADDMNMX.RELU) where the
// symmetric clamp takes two.  The multiply-accumulates take the biased operands as they are: the bias times a
// twiddle is a constant per twiddle, kept beside it (kr / ki) and added where the rounding constant is added
// anyway; 0x8000 * (2 cos) has a zero low word, so the tie rule still reads the low word of the second product.
struct DcsbTw4 { int c2, s2, kr, ki; };
struct DcsbTw94 {
    DcsbTw4 tw[64];                // butterfly twiddles, reference table order (partition index)
    int pre_c0[64], pre_c1[64];    // pre-pass coefficients, natural order i
};
#define DCSB_BIAS 0x8000
// fill element i of a (shared-memory) copy from the context's tables
DCSB_HD void dcsb_tw94_fill(DcsbTw94 *dst, const DcsbTables *tab, int i)
{
    const int c2 = tab->tw_c2[i], s2 = tab->tw_s2[i];
    dst->tw[i].c2 = c2;
    dst->tw[i].s2 = s2;
    dst->tw[i].kr = (int)(((uint32_t)(c2 - s2) << 15) - 0x8000u);       // tr: q = ai' s2 + kr, r = ar' c2 - q
    dst->tw[i].ki = (int)(0x8000u - ((uint32_t)(c2 + s2) << 15));       // ti: q = ar' s2 + ki, r = ai' c2 + q
    dst->pre_c0[i] = tab->pre_c0[i];
    dst->pre_c1[i] = tab->pre_c1[i];
}

// sext16( MR1( a*b2 -/+ c*d2 + rounding ) ), rounding tie rule on the low word of c*d2 (:3503-3554); see dcsb_mac_hi
template <bool SUB>
DCSB_HD int dcsb_mac2(int a, int b2, int c, int d2) { return (int)dcsb_mac_hi<SUB>(a, b2, c, d2) >> 16; }
// the same on biased a, c: k carries the rounding constant and the bias terms
template <bool SUB>
DCSB_HD int dcsb_mac2k(int a, int b2, int c, int d2, int k)
{
    const uint32_t q = (uint32_t)c * (uint32_t)d2 + (uint32_t)k;
    const uint32_t r = SUB ? (uint32_t)a * (uint32_t)b2 - q : (uint32_t)a * (uint32_t)b2 + q;
    return (int)(r & (0xFFFE0000u | (0u - (q & 0xFFFFu)))) >> 16;
}
// saturating add of a biased element and a plain value, biased
DCSB_HD int dcsb_addb(int xb, int t)
{
#if DCSB_DEVICE_PASS
    return __viaddmin_s32_relu(xb, t, 0xFFFF);          // max(min(xb + t, 0xFFFF), 0): VIADDMNMX.RELU
#else
    const int v = xb + t;
    return v > 0xFFFF ? 0xFFFF : (v < 0 ? 0 : v);
#endif
}
DCSB_HD int dcsb_negw(int v) { return dcsb_s16((uint32_t)-v); }      // MulSS(v, 0x8000): wrap16(-v)

// pre-pass on the pair (element i = x, element 128-i = y) (:405-456)
DCSB_HD void dcsb_prepair94(int &xr, int &xi, int &yr, int &yi, int c0x2, int c1x2)
{
    const int p0r = dcsb_negw(dcsb_sat16(xr + yr)), p1r = dcsb_negw(dcsb_sat16(xr - yr));
    const int p0i = dcsb_negw(dcsb_sat16(xi - yi)), p1i = dcsb_negw(dcsb_sat16(xi + yi));
    const int prod0 = dcsb_mac2<true>(p1i, c1x2, p1r, c0x2);
    const int prod1 = dcsb_mac2<false>(p1i, c0x2, p1r, c1x2);
    xr = dcsb_sat16(prod1 + p0r);
    xi = dcsb_sat16(prod0 + p0i);
    yr = dcsb_sat16(p0r - prod1);
    yi = dcsb_sat16(prod0 - p0i);
}
// half fold (:458-471); plain in, BIASED out
DCSB_HD void dcsb_fold94(int &ur, int &ui, int &ar, int &ai)
{
    const int ubr = ur + DCSB_BIAS, ubi = ui + DCSB_BIAS;
    const int sr = dcsb_addb(ubr, ar), si = dcsb_addb(ubi, ai);
    const int dr = dcsb_addb(ubr, -ar), di = dcsb_addb(ubi, -ai);
    ur = sr; ui = si; ar = dr; ai = di;
}
// radix-2 butterfly, saturating (:480-524): u' = u - t, a' = u + t, t = a * (cos + i sin); biased in and out
DCSB_HD void dcsb_bfly94(int &ur, int &ui, int &ar, int &ai, const DcsbTw4 &w)
{
    const int tr = dcsb_mac2k<true>(ar, w.c2, ai, w.s2, w.kr);
    const int ti = dcsb_mac2k<false>(ai, w.c2, ar, w.s2, w.ki);
    const int nur = dcsb_addb(ur, -tr), nui = dcsb_addb(ui, -ti);
    ar = dcsb_addb(ur, tr);
    ai = dcsb_addb(ui, ti);
    ur = nur;
    ui = nui;
}
// three radix-2 stages on 8 points held in registers; twiddle indices iA, iB..iB+1, iC..iC+3
DCSB_HD void dcsb_fft8_94(int *xr, int *xi, const DcsbTw94 *tw, int iA, int iB, int iC)
{
    {
        const DcsbTw4 w = tw->tw[iA];
#pragma unroll
        for (int k = 0; k < 4; ++k) dcsb_bfly94(xr[k], xi[k], xr[k + 4], xi[k + 4], w);
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const DcsbTw4 w = tw->tw[iB + q];
#pragma unroll
        for (int k = 0; k < 2; ++k) dcsb_bfly94(xr[4 * q + k], xi[4 * q + k], xr[4 * q + k + 2], xi[4 * q + k + 2], w);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const DcsbTw4 w = tw->tw[iC + q];
        dcsb_bfly94(xr[2 * q], xi[2 * q], xr[2 * q + 1], xi[2 * q + 1], w);
    }
}

DCSB_HD int dcsb_rev3(int x) { return ((x & 1) << 2) | (x & 2) | ((x >> 2) & 1); }

DCSB_HD void dcsb_lane_transform94(int16_t *r16, const DcsbTw94 *tw, int vs)
{
    int16_t *re = r16, *im = r16 + 128;
    // ---- pre-pass + half fold on the closed element sets {i, 128-i, 64-i, 64+i}
    {   // i = 0: element 128 is the always-zero phantom, element 64 has its real part negated (:403-404)
        int xr = re[0], xi = im[0], yr = 0, yi = 0;
        dcsb_prepair94(xr, xi, yr, yi, tw->pre_c0[0], tw->pre_c1[0]);
        const int a64 = 1;                                   // bitrev7(64)
        int mr = dcsb_negw((int)re[a64]), mi = im[a64];
        dcsb_fold94(xr, xi, mr, mi);
        re[0] = (int16_t)xr; im[0] = (int16_t)xi;
        re[a64] = (int16_t)mr; im[a64] = (int16_t)mi;
    }
    {   // i = 32: the set degenerates to the pair (32, 96)
        const int a32 = 2, a96 = 3;                          // bitrev7(32), bitrev7(96)
        int xr = re[a32], xi = im[a32], yr = re[a96], yi = im[a96];
        dcsb_prepair94(xr, xi, yr, yi, tw->pre_c0[32], tw->pre_c1[32]);
        dcsb_fold94(xr, xi, yr, yi);
        re[a32] = (int16_t)xr; im[a32] = (int16_t)xi;
        re[a96] = (int16_t)yr; im[a96] = (int16_t)yi;
    }
#pragma unroll 1
    for (int i = 1; i < 32; ++i) {
        const int aA = (int)(dcsb_brev((uint32_t)i) >> 25), aB = (int)(dcsb_brev((uint32_t)(128 - i)) >> 25);
        const int aC = (int)(dcsb_brev((uint32_t)(64 - i)) >> 25), aD = (int)(dcsb_brev((uint32_t)(64 + i)) >> 25);
        int Ar = re[aA], Ai = im[aA], Br = re[aB], Bi = im[aB];
        int Cr = re[aC], Ci = im[aC], Dr = re[aD], Di = im[aD];
        dcsb_prepair94(Ar, Ai, Br, Bi, tw->pre_c0[i], tw->pre_c1[i]);
        dcsb_prepair94(Cr, Ci, Dr, Di, tw->pre_c0[64 - i], tw->pre_c1[64 - i]);
        dcsb_fold94(Ar, Ai, Dr, Di);                         // (i, i + 64)
        dcsb_fold94(Cr, Ci, Br, Bi);                         // (64 - i, 128 - i)
        re[aA] = (int16_t)Ar; im[aA] = (int16_t)Ai; re[aB] = (int16_t)Br; im[aB] = (int16_t)Bi;
        re[aC] = (int16_t)Cr; im[aC] = (int16_t)Ci; re[aD] = (int16_t)Dr; im[aD] = (int16_t)Di;
    }
    // ---- six radix-2 stages on the two 64-point halves, as two passes of three stages
    // pass 1: stages 0..2 on elements [h | k | g], k = 0..7, at rev3(g)*16 + rev3(k)*2 + h
#pragma unroll 1
    for (int hg = 0; hg < 16; ++hg) {
        const int h = hg & 1, rg = hg >> 1;
        const int base = rg * 16 + h;
        int xr[8], xi[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { xr[k] = (uint16_t)re[base + 2 * dcsb_rev3(k)]; xi[k] = (uint16_t)im[base + 2 * dcsb_rev3(k)]; }
        dcsb_fft8_94(xr, xi, tw, h, 2 * h, 4 * h);
#pragma unroll
        for (int k = 0; k < 8; ++k) { re[base + 2 * dcsb_rev3(k)] = (int16_t)xr[k]; im[base + 2 * dcsb_rev3(k)] = (int16_t)xi[k]; }
    }
    // pass 2: stages 3..5 on elements [h | g | b], b = 0..7, at rev3(b)*16 + rev3(g)*2 + h;
    // then the volume shift (:532-534).  Element e lands on sample pair bitrev7(e) = its address.
#pragma unroll 1
    for (int hg = 0; hg < 16; ++hg) {
        const int h = hg & 1, rg = hg >> 1, g = dcsb_rev3(rg);
        const int base = rg * 2 + h;
        int xr[8], xi[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) { xr[b] = (uint16_t)re[base + 16 * dcsb_rev3(b)]; xi[b] = (uint16_t)im[base + 16 * dcsb_rev3(b)]; }
        dcsb_fft8_94(xr, xi, tw, 8 * h + g, 16 * h + 2 * g, 32 * h + 4 * g);
#pragma unroll
        for (int b = 0; b < 8; ++b) {           // bias off and volume shift in one: ((x' << 16) - 2^31) >> (16 + vs)
            re[base + 16 * dcsb_rev3(b)] = (int16_t)((int)((uint32_t)xr[b] * 0x10000u + 0x80000000u) >> (16 + vs));
            im[base + 16 * dcsb_rev3(b)] = (int16_t)((int)((uint32_t)xi[b] * 0x10000u + 0x80000000u) >> (16 + vs));
        }
    }
}

// =======================================================================================
// Overlap-add (:538-555), lane-private: mixes the previous frame's samples 240..255 (its row
// positions 120..127, untouched by this step) into this frame's first 16 samples in place.
DCSB_HD void dcsb_lane_overlap94(int16_t *r16, const int16_t *prev_re120, const int16_t *prev_im120, const DcsbTables *tab)
{
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        r16[m] = (int16_t)dcsb_overlap_mix(r16[m], prev_re120[m], tab->overlap[2 * m], tab->overlap[15 - 2 * m]);
        r16[128 + m] = (int16_t)dcsb_overlap_mix(r16[128 + m], prev_im120[m], tab->overlap[2 * m + 1], tab->overlap[14 - 2 * m]);
    }
}

// K2 phase C for one finished frame row: interleave the planes, coalesced 8-byte PCM stores,
// checksum partials (sum over samples i of (uint16)s[i] * (2i + 1)).
#if DCSB_DEVICE_PASS
typedef uint32_t DcsbAcc;       // per-lane partial sums stay below 2^32 (at most 2 word pairs per lane and frame)
#else
typedef uint64_t DcsbAcc;       // the simulator's single "lane" sums the whole frame
#endif
// a.lo16 * b.byte0 + a.hi16 * b.byte1 + c, all unsigned (IDP.2A on the device)
DCSB_HD uint32_t dcsb_dp2a(uint32_t a, uint32_t b, uint32_t c)
{
#if DCSB_DEVICE_PASS
    return __dp2a_lo(a, b, c);
#else
    return c + (a & 0xFFFFu) * (b & 0xFFu) + (a >> 16) * ((b >> 8) & 0xFFu);
#endif
}

DCSB_HD void dcsb_frame_output94(const int16_t *r16, uint32_t *pcm32, uint32_t frame, unsigned long long &csum)
{
    DcsbAcc sa = 0, sb = 0;
    const uint32_t *r32 = reinterpret_cast<const uint32_t *>(r16);
    DCSB_FOR_LANES(q, 60) {                    // word pair q = samples 4q .. 4q+3
        const uint32_t rr = r32[q], ii = r32[64 + q];
        const uint32_t w0 = (rr & 0xFFFFu) | (ii << 16), w1 = (rr >> 16) | (ii & 0xFFFF0000u);
        *reinterpret_cast<uint2 *>(pcm32 + (size_t)frame * 120 + 2 * q) = make_uint2(w0, w1);
        const uint32_t t = dcsb_dp2a(w1, 0x0101u, dcsb_dp2a(w0, 0x0101u, 0u));        // u0 + u1 + u2 + u3
        sa += t;
        sb += (DcsbAcc)t * (uint32_t)(8 * q + 1) + dcsb_dp2a(w1, 0x0604u, dcsb_dp2a(w0, 0x0200u, 0u));   // + 2 u1 + 4 u2 + 6 u3
    }
    csum += (unsigned long long)sa * (480ull * frame) + sb;
}

// =======================================================================================
// K2 body: one warp renders one work item (count consecutive output frames of a stream) tile
// by tile; lane l of a tile owns frame tb + l.  rows = 32 x 129 words + 8 words of carried tail.
// Returns this lane's share of the item checksum (device) / the whole share (simulator).
#define DCSB_ROW94_WORDS 129
#define DCSB_WARP94_WORDS (32 * DCSB_ROW94_WORDS + 8)

DCSB_HD unsigned long long dcsb_decode94_item(const uint8_t *slab, const DcsbStreamRec *streams, DcsbTile it,
                                              const DcsbTables *tab, const uint16_t *lut, const DcsbTw94 *tw,
                                              const uint8_t *hdr, const DcsbScanOut &scan, uint32_t nplay, int stopband,
                                              int16_t *pcm, uint32_t *rows)
{
    const DcsbStreamRec *sp = streams + it.stream;
    const uint32_t fb = sp->frame_base;
    const uint32_t fend = it.first + it.count;
    int16_t *tail = reinterpret_cast<int16_t *>(rows + 32 * DCSB_ROW94_WORDS);   // re[0..8), im[8..16)
    uint32_t *pcm32 = reinterpret_cast<uint32_t *>(pcm + sp->pcm_off);
    unsigned long long csum = 0;

    uint32_t cur = it.first;
    bool have_tail = cur == 0;                     // a fresh decoder's overlap buffer is zero
    if (have_tail) {
        DCSB_FOR_LANES(i, 8) rows[32 * DCSB_ROW94_WORDS + i] = 0;
        DCSB_SYNCWARP();
    }
    while (cur < fend) {
        const uint32_t tb = have_tail ? cur : cur - 1;          // frame of lane 0 (warm-up frame if no tail yet)
        const int out_from = have_tail ? 0 : 1;
        const int nfr = (int)(fend - tb < 32u ? fend - tb : 32u);
        // ---- phases A + B, lane-private
        DCSB_FOR_LANES(l, 32) {
            const uint32_t f = tb + (uint32_t)l;
            uint32_t *row = rows + l * DCSB_ROW94_WORDS;
            if (l < nfr) {
                for (int i = 0; i < 128; ++i) row[i] = 0;
                if (f < nplay) {
                    int16_t *r16 = reinterpret_cast<int16_t *>(row);
                    const uint2 bp = DCSB_LDCG(scan.bt + fb + f), bc = DCSB_LDCG(scan.bt + fb + f + 1);
                    DcsbWin win = dcsb_make_window(slab, *sp, DCSB_LDCG(scan.bitpos + fb + f) + DCSB_LDCG(scan.hdrbits + fb + f));
                    const int zero_from = (f == nplay - 1 && stopband != 0xFF) ? stopband : 16;
                    dcsb_lane_decode94<false>(hdr, lut, win, ((uint64_t)bp.y << 32) | bp.x, ((uint64_t)bc.y << 32) | bc.x,
                                              f == 0 ? sp->mult0 : sp->mult1, zero_from, r16);
                    dcsb_lane_transform94(r16, tw, f == 0 ? sp->vs0 : sp->vs1);
                }
            }
        }
        DCSB_SYNCWARP();
        // ---- overlap-add, lane-private (reads the neighbour row's tail, writes its own head)
        DCSB_FOR_LANES(l, 32) {
            if (l >= out_from && l < nfr) {
                int16_t *r16 = reinterpret_cast<int16_t *>(rows + l * DCSB_ROW94_WORDS);
                const int16_t *pr = l ? r16 - 2 * DCSB_ROW94_WORDS + 120 : tail;
                const int16_t *pi = l ? r16 - 2 * DCSB_ROW94_WORDS + 248 : tail + 8;
                dcsb_lane_overlap94(r16, pr, pi, tab);
            }
        }
        DCSB_SYNCWARP();
        // ---- phase C, warp-cooperative per frame
        for (int k = out_from; k < nfr; ++k)
            dcsb_frame_output94(reinterpret_cast<const int16_t *>(rows + k * DCSB_ROW94_WORDS), pcm32, tb + (uint32_t)k, csum);
        DCSB_SYNCWARP();
        // carry the last frame's samples 240..255 into the next tile
        {
            const int16_t *last = reinterpret_cast<const int16_t *>(rows + (nfr - 1) * DCSB_ROW94_WORDS);
            DCSB_FOR_LANES(i, 16) tail[i] = i < 8 ? last[120 + i] : last[248 + i - 8];
            DCSB_SYNCWARP();
        }
        cur = tb + (uint32_t)nfr;
        have_tail = true;
    }
    return csum;
}
