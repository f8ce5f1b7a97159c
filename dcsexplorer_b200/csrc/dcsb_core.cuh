// dcsb200 device core: bit reader, per-frame walkers (scan = lengths only, decode = bins)
// and the fixed-point inverse transforms.  Everything here is written once and compiled
// twice: by nvcc for sm_100a (the product) and by g++ as plain C++ for the host-side
// kernel simulator used only by the CPU test-suite (tests/hostsim) -- that build replaces
// the warp by a loop over lanes and is never part of the product library.
//
// Format/arith sources: DCSDecoder/DCSDecoderNative.cpp (1994 frames :1679-2261, 1993 frames
// :2293-2684, OS93a type 1 :2831-3032, transforms :397-576 / :614-813, MAC rounding :3503-3580).
#pragma once
#include <stdint.h>
#include "dcsb_internal.h"

#if defined(__CUDACC__)
#define DCSB_HD __host__ __device__ __forceinline__
#else
#define DCSB_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define DCSB_DEVICE_PASS 1
#else
#define DCSB_DEVICE_PASS 0
#endif

// ---------------------------------------------------------------------------------------
// Lane loops: on the device `lane` strides by 32; in the host simulator one "thread" plays
// all 32 lanes of a step, steps being separated by warp barriers.
#if DCSB_DEVICE_PASS
#define DCSB_FOR_LANES(var, n) for (int var = (int)(threadIdx.x & 31); var < (n); var += 32)
#define DCSB_SYNCWARP() __syncwarp()
#define DCSB_LDG(p) __ldg(p)
#else
#define DCSB_FOR_LANES(var, n) for (int var = 0; var < (n); ++var)
#define DCSB_SYNCWARP() ((void)0)
#define DCSB_LDG(p) (*(p))
#endif

// ---------------------------------------------------------------------------------------
// MSB-first bit reader over 32-bit words (ROMBitPointer, DCSDecoderNative.h:229-289, as a
// pure function of the bit position).  `w` is 4-byte aligned; `bias` is the bit offset of
// the stream's first data bit inside w[0].
struct DcsbBits {
    const uint32_t *w;
    uint32_t bias;
    DCSB_HD static uint32_t be(uint32_t v)
    {
#if DCSB_DEVICE_PASS
        return __byte_perm(v, 0, 0x0123);
#else
        return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24);
#endif
    }
    // next n bits (1 <= n <= 24) at bit position pos
    DCSB_HD uint32_t peek(uint32_t pos, int n) const
    {
        uint32_t p = pos + bias;
        uint32_t i = p >> 5, s = p & 31;
#if DCSB_DEVICE_PASS
        uint32_t a = be(__ldg(w + i)), b = be(__ldg(w + i + 1));
        return __funnelshift_l(b, a, s) >> (32 - n);
#else
        uint64_t v = ((uint64_t)be(w[i]) << 32) | be(w[i + 1]);
        return (uint32_t)((v << s) >> (64 - n));
#endif
    }
};

DCSB_HD int dcsb_sext(uint32_t v, int n) { return (int)(v << (32 - n)) >> (32 - n); }
// 16-bit saturation.  On the device the clamp goes through opaque min/max PTX: written as
// plain C++ on values that came from int16, nvcc 12.9 narrows `sat16(x + (-sat16(y)))` to a
// 16-bit ssub.sat whose negated operand wraps at -32768 (wrong PCM on loud frames; the -G
// build was right).  ptxas still fuses add + min into VIADDMNMX.
DCSB_HD int dcsb_sat16(int v)
{
#if DCSB_DEVICE_PASS
    int r;
    asm("min.s32 %0, %1, 32767;\n\tmax.s32 %0, %0, -32768;" : "=r"(r) : "r"(v));
    return r;
#else
    return v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
#endif
}
DCSB_HD int dcsb_s16(uint32_t v) { return (int)(int16_t)(uint16_t)v; }

// ADSP-2105 multiply/accumulate + round in 32-bit wrap-around arithmetic:
// returns sext16( MR1( 2ab -/+ 2cd + rounding ) ) where the "unbiased rounding" rule
// clears bit 16 when the low word of the SECOND product is exactly 0x8000 (:3503-3554).
template <bool SUB>
DCSB_HD int dcsb_mac_round(int a, int b, int c, int d)
{
    uint32_t p1 = (uint32_t)(a * b) << 1;
    uint32_t p2 = (uint32_t)(c * d) << 1;
    uint32_t r = (SUB ? p1 - p2 : p1 + p2) + 0x8000u;
    if ((p2 & 0xFFFFu) == 0x8000u) r &= ~0x10000u;
    return (int)r >> 16;
}

// The same value as dcsb_mac_round(a, b, c, d) for b2 = 2b, d2 = 2d, left in bits 16..31 of the result (the low half is
// unspecified).  The rounding constant rides on the second product (an IMAD addend), so the tie -- low word of the
// second product exactly 0x8000 -- shows as a zero low word of q, and clearing bit 16 is one AND with a mask made
// from it: 0xFFFE0000 | -low has bit 16 set for every low in 1..0xFFFF.  No shift, no predicate.
template <bool SUB>
DCSB_HD uint32_t dcsb_mac_hi(int a, int b2, int c, int d2)
{
    const uint32_t q = (uint32_t)c * (uint32_t)d2 + (SUB ? 0xFFFF8000u : 0x8000u);
    const uint32_t r = SUB ? (uint32_t)a * (uint32_t)b2 - q : (uint32_t)a * (uint32_t)b2 + q;
    return r & (0xFFFE0000u | (0u - (q & 0xFFFFu)));
}
// (high half of lo) | (high half of hi) << 16
DCSB_HD uint32_t dcsb_hi_pair(uint32_t lo, uint32_t hi)
{
#if DCSB_DEVICE_PASS
    return __byte_perm(lo, hi, 0x7632);
#else
    return (lo >> 16) | (hi & 0xFFFF0000u);
#endif
}
// halfword-wise wrapping add / subtract
DCSB_HD uint32_t dcsb_add2(uint32_t x, uint32_t y)
{
#if DCSB_DEVICE_PASS
    return __vadd2(x, y);
#else
    return ((x + y) & 0xFFFFu) | (((x >> 16) + (y >> 16)) << 16);
#endif
}
DCSB_HD uint32_t dcsb_sub2(uint32_t x, uint32_t y)
{
#if DCSB_DEVICE_PASS
    return __vsub2(x, y);
#else
    return ((x - y) & 0xFFFFu) | (((x >> 16) - (y >> 16)) << 16);
#endif
}

// scale mantissa table {0x8000,0x9838,0xb505,0xd745} >> (15 - exponent)  (:1978-1979, :2337-2343)
DCSB_HD uint32_t dcsb_scale_factor(int code)
{
    uint32_t m = (code & 2) ? ((code & 1) ? 0xd745u : 0xb505u) : ((code & 1) ? 0x9838u : 0x8000u);
    return m >> (15 - ((code >> 2) & 15));
}

// add one dequantised sample into a bin (:2244-2250, :2434-2443)
DCSB_HD void dcsb_add_bin(int16_t *row, int idx, int sample, uint32_t scale, uint32_t mult)
{
    uint32_t ss = ((uint32_t)sample * scale) & 0xFFFFu;
    int c = ((int)ss + dcsb_s16(ss) * (int)mult) >> 16;
    row[idx] = (int16_t)(row[idx] + c);
}

DCSB_HD void dcsb_fix_bin01(int16_t *row, int old1)
{
    int delta = dcsb_sat16((int)row[1] - old1);
    row[0] = (int16_t)dcsb_sat16(delta + (int)row[0]);
    row[1] = (int16_t)old1;
}

// prefix codes longer than the 8-bit peek LUT: match bit-serially against the long list
DCSB_HD int dcsb_long_code(const DcsbBits &rd, uint32_t &pos, const DcsbLongCode *lc, int n)
{
    for (int i = 0; i < n; ++i) {
        int len = lc[i].len;
        uint32_t v = len <= 24 ? rd.peek(pos, len)
                               : ((rd.peek(pos, 24) << (len - 24)) | rd.peek(pos + 24, len - 24));
        if (v == lc[i].code) { pos += len; return lc[i].val; }
    }
    return -1;
}

// ---------------------------------------------------------------------------------------
// scan -> decode hand-off (both kernels resident at the same time): the scan publishes how many
// checkpoints of a stream are valid, the decode warps wait for the ones they need and read them
// past L1 (a line fetched earlier may predate the scan's later writes to it).
#if DCSB_DEVICE_PASS
DCSB_HD void dcsb_publish(uint32_t *progress, int si, uint32_t v)
{
    if (!progress) return;
    __threadfence();
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(progress + si), "r"(v) : "memory");
}
// returns true when the stream's scan is finished (nplay / stopband are final)
__device__ __forceinline__ bool dcsb_await(const uint32_t *progress, uint32_t stream, uint32_t need, uint32_t *errw = nullptr)
{
    if (!progress) return true;
    uint32_t v = 0;
    if ((threadIdx.x & 31) == 0) {
        const long long t0 = clock64();
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(progress + stream) : "memory");
            if ((v & DCSB_SCAN_DONE) || v >= need) break;
            if (clock64() - t0 > 4000000000ll) {             // ~2 s: never hang the GPU on a lost producer ...
                if (errw) atomicOr(errw, DCSB_DEV_E_AWAIT);  // ... and never let it pass for a success
                break;
            }
            __nanosleep(256);
        }
    }
    v = __shfl_sync(0xffffffffu, v, 0);
    return (v & DCSB_SCAN_DONE) != 0;
}
#define DCSB_LDCG(p) __ldcg(p)
// append the stream's work items covering output frames [from, to) to the ready queue
__device__ __forceinline__ void dcsb_queue_push(const DcsbScanOut &out, int si, uint32_t from, uint32_t to, bool fin)
{
    if (!out.queue || from >= to) return;
    __threadfence();
    const uint32_t n = (to - from + DCSB_QITEM - 1) / DCSB_QITEM;
    const uint32_t slot = atomicAdd(out.qctl, n);
    for (uint32_t k = 0; k < n; ++k) {
        const unsigned long long e = DCSB_Q_VALID | (fin ? DCSB_Q_FINAL : 0ull) | ((unsigned long long)(uint32_t)si << 24) |
                                     (unsigned long long)(from + k * DCSB_QITEM);
        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(out.queue + slot + k), "l"(e) : "memory");
    }
}
#else
struct DcsbScanOut;
DCSB_HD void dcsb_queue_push(const DcsbScanOut &, int, uint32_t, uint32_t, bool) {}
DCSB_HD void dcsb_publish(uint32_t *, int, uint32_t) {}
DCSB_HD bool dcsb_await(const uint32_t *, uint32_t, uint32_t, uint32_t * = nullptr) { return true; }
#define DCSB_LDCG(p) (*(p))
#endif

#define DCSB_WALK_OK        0
#define DCSB_WALK_BANDTYPE -3

struct DcsbWalkCtx {
    DcsbBits rd;
    const uint8_t *hdr;        // 16 header bytes
    const uint16_t *lut;       // DCSB_LUT_* block
    const DcsbTables *tab;     // full tables (global memory) for the rare paths
    uint32_t mult;             // effective channel multiplier
    int zero_from;             // bands >= this contribute nothing (the reference's error path)
};

// ======================= 1993 frame (:2293-2684) =======================================
// `row` is only touched in DECODE mode.
template <bool DECODE>
DCSB_HD int dcsb_walk93(const DcsbWalkCtx &cx, uint32_t &pos, uint64_t &bt, int16_t *row)
{
    const uint8_t *hdr = cx.hdr;
    const uint16_t *lut = cx.lut;
    const int type1 = hdr[0] >> 7;
    int subtype = type1 ? 0 : 2;                                         // :2309
    uint32_t prv = 0, prvd = 0;                                          // 16-bit wrap-around state
    int reuse = 0, code = 0;
    int old1 = 0;
    if (DECODE) old1 = row[1];
    int idx = 1;
    for (int band = 0; band < 16; ++band) {
        int hb = hdr[band] & 0x7F;
        if (hb == 0x7F) break;
        const uint32_t scale = dcsb_scale_factor(hb);                    // :2337-2343
        const int stride2 = hb >> 6;
        int n, inc, fixup, stride;
        if (!type1) {                                                    // :2351-2368
            if (!stride2) { n = 16; inc = 1; fixup = 0; stride = 16; }
            else { ++idx; n = 16; inc = 2; fixup = -1; stride = 31; }
        } else {                                                         // :2369-2383
            if (!stride2) { inc = 1; fixup = 0; n = stride = (band == 0 ? 15 : 16); }
            else { inc = 2; fixup = 0; n = stride = 8; }
        }
        if (reuse) { reuse = (int)cx.rd.peek(pos, 1); pos += 1; }        // :2388-2389
        if (!reuse) {
            if (!type1) {                                                // :2396-2419
                uint32_t v = cx.rd.peek(pos, 6);
                if (v & 0x20) {
                    subtype = (v & 0x10) ? (subtype == 2 ? 0 : subtype + 1) : (subtype == 0 ? 2 : subtype - 1);
                    code = (int)(v & 15);
                    pos += 6;
                } else {
                    code = (int)((v >> 1) & 15);
                    pos += 5;
                }
            } else {                                                     // :2420-2430, :2618-2684
                uint32_t e = lut[DCSB_LUT_HDR93 + cx.rd.peek(pos, 8)];
                int v;
                if (e) { pos += e >> 8; v = (int)(e & 0xFF); }
                else v = dcsb_long_code(cx.rd, pos, cx.tab->long93, cx.tab->n_long93);
                int delta;
                if (v < 0x1E) delta = v - 0x0F;
                else { delta = v - 0x2E; subtype = subtype ? 0 : 1; }
                int nbt = (int)((bt >> (4 * band)) & 15) + delta;
                if (v < 0 || nbt < 0 || nbt > 15) return DCSB_WALK_BANDTYPE;
                bt = (bt & ~(15ull << (4 * band))) | ((uint64_t)nbt << (4 * band));
                code = nbt;
            }
        }
        if (code == 0) {                                                 // :2446-2547
            reuse = 1;
            if (subtype == 0) { idx += stride; prv = 0; prvd = 0; }
            else if (subtype == 1) {
                if (DECODE) {
                    // running 16.16 accumulator whose low word carries between samples (:2513-2535)
                    uint32_t low = ((uint32_t)dcsb_s16(prv) * scale) & 0xFFFFu;
                    const int step = dcsb_s16(low) * (int)cx.mult;
                    for (int i = 0; i < n; ++i) {
                        if (idx < 512) {
                            int t = (int)low + step;     // |t| < 2^31: low < 2^16, |step| <= 2^15 * (2^16-1)
                            row[idx] = (int16_t)(row[idx] + (t >> 16));
                            low = (uint32_t)t & 0xFFFFu;
                        }
                        idx += inc;
                    }
                } else idx += n * inc;
                prvd = 0;
                idx += fixup;
            } else {
                for (int i = 0; i < n; ++i) {
                    prv = (prv + prvd) & 0xFFFFu;
                    if (DECODE && idx < 512) dcsb_add_bin(row, idx, dcsb_s16(prv), scale, cx.mult);
                    idx += inc;
                }
                idx += fixup;
            }
        } else {                                                         // :2548-2603
            const int width = code + (type1 ? 0 : 1);
            if (!DECODE) {
                // length-only walk (the scan): the band is n samples of `width` bits, and no sample value steers the
                // walk -- band types, subtype changes and reuse flags all come from header codes
                pos += (uint32_t)(width * n);
                idx += n * inc + fixup;
                continue;
            }
            uint32_t last = 0, last2 = 0;
            for (int i = 0; i < n; ++i) {
                uint32_t in = (uint32_t)dcsb_sext(cx.rd.peek(pos, width), width) & 0xFFFFu;
                pos += width;
                uint32_t out;
                if (subtype == 0) { out = in; last2 = last; last = in; }
                else if (subtype == 1) { prvd = in; prv = (prv + prvd) & 0xFFFFu; out = prv; }
                else { prvd = (prvd + in) & 0xFFFFu; prv = (prv + prvd) & 0xFFFFu; out = prv; }
                if (DECODE && idx < 512) dcsb_add_bin(row, idx, dcsb_s16(out), scale, cx.mult);
                idx += inc;
            }
            if (subtype == 0) { prv = last; prvd = (last - last2) & 0xFFFFu; }
            idx += fixup;
        }
    }
    if (DECODE) dcsb_fix_bin01(row, old1);
    return DCSB_WALK_OK;
}

// ======================= OS93a type-1 frame (:2831-3032) ===============================
template <bool DECODE>
DCSB_HD int dcsb_walk93a1(const DcsbWalkCtx &cx, uint32_t &pos, int16_t *row)
{
    const uint16_t *lut = cx.lut;
    const int hb = cx.hdr[0];
    const int sel = (hb & 0x60) >> 5, nbands = hb & 0x1F;
    int prvscale = 0x1A, idx = 0;
    for (int b = 0; b < nbands && b < 18; ++b) {
        // inputs per band {2,2,2,2,3,4,5,6,5,6,7,9,11,14,12,12,12,13} (:2865)
        const int ninputs = (int)((b < 16 ? (0xcceb976565432222ull >> (4 * b)) : (0xdcull >> (4 * (b - 16)))) & 15);
        uint32_t e = lut[DCSB_LUT_BB93A + sel * 16 + cx.rd.peek(pos, 4)];
        pos += e >> 8;
        const int bits = (int)(e & 0xFF);
        if (bits == 0xFF) break;                                         // :2922
        if (bits == 0) { idx += ninputs * 2; continue; }
        e = lut[DCSB_LUT_SC93A + cx.rd.peek(pos, 8)];                    // :2938-2970
        pos += e >> 8;
        int sc = prvscale + (int)(e & 0xFF) - 1 + bits * 2;              // :2975-2981
        if (sc > 0x39) sc -= 0x36;
        prvscale = sc - bits * 2;
        if (DECODE) {
            uint32_t sf = 0x8000u;
            for (int i = 0; i < (sc & 3); ++i) sf = (sf * 0x9838u) >> 15;    // :2986-2991
            sf <<= ((sc >> 2) & 31);    // count >= 32 only on malformed streams: follow the x86 build of the reference
            sf = ((sf >> 16) * cx.mult) >> 15;                           // :2995
            const int sfs = dcsb_s16(sf);
            const uint16_t *base = cx.tab->pairs93a + (2 << bits);
            for (int i = 0; i < ninputs; ++i) {
                uint32_t smp = cx.rd.peek(pos, bits);
                pos += bits;
                for (int k = 0; k < 2; ++k, ++idx) {
                    if (idx >= 512) continue;
                    // MultiplyRoundAdd on MR = bin << 16 (:3010-3015, :3540-3546)
                    uint32_t p = (uint32_t)(dcsb_s16(base[smp * 2 + k]) * sfs) << 1;
                    uint32_t r = ((uint32_t)(uint16_t)row[idx] << 16) + p + 0x8000u;
                    if ((p & 0xFFFFu) == 0x8000u) r &= ~0x10000u;
                    row[idx] = (int16_t)(r >> 16);
                }
            }
        } else {
            pos += bits * ninputs;
            idx += ninputs * 2;
        }
    }
    return DCSB_WALK_OK;
}

template <bool DECODE>
DCSB_HD int dcsb_walk(int fmt, const DcsbWalkCtx &cx, uint32_t &pos, uint64_t &bt, int16_t *row, int &stop_band)
{
    // (the 1994 layout never comes here: dcsb_fast94.cuh)
    if (fmt == DCSB_FMT_93) return dcsb_walk93<DECODE>(cx, pos, bt, row);
    return dcsb_walk93a1<DECODE>(cx, pos, row);
}

// ---------------------------------------------------------------------------------------
// Transforms.  A frame lives in shared memory as 16-bit bins; complex element k is the
// 32-bit word k = (re = bin 2k in the low half, im = bin 2k+1 in the high half).
DCSB_HD uint32_t dcsb_pack(int re, int im) { return ((uint32_t)re & 0xFFFFu) | ((uint32_t)im << 16); }
DCSB_HD int dcsb_re(uint32_t w) { return (int)(int16_t)(w & 0xFFFFu); }
DCSB_HD int dcsb_im(uint32_t w) { return (int)w >> 16; }
DCSB_HD int dcsb_rev7(int x)
{
#if DCSB_DEVICE_PASS
    return (int)(__brev((unsigned)x) >> 25);
#else
    int r = 0;
    for (int i = 0; i < 7; ++i) r |= ((x >> i) & 1) << (6 - i);
    return r;
#endif
}
DCSB_HD int dcsb_rev8(int x)
{
#if DCSB_DEVICE_PASS
    return (int)(__brev((unsigned)x) >> 24);
#else
    int r = 0;
    for (int i = 0; i < 8; ++i) r |= ((x >> i) & 1) << (7 - i);
    return r;
#endif
}

// overlap-add of one of the first 16 samples (:538-555, :787-802)
DCSB_HD int dcsb_overlap_mix(int cur, int prev, uint32_t wcur, uint32_t wprev)
{
    uint32_t a = (uint32_t)(cur * (int)wcur) << 1;
    uint32_t b = (uint32_t)(prev * (int)wprev) << 1;
    return (int)(a + b + 0x8000u) >> 16;
}

// radix-2 butterfly with the reference's rounding; SAT selects the 1994 (saturating) or
// 1993 (wrapping) flavour.  u' = u - t, a' = u + t, t = a * (cos + i sin) (:480-524, :742-778)
template <bool SAT>
DCSB_HD void dcsb_butterfly(uint32_t &u, uint32_t &a, uint32_t tw)
{
    const int sv = dcsb_re(tw), cv = dcsb_im(tw);
    const int ar = dcsb_re(a), ai = dcsb_im(a), ur = dcsb_re(u), ui = dcsb_im(u);
    const int tr = dcsb_mac_round<true>(ar, cv, ai, sv);
    const int ti = dcsb_mac_round<false>(ai, cv, ar, sv);
    if (SAT) {
        u = dcsb_pack(dcsb_sat16(ur - tr), dcsb_sat16(ui - ti));
        a = dcsb_pack(dcsb_sat16(ur + tr), dcsb_sat16(ui + ti));
    } else {
        u = dcsb_pack(ur - tr, ui - ti);
        a = dcsb_pack(ur + tr, ui + ti);
    }
}

// The 1993 (wrapping) butterfly on packed elements with pre-doubled twiddles: the products stay in the high halves,
// one byte permute packs t, the two results are halfword-wise add / subtract (:742-778)
DCSB_HD void dcsb_butterfly93(uint32_t &u, uint32_t &a, int c2, int s2)
{
    const int ar = dcsb_re(a), ai = dcsb_im(a);
    const uint32_t t = dcsb_hi_pair(dcsb_mac_hi<true>(ar, c2, ai, s2), dcsb_mac_hi<false>(ai, c2, ar, s2));
    const uint32_t nu = dcsb_sub2(u, t);
    a = dcsb_add2(u, t);
    u = nu;
}

// 1994 transform, one warp per frame, in place on c[0..128] (word 128 is the always-zero
// phantom element the reference reads at frameBuffer[0x100], :405-418).  After the call
// c[k] holds complex element k of the finished IFFT (before volume shift / reordering).
DCSB_HD void dcsb_transform94_warp(uint32_t *c, const DcsbTables *tab)
{
    // pairing pass + twiddle pass, fused: both work on the pairs (i, 128-i) (:403-456)
    DCSB_FOR_LANES(i, 64) {
        uint32_t A = c[i], B = (i == 0) ? 0u : c[128 - i];
        const int x0 = dcsb_re(A), x1 = dcsb_im(A), y0 = dcsb_re(B), y1 = dcsb_im(B);
        // MulSS(v, 0x8000) == wrap16(-v)
        const int p0r = dcsb_s16((uint32_t)-dcsb_sat16(x0 + y0)), p1r = dcsb_s16((uint32_t)-dcsb_sat16(x0 - y0));
        const int p0i = dcsb_s16((uint32_t)-dcsb_sat16(x1 - y1)), p1i = dcsb_s16((uint32_t)-dcsb_sat16(x1 + y1));
        const uint32_t tw = tab->pretw[i];
        const int c1 = dcsb_re(tw), c0 = dcsb_im(tw);
        const int prod0 = dcsb_mac_round<true>(p1i, c1, p1r, c0);
        const int prod1 = dcsb_mac_round<false>(p1i, c0, p1r, c1);
        c[i] = dcsb_pack(dcsb_sat16(prod1 + p0r), dcsb_sat16(prod0 + p0i));
        if (i) c[128 - i] = dcsb_pack(dcsb_sat16(p0r - prod1), dcsb_sat16(prod0 - p0i));
        else {
            // element 64 is only negated in its real part (:403-404)
            uint32_t M = c[64];
            c[64] = dcsb_pack(-dcsb_re(M), dcsb_im(M));
        }
    }
    DCSB_SYNCWARP();
    // half fold (:458-471): complex k and k+64, saturating, no twiddle
    DCSB_FOR_LANES(k, 64) {
        uint32_t u = c[k], a = c[k + 64];
        const int ur = dcsb_re(u), ui = dcsb_im(u), ar = dcsb_re(a), ai = dcsb_im(a);
        c[k] = dcsb_pack(dcsb_sat16(ur + ar), dcsb_sat16(ui + ai));
        c[k + 64] = dcsb_pack(dcsb_sat16(ur - ar), dcsb_sat16(ui - ai));
    }
    DCSB_SYNCWARP();
    // six radix-2 stages over the two 64-point halves (:480-524)
    for (int st = 0; st < 6; ++st) {
        const int span = 32 >> st;                  // complex elements between butterfly legs
        DCSB_FOR_LANES(b, 64) {
            const int p = b >> (5 - st), j = b & (span - 1);
            const int e0 = p * 2 * span + j;
            uint32_t u = c[e0], a = c[e0 + span];
            dcsb_butterfly<true>(u, a, tab->twiddle[p]);
            c[e0] = u;
            c[e0 + span] = a;
        }
        DCSB_SYNCWARP();
    }
}

// magnitude of (bin0 + i bin1) by the 1.15 Taylor series (:633-710); returns the new bin 0
DCSB_HD uint32_t dcsb_magnitude93(uint32_t c0)
{
    uint32_t AR = c0 & 0xFFFFu;
    const int b1 = dcsb_im(c0);
    const bool neg = dcsb_s16(AR) < 0;
    if (neg) AR = (uint32_t)(-dcsb_s16(AR)) & 0xFFFFu;
    long long MR = (((long long)b1 * b1) << 1) + (((long long)dcsb_s16(AR) * dcsb_s16(AR)) << 1);
    uint32_t SR = (uint32_t)(MR & 0xFFFFFFFFll);
    int exponent = 0;                                              // CalcExp32 (:3447-3459)
    {
        uint32_t x = SR;
        if (x & 0x80000000u) { for (; x & 0x40000000u; --exponent, x <<= 1) ; }
        else { for (; exponent > -31 && !(x & 0x40000000u); --exponent, x <<= 1) ; }
    }
    if (exponent < 0) SR <<= -exponent;
    AR = SR >> 16;
    if (AR != 0) {
        const int k[5] = { 0x5D1D, -22035, 0x46D6, -8790, 0x072D };
        unsigned long long mr = 0x0D490000ull;
        int mf = dcsb_s16(AR);
        for (int t = 0; t < 5; ++t) {
            mr += (unsigned long long)((long long)k[t] * (long long)mf * 2);
            if (t < 4) mf = dcsb_mac_round<false>(0, 0, dcsb_s16(AR), mf);
        }
        if (exponent & 1) {
            int prod = (int)((uint32_t)(dcsb_s16((uint32_t)(mr >> 16)) * 0x5A82) << 1);
            long long r = (long long)prod + 0x8000;
            if ((prod & 0xFFFF) == 0x8000) r &= ~0x10000ll;
            mr = (unsigned long long)r;
            exponent += 1;
        }
        exponent = exponent / 2 + 1;
        const int v = (int)(uint32_t)(mr & 0xFFFFFFFFull);
        uint32_t sr;                                                // BitShiftSigned32 (:3486-3501)
        if (exponent >= 0) sr = (uint32_t)v << exponent;
        else if (v >= 0) sr = (uint32_t)v >> -exponent;
        else sr = ((uint32_t)v >> -exponent) | (~0u << (32 + exponent));
        AR = sr >> 16;
        if (neg) AR = (uint32_t)(-dcsb_s16(AR)) & 0xFFFFu;
    }
    return AR;
}

// 1993 transform on c[0..257] (256 complex points + the wrap-around element 128 the
// reference writes at frameBuffer[0x100]) (:614-813).  Leaves the IFFT in c[0..255].
// One warp per frame (the first path's form; the tile kernel now uses the lane-private one below).
DCSB_HD void dcsb_transform93_warp(uint32_t *c, const DcsbTables *tab)
{
    DCSB_FOR_LANES(l, 1) {
        const uint32_t AR = dcsb_magnitude93(c[0]);
        c[0] = AR;          // imaginary part zero
        c[128] = AR;
    }
    DCSB_SYNCWARP();
    // 256 -> 512 expansion with wrap-around arithmetic (:714-732): complex k=1+i and 127-i
    DCSB_FOR_LANES(i, 64) {
        const int k0 = 1 + i, k1 = 127 - i;
        const uint32_t X = c[k0], Y = c[k1];
        const int xr = dcsb_re(X), xi = dcsb_im(X), yr = dcsb_re(Y), yi = dcsb_im(Y);
        c[129 + i] = dcsb_pack(xr - yr, xi + yi);
        c[255 - i] = dcsb_pack(yr - xr, xi + yi);
        c[k0] = dcsb_pack(xr + yr, xi - yi);
        c[k1] = dcsb_pack(xr + yr, yi - xi);        // i == 63: k0 == k1 and both parts agree (imaginary 0)
    }
    DCSB_SYNCWARP();
    // seven radix-2 stages over 256 complex points (:742-778)
    for (int st = 0; st < 7; ++st) {
        const int span = 64 >> st;
        DCSB_FOR_LANES(b, 128) {
            const int p = b >> (6 - st), j = b & (span - 1);
            const int e0 = p * 2 * span + j;
            uint32_t u = c[e0], a = c[e0 + span];
            const DcsbTw2 tw = tab->tw93[p];
            dcsb_butterfly93(u, a, tw.c2, tw.s2);
            c[e0] = u;
            c[e0 + span] = a;
        }
        DCSB_SYNCWARP();
    }
}

// The same transform with one LANE per frame: every lane works on its own row (row stride 257 words,
// so equal offsets in different rows hit different banks), no warp synchronisation; three radix-2
// stages at a time on 8 points held in registers (stages 0-2, 3-5), then stage 6.
DCSB_HD void dcsb_transform93_lane(uint32_t *c, const DcsbTw2 *tw2)
{
    {
        const uint32_t AR = dcsb_magnitude93(c[0]);
        c[0] = AR;
        c[128] = AR;
    }
    for (int i = 0; i < 64; ++i) {
        const int k0 = 1 + i, k1 = 127 - i;
        const uint32_t X = c[k0], Y = c[k1];
        const int xr = dcsb_re(X), xi = dcsb_im(X), yr = dcsb_re(Y), yi = dcsb_im(Y);
        c[129 + i] = dcsb_pack(xr - yr, xi + yi);
        c[255 - i] = dcsb_pack(yr - xr, xi + yi);
        c[k0] = dcsb_pack(xr + yr, xi - yi);
        c[k1] = dcsb_pack(xr + yr, yi - xi);
    }
    // stage st pairs e and e + (64 >> st) inside partitions of 128 >> st elements; twiddle = partition index.
    // Pass over stages s0..s0+2: elements  base + k * (span >> 2), k = 0..7, span = 64 >> s0 (the first stage's
    // distance), for every base = partition start + j, j < span >> 2.
#pragma unroll 1
    for (int s0 = 0; s0 < 6; s0 += 3) {
        const int span = 64 >> s0, step = span >> 2;          // 64,16 / 8,2
        const int nparts = 1 << s0;                           // partitions at stage s0
#pragma unroll 1
        for (int g = 0; g < 32; ++g) {                        // 32 groups of 8 points
            const int p = g / step, j = g - p * step;         // partition at stage s0, offset inside the quarter-span
            const int base = p * 2 * span + j;
            uint32_t x[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = c[base + k * step];
            {   // stage s0: pairs (k, k + 4), twiddle p
                const DcsbTw2 tw = tw2[p];
#pragma unroll
                for (int k = 0; k < 4; ++k) dcsb_butterfly93(x[k], x[k + 4], tw.c2, tw.s2);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {   // stage s0 + 1: pairs (4q + k, 4q + k + 2), twiddle 2p + q
                const DcsbTw2 tw = tw2[2 * p + q];
#pragma unroll
                for (int k = 0; k < 2; ++k) dcsb_butterfly93(x[4 * q + k], x[4 * q + k + 2], tw.c2, tw.s2);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {   // stage s0 + 2: pairs (2q, 2q + 1), twiddle 4p + q
                const DcsbTw2 tw = tw2[4 * p + q];
                dcsb_butterfly93(x[2 * q], x[2 * q + 1], tw.c2, tw.s2);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) c[base + k * step] = x[k];
            (void)nparts;
        }
    }
    // stage 6: pairs (2p, 2p + 1), twiddle p
#pragma unroll 4
    for (int p = 0; p < 128; ++p) {
        uint32_t u = c[2 * p], a = c[2 * p + 1];
        const DcsbTw2 tw = tw2[p];
        dcsb_butterfly93(u, a, tw.c2, tw.s2);
        c[2 * p] = u;
        c[2 * p + 1] = a;
    }
}

// =======================================================================================
// Kernel bodies.  The __global__ wrappers in dcsb_kernels.cu only compute indices and
// carve shared memory; everything observable happens here.

DCSB_HD DcsbBits dcsb_make_reader(const uint8_t *slab, const DcsbStreamRec &s)
{
    const uint64_t start = s.data_off + 2 + s.hdr_len;
    DcsbBits rd;
    rd.w = reinterpret_cast<const uint32_t *>(slab + (start & ~3ull));
    rd.bias = (uint32_t)(start & 3) * 8;
    return rd;
}

// K1 body (1993 family; the 1994 layout has its own fast walker in dcsb_fast94.cuh): one
// thread walks one stream (lengths only) and writes a checkpoint per frame + the end entry.
// [f0, f1): the frames this call walks; f0 > 0 resumes from the checkpoint an earlier call left
// (see dcsb_scan94_stream)
DCSB_HD void dcsb_scan_stream(const uint8_t *slab, const DcsbStreamRec *streams, int si,
                              const DcsbTables *tab, const uint16_t *lut, const DcsbScanOut &out,
                              uint32_t f0 = 0, uint32_t f1 = 0xFFFFFFFFu)
{
    if (f0 && out.status[si] != DCSB_SCAN_RUNNING) return;
    const DcsbStreamRec s = streams[si];
    int status = 0;
    uint32_t nplay = 0, pos = 0;
    int stopband = 0xFF;
    if (s.nframes == 0) {
        status = -1;                               // DCSB_E_EMPTY (the host refines DCSB_E_SHORT)
    } else {
        DcsbWalkCtx cx;
        cx.rd = dcsb_make_reader(slab, s);
        cx.hdr = streams[si].hdr;
        cx.lut = lut;
        cx.tab = tab;
        cx.mult = 0;
        cx.zero_from = 16;
        const uint32_t nbits = (s.nbytes - 2 - s.hdr_len) * 8u;
        uint64_t bt = 0;                           // InitStreamPlayback zeroes the band types (:1640)
        nplay = s.nframes;
        uint32_t f = f0;
        if (f0) {
            pos = out.bitpos[s.frame_base + f0];
            const uint2 b2 = out.bt[s.frame_base + f0];
            bt = ((uint64_t)b2.y << 32) | b2.x;
        }
        const uint32_t fe = f1 < s.nframes ? f1 : s.nframes;
        for (; f < fe; ++f) {
            out.bitpos[s.frame_base + f] = pos;
            out.bt[s.frame_base + f] = make_uint2((uint32_t)bt, (uint32_t)(bt >> 32));
            out.hdrbits[s.frame_base + f] = 0;
            int sb = 99;
            int rc = dcsb_walk<false>(s.fmt, cx, pos, bt, nullptr, sb);
            if (pos > nbits) rc = -2;              // DCSB_E_TRUNCATED (whatever the bytes behind the stream made of the frame)
            if (rc) { status = rc; nplay = f; break; }
            if (sb != 99) { status = -5; nplay = f + 1; stopband = sb; ++f; break; }   // DCSB_E_STOPPED
            if ((f & 15) == 15) {
                out.bitpos[s.frame_base + f + 1] = pos;
                out.bt[s.frame_base + f + 1] = make_uint2((uint32_t)bt, (uint32_t)(bt >> 32));
                dcsb_publish(out.progress, si, f + 2);
            }
        }
        if (status == 0 || status == -5) {
            out.bitpos[s.frame_base + f] = pos;
            out.bt[s.frame_base + f] = make_uint2((uint32_t)bt, (uint32_t)(bt >> 32));
        }
        if (status == 0 && f < s.nframes) status = DCSB_SCAN_RUNNING;   // the next slice carries on from checkpoint f
    }
    out.status[si] = status;
    out.nplay[si] = nplay;
    out.endbits[si] = status == -2 ? (s.nbytes - 2 - s.hdr_len) * 8u : pos;     // a truncated stream occupies all of its bytes
    out.stopband[si] = (uint8_t)stopband;
    if (status == DCSB_SCAN_RUNNING) return;
    dcsb_publish(out.progress, si, DCSB_SCAN_DONE);
}

template <bool T93> struct DcsbRow { static constexpr int WORDS = T93 ? 257 : 129; };
// shared-memory words one warp needs: 32 frame rows + two 8-word overlap buffers
template <bool T93> struct DcsbWarpSmem { static constexpr int WORDS = 32 * DcsbRow<T93>::WORDS + 16; };

// K2 body: one warp renders one tile.  Returns this lane's share of the tile checksum.
template <bool T93>
DCSB_HD unsigned long long dcsb_decode_tile(const uint8_t *slab, const DcsbStreamRec *streams, DcsbTile tl,
                                            const DcsbTables *tab, const uint16_t *lut, const DcsbScanOut &scan,
                                            int16_t *pcm, uint32_t *rows)
{
    constexpr int ROWW = DcsbRow<T93>::WORDS;
    const DcsbStreamRec *sp = streams + tl.stream;
    const long long out_frames = sp->out_frames;
    // progress f + 2 = frame f walked and found whole (its checkpoint alone says nothing about the frame itself: a
    // stream cut short inside frame f has a checkpoint for it); the last frame of a stream is only known with DONE
    const bool fin = dcsb_await(scan.progress, tl.stream, tl.first + tl.count + 1 < sp->nframes + 1 ? tl.first + tl.count + 1 : sp->nframes + 1, scan.qctl ? scan.qctl + 2 : nullptr);
    const long long nplay = fin ? (long long)DCSB_LDCG(scan.nplay + tl.stream) : (long long)sp->nframes;
    const int fmt = sp->fmt;
    uint32_t *tails = rows + 32 * ROWW;             // two 8-word overlap buffers (ping-pong)

    DCSB_FOR_LANES(i, DcsbWarpSmem<T93>::WORDS) rows[i] = 0;
    DCSB_SYNCWARP();

    // ---- phase A: lane l decodes frame first-1+l (lane 0 = warm-up frame for the overlap)
    DCSB_FOR_LANES(l, 32) {
        const long long f = (long long)tl.first - 1 + l;
        if (f >= 0 && f < nplay && l <= (int)tl.count) {
            DcsbWalkCtx cx;
            cx.rd = dcsb_make_reader(slab, *sp);
            cx.hdr = sp->hdr;
            cx.lut = lut;
            cx.tab = tab;
            cx.mult = f == 0 ? sp->mult0 : sp->mult1;
            cx.zero_from = 16;
            if (f == nplay - 1) {
                const int sb = fin ? DCSB_LDCG(scan.stopband + tl.stream) : 0xFF;
                if (sb != 0xFF) cx.zero_from = sb;
            }
            uint32_t pos = DCSB_LDCG(scan.bitpos + sp->frame_base + (uint32_t)f);
            const uint2 b2 = DCSB_LDCG(scan.bt + sp->frame_base + (uint32_t)f);
            uint64_t bt = ((uint64_t)b2.y << 32) | b2.x;
            int sb = 99;
            dcsb_walk<true>(fmt, cx, pos, bt, reinterpret_cast<int16_t *>(rows + l * ROWW), sb);
        }
    }
    DCSB_SYNCWARP();

    // ---- phase B (1993 layouts): every lane transforms its own frame
    if (T93) {
        DCSB_FOR_LANES(l, 32) {
            const long long f = (long long)tl.first - 1 + l;
            if (f >= 0 && f < nplay && f < out_frames && l <= (int)tl.count) dcsb_transform93_lane(rows + l * ROWW, tab->tw93);
        }
        DCSB_SYNCWARP();
    }
    // ---- phase C: the frames in order, carrying the 16-sample tail
    unsigned long long csum = 0;
    const int vs0 = sp->vs0, vs1 = sp->vs1, vsi = sp->vs_idle;
    uint32_t *pcm32 = reinterpret_cast<uint32_t *>(pcm + sp->pcm_off);
    for (int k = 0; k < 32; ++k) {
        const long long f = (long long)tl.first - 1 + k;
        if (f < 0) continue;                          // first tile of a stream: the overlap buffer starts at zero
        if (f >= out_frames || k > (int)tl.count) break;
        uint32_t *c = rows + k * ROWW;
        if (!T93 && f < nplay) dcsb_transform94_warp(c, tab);       // silent frames transform to silence
        const int vs = f == 0 ? vs0 : (f < nplay ? vs1 : vsi);
        const uint32_t *tin = tails + (k & 1) * 8;
        uint32_t *tout = tails + ((k + 1) & 1) * 8;
        DCSB_FOR_LANES(m, 128) {
            int s0, s1;
            if (T93) {
                // sample n is the real part of element bitrev8(n) (:782-785)
                s0 = dcsb_re(c[dcsb_rev8(2 * m)]) >> vs;
                s1 = dcsb_re(c[dcsb_rev8(2 * m + 1)]) >> vs;
            } else {
                // samples (2m, 2m+1) are (re, im) of element bitrev7(m) (:559-565)
                const uint32_t w = c[dcsb_rev7(m)];
                s0 = dcsb_re(w) >> vs;
                s1 = dcsb_im(w) >> vs;
            }
            if (m < 8) {
                const uint32_t t = tin[m];
                s0 = dcsb_overlap_mix(s0, dcsb_re(t), tab->overlap[2 * m], tab->overlap[15 - 2 * m]);
                s1 = dcsb_overlap_mix(s1, dcsb_im(t), tab->overlap[2 * m + 1], tab->overlap[14 - 2 * m]);
            }
            const uint32_t w = dcsb_pack(s0, s1);
            if (m >= 120) tout[m - 120] = w;
            else if (k > 0) {
                pcm32[(size_t)f * 120 + m] = w;
                const unsigned long long i0 = (unsigned long long)f * 240 + 2 * m;
                csum += (unsigned long long)(w & 0xFFFFu) * (2 * i0 + 1) + (unsigned long long)(w >> 16) * (2 * i0 + 3);
            }
        }
        DCSB_SYNCWARP();
    }
    return csum;
}
