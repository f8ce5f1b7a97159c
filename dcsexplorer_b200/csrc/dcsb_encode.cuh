// dcsb200 forward path: kernel bodies and host helpers (see dcsb_encode.cu for what is restated and how it is checked).
// Everything a kernel does lives here as a function of its thread index, compiled by nvcc for the device and -- the same
// text -- by g++ for the CPU simulator of tests/hostsim, where a loop plays the grid.  Float operations go through
// enc_mul / enc_add / enc_sub / enc_div: single round-to-nearest operations on the device (no fused multiply-add), plain
// operators on the host (x86-64 without FMA contraction: tests/hostsim/Makefile passes -ffp-contract=off).
#ifndef DCSB_ENCODE_CUH
#define DCSB_ENCODE_CUH
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/dcsb200.h"
#include "dcs_tables.h"
#if defined(__CUDACC__)
#define ENC_HD __host__ __device__ __forceinline__
#else
#define ENC_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define ENC_DEV 1
#else
#define ENC_DEV 0
#endif
ENC_HD float enc_mul(float a, float b)
{
#if ENC_DEV
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
ENC_HD float enc_add(float a, float b)
{
#if ENC_DEV
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
ENC_HD float enc_sub(float a, float b)
{
#if ENC_DEV
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
ENC_HD float enc_div(float a, float b)
{
#if ENC_DEV
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
ENC_HD int enc_rev9(int idx)            // 9-bit reversal (bitRev9, DCSEncoder.cpp:147-158)
{
    int r = 0;
    for (int i = 0; i < 9; ++i) r |= ((idx >> i) & 1) << (8 - i);
    return r;
}
// OR a big-endian 32-bit word into the stream data (frames share their first and last word with their neighbours)
ENC_HD void enc_or_word(uint32_t *p, uint32_t w)
{
    const uint32_t v = (w >> 24) | ((w >> 8) & 0xFF00u) | ((w << 8) & 0xFF0000u) | (w << 24);
#if ENC_DEV
    atomicOr(p, v);
#else
    *p |= v;
#endif
}

#define ENC_THREADS 128
#define ENC_NV 5                    // scale pre-adjustment alternatives a type-1 band 0..2 can meet (Enc.cpp:710-716)

struct EncStream {                  // per stream, device copy
    uint64_t pcm_off;               // first sample in the concatenated PCM
    uint64_t n_samples;
    uint32_t frame0;                // first frame in the frame arrays
    uint32_t n_frames;
    int32_t type, subtype;
    float max_err2;                 // maximumQuantizationError squared
    float min_range;
    uint8_t hdr[16];                // stream header as stored (scale codes, 0xFF behind the kept bands, flag bits)
    int32_t bands;                  // bands kept
    int32_t fmt93;                  // 1: 1993 layout (16 bands of 16 samples, CompressFrame93b), stream type 0
};

struct EncTables {
    float coeff[896];               // DualFFT twiddles in the reference's order (:1384-1398), made on the host with its libm
    float tw2[128];                 // second twiddle table of DFTAlgorithmOrig (:1263-1280)
    float window[16];               // :1010-1013
    int scale[64];                  // scalingFactors (:78-143) = the decoder's scale table
    uint16_t xlat[48];              // type-1 band translation (:1871-1917): width << 8 | scale adjustment
    uint32_t cb_code[6][64];        // sample codebooks 1..6 by stored value: code word, length
    uint8_t cb_len[6][64];
    uint32_t dz_code[6];            // 'two zeros' codeword
    uint8_t dz_len[6];
    uint32_t hdr_code[31];          // header delta codes, delta + 16
    uint8_t hdr_len[31];
    uint32_t h93_code[2][32];       // 1993b type-1 band-type delta codes, [0] same subtype / [1] subtype inverts, delta + 16 (:2058-2127)
    uint8_t h93_len[2][32];
};
#if defined(__CUDACC__)
static __constant__ EncTables c_enc;
#endif
// the tables as the running pass sees them: constant memory on the device, the host copy elsewhere (the CPU simulator
// of tests/hostsim and nvcc's host pass)
static EncTables g_enc_host;
#if ENC_DEV
#define ENC_TAB c_enc
#else
#define ENC_TAB g_enc_host
#endif

ENC_HD int enc_band_count(int b) { return b == 0 ? 7 : (b == 1 ? 8 : (b == 15 ? 32 : 16)); }
ENC_HD int enc_band_first(int b) { return b == 0 ? 0 : (b == 1 ? 7 : 15 + 16 * (b - 2)); }
ENC_HD int enc_band_count_f(int fmt93, int b) { return fmt93 ? 16 : enc_band_count(b); }
ENC_HD float enc_half_sum(float a, float b) { return enc_mul(enc_add(a, b), 0.5f); }
ENC_HD float enc_half_diff(float a, float b) { return enc_mul(enc_sub(a, b), 0.5f); }

// ---------------------------------------------------------------------------------------------------------------
// K6a: one frame per thread.  f[frame][0..255] = the reference's Stream::Frame::f; power / lo / hi per band.
ENC_HD void dcsb_enc_transform_body(uint32_t tid, const float *__restrict__ pcm, const EncStream *__restrict__ streams, const uint32_t *__restrict__ frame_stream,
                          uint32_t n_frames_total, float *__restrict__ f_out, float *__restrict__ power, float *__restrict__ lo, float *__restrict__ hi)
{
    const uint32_t fr = tid;
    if (fr >= n_frames_total) return;
    const EncStream s = streams[frame_stream[fr]];
    const uint32_t k = fr - s.frame0;
    float in[256];
    float buf[258];
    // framing: 16 samples of overlap (raw, zero in front of the first frame), 240 new ones, zeros behind the end (:693-703, :724-737)
    const long long base = (long long)k * 240 - 16;
    for (int j = 0; j < 256; ++j) {
        const long long x = base + j;
        in[j] = (x >= 0 && (uint64_t)x < s.n_samples) ? pcm[s.pcm_off + (uint64_t)x] : 0.0f;
    }
    for (int i = 0; i < 16; ++i) {                                           // window (:1014-1018)
        in[i] = enc_mul(in[i], ENC_TAB.window[i]);
        in[255 - i] = enc_mul(in[255 - i], ENC_TAB.window[i]);
    }
    // DualFFT: bit-reversed load of (re, im) pairs, six radix-2 stages on 128 complex points (:1362-1471)
    for (int i = 0; i < 128; ++i) {
        const int idx = 2 * i;
        const int bi = enc_rev9(idx);
        buf[bi] = in[idx];
        buf[bi + 1] = in[idx + 1];
    }
    {
        int cp = 0;
        for (int st = 1; st <= 6; ++st) {
            const int m = 1 << st;
            for (int kk = 0; kk < 128; kk += m)
                for (int j = 0; j < m / 2; ++j) {
                    const float c = ENC_TAB.coeff[cp], sn = ENC_TAB.coeff[cp + 1];
                    cp += 2;
                    const int t = (kk + j + m / 2) * 2, u = (kk + j) * 2;
                    const float ar = buf[t], ai = buf[t + 1];
                    const float tr = enc_sub(enc_mul(ar, c), enc_mul(ai, sn));
                    const float ti = enc_add(enc_mul(ar, sn), enc_mul(ai, c));
                    const float ur = buf[u], ui = buf[u + 1];
                    buf[u] = enc_add(tr, ur);
                    buf[u + 1] = enc_add(ti, ui);
                    buf[t] = enc_sub(ur, tr);
                    buf[t + 1] = enc_sub(ui, ti);
                }
        }
        // the rotation of the seventh stage on the upper half only (:1473-1490), then the 1/64 scale
        cp = 896 - 126;
        for (int j = 1; j < 64; ++j) {
            const float c = ENC_TAB.coeff[cp], sn = ENC_TAB.coeff[cp + 1];
            cp += 2;
            const int t = 128 + 2 * j;
            const float ar = buf[t], ai = buf[t + 1];
            buf[t] = enc_sub(enc_mul(ar, c), enc_mul(ai, sn));
            buf[t + 1] = enc_add(enc_mul(ar, sn), enc_mul(ai, c));
        }
        for (int i = 0; i < 256; ++i) buf[i] = enc_mul(buf[i], 1.0f / 64.0f);
    }
    // DFTAlgorithmOrig: the decoder's pre-pass steps undone in float (:1220-1356)
    buf[1] = enc_half_sum(buf[0], buf[0x80]);
    buf[0x81] = buf[1];
    buf[0x100] = buf[1];
    buf[0x101] = buf[1];
    for (int i = 0; i < 64; ++i) {
        const int p0 = 2 * i, p1 = 0x80 + 2 * i;
        const float x0 = buf[p0], y0 = buf[p0 + 1], x1 = buf[p1], y1 = buf[p1 + 1];
        buf[p0] = enc_half_sum(x0, x1);
        buf[p0 + 1] = enc_half_sum(y0, y1);
        buf[p1] = enc_half_diff(x0, x1);
        buf[p1 + 1] = enc_half_diff(y0, y1);
    }
    for (int i = 0; i < 64; ++i) {
        const int p0 = 2 * i, p1 = 0x100 - 2 * i;
        const float x0 = buf[p0], y0 = buf[p0 + 1], x1 = buf[p1], y1 = buf[p1 + 1];
        const float xs = enc_half_diff(x0, x1), ys = enc_half_sum(y0, y1);
        const float c = ENC_TAB.tw2[2 * i], sn = ENC_TAB.tw2[2 * i + 1];
        buf[p0] = enc_half_sum(x0, x1);
        buf[p0 + 1] = enc_half_diff(y0, y1);
        buf[p1] = enc_sub(enc_mul(xs, sn), enc_mul(ys, c));
        buf[p1 + 1] = enc_add(enc_mul(xs, c), enc_mul(ys, sn));
    }
    for (int i = 0; i < 64; ++i) {
        const int p0 = 2 * i, p1 = 0x100 - 2 * i;
        const float x0 = -buf[p0], y0 = -buf[p0 + 1], x1 = -buf[p1], y1 = -buf[p1 + 1];
        buf[p0] = enc_half_sum(x0, x1);
        buf[p0 + 1] = enc_half_sum(y0, y1);
        buf[p1] = enc_half_diff(x0, x1);
        buf[p1 + 1] = enc_half_diff(y0, y1);
    }
    buf[0x80] = -buf[0x80];
    buf[0x81] = -buf[0x81];
    for (int i = 129; i < 256; i += 2) buf[i] = -buf[i];
    buf[1] = buf[0];                                                         // TransformFrame :1045
    // the frame as kept: fbuf[1..256]; band power and range (:2545-2564)
    float *fo = f_out + (size_t)fr * 256;
    for (int i = 0; i < 256; ++i) fo[i] = buf[1 + i];
    int p = 1;
    for (int b = 0; b < 16; ++b) {
        float l = buf[p], h = l, pw = enc_mul(l, l);
        ++p;
        for (int j = enc_band_count_f(s.fmt93, b); j > 1; --j) {
            const float v = buf[p++];
            pw = enc_add(pw, enc_mul(v, v));
            if (v < l) l = v;
            if (v > h) h = v;
        }
        power[(size_t)fr * 16 + b] = pw;
        lo[(size_t)fr * 16 + b] = l;
        hi[(size_t)fr * 16 + b] = h;
    }
}

// K6s: per stream and band, over the frames in order: power sum (float adds in frame order, :1056-1063), extremes
ENC_HD void dcsb_enc_stats_body(uint32_t tid, const EncStream *__restrict__ streams, int n, const float *__restrict__ power, const float *__restrict__ lo,
                                      const float *__restrict__ hi, float *__restrict__ out /* n x 48: power sum, lo, hi */)
{
    const int t = (int)tid;
    if (t >= n * 16) return;
    const int si = t >> 4, b = t & 15;
    const EncStream s = streams[si];
    float ps = 0.0f, l = 0.0f, h = 0.0f;
    for (uint32_t k = 0; k < s.n_frames; ++k) {
        const size_t i = (size_t)(s.frame0 + k) * 16 + b;
        ps = enc_add(ps, power[i]);
        const float fl = lo[i], fh = hi[i];
        if (k == 0 || fl < l) l = fl;
        if (k == 0 || fh > h) h = fh;
    }
    out[(size_t)si * 48 + b] = ps;
    out[(size_t)si * 48 + 16 + b] = l;
    out[(size_t)si * 48 + 32 + b] = h;
}

// what a band type code means for band `band` of a stream (InterpretBandTypeCode, :1840-1921): width, scale code
ENC_HD void enc_interpret(const EncStream &s, int band, int code, int padj, int &width, int &scale_code)
{
    const int sc = s.hdr[band] & 0x3F;
    if (s.type == 0) { width = code; scale_code = sc; return; }
    const uint32_t x = ENC_TAB.xlat[(band < 3 ? 0 : (band < 6 ? 16 : 32)) + code];
    width = (int)(x >> 8);
    scale_code = sc + (int)(x & 0xFF) + (band < 3 ? padj : 0);
}
ENC_HD float enc_scale(int scale_code) { return (float)ENC_TAB.scale[scale_code < 0 ? 0 : (scale_code > 63 ? 63 : scale_code)]; }
ENC_HD int enc_quant(float v, float scale) { return (int)roundf(enc_div(enc_mul(v, 32768.0f), scale)); }

// K6b: the search of FindBestBandEncoding for one band of one frame, tabulated for every scale pre-adjustment the band
// can meet (type-1 streams, bands 0..2) and for "code 15 allowed / not allowed" (the delta code reaches old + 14 only).
// best[frame][band][v][a]; 0 = the band's dynamic range is below the threshold (:1951-1955)
ENC_HD void dcsb_enc_search_body(uint32_t tid, const EncStream *__restrict__ streams, const uint32_t *__restrict__ frame_stream, uint32_t n_frames_total,
                       const float *__restrict__ f, const float *__restrict__ lo, const float *__restrict__ hi, uint8_t *__restrict__ best)
{
    // (band-major: the threads of a warp work on the SAME band of 32 consecutive frames -- same sample count, same
    // number of alternatives -- instead of on the 16 different bands of two frames: 6.9 -> ~30 active threads per instruction)
    const uint32_t t = tid;
    if (t >= n_frames_total * 16u) return;
    const uint32_t fr = t % n_frames_total;
    const int band = (int)(t / n_frames_total);
    const EncStream s = streams[frame_stream[fr]];
    uint8_t *bo = best + ((size_t)fr * 16 + band) * (ENC_NV * 2);
    for (int i = 0; i < ENC_NV * 2; ++i) bo[i] = 0;
    if (band >= s.bands || s.fmt93) return;
    if (enc_sub(hi[(size_t)fr * 16 + band], lo[(size_t)fr * 16 + band]) < s.min_range) return;
    const int n = enc_band_count(band);
    const float *x = f + (size_t)fr * 256 + enc_band_first(band);
    const float err_max = enc_mul(s.max_err2, (float)n);
    const int nv = (s.type != 0 && band < 3) ? (s.subtype == 0 ? 2 : ENC_NV) : 1;
    for (int v = 0; v < nv; ++v) {
        float err[16];
        int wid[16];
        bool pass[16];
        for (int code = 1; code <= 15; ++code) {
            int width, sc;
            enc_interpret(s, band, code, v, width, sc);
            const float scale = enc_scale(sc);
            const int ref = width != 0 ? 1 << (width - 1) : 0;
            const int mask = 0xFFFF >> (16 - width);
            float sum = 0.0f;
            for (int i = 0; i < n; ++i) {
                const float o = x[i];
                const int stored = (enc_quant(o, scale) + ref) & mask;
                const float rec = enc_div(enc_mul((float)(stored - ref), scale), 32768.0f);
                const float e = enc_sub(rec, o);
                sum = enc_add(sum, enc_mul(e, e));
            }
            err[code] = sum;
            wid[code] = width;
            pass[code] = sum <= err_max;
        }
        for (int a = 0; a < 2; ++a) {                   // FindBestResult (:1574-1621) over codes 1..15 / 1..14
            const int top = a ? 14 : 15;
            int narrow = -1;
            for (int c = 1; c <= top; ++c)
                if (pass[c] && (narrow == -1 || wid[c] < narrow)) narrow = wid[c];
            float min_err = -1.0f;
            int pick = 0;
            for (int c = 1; c <= top; ++c)
                if (narrow == -1 || wid[c] == narrow)
                    if (min_err < 0 || err[c] < min_err) { pick = c; min_err = err[c]; }
            bo[v * 2 + a] = (uint8_t)pick;
        }
    }
}

ENC_HD int enc_preadj(const EncStream &s, int old_code)       // preAdjMap0 / preAdjMap3 (:710-716)
{
    if (s.type == 0) return 0;
    if (s.subtype == 0) return old_code < 4 ? 0 : 1;
    return old_code < 4 ? 0 : (old_code > 7 ? 4 : old_code - 3);
}

// K6c: the band type codes of every frame, in order (the previous frame's code picks the alternative)
ENC_HD void dcsb_enc_resolve_body(uint32_t tid, const EncStream *__restrict__ streams, int n, const uint8_t *__restrict__ best,
                                        uint8_t *__restrict__ codes /* frame x 16 */, uint8_t *__restrict__ padj /* frame x 4 */)
{
    const int si = (int)tid;
    if (si >= n) return;
    const EncStream s = streams[si];
    if (s.fmt93) return;
    int old[16];
    for (int b = 0; b < 16; ++b) old[b] = 0;
    for (uint32_t k = 0; k < s.n_frames; ++k) {
        const size_t fr = (size_t)s.frame0 + k;
        int pa[3];
        for (int b = 0; b < 3; ++b) { pa[b] = enc_preadj(s, old[b]); padj[fr * 4 + b] = (uint8_t)pa[b]; }
        padj[fr * 4 + 3] = 0;
        for (int b = 0; b < 16; ++b) {
            int c = 0;
            if (b < s.bands) {
                const int v = b < 3 ? pa[b] : 0;
                c = best[(fr * 16 + b) * (ENC_NV * 2) + v * 2 + (old[b] == 0 ? 1 : 0)];
            }
            codes[fr * 16 + b] = (uint8_t)c;
            old[b] = c;
        }
    }
}

// One frame's bits (CompressFrame94 :1936-2048).  WRITE = false: count only.
struct EncBitSink {
    uint32_t *words;            // the stream's data words (big-endian byte order in memory)
    unsigned long long acc;
    int nacc;
    uint64_t widx;
    uint64_t count;
    template <bool WRITE> ENC_HD void put(uint32_t code, int len)
    {
        count += (uint64_t)len;
        if (!WRITE) return;
        acc = (acc << len) | (unsigned long long)code;
        nacc += len;
        while (nacc >= 32) {
            const uint32_t w = (uint32_t)(acc >> (nacc - 32));
            enc_or_word(words + widx, w);
            ++widx;
            nacc -= 32;
            acc &= (1ull << nacc) - 1ull;
        }
    }
    ENC_HD void flush()
    {
        if (nacc > 0) enc_or_word(words + widx, (uint32_t)(acc << (32 - nacc)));
    }
};

template <bool WRITE>
ENC_HD void dcsb_enc_emit_body(uint32_t tid, const EncStream *__restrict__ streams, const uint32_t *__restrict__ frame_stream, uint32_t n_frames_total,
                     const float *__restrict__ f, const uint8_t *__restrict__ codes, const uint8_t *__restrict__ padj,
                     uint32_t *__restrict__ frame_bits, const uint64_t *__restrict__ frame_pos, uint32_t *__restrict__ out_words,
                     const uint64_t *__restrict__ stream_word0)
{
    const uint32_t fr = tid;
    if (fr >= n_frames_total) return;
    const uint32_t si = frame_stream[fr];
    const EncStream s = streams[si];
    if (s.fmt93) return;
    EncBitSink sink;
    sink.count = 0;
    sink.acc = 0;
    sink.nacc = 0;
    sink.widx = 0;
    sink.words = nullptr;
    if (WRITE) {
        const uint64_t p = frame_pos[fr];
        sink.words = out_words + stream_word0[si];
        sink.widx = p >> 5;
        sink.nacc = (int)(p & 31);          // (leading zero bits: the words line up with the stream)
    }
    const bool first = fr == s.frame0;
    for (int b = 0; b < s.bands; ++b) {
        const int oldc = first ? 0 : codes[(size_t)(fr - 1) * 16 + b];
        const int d = (int)codes[(size_t)fr * 16 + b] - oldc + 16;
        sink.put<WRITE>(ENC_TAB.hdr_code[d], ENC_TAB.hdr_len[d]);
    }
    for (int b = 0; b < s.bands; ++b) {
        const int code = codes[(size_t)fr * 16 + b];
        if (code == 0) continue;
        int width, sc;
        enc_interpret(s, b, code, b < 3 ? padj[(size_t)fr * 4 + b] : 0, width, sc);
        if (width == 0) continue;
        const float scale = enc_scale(sc);
        const int mask = 0xFFFF >> (16 - width);
        const bool book = width <= 6;
        const int ref = book ? 1 << (width - 1) : 0;
        const int n = enc_band_count(b);
        const float *x = f + (size_t)fr * 256 + enc_band_first(b);
        int q[32];
        for (int i = 0; i < n; ++i) q[i] = enc_quant(x[i], scale);
        for (int i = 0; i < n; ++i) {
            if (book && q[i] == 0 && i + 1 < n && q[i + 1] == 0) {
                sink.put<WRITE>(ENC_TAB.dz_code[width - 1], ENC_TAB.dz_len[width - 1]);
                ++i;
            } else {
                const int v = (q[i] + ref) & mask;
                if (book) sink.put<WRITE>(ENC_TAB.cb_code[width - 1][v], ENC_TAB.cb_len[width - 1][v]);
                else sink.put<WRITE>((uint32_t)v, width);
            }
        }
    }
    if (WRITE) sink.flush();
    else frame_bits[fr] = (uint32_t)sink.count;
}

// ---- the 1993 layouts (CompressFrame93b, :2053-2473) -----------------------------------------------------------------
// The bands of a frame hang together: every band may be stored as values, first or (type 0) second differences against
// the samples before it, takes the narrowest of them, and says so relative to the band before.  Type 0 carries nothing
// from frame to frame, so a thread takes a frame and decides as it goes.  Type 1 delta-codes a band's type against the
// SAME band of the frame before (and the search's upper limit follows that old code), so its decisions are made by a
// thread per stream walking the frames (dcsb_enc_resolve93_kernel) from a table of the two possible search outcomes
// (dcsb_enc_search93_kernel), and the frame kernel only writes what was decided.
ENC_HD int enc93_count(const EncStream &s, int band) { return (s.type == 1 && band == 0) ? 15 : 16; }
ENC_HD int enc93_first(const EncStream &s, int band) { return s.type == 1 ? (band == 0 ? 0 : 16 * band - 1) : 16 * band; }

// FindBestBandEncoding for one band of a 1993-layout frame: codes 1..top, width = code + wadd
ENC_HD int enc93_search(const float *x, int n, float scale, float err_max, int wadd, int top)
{
    float err[16];
    bool pass[16];
    for (int code = 1; code <= 15; ++code) {
        const int width = code + wadd, ref = 1 << (width - 1), mask = 0xFFFF >> (16 - width);
        float sum = 0.0f;
        for (int i = 0; i < n; ++i) {
            const float o = x[i];
            const int stored = (enc_quant(o, scale) + ref) & mask;
            const float rec = enc_div(enc_mul((float)(stored - ref), scale), 32768.0f);
            const float e = enc_sub(rec, o);
            sum = enc_add(sum, enc_mul(e, e));
        }
        err[code] = sum;
        pass[code] = sum <= err_max;
    }
    int narrow = -1, pick = 0;
    for (int c = 1; c <= top; ++c)
        if (pass[c] && (narrow == -1 || c + wadd < narrow)) narrow = c + wadd;
    float min_err = -1.0f;
    for (int c = 1; c <= top; ++c)
        if (narrow == -1 || c + wadd == narrow)
            if (min_err < 0 || err[c] < min_err) { pick = c; min_err = err[c]; }
    return pick;
}
// the band type a run of differences needs (GetDeltaBandCode, :2224-2256)
ENC_HD int enc93_delta_code(const int *b, int n, int type)
{
    int lo = b[0], hi = b[0];
    for (int i = 1; i < n; ++i) { lo = b[i] < lo ? b[i] : lo; hi = b[i] > hi ? b[i] : hi; }
    if (hi < 0) hi = -hi;
    if (lo < 0) lo = -lo;
    if (lo > hi) hi = lo;
    if (hi == 0) return 0;
    int nb = 1;
    for (; hi != 0; hi >>= 1) ++nb;
    return nb - (type == 0 ? 1 : 0);
}

// type 1: the two outcomes of the search a band can have (code 15 within reach of the delta code or not)
ENC_HD void dcsb_enc_search93_body(uint32_t tid, const EncStream *__restrict__ streams, const uint32_t *__restrict__ frame_stream, uint32_t n_frames_total,
                         const float *__restrict__ f, uint8_t *__restrict__ best)
{
    const uint32_t t = tid;
    if (t >= n_frames_total * 16u) return;
    const uint32_t fr = t % n_frames_total;             // band-major, as in dcsb_enc_search_kernel
    const int band = (int)(t / n_frames_total);
    const EncStream s = streams[frame_stream[fr]];
    if (!s.fmt93 || s.type != 1 || band >= s.bands) return;
    const int n = enc93_count(s, band);
    const float *x = f + (size_t)fr * 256 + enc93_first(s, band);
    const float scale = enc_scale(s.hdr[band] & 0x3F), err_max = enc_mul(s.max_err2, (float)n);
    uint8_t *bo = best + ((size_t)fr * 16 + band) * (ENC_NV * 2);
    bo[0] = (uint8_t)enc93_search(x, n, scale, err_max, 0, 15);
    bo[1] = (uint8_t)enc93_search(x, n, scale, err_max, 0, 14);
}

// type 1: decisions of every band of every frame, in order.  dec[frame][band] = code | subtype << 4 | "same again" << 5 |
// (delta + 16) << 8
ENC_HD void dcsb_enc_resolve93_body(uint32_t tid, const EncStream *__restrict__ streams, int n, const float *__restrict__ f,
                                          const uint8_t *__restrict__ best, uint16_t *__restrict__ dec)
{
    const int si = (int)tid;
    if (si >= n) return;
    const EncStream s = streams[si];
    if (!s.fmt93 || s.type != 1) return;
    int old[16];
    for (int b = 0; b < 16; ++b) old[b] = 0;
    for (uint32_t k = 0; k < s.n_frames; ++k) {
        const size_t fr = (size_t)s.frame0 + k;
        int last_code = -1, last_sub = 0, prv = 0;
        for (int band = 0; band < s.bands; ++band) {
            const int cnt = enc93_count(s, band);
            const float *x = f + fr * 256 + enc93_first(s, band);
            const float scale = enc_scale(s.hdr[band] & 0x3F);
            const int band_prv = prv;
            int b1[16];
            for (int i = 0; i < cnt; ++i) {
                const int cur = enc_quant(x[i], scale);
                b1[i] = cur - prv;
                prv = cur;
            }
            // values: the search reaches old + 14 when the subtype stays 0, old + 15 when it changes to 0 (:2151-2172)
            const int top_is_14 = (last_sub == 0 && old[band] == 0) ? 1 : 0;
            int code = best[(fr * 16 + band) * (ENC_NV * 2) + top_is_14], sub = 0;
            const int c1 = enc93_delta_code(b1, cnt, 1);
            if (c1 < code || (c1 == code && last_sub == 1)) { sub = 1; code = c1; }
            uint16_t d;
            if (last_code == 0 && code == 0 && last_sub == sub) d = (uint16_t)(code | (sub << 4) | 0x20);
            else {
                int delta = code - old[band] + 16;
                delta = delta < 0 ? 0 : (delta > 31 ? 31 : delta);
                d = (uint16_t)(code | (sub << 4) | (delta << 8));
                old[band] = code;
                if (code == 0) prv = sub == 0 ? 0 : band_prv;
            }
            dec[fr * 16 + band] = d;
            last_code = code;
            last_sub = sub;
        }
    }
}

template <bool WRITE>
ENC_HD void dcsb_enc_frame93_body(uint32_t tid, const EncStream *__restrict__ streams, const uint32_t *__restrict__ frame_stream, uint32_t n_frames_total,
                        const float *__restrict__ f, const uint16_t *__restrict__ dec, uint32_t *__restrict__ frame_bits,
                        const uint64_t *__restrict__ frame_pos, uint32_t *__restrict__ out_words, const uint64_t *__restrict__ stream_word0)
{
    const uint32_t fr = tid;
    if (fr >= n_frames_total) return;
    const uint32_t si = frame_stream[fr];
    const EncStream s = streams[si];
    if (!s.fmt93) return;
    EncBitSink sink;
    sink.count = 0;
    sink.acc = 0;
    sink.nacc = 0;
    sink.widx = 0;
    sink.words = nullptr;
    if (WRITE) {
        const uint64_t p = frame_pos[fr];
        sink.words = out_words + stream_word0[si];
        sink.widx = p >> 5;
        sink.nacc = (int)(p & 31);
    }
    const int type = s.type;
    int last_code = -1, last_sub = type == 1 ? 0 : 2, prv = 0, prvd = 0;
    for (int band = 0; band < s.bands; ++band) {
        const int cnt = enc93_count(s, band);
        const float *x = f + (size_t)fr * 256 + enc93_first(s, band);
        const float scale = enc_scale(s.hdr[band] & 0x3F);
        const int band_prv = prv, band_prvd = prvd;
        int b0[16], b1[16], b2[16];
        for (int i = 0; i < cnt; ++i) {
            const int cur = enc_quant(x[i], scale);
            b0[i] = cur;
            b1[i] = cur - prv;
            b2[i] = cur - prv - prvd;
            prvd = b1[i];
            prv = cur;
        }
        int code, sub, hidx = 0;
        bool same;
        if (type == 0) {
            code = enc93_search(x, cnt, scale, enc_mul(s.max_err2, (float)cnt), 1, 15);
            sub = 0;
            const int c1 = enc93_delta_code(b1, cnt, 0), c2 = enc93_delta_code(b2, cnt, 0);
            if (c1 < code || (c1 == code && last_sub == 1)) { sub = 1; code = c1; }
            if (c2 < code) { sub = 2; code = c2; }
            same = last_code == 0 && code == 0 && last_sub == sub;
        } else {
            const uint16_t d = dec[(size_t)fr * 16 + band];
            code = d & 15;
            sub = (d >> 4) & 1;
            same = (d & 0x20) != 0;
            hidx = d >> 8;
        }
        if (same) {
            sink.put<WRITE>(1u, 1);                             // "the same again" (:2283-2288)
        } else {
            if (last_code == 0) sink.put<WRITE>(0u, 1);
            if (type == 0) {
                if (sub == last_sub) sink.put<WRITE>(0u, 1);
                else {
                    sink.put<WRITE>(1u, 1);
                    sink.put<WRITE>((uint32_t)(((sub - last_sub + 3) % 3) == 1 ? 1 : 0), 1);     // up / down modulo 3 (:2307-2311)
                }
                sink.put<WRITE>((uint32_t)code, 4);
            } else {
                const int inv = sub == last_sub ? 0 : 1;
                sink.put<WRITE>(ENC_TAB.h93_code[inv][hidx], ENC_TAB.h93_len[inv][hidx]);
            }
            if (code == 0) {
                if (sub == 0) { prv = 0; prvd = 0; }
                else if (sub == 1) { prv = band_prv; prvd = 0; }
                else { prv = band_prv; prvd = band_prvd; }
            } else {
                const int nb = code + (type == 0 ? 1 : 0), mask = (1 << nb) - 1;
                const int *b = sub == 0 ? b0 : (sub == 1 ? b1 : b2);
                for (int i = 0; i < cnt; ++i) sink.put<WRITE>((uint32_t)(b[i] & mask), nb);
            }
        }
        last_code = code;
        last_sub = sub;
    }
    if (WRITE) sink.flush();
    else frame_bits[fr] = (uint32_t)sink.count;
}

// per stream: bit position of every frame, total bits
ENC_HD void dcsb_enc_scan_body(uint32_t tid, const EncStream *__restrict__ streams, int n, const uint32_t *__restrict__ frame_bits,
                                     uint64_t *__restrict__ frame_pos, uint64_t *__restrict__ stream_bits)
{
    const int si = (int)tid;
    if (si >= n) return;
    const EncStream s = streams[si];
    uint64_t p = 0;
    for (uint32_t k = 0; k < s.n_frames; ++k) {
        frame_pos[s.frame0 + k] = p;
        p += frame_bits[s.frame0 + k];
    }
    stream_bits[si] = p;
}

// ---------------------------------------------------------------------------------------------------------------
// host side
inline void enc_build_tables(EncTables *t)
{
    // DualFFT twiddles exactly as the reference makes them: float theta, cosf / sinf (:1384-1398)
    {
        const float PI = 3.1415926536f;
        float *cp = t->coeff;
        for (int s = 1; s <= 7; ++s) {
            const int m = 1 << s;
            for (int k = 0; k < 128; k += m)
                for (int j = 0; j < m / 2; ++j) {
                    const float theta = -2 * PI * static_cast<float>(j) / static_cast<float>(m);
                    *cp++ = cosf(theta);
                    *cp++ = sinf(theta);
                }
        }
    }
    // the second table is the decoder's 1.15 twiddles printed with seven decimals (:1263-1280): -cos, -sin of i pi / 128
    for (int i = 0; i < 64; ++i) {
        const double th = 3.14159265358979323846 * i / 128.0;
        const long c = lround(cos(th) * 32768.0), s = lround(sin(th) * 32768.0);
        char tmp[32];
        snprintf(tmp, sizeof(tmp), "%.7f", -(double)c / 32768.0);
        t->tw2[2 * i] = strtof(tmp, nullptr);
        snprintf(tmp, sizeof(tmp), "%.7f", -(double)s / 32768.0);
        t->tw2[2 * i + 1] = strtof(tmp, nullptr);
    }
    static const float window[16] = { 0.010179f, 0.040507f, 0.090368f, 0.158746f, 0.244250f, 0.345139f, 0.459359f, 0.584585f,
                                      0.647178f, 0.752018f, 0.829799f, 0.888221f, 0.932184f, 0.964581f, 0.986700f, 0.998439f };
    memcpy(t->window, window, sizeof(window));
    for (int j = 0; j < 64; ++j) {
        const uint32_t m = (j & 2) ? ((j & 1) ? 0xd745u : 0xb505u) : ((j & 1) ? 0x9838u : 0x8000u);
        t->scale[j] = (int)(m >> (15 - ((j >> 2) & 15)));
    }
    static const uint16_t xl[48] = {
        0x0000, 0x0100, 0x0200, 0x0300, 0x0400, 0x0402, 0x0405, 0x0505, 0x0509, 0x050d, 0x060d, 0x0611, 0x0615, 0x0719, 0x071d, 0x081d,
        0x0000, 0x0100, 0x0200, 0x0300, 0x0400, 0x0402, 0x0407, 0x040b, 0x050b, 0x050f, 0x0513, 0x0517, 0x0617, 0x061b, 0x061f, 0x071f,
        0x0000, 0x0100, 0x0200, 0x0300, 0x0302, 0x0402, 0x0407, 0x040b, 0x050b, 0x050f, 0x0513, 0x0517, 0x0617, 0x061b, 0x061f, 0x0723 };
    memcpy(t->xlat, xl, sizeof(xl));
    const dcs_code_t *cbs[6] = { dcs94_cb1, dcs94_cb2, dcs94_cb3, dcs94_cb4, dcs94_cb5, dcs94_cb6 };
    const int ncb[6] = { 3, 5, 9, 17, 33, 65 };
    memset(t->cb_code, 0, sizeof(t->cb_code));
    memset(t->cb_len, 0, sizeof(t->cb_len));
    for (int k = 0; k < 6; ++k)
        for (int i = 0; i < ncb[k]; ++i) {
            const dcs_code_t &e = cbs[k][i];
            if (e.val == 0x80) { t->dz_code[k] = e.code; t->dz_len[k] = e.len; }
            else { t->cb_code[k][e.val & 63] = e.code; t->cb_len[k][e.val & 63] = e.len; }
        }
    memset(t->h93_code, 0, sizeof(t->h93_code));
    memset(t->h93_len, 0, sizeof(t->h93_len));
    for (int i = 0; i < 62; ++i) {
        const dcs_code_t &e = dcs93_hdr[i];
        const int inv = e.val >= 0x1E ? 1 : 0, d = (int)e.val - (inv ? 0x2E : 0x0F) + 16;
        if (d >= 0 && d < 32) { t->h93_code[inv][d] = e.code; t->h93_len[inv][d] = e.len; }
    }
    for (int i = 0; i < 31; ++i) {
        const dcs_code_t &e = dcs94_hdr[i];
        const int d = (int)e.val - 0x2E + 16;
        t->hdr_code[d] = e.code;
        t->hdr_len[d] = e.len;
    }
}

// CloseStream's power cut (:738-770) and CompressStream's header (:866-974) for one stream
inline void enc_stream_header(const float *stats /* 48 */, const dcsb_encode_params &pr, const EncTables &tab, EncStream *s)
{
    const bool f93 = s->fmt93 != 0;
    static const float norm94[16] = { 16.0f / 7, 16.0f / 8, 16.0f / 16, 16.0f / 16, 16.0f / 16, 16.0f / 16, 16.0f / 16, 16.0f / 16,
                                    16.0f / 16, 16.0f / 16, 16.0f / 16, 16.0f / 16, 16.0f / 16, 16.0f / 16, 16.0f / 16, 16.0f / 32 };
    static const int counts94[16] = { 7, 8, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 32 };
    float norm[16];
    int counts[16];
    for (int i = 0; i < 16; ++i) { norm[i] = f93 ? 1.0f : norm94[i]; counts[i] = f93 ? 16 : counts94[i]; }
    if (f93 && pr.stream_type == 1) counts[0] = 15;          // bandSampleCounts93b_Type1 (:53-55, :866-868)
    float rms[16], total = 0.0f;
    for (int i = 0; i < 16; ++i) {
        rms[i] = sqrtf(stats[i] * norm[i]);
        total += rms[i];
    }
    const float pn = 1.0f / total;
    int keep = 16;
    if (total != 0.0f) {
        float below = 0.0f;
        for (int i = 0; i < 16; ++i) {
            below += rms[i] * pn;
            if (below >= pr.power_band_cutoff) { keep = i; break; }
        }
    }
    const float fps = 31250.0f / 240.0f;
    const float bpf = static_cast<float>(pr.target_bit_rate) / fps;
    static const int share[16] = { 16, 14, 12, 10, 9, 8, 6, 5, 4, 4, 3, 3, 3, 3, 2, 2 };
    float share_norm = 0;
    for (int i = 0; i < keep; ++i) share_norm += static_cast<float>(share[i] * counts[i]);
    uint8_t *h = s->hdr;
    for (int band = 0; band < keep; ++band) {
        const int bits = static_cast<int>(static_cast<float>(share[band]) / share_norm * bpf);
        float lo = stats[16 + band] * -32768.0f, hi = stats[32 + band] * 32768.0f;
        if (lo < 0) lo = 0;
        if (hi < 0) hi = 0;
        const float full = hi > lo ? hi : lo;
        const int divider = 1 << bits;
        const int target = full != 0 ? static_cast<int>(ceil(full / divider)) : 1;
        h[band] = 0;
        for (int j = 0; j < 64; ++j) {
            if (tab.scale[j] < target) h[band] = (uint8_t)j;
            else break;
        }
        if (!f93 && pr.stream_type == 1) {
            int adjust = (band < 3) ? 0x0d : 0x17;
            adjust += pr.stream_subtype == 0 ? 1 : 3;
            if (h[band] > adjust) h[band] = (uint8_t)(h[band] - adjust);
            else h[band] = 0;
        }
    }
    for (int band = keep; band < 16; ++band) h[band] = 0xFF;
    if (pr.stream_type != 0) h[0] |= 0x80;
    h[1] |= (uint8_t)((pr.stream_subtype & 0x02) << 6);
    h[2] |= (uint8_t)((pr.stream_subtype & 0x01) << 7);
    // (the frame compressor stops at the first band whose low seven bits are all set, :1936)
    int bands = 0;
    while (bands < 16 && (h[bands] & 0x7F) != 0x7F) ++bands;
    s->bands = bands;
}

#endif
