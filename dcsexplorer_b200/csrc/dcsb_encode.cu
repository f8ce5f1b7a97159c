// dcsb200 forward path: PCM -> 1994-layout DCS streams, in batch, on the GPU.
//
// What it restates is the reference's DCSEncoder minus its resampler (DCSEncoder/DCSEncoder.cpp; "Enc.cpp" below):
// frames of 16 + 240 samples, window, the float transform DFTAlgorithmOrig (:1218-1357) over DualFFT (:1360-1499),
// per-band power / range (:2535-2565), the power cut and the header's scale codes (CloseStream :738-770,
// CompressStream :859-974), the per-band search for the narrowest band type within the quantisation-error limit
// (FindBestBandEncoding :1502-1621), header delta codes and sample codes with the 'two zeros' codeword
// (CompressFrame94 :1623-2051), the bit packing (BitWriter :2589-2705).  Every float operation is done in the
// reference's order with round-to-nearest single operations (no fused multiply-add), so the decisions -- and with them
// the stream BYTES -- are the reference's for the same frames (tests/test_gpu_encode.py compares with
// oracle/_ref's encoder fed the same framing).
//
// Kernels: K6a transform (a thread per frame), K6s per-stream band statistics, K6b band-type search (a thread per
// frame and band), K6c code resolution (a thread per stream: a band's choice depends on the previous frame's code only
// through two small rules, so K6b tabulates the alternatives and K6c picks), K6d frame sizes, K6e bit packing.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <thread>
#include <vector>
#include "dcsb_internal.h"
#include "dcsb_ctx.h"
#include "dcs_tables.h"

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
__constant__ EncTables c_enc;

__device__ __forceinline__ int enc_band_count(int b) { return b == 0 ? 7 : (b == 1 ? 8 : (b == 15 ? 32 : 16)); }
__device__ __forceinline__ int enc_band_first(int b) { return b == 0 ? 0 : (b == 1 ? 7 : 15 + 16 * (b - 2)); }
__device__ __forceinline__ int enc_band_count_f(int fmt93, int b) { return fmt93 ? 16 : enc_band_count(b); }
__device__ __forceinline__ float enc_half_sum(float a, float b) { return __fmul_rn(__fadd_rn(a, b), 0.5f); }
__device__ __forceinline__ float enc_half_diff(float a, float b) { return __fmul_rn(__fsub_rn(a, b), 0.5f); }

// ---------------------------------------------------------------------------------------------------------------
// K6a: one frame per thread.  f[frame][0..255] = the reference's Stream::Frame::f; power / lo / hi per band.
__global__ void __launch_bounds__(ENC_THREADS)
dcsb_enc_transform_kernel(const float *__restrict__ pcm, const EncStream *__restrict__ streams, const uint32_t *__restrict__ frame_stream,
                          uint32_t n_frames_total, float *__restrict__ f_out, float *__restrict__ power, float *__restrict__ lo, float *__restrict__ hi)
{
    const uint32_t fr = blockIdx.x * blockDim.x + threadIdx.x;
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
        in[i] = __fmul_rn(in[i], c_enc.window[i]);
        in[255 - i] = __fmul_rn(in[255 - i], c_enc.window[i]);
    }
    // DualFFT: bit-reversed load of (re, im) pairs, six radix-2 stages on 128 complex points (:1362-1471)
    for (int i = 0; i < 128; ++i) {
        const int idx = 2 * i;
        const int bi = (int)(__brev((unsigned)idx) >> 23);
        buf[bi] = in[idx];
        buf[bi + 1] = in[idx + 1];
    }
    {
        int cp = 0;
        for (int st = 1; st <= 6; ++st) {
            const int m = 1 << st;
            for (int kk = 0; kk < 128; kk += m)
                for (int j = 0; j < m / 2; ++j) {
                    const float c = c_enc.coeff[cp], sn = c_enc.coeff[cp + 1];
                    cp += 2;
                    const int t = (kk + j + m / 2) * 2, u = (kk + j) * 2;
                    const float ar = buf[t], ai = buf[t + 1];
                    const float tr = __fsub_rn(__fmul_rn(ar, c), __fmul_rn(ai, sn));
                    const float ti = __fadd_rn(__fmul_rn(ar, sn), __fmul_rn(ai, c));
                    const float ur = buf[u], ui = buf[u + 1];
                    buf[u] = __fadd_rn(tr, ur);
                    buf[u + 1] = __fadd_rn(ti, ui);
                    buf[t] = __fsub_rn(ur, tr);
                    buf[t + 1] = __fsub_rn(ui, ti);
                }
        }
        // the rotation of the seventh stage on the upper half only (:1473-1490), then the 1/64 scale
        cp = 896 - 126;
        for (int j = 1; j < 64; ++j) {
            const float c = c_enc.coeff[cp], sn = c_enc.coeff[cp + 1];
            cp += 2;
            const int t = 128 + 2 * j;
            const float ar = buf[t], ai = buf[t + 1];
            buf[t] = __fsub_rn(__fmul_rn(ar, c), __fmul_rn(ai, sn));
            buf[t + 1] = __fadd_rn(__fmul_rn(ar, sn), __fmul_rn(ai, c));
        }
        for (int i = 0; i < 256; ++i) buf[i] = __fmul_rn(buf[i], 1.0f / 64.0f);
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
        const float c = c_enc.tw2[2 * i], sn = c_enc.tw2[2 * i + 1];
        buf[p0] = enc_half_sum(x0, x1);
        buf[p0 + 1] = enc_half_diff(y0, y1);
        buf[p1] = __fsub_rn(__fmul_rn(xs, sn), __fmul_rn(ys, c));
        buf[p1 + 1] = __fadd_rn(__fmul_rn(xs, c), __fmul_rn(ys, sn));
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
        float l = buf[p], h = l, pw = __fmul_rn(l, l);
        ++p;
        for (int j = enc_band_count_f(s.fmt93, b); j > 1; --j) {
            const float v = buf[p++];
            pw = __fadd_rn(pw, __fmul_rn(v, v));
            if (v < l) l = v;
            if (v > h) h = v;
        }
        power[(size_t)fr * 16 + b] = pw;
        lo[(size_t)fr * 16 + b] = l;
        hi[(size_t)fr * 16 + b] = h;
    }
}

// K6s: per stream and band, over the frames in order: power sum (float adds in frame order, :1056-1063), extremes
__global__ void dcsb_enc_stats_kernel(const EncStream *__restrict__ streams, int n, const float *__restrict__ power, const float *__restrict__ lo,
                                      const float *__restrict__ hi, float *__restrict__ out /* n x 48: power sum, lo, hi */)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 16) return;
    const int si = t >> 4, b = t & 15;
    const EncStream s = streams[si];
    float ps = 0.0f, l = 0.0f, h = 0.0f;
    for (uint32_t k = 0; k < s.n_frames; ++k) {
        const size_t i = (size_t)(s.frame0 + k) * 16 + b;
        ps = __fadd_rn(ps, power[i]);
        const float fl = lo[i], fh = hi[i];
        if (k == 0 || fl < l) l = fl;
        if (k == 0 || fh > h) h = fh;
    }
    out[(size_t)si * 48 + b] = ps;
    out[(size_t)si * 48 + 16 + b] = l;
    out[(size_t)si * 48 + 32 + b] = h;
}

// what a band type code means for band `band` of a stream (InterpretBandTypeCode, :1840-1921): width, scale code
__device__ __forceinline__ void enc_interpret(const EncStream &s, int band, int code, int padj, int &width, int &scale_code)
{
    const int sc = s.hdr[band] & 0x3F;
    if (s.type == 0) { width = code; scale_code = sc; return; }
    const uint32_t x = c_enc.xlat[(band < 3 ? 0 : (band < 6 ? 16 : 32)) + code];
    width = (int)(x >> 8);
    scale_code = sc + (int)(x & 0xFF) + (band < 3 ? padj : 0);
}
__device__ __forceinline__ float enc_scale(int scale_code) { return (float)c_enc.scale[scale_code < 0 ? 0 : (scale_code > 63 ? 63 : scale_code)]; }
__device__ __forceinline__ int enc_quant(float v, float scale) { return (int)roundf(__fdiv_rn(__fmul_rn(v, 32768.0f), scale)); }

// K6b: the search of FindBestBandEncoding for one band of one frame, tabulated for every scale pre-adjustment the band
// can meet (type-1 streams, bands 0..2) and for "code 15 allowed / not allowed" (the delta code reaches old + 14 only).
// best[frame][band][v][a]; 0 = the band's dynamic range is below the threshold (:1951-1955)
__global__ void __launch_bounds__(ENC_THREADS)
dcsb_enc_search_kernel(const EncStream *__restrict__ streams, const uint32_t *__restrict__ frame_stream, uint32_t n_frames_total,
                       const float *__restrict__ f, const float *__restrict__ lo, const float *__restrict__ hi, uint8_t *__restrict__ best)
{
    // (band-major: the threads of a warp work on the SAME band of 32 consecutive frames -- same sample count, same
    // number of alternatives -- instead of on the 16 different bands of two frames: 6.9 -> ~30 active threads per instruction)
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_frames_total * 16u) return;
    const uint32_t fr = t % n_frames_total;
    const int band = (int)(t / n_frames_total);
    const EncStream s = streams[frame_stream[fr]];
    uint8_t *bo = best + ((size_t)fr * 16 + band) * (ENC_NV * 2);
    for (int i = 0; i < ENC_NV * 2; ++i) bo[i] = 0;
    if (band >= s.bands || s.fmt93) return;
    if (__fsub_rn(hi[(size_t)fr * 16 + band], lo[(size_t)fr * 16 + band]) < s.min_range) return;
    const int n = enc_band_count(band);
    const float *x = f + (size_t)fr * 256 + enc_band_first(band);
    const float err_max = __fmul_rn(s.max_err2, (float)n);
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
                const float rec = __fdiv_rn(__fmul_rn((float)(stored - ref), scale), 32768.0f);
                const float e = __fsub_rn(rec, o);
                sum = __fadd_rn(sum, __fmul_rn(e, e));
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

__device__ __forceinline__ int enc_preadj(const EncStream &s, int old_code)       // preAdjMap0 / preAdjMap3 (:710-716)
{
    if (s.type == 0) return 0;
    if (s.subtype == 0) return old_code < 4 ? 0 : 1;
    return old_code < 4 ? 0 : (old_code > 7 ? 4 : old_code - 3);
}

// K6c: the band type codes of every frame, in order (the previous frame's code picks the alternative)
__global__ void dcsb_enc_resolve_kernel(const EncStream *__restrict__ streams, int n, const uint8_t *__restrict__ best,
                                        uint8_t *__restrict__ codes /* frame x 16 */, uint8_t *__restrict__ padj /* frame x 4 */)
{
    const int si = blockIdx.x * blockDim.x + threadIdx.x;
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
    template <bool WRITE> __device__ __forceinline__ void put(uint32_t code, int len)
    {
        count += (uint64_t)len;
        if (!WRITE) return;
        acc = (acc << len) | (unsigned long long)code;
        nacc += len;
        while (nacc >= 32) {
            const uint32_t w = (uint32_t)(acc >> (nacc - 32));
            atomicOr(words + widx, __byte_perm(w, 0, 0x0123));
            ++widx;
            nacc -= 32;
            acc &= (1ull << nacc) - 1ull;
        }
    }
    __device__ __forceinline__ void flush()
    {
        if (nacc > 0) atomicOr(words + widx, __byte_perm((uint32_t)(acc << (32 - nacc)), 0, 0x0123));
    }
};

template <bool WRITE>
__global__ void __launch_bounds__(ENC_THREADS)
dcsb_enc_emit_kernel(const EncStream *__restrict__ streams, const uint32_t *__restrict__ frame_stream, uint32_t n_frames_total,
                     const float *__restrict__ f, const uint8_t *__restrict__ codes, const uint8_t *__restrict__ padj,
                     uint32_t *__restrict__ frame_bits, const uint64_t *__restrict__ frame_pos, uint32_t *__restrict__ out_words,
                     const uint64_t *__restrict__ stream_word0)
{
    const uint32_t fr = blockIdx.x * blockDim.x + threadIdx.x;
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
        sink.put<WRITE>(c_enc.hdr_code[d], c_enc.hdr_len[d]);
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
                sink.put<WRITE>(c_enc.dz_code[width - 1], c_enc.dz_len[width - 1]);
                ++i;
            } else {
                const int v = (q[i] + ref) & mask;
                if (book) sink.put<WRITE>(c_enc.cb_code[width - 1][v], c_enc.cb_len[width - 1][v]);
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
__device__ __forceinline__ int enc93_count(const EncStream &s, int band) { return (s.type == 1 && band == 0) ? 15 : 16; }
__device__ __forceinline__ int enc93_first(const EncStream &s, int band) { return s.type == 1 ? (band == 0 ? 0 : 16 * band - 1) : 16 * band; }

// FindBestBandEncoding for one band of a 1993-layout frame: codes 1..top, width = code + wadd
__device__ __forceinline__ int enc93_search(const float *x, int n, float scale, float err_max, int wadd, int top)
{
    float err[16];
    bool pass[16];
    for (int code = 1; code <= 15; ++code) {
        const int width = code + wadd, ref = 1 << (width - 1), mask = 0xFFFF >> (16 - width);
        float sum = 0.0f;
        for (int i = 0; i < n; ++i) {
            const float o = x[i];
            const int stored = (enc_quant(o, scale) + ref) & mask;
            const float rec = __fdiv_rn(__fmul_rn((float)(stored - ref), scale), 32768.0f);
            const float e = __fsub_rn(rec, o);
            sum = __fadd_rn(sum, __fmul_rn(e, e));
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
__device__ __forceinline__ int enc93_delta_code(const int *b, int n, int type)
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
__global__ void __launch_bounds__(ENC_THREADS)
dcsb_enc_search93_kernel(const EncStream *__restrict__ streams, const uint32_t *__restrict__ frame_stream, uint32_t n_frames_total,
                         const float *__restrict__ f, uint8_t *__restrict__ best)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_frames_total * 16u) return;
    const uint32_t fr = t % n_frames_total;             // band-major, as in dcsb_enc_search_kernel
    const int band = (int)(t / n_frames_total);
    const EncStream s = streams[frame_stream[fr]];
    if (!s.fmt93 || s.type != 1 || band >= s.bands) return;
    const int n = enc93_count(s, band);
    const float *x = f + (size_t)fr * 256 + enc93_first(s, band);
    const float scale = enc_scale(s.hdr[band] & 0x3F), err_max = __fmul_rn(s.max_err2, (float)n);
    uint8_t *bo = best + ((size_t)fr * 16 + band) * (ENC_NV * 2);
    bo[0] = (uint8_t)enc93_search(x, n, scale, err_max, 0, 15);
    bo[1] = (uint8_t)enc93_search(x, n, scale, err_max, 0, 14);
}

// type 1: decisions of every band of every frame, in order.  dec[frame][band] = code | subtype << 4 | "same again" << 5 |
// (delta + 16) << 8
__global__ void dcsb_enc_resolve93_kernel(const EncStream *__restrict__ streams, int n, const float *__restrict__ f,
                                          const uint8_t *__restrict__ best, uint16_t *__restrict__ dec)
{
    const int si = blockIdx.x * blockDim.x + threadIdx.x;
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
__global__ void __launch_bounds__(ENC_THREADS)
dcsb_enc_frame93_kernel(const EncStream *__restrict__ streams, const uint32_t *__restrict__ frame_stream, uint32_t n_frames_total,
                        const float *__restrict__ f, const uint16_t *__restrict__ dec, uint32_t *__restrict__ frame_bits,
                        const uint64_t *__restrict__ frame_pos, uint32_t *__restrict__ out_words, const uint64_t *__restrict__ stream_word0)
{
    const uint32_t fr = blockIdx.x * blockDim.x + threadIdx.x;
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
            code = enc93_search(x, cnt, scale, __fmul_rn(s.max_err2, (float)cnt), 1, 15);
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
                sink.put<WRITE>(c_enc.h93_code[inv][hidx], c_enc.h93_len[inv][hidx]);
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
__global__ void dcsb_enc_scan_kernel(const EncStream *__restrict__ streams, int n, const uint32_t *__restrict__ frame_bits,
                                     uint64_t *__restrict__ frame_pos, uint64_t *__restrict__ stream_bits)
{
    const int si = blockIdx.x * blockDim.x + threadIdx.x;
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
static void enc_build_tables(EncTables *t)
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
static void enc_stream_header(const float *stats /* 48 */, const dcsb_encode_params &pr, const EncTables &tab, EncStream *s)
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

// device / pinned buffers kept in the context between calls (no allocation in the steady state)
struct EncCache {
    DcsbBuf pcm, f, power, lo, hi, stats, streams, frame_stream, frame_bits, words, best, codes, padj, frame_pos, stream_bits, word0, dec;
    DcsbBuf h_words, h_pcm;          // pinned staging: stream data on its way out, PCM on its way in
    bool tables_up = false;
};
static void enc_cache_free(void *p)
{
    EncCache *c = static_cast<EncCache *>(p);
    if (!c) return;
    for (DcsbBuf *b : { &c->pcm, &c->f, &c->power, &c->lo, &c->hi, &c->stats, &c->streams, &c->frame_stream, &c->frame_bits, &c->words,
                        &c->best, &c->codes, &c->padj, &c->frame_pos, &c->stream_bits, &c->word0, &c->dec }) b->release(false);
    c->h_words.release(true);
    c->h_pcm.release(true);
    delete c;
}

#define CKE(x, what) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = fail(ctx, DCSB_E_CUDA, what, e_); goto done; } } while (0)

extern "C" uint64_t dcsb_encode_bound(uint64_t n_samples)
{
    const uint64_t frames = (n_samples + 239) / 240;
    return 18 + frames * ((16 * 23 + 255 * 15 + 7) / 8 + 1) + 8;
}

// the call for explicit stream types; consecutive entries with the same pcm pointer and length share one upload
static int encode_impl(dcsb_ctx *ctx, const float *const *pcm, const uint64_t *n_samples, size_t n,
                       const dcsb_encode_params *params, uint8_t *out, uint64_t out_capacity, uint64_t *out_offsets,
                       float *frames_out)
{
    if (!ctx || (n && (!pcm || !n_samples || !params || !out || !out_offsets))) return fail(ctx, DCSB_E_ARG, "dcsb_encode_streams: bad argument");
    if (n == 0) return DCSB_OK;
    uint64_t total_samples = 0, total_frames = 0;
    bool any93 = false, any93t1 = false;
    std::vector<EncStream> hs(n);
    for (size_t i = 0; i < n; ++i) {
        const dcsb_encode_params &p = params[i];
        if (!pcm[i] || n_samples[i] == 0) return fail(ctx, DCSB_E_ARG, "dcsb_encode_streams: empty clip");
        if ((p.stream_type != 0 && p.stream_type != 1) || (p.stream_subtype != 0 && p.stream_subtype != 3) || p.target_bit_rate <= 0)
            return fail(ctx, DCSB_E_ARG, "dcsb_encode_streams: stream type must be 0 or 1, subtype 0 or 3, bit rate positive");
        const bool f93 = p.format_version == DCSB_OS93A || p.format_version == DCSB_OS93B;
        if (p.format_version != 0 && p.format_version != DCSB_OS94 && !f93)
            return fail(ctx, DCSB_E_ARG, "dcsb_encode_streams: format version must be 0 / $9400, $9301 or $9302");
        if (f93 && (p.stream_subtype != 0 || (p.stream_type == 1 && p.format_version != DCSB_OS93B)))
            return fail(ctx, DCSB_E_ARG, "dcsb_encode_streams: 1993 layouts: subtype 0; stream type 1 only with format version $9302");
        const uint64_t nf = (n_samples[i] + 239) / 240;
        if (nf > 65535) return fail(ctx, DCSB_E_ARG, "dcsb_encode_streams: a stream holds at most 65535 frames");
        EncStream &s = hs[i];
        memset(&s, 0, sizeof(s));
        const bool shared = i > 0 && pcm[i] == pcm[i - 1] && n_samples[i] == n_samples[i - 1];
        s.pcm_off = shared ? hs[i - 1].pcm_off : total_samples;
        s.n_samples = n_samples[i];
        s.frame0 = (uint32_t)total_frames;
        s.n_frames = (uint32_t)nf;
        s.type = p.stream_type;
        s.subtype = p.stream_subtype;
        s.max_err2 = p.max_quantization_error * p.max_quantization_error;
        s.min_range = p.min_dynamic_range;
        s.fmt93 = f93 ? 1 : 0;
        any93 = any93 || f93;
        any93t1 = any93t1 || (f93 && p.stream_type == 1);
        if (!shared) total_samples += n_samples[i];
        total_frames += nf;
    }
    if (total_frames >= 0x0FFFFFFFull) return fail(ctx, DCSB_E_ARG, "dcsb_encode_streams: more than 2^28 frames in one call");
    int rc = DCSB_OK;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, DCSB_E_CUDA, "cudaSetDevice");
    const bool trace = getenv("DCSB_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {            // DCSB_TRACE=1: host-side phases of the call
        if (!trace) return;
        cudaDeviceSynchronize();
        fprintf(stderr, "[dcsb trace] encode_streams %-28s at %8.3f ms\n", what,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };
    static EncTables tab;
    static bool tab_ready = false;
    if (!tab_ready) { enc_build_tables(&tab); tab_ready = true; }
    if (!ctx->encode_cache) { ctx->encode_cache = new EncCache(); ctx->encode_cache_free = enc_cache_free; }
    EncCache &ec = *static_cast<EncCache *>(ctx->encode_cache);
    std::vector<uint32_t> frame_stream(total_frames);
    std::vector<float> stats(n * 48);
    std::vector<uint64_t> sbits(n), word0(n + 1, 0);
    const uint32_t nfr = (uint32_t)total_frames;
    const unsigned gf = (nfr + ENC_THREADS - 1) / ENC_THREADS, gfb = (unsigned)(((uint64_t)nfr * 16 + ENC_THREADS - 1) / ENC_THREADS);
    const unsigned gs = (unsigned)((n + 63) / 64), gsb = (unsigned)((n * 16 + 63) / 64);
#define ENSE(buf, bytes, host, what) do { cudaError_t e_ = (buf).ensure((bytes), (host)); if (e_ != cudaSuccess) { rc = fail(ctx, DCSB_E_NOMEM, what, e_); goto done; } } while (0)
    ENSE(ec.pcm, total_samples * sizeof(float), false, "cudaMalloc(pcm)");
    ENSE(ec.streams, n * sizeof(EncStream), false, "cudaMalloc(streams)");
    ENSE(ec.frame_stream, total_frames * 4, false, "cudaMalloc(frame map)");
    ENSE(ec.f, total_frames * 256 * sizeof(float), false, "cudaMalloc(frames)");
    ENSE(ec.power, total_frames * 16 * sizeof(float), false, "cudaMalloc(power)");
    ENSE(ec.lo, total_frames * 16 * sizeof(float), false, "cudaMalloc(lo)");
    ENSE(ec.hi, total_frames * 16 * sizeof(float), false, "cudaMalloc(hi)");
    ENSE(ec.stats, n * 48 * sizeof(float), false, "cudaMalloc(stats)");
    ENSE(ec.best, total_frames * 16 * ENC_NV * 2, false, "cudaMalloc(search table)");
    ENSE(ec.codes, total_frames * 16, false, "cudaMalloc(codes)");
    ENSE(ec.padj, total_frames * 4, false, "cudaMalloc(pre-adjustments)");
    if (any93t1) ENSE(ec.dec, total_frames * 32, false, "cudaMalloc(band decisions)");
    ENSE(ec.frame_bits, total_frames * 4, false, "cudaMalloc(frame sizes)");
    ENSE(ec.frame_pos, total_frames * 8, false, "cudaMalloc(frame positions)");
    ENSE(ec.stream_bits, n * 8, false, "cudaMalloc(stream sizes)");
    ENSE(ec.word0, (n + 1) * 8, false, "cudaMalloc(stream offsets)");
    {
        float *d_pcm = static_cast<float *>(ec.pcm.p), *d_f = static_cast<float *>(ec.f.p), *d_power = static_cast<float *>(ec.power.p);
        float *d_lo = static_cast<float *>(ec.lo.p), *d_hi = static_cast<float *>(ec.hi.p), *d_stats = static_cast<float *>(ec.stats.p);
        EncStream *d_streams = static_cast<EncStream *>(ec.streams.p);
        uint32_t *d_frame_stream = static_cast<uint32_t *>(ec.frame_stream.p), *d_frame_bits = static_cast<uint32_t *>(ec.frame_bits.p);
        uint8_t *d_best = static_cast<uint8_t *>(ec.best.p), *d_codes = static_cast<uint8_t *>(ec.codes.p), *d_padj = static_cast<uint8_t *>(ec.padj.p);
        uint64_t *d_frame_pos = static_cast<uint64_t *>(ec.frame_pos.p), *d_stream_bits = static_cast<uint64_t *>(ec.stream_bits.p);
        uint64_t *d_word0 = static_cast<uint64_t *>(ec.word0.p);
        for (size_t i = 0; i < n; ++i)
            for (uint32_t k = 0; k < hs[i].n_frames; ++k) frame_stream[hs[i].frame0 + k] = (uint32_t)i;
        if (!ec.tables_up) { CKE(cudaMemcpyToSymbol(c_enc, &tab, sizeof(tab)), "H2D encoder tables"); ec.tables_up = true; }
        // PCM through two pinned staging halves: the host copies clip data into one while the other is on the link
        {
            const size_t half = (size_t)8 << 20;                // samples per half (32 MB)
            ENSE(ec.h_pcm, 2 * half * sizeof(float), true, "cudaMallocHost(pcm staging)");
            float *hp = static_cast<float *>(ec.h_pcm.p);
            cudaEvent_t evh[2];
            CKE(cudaEventCreateWithFlags(&evh[0], cudaEventDisableTiming), "cudaEventCreate");
            CKE(cudaEventCreateWithFlags(&evh[1], cudaEventDisableTiming), "cudaEventCreate");
            uint64_t done_samples = 0;
            size_t ci = 0, co = 0;                               // clip, offset inside it
            int hb = 0;
            bool used[2] = { false, false };
            cudaError_t ee = cudaSuccess;
            while (done_samples < total_samples && ee == cudaSuccess) {
                if (used[hb]) ee = cudaEventSynchronize(evh[hb]);
                size_t fill = 0;
                struct Seg { float *dst; const float *src; size_t n; };
                std::vector<Seg> segs;
                while (fill < half && ci < n) {
                    if (co == 0 && ci > 0 && hs[ci].pcm_off == hs[ci - 1].pcm_off) { ++ci; continue; }      // shares the previous clip's samples
                    const size_t take = std::min<size_t>(half - fill, (size_t)(n_samples[ci] - co));
                    // (pieces of at most 1 M samples, so that the copy threads below share the work evenly)
                    for (size_t o = 0; o < take; o += (size_t)1 << 20)
                        segs.push_back(Seg{ hp + hb * half + fill + o, pcm[ci] + co + o, std::min<size_t>((size_t)1 << 20, take - o) });
                    fill += take;
                    co += take;
                    if (co == n_samples[ci]) { ++ci; co = 0; }
                }
                {
                    // pageable -> pinned is a plain memcpy and one core does ~9 GB/s of it: four threads keep the link busier
                    const unsigned nth = (unsigned)std::min<size_t>(4, std::max<size_t>(1, fill >> 21));
                    if (nth <= 1) for (const Seg &g : segs) memcpy(g.dst, g.src, g.n * sizeof(float));
                    else {
                        std::vector<std::thread> th;
                        for (unsigned j = 0; j < nth; ++j)
                            th.emplace_back([&segs, j, nth] { for (size_t k = j; k < segs.size(); k += nth) memcpy(segs[k].dst, segs[k].src, segs[k].n * sizeof(float)); });
                        for (auto &x : th) x.join();
                    }
                }
                if (ee == cudaSuccess) ee = cudaMemcpyAsync(d_pcm + done_samples, hp + hb * half, fill * sizeof(float), cudaMemcpyHostToDevice, 0);
                if (ee == cudaSuccess) ee = cudaEventRecord(evh[hb], 0);
                used[hb] = true;
                done_samples += fill;
                hb ^= 1;
            }
            if (ee == cudaSuccess) ee = cudaStreamSynchronize(0);
            cudaEventDestroy(evh[0]);
            cudaEventDestroy(evh[1]);
            CKE(ee, "H2D pcm");
        }
        lap("pcm uploaded");
        CKE(cudaMemcpy(d_streams, hs.data(), n * sizeof(EncStream), cudaMemcpyHostToDevice), "H2D streams");
        CKE(cudaMemcpy(d_frame_stream, frame_stream.data(), total_frames * 4, cudaMemcpyHostToDevice), "H2D frame map");
        dcsb_enc_transform_kernel<<<gf, ENC_THREADS>>>(d_pcm, d_streams, d_frame_stream, nfr, d_f, d_power, d_lo, d_hi);
        dcsb_enc_stats_kernel<<<gsb, 64>>>(d_streams, (int)n, d_power, d_lo, d_hi, d_stats);
        CKE(cudaGetLastError(), "encoder kernel launch");
        CKE(cudaMemcpy(stats.data(), d_stats, n * 48 * sizeof(float), cudaMemcpyDeviceToHost), "D2H band statistics");
        lap("transform + statistics");
        if (frames_out) CKE(cudaMemcpy(frames_out, d_f, total_frames * 256 * sizeof(float), cudaMemcpyDeviceToHost), "D2H frames");
        for (size_t i = 0; i < n; ++i) enc_stream_header(&stats[i * 48], params[i], tab, &hs[i]);
        CKE(cudaMemcpy(d_streams, hs.data(), n * sizeof(EncStream), cudaMemcpyHostToDevice), "H2D streams");
        dcsb_enc_search_kernel<<<gfb, ENC_THREADS>>>(d_streams, d_frame_stream, nfr, d_f, d_lo, d_hi, d_best);
        dcsb_enc_resolve_kernel<<<gs, 64>>>(d_streams, (int)n, d_best, d_codes, d_padj);
        dcsb_enc_emit_kernel<false><<<gf, ENC_THREADS>>>(d_streams, d_frame_stream, nfr, d_f, d_codes, d_padj, d_frame_bits, nullptr, nullptr, nullptr);
        if (any93t1) {
            dcsb_enc_search93_kernel<<<gfb, ENC_THREADS>>>(d_streams, d_frame_stream, nfr, d_f, d_best);
            dcsb_enc_resolve93_kernel<<<gs, 64>>>(d_streams, (int)n, d_f, d_best, static_cast<uint16_t *>(ec.dec.p));
        }
        if (any93) dcsb_enc_frame93_kernel<false><<<gf, ENC_THREADS>>>(d_streams, d_frame_stream, nfr, d_f, static_cast<const uint16_t *>(ec.dec.p), d_frame_bits, nullptr, nullptr, nullptr);
        dcsb_enc_scan_kernel<<<gs, 64>>>(d_streams, (int)n, d_frame_bits, d_frame_pos, d_stream_bits);
        CKE(cudaGetLastError(), "encoder kernel launch");
        CKE(cudaMemcpy(sbits.data(), d_stream_bits, n * 8, cudaMemcpyDeviceToHost), "D2H stream sizes");
        lap("search + resolve + sizes");
        {
            uint64_t need = 0;
            for (size_t i = 0; i < n; ++i) {
                word0[i + 1] = word0[i] + (sbits[i] + 31) / 32 + 1;
                need += 18 + (sbits[i] + 7) / 8;
            }
            if (need > out_capacity) { rc = fail(ctx, DCSB_E_NOMEM, "dcsb_encode_streams: output buffer too small (see dcsb_encode_bound)"); goto done; }
        }
        ENSE(ec.words, word0[n] * 4, false, "cudaMalloc(stream data)");
        ENSE(ec.h_words, word0[n] * 4, true, "cudaMallocHost(stream data)");
        uint32_t *d_words = static_cast<uint32_t *>(ec.words.p);
        CKE(cudaMemset(d_words, 0, word0[n] * 4), "memset stream data");
        CKE(cudaMemcpy(d_word0, word0.data(), (n + 1) * 8, cudaMemcpyHostToDevice), "H2D stream offsets");
        dcsb_enc_emit_kernel<true><<<gf, ENC_THREADS>>>(d_streams, d_frame_stream, nfr, d_f, d_codes, d_padj, nullptr, d_frame_pos, d_words, d_word0);
        if (any93) dcsb_enc_frame93_kernel<true><<<gf, ENC_THREADS>>>(d_streams, d_frame_stream, nfr, d_f, static_cast<const uint16_t *>(ec.dec.p), nullptr, d_frame_pos, d_words, d_word0);
        CKE(cudaGetLastError(), "encoder kernel launch");
        lap("packed");
        CKE(cudaMemcpy(ec.h_words.p, d_words, word0[n] * 4, cudaMemcpyDeviceToHost), "D2H stream data");
        {
            const uint32_t *words = static_cast<const uint32_t *>(ec.h_words.p);
            uint64_t o = 0;
            for (size_t i = 0; i < n; ++i) { out_offsets[i] = o; o += 18 + (sbits[i] + 7) / 8; }
            out_offsets[n] = o;
            auto store = [&](size_t i) {                      // BitWriter::Store (:2665-2703): frame count, header, data
                uint8_t *q = out + out_offsets[i];
                q[0] = (uint8_t)(hs[i].n_frames >> 8);
                q[1] = (uint8_t)(hs[i].n_frames & 0xFF);
                memcpy(q + 2, hs[i].hdr, 16);
                memcpy(q + 18, reinterpret_cast<const uint8_t *>(words + word0[i]), (sbits[i] + 7) / 8);
            };
            const unsigned nth = (unsigned)std::min<uint64_t>(4, std::max<uint64_t>(1, o >> 22));
            if (nth <= 1) for (size_t i = 0; i < n; ++i) store(i);
            else {
                std::vector<std::thread> th;
                for (unsigned j = 0; j < nth; ++j) th.emplace_back([&, j] { for (size_t i = j; i < n; i += nth) store(i); });
                for (auto &x : th) x.join();
            }
        }
        lap("streams stored");
    }
done:
#undef ENSE
    return rc;
}

// Public entry: explicit stream types go straight through; -1 in stream_type / stream_subtype is the reference's
// wildcard (CloseStream :779-836): the clip is encoded with every matching format of {0.0, 0.3, 1.0, 1.3} in that
// order -- one upload of its samples -- and the first of the smallest streams is kept.
extern "C" int dcsb_encode_streams(dcsb_ctx *ctx, const float *const *pcm, const uint64_t *n_samples, size_t n,
                                   const dcsb_encode_params *params, uint8_t *out, uint64_t out_capacity, uint64_t *out_offsets,
                                   float *frames_out)
{
    if (!ctx || (n && (!pcm || !n_samples || !params || !out || !out_offsets))) return fail(ctx, DCSB_E_ARG, "dcsb_encode_streams: bad argument");
    bool wild = false;
    for (size_t i = 0; i < n; ++i) {
        const dcsb_encode_params &p = params[i];
        if (p.stream_type < -1 || p.stream_type > 1 || (p.stream_subtype != -1 && p.stream_subtype != 0 && p.stream_subtype != 3))
            return fail(ctx, DCSB_E_ARG, "dcsb_encode_streams: stream type must be -1, 0 or 1, subtype -1, 0 or 3");
        wild = wild || p.stream_type < 0 || p.stream_subtype < 0;
    }
    if (!wild) return encode_impl(ctx, pcm, n_samples, n, params, out, out_capacity, out_offsets, frames_out);
    if (frames_out) return fail(ctx, DCSB_E_ARG, "dcsb_encode_streams: frames_out needs explicit stream types");
    static const int formats[4][2] = { { 0, 0 }, { 0, 3 }, { 1, 0 }, { 1, 3 } };
    std::vector<const float *> jp;
    std::vector<uint64_t> jn;
    std::vector<dcsb_encode_params> jpar;
    std::vector<size_t> first(n + 1, 0);
    uint64_t cap = 0;
    for (size_t i = 0; i < n; ++i) {
        first[i] = jp.size();
        for (const auto &f : formats) {
            const dcsb_encode_params &p = params[i];
            const bool f93 = p.format_version == DCSB_OS93A || p.format_version == DCSB_OS93B;
            if (f93 && (f[1] != 0 || (f[0] == 1 && p.format_version != DCSB_OS93B))) continue;       // (no subtypes there; no encoder for OS93a type 1, :798-815)
            if ((p.stream_type >= 0 && p.stream_type != f[0]) || (!f93 && p.stream_subtype >= 0 && p.stream_subtype != f[1])) continue;
            dcsb_encode_params q = p;
            q.stream_type = f[0];
            q.stream_subtype = f[1];
            jp.push_back(pcm[i]);
            jn.push_back(n_samples[i]);
            jpar.push_back(q);
            cap += dcsb_encode_bound(n_samples[i]);
        }
    }
    first[n] = jp.size();
    std::vector<uint8_t> tmp(cap + 64);
    std::vector<uint64_t> joffs(jp.size() + 1, 0);
    const int rc = encode_impl(ctx, jp.data(), jn.data(), jp.size(), jpar.data(), tmp.data(), tmp.size(), joffs.data(), nullptr);
    if (rc != DCSB_OK) return rc;
    uint64_t o = 0;
    for (size_t i = 0; i < n; ++i) {
        size_t best = first[i];
        for (size_t j = first[i] + 1; j < first[i + 1]; ++j)
            if (joffs[j + 1] - joffs[j] < joffs[best + 1] - joffs[best]) best = j;
        const uint64_t nb = joffs[best + 1] - joffs[best];
        if (o + nb > out_capacity) return fail(ctx, DCSB_E_NOMEM, "dcsb_encode_streams: output buffer too small (see dcsb_encode_bound)");
        out_offsets[i] = o;
        memcpy(out + o, tmp.data() + joffs[best], nb);
        o += nb;
    }
    out_offsets[n] = o;
    return DCSB_OK;
}
