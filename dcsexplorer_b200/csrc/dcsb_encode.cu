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

#include "dcsb_encode.cuh"

#define ENC_TID (blockIdx.x * blockDim.x + threadIdx.x)
__global__ void __launch_bounds__(ENC_THREADS)
dcsb_enc_transform_kernel(const float *__restrict__ pcm, const EncStream *__restrict__ streams, const uint32_t *__restrict__ frame_stream,
                          uint32_t n_frames_total, float *__restrict__ f_out, float *__restrict__ power, float *__restrict__ lo, float *__restrict__ hi)
{
    dcsb_enc_transform_body(ENC_TID, pcm, streams, frame_stream, n_frames_total, f_out, power, lo, hi);
}
__global__ void dcsb_enc_stats_kernel(const EncStream *__restrict__ streams, int n, const float *__restrict__ power, const float *__restrict__ lo,
                                      const float *__restrict__ hi, float *__restrict__ out)
{
    dcsb_enc_stats_body(ENC_TID, streams, n, power, lo, hi, out);
}
__global__ void __launch_bounds__(ENC_THREADS)
dcsb_enc_search_kernel(const EncStream *__restrict__ streams, const uint32_t *__restrict__ frame_stream, uint32_t n_frames_total,
                       const float *__restrict__ f, const float *__restrict__ lo, const float *__restrict__ hi, uint8_t *__restrict__ best)
{
    dcsb_enc_search_body(ENC_TID, streams, frame_stream, n_frames_total, f, lo, hi, best);
}
__global__ void dcsb_enc_resolve_kernel(const EncStream *__restrict__ streams, int n, const uint8_t *__restrict__ best,
                                        uint8_t *__restrict__ codes, uint8_t *__restrict__ padj)
{
    dcsb_enc_resolve_body(ENC_TID, streams, n, best, codes, padj);
}
template <bool WRITE>
__global__ void __launch_bounds__(ENC_THREADS)
dcsb_enc_emit_kernel(const EncStream *__restrict__ streams, const uint32_t *__restrict__ frame_stream, uint32_t n_frames_total,
                     const float *__restrict__ f, const uint8_t *__restrict__ codes, const uint8_t *__restrict__ padj,
                     uint32_t *__restrict__ frame_bits, const uint64_t *__restrict__ frame_pos, uint32_t *__restrict__ out_words,
                     const uint64_t *__restrict__ stream_word0)
{
    dcsb_enc_emit_body<WRITE>(ENC_TID, streams, frame_stream, n_frames_total, f, codes, padj, frame_bits, frame_pos, out_words, stream_word0);
}
__global__ void __launch_bounds__(ENC_THREADS)
dcsb_enc_search93_kernel(const EncStream *__restrict__ streams, const uint32_t *__restrict__ frame_stream, uint32_t n_frames_total,
                         const float *__restrict__ f, uint8_t *__restrict__ best)
{
    dcsb_enc_search93_body(ENC_TID, streams, frame_stream, n_frames_total, f, best);
}
__global__ void dcsb_enc_resolve93_kernel(const EncStream *__restrict__ streams, int n, const float *__restrict__ f,
                                          const uint8_t *__restrict__ best, uint16_t *__restrict__ dec)
{
    dcsb_enc_resolve93_body(ENC_TID, streams, n, f, best, dec);
}
template <bool WRITE>
__global__ void __launch_bounds__(ENC_THREADS)
dcsb_enc_frame93_kernel(const EncStream *__restrict__ streams, const uint32_t *__restrict__ frame_stream, uint32_t n_frames_total,
                        const float *__restrict__ f, const uint16_t *__restrict__ dec, uint32_t *__restrict__ frame_bits,
                        const uint64_t *__restrict__ frame_pos, uint32_t *__restrict__ out_words, const uint64_t *__restrict__ stream_word0)
{
    dcsb_enc_frame93_body<WRITE>(ENC_TID, streams, frame_stream, n_frames_total, f, dec, frame_bits, frame_pos, out_words, stream_word0);
}
__global__ void dcsb_enc_scan_kernel(const EncStream *__restrict__ streams, int n, const uint32_t *__restrict__ frame_bits,
                                     uint64_t *__restrict__ frame_pos, uint64_t *__restrict__ stream_bits)
{
    dcsb_enc_scan_body(ENC_TID, streams, n, frame_bits, frame_pos, stream_bits);
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
