// dcsb200 internal: context / batch objects shared by the C-ABI translation units
// (dcsb_api.cu: streams and batches, dcsb_player.cu: ROM sets, players, timelines).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/dcsb200.h"
#include "dcsb_internal.h"

// ======================================================================================
// grow-only buffer (device or pinned host) owned by a context
struct DcsbBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes, bool host)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) { if (host) cudaFreeHost(p); else cudaFree(p); p = nullptr; cap = 0; }
        const size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = host ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
        if (e == cudaSuccess && !host) e = cudaMemset(p, 0, want);     // no buffer is ever read uninitialised
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release(bool host) { if (p) { if (host) cudaFreeHost(p); else cudaFree(p); } p = nullptr; cap = 0; }
};

// One pipeline lane of dcsb_decode_streams: a CUDA stream plus everything one chunk of
// streams needs, kept across calls so that the steady state does no allocation.
#define DCSB_MAX_LANES 8
#define DCSB_DEFAULT_LANES 6     // + the upload and the download stream = 8 streams: one per hardware queue of the device
struct DcsbLane {
    cudaStream_t st = nullptr, aux = nullptr;                // aux: the scan runs beside the decode
    cudaEvent_t ev_go = nullptr, ev_scan = nullptr;
    DcsbBuf d_progress, d_queue, d_order;
    DcsbBuf h_slab, h_res, h_meta;                           // pinned
    DcsbBuf d_slab, d_recs, d_tiles, d_bitpos, d_bt, d_hdrbits, d_status, d_nplay, d_endbits, d_stopband, d_csum, d_pcm;
    DcsbPrepared prep;
    std::vector<DcsbTile> slice_tiles;                       // time-sliced chunk: work items of every slice, slice after slice
    std::vector<cudaEvent_t> ev_slices;                      // slice k decoded (the PCM copy of the slice waits for it)
    std::vector<size_t> sl_off;                              // first work item of slice k in slice_tiles: 2 entries per slice (1994, 1993 family) + end
    uint32_t slice = 0, nslices = 1;                         // frames per time slice (0 = not sliced)
    std::vector<uint32_t> sl_bound;                          // slice k = output frames [sl_bound[k], sl_bound[k + 1])
    size_t first = 0, count = 0;
    uint64_t pcm_base = 0;                                   // sample offset of the chunk in the packed output
    bool direct_pcm = false;                                 // PCM is copied straight into the caller's (pinned) buffer
    bool copy2d = false;                                     // ... one strided copy per slice (else one batched copy of a piece per stream)
    std::vector<uint64_t> dst_off;                           // batched copies: sample offset of each stream in the caller's buffer
    std::vector<void *> cp_dst, cp_src;                      // operands of the batched copy being submitted
    std::vector<size_t> cp_size;
};

// DCSB_TRACE=1: device-side timeline of one dcsb_decode_streams call (CUDA events on the lanes'
// streams, printed to stderr after the call) -- a tuning aid, off by default
struct DcsbTrace {
    bool on = false;
    cudaEvent_t t0 = nullptr;
    struct Mark { int lane; int slice; const char *what; cudaEvent_t ev; };
    std::vector<Mark> marks;
    void mark(int lane, int slice, const char *what, cudaStream_t st)
    {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        marks.push_back(Mark{ lane, slice, what, e });
    }
    void dump()
    {
        if (!on) return;
        for (const Mark &m : marks) {
            float ms = 0;
            cudaEventElapsedTime(&ms, t0, m.ev);
            fprintf(stderr, "[dcsb trace] lane %d slice %2d %-10s %8.3f ms\n", m.lane, m.slice, m.what, ms);
            cudaEventDestroy(m.ev);
        }
        marks.clear();
        cudaEventDestroy(t0);
    }
};
struct dcsb_ctx {
    int device = 0;
    DcsbTables *d_tables = nullptr;
    cudaStream_t aux = nullptr;              // resident batches: stream the scan runs on beside the decode
    cudaStream_t up = nullptr, down = nullptr;   // dcsb_decode_streams: all uploads / all PCM downloads, in submission order
    bool overlap = true;                     // scan and decode kernels resident together (dcsb_set_overlap)
    int max_chunks = 0;                      // dcsb_set_pipeline: chunks of dcsb_decode_streams (0 = choose)
    int slice_frames = 0;                    // dcsb_set_pipeline: frames per time slice (0 = choose, < 0 = never slice)
    DcsbLane lanes[DCSB_MAX_LANES];
    DcsbTrace trace;
    // dcsb_render_timelines: stream, device buffers and pinned staging kept for the next call (dcsb_player.cu owns the type)
    void *timeline_cache = nullptr;
    void (*timeline_cache_free)(void *) = nullptr;
    void *encode_cache = nullptr;                // buffers of dcsb_encode_streams, kept between calls (dcsb_encode.cu)
    void (*encode_cache_free)(void *) = nullptr;
    std::string err;
};

struct dcsb_batch {
    dcsb_ctx *ctx = nullptr;
    size_t n = 0;
    std::vector<DcsbStreamRec> recs;
    std::vector<int32_t> host_status;        // host-side rejections (0 = let the scan decide)
    std::vector<DcsbTile> tiles;
    int ntiles94 = 0, ntiles93 = 0;
    uint64_t total_frames_in = 0;            // stream frames (checkpoint entries)
    uint64_t total_out_frames = 0;
    uint64_t compressed_bytes = 0;
    size_t slab_bytes = 0;
    // device
    uint8_t *d_slab = nullptr;
    DcsbStreamRec *d_recs = nullptr;
    DcsbTile *d_tiles = nullptr;
    uint32_t *d_order = nullptr;             // scan order (dcsb_prepare): the n94 streams of the 1994 layout first
    size_t n94 = 0;
    DcsbScanOut scan{};
    int16_t *d_pcm = nullptr;                // internal PCM buffer (lazy)
    unsigned long long *d_checksums = nullptr;
    int nqueue94 = 0;
    uint32_t *d_progress = nullptr;          // [n] scan progress + [1] resident scan CTAs + [2] queue tail / head
    unsigned long long *d_queue = nullptr;   // [nqueue94] ready queue scan -> decode
    cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };    // step start, scan end, decode end, decode start
    bool timed = false;
};

static inline int fail(dcsb_ctx *ctx, int code, const char *what, cudaError_t e = cudaSuccess)
{
    if (ctx) {
        char buf[512];
        if (e != cudaSuccess) snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
        else snprintf(buf, sizeof(buf), "%s", what);
        ctx->err = buf;
    }
    return code;
}
#define CK(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, DCSB_E_CUDA, what, e_); } while (0)


int dcsb_batch_create_impl(dcsb_ctx *ctx, const dcsb_stream_desc *descs, size_t n, const uint8_t *in_place_base,
                           size_t in_place_span, dcsb_batch **out);
