// dcsb200 C-ABI implementation: host-side control plane (stream validation, gain staging,
// slab packing, tiling) and kernel orchestration.  See include/dcsb200.h.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <thread>
#include <vector>
#include <algorithm>
#include "../../include/dcsb200.h"
#include "dcsb_internal.h"

// ======================================================================================
struct dcsb_ctx {
    int device = 0;
    DcsbTables *d_tables = nullptr;
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
    DcsbScanOut scan{};
    int16_t *d_pcm = nullptr;                // internal PCM buffer (lazy)
    unsigned long long *d_checksums = nullptr;
    cudaEvent_t ev[3] = { nullptr, nullptr, nullptr };
    bool timed = false;
};

static int fail(dcsb_ctx *ctx, int code, const char *what, cudaError_t e = cudaSuccess)
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

extern "C" const char *dcsb_version(void) { return "dcsb200 0.1 (sm_100a)"; }

extern "C" int dcsb_create(int device, dcsb_ctx **out)
{
    if (!out) return DCSB_E_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return DCSB_E_CUDA;   // no CPU fallback
    dcsb_ctx *ctx = new dcsb_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return DCSB_E_CUDA; }
    DcsbTables *h = new DcsbTables();
    dcsb_build_tables(h);
    e = cudaMalloc(&ctx->d_tables, sizeof(DcsbTables));
    if (e == cudaSuccess) e = cudaMemcpy(ctx->d_tables, h, sizeof(DcsbTables), cudaMemcpyHostToDevice);
    delete h;
    if (e != cudaSuccess) { cudaFree(ctx->d_tables); delete ctx; return DCSB_E_CUDA; }
    *out = ctx;
    return DCSB_OK;
}

extern "C" void dcsb_destroy(dcsb_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaFree(ctx->d_tables);
    delete ctx;
}

extern "C" const char *dcsb_last_error(const dcsb_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }

extern "C" void dcsb_batch_destroy(dcsb_batch *b)
{
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaFree(b->d_slab); cudaFree(b->d_recs); cudaFree(b->d_tiles);
    cudaFree(b->scan.bitpos); cudaFree(b->scan.bt); cudaFree(b->scan.status); cudaFree(b->scan.nplay);
    cudaFree(b->scan.endbits); cudaFree(b->scan.stopband);
    cudaFree(b->d_pcm); cudaFree(b->d_checksums);
    for (auto &e : b->ev) if (e) cudaEventDestroy(e);
    delete b;
}

extern "C" int dcsb_batch_create(dcsb_ctx *ctx, const dcsb_stream_desc *descs, size_t n, dcsb_batch **out)
{
    if (!ctx || !out || (!descs && n)) return fail(ctx, DCSB_E_ARG, "dcsb_batch_create: bad argument");
    *out = nullptr;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    dcsb_batch *b = new dcsb_batch();
    b->ctx = ctx;
    b->n = n;

    DcsbPrepared prep;
    {
        int rc = dcsb_prepare(descs, n, &prep);
        if (rc != DCSB_OK) { delete b; return fail(ctx, rc, "dcsb_batch_create: unknown os_version or batch too large"); }
    }
    b->recs = prep.recs;
    b->host_status = prep.host_status;
    b->tiles = prep.tiles;
    b->ntiles94 = prep.ntiles94;
    b->ntiles93 = prep.ntiles93;
    b->total_frames_in = prep.total_frames_in;
    b->total_out_frames = prep.total_out_frames;
    b->compressed_bytes = prep.compressed_bytes;
    b->slab_bytes = prep.slab_bytes;
    const uint64_t frames = b->total_frames_in;

    // pack the compressed slab in pinned memory (multi-threaded), one H2D copy
    uint8_t *h_slab = nullptr;
    cudaError_t e = cudaMallocHost(&h_slab, b->slab_bytes);
    if (e != cudaSuccess) { delete b; return fail(ctx, DCSB_E_NOMEM, "cudaMallocHost(slab)", e); }
    dcsb_pack_slab(descs, n, &prep, h_slab);
#define CKB(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cudaFreeHost(h_slab); dcsb_batch_destroy(b); return fail(ctx, DCSB_E_CUDA, what, e_); } } while (0)
    CKB(cudaMalloc(&b->d_slab, b->slab_bytes), "cudaMalloc(slab)");
    CKB(cudaMemcpy(b->d_slab, h_slab, b->slab_bytes, cudaMemcpyHostToDevice), "H2D slab");
    cudaFreeHost(h_slab);
    h_slab = nullptr;
    CKB(cudaMalloc(&b->d_recs, std::max<size_t>(1, n) * sizeof(DcsbStreamRec)), "cudaMalloc(recs)");
    CKB(cudaMemcpy(b->d_recs, b->recs.data(), n * sizeof(DcsbStreamRec), cudaMemcpyHostToDevice), "H2D recs");
    CKB(cudaMalloc(&b->d_tiles, std::max<size_t>(1, b->tiles.size()) * sizeof(DcsbTile)), "cudaMalloc(tiles)");
    CKB(cudaMemcpy(b->d_tiles, b->tiles.data(), b->tiles.size() * sizeof(DcsbTile), cudaMemcpyHostToDevice), "H2D tiles");
    CKB(cudaMalloc(&b->scan.bitpos, std::max<uint64_t>(1, frames) * sizeof(uint32_t)), "cudaMalloc(bitpos)");
    CKB(cudaMalloc(&b->scan.bt, std::max<uint64_t>(1, frames) * sizeof(uint2)), "cudaMalloc(bt)");
    CKB(cudaMalloc(&b->scan.status, std::max<size_t>(1, n) * sizeof(int32_t)), "cudaMalloc(status)");
    CKB(cudaMalloc(&b->scan.nplay, std::max<size_t>(1, n) * sizeof(uint32_t)), "cudaMalloc(nplay)");
    CKB(cudaMalloc(&b->scan.endbits, std::max<size_t>(1, n) * sizeof(uint32_t)), "cudaMalloc(endbits)");
    CKB(cudaMalloc(&b->scan.stopband, std::max<size_t>(1, n)), "cudaMalloc(stopband)");
    CKB(cudaMalloc(&b->d_checksums, std::max<size_t>(1, n) * sizeof(unsigned long long)), "cudaMalloc(checksums)");
    for (auto &ev : b->ev) CKB(cudaEventCreate(&ev), "cudaEventCreate");
#undef CKB
    *out = b;
    return DCSB_OK;
}

extern "C" uint64_t dcsb_batch_total_samples(const dcsb_batch *b) { return b ? b->total_out_frames * 240 : 0; }
extern "C" uint64_t dcsb_batch_total_frames(const dcsb_batch *b) { return b ? b->total_out_frames : 0; }
extern "C" uint64_t dcsb_batch_compressed_bytes(const dcsb_batch *b) { return b ? b->compressed_bytes : 0; }
extern "C" uint64_t dcsb_batch_pcm_offset(const dcsb_batch *b, size_t i) { return (b && i < b->n) ? b->recs[i].pcm_off : 0; }
extern "C" void *dcsb_batch_device_pcm(dcsb_batch *b) { return b ? b->d_pcm : nullptr; }
extern "C" int dcsb_batch_launches(const dcsb_batch *b)
{
    if (!b) return 0;
    return (b->n ? 1 : 0) + (b->ntiles94 ? 1 : 0) + (b->ntiles93 ? 1 : 0);   // scan + one decode launch per transform family
}

extern "C" int dcsb_batch_decode(dcsb_batch *b, void *d_pcm, void *cuda_stream)
{
    if (!b) return DCSB_E_ARG;
    dcsb_ctx *ctx = b->ctx;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int16_t *pcm = (int16_t *)d_pcm;
    if (!pcm) {
        if (!b->d_pcm) CK(cudaMalloc(&b->d_pcm, std::max<uint64_t>(2, b->total_out_frames * 480)), "cudaMalloc(pcm)");
        pcm = b->d_pcm;
    }
    if (b->n == 0) return DCSB_OK;
    CK(cudaMemsetAsync(b->d_checksums, 0, b->n * sizeof(unsigned long long), st), "memset checksums");
    CK(cudaEventRecord(b->ev[0], st), "event");
    CK(dcsb_launch_scan(b->d_slab, b->d_recs, (int)b->n, ctx->d_tables, b->scan, st), "scan kernel launch");
    CK(cudaEventRecord(b->ev[1], st), "event");
    CK(dcsb_launch_decode(b->d_slab, b->d_recs, b->d_tiles, b->ntiles94, b->ntiles93, ctx->d_tables, b->scan,
                          pcm, b->d_checksums, st), "decode kernel launch");
    CK(cudaEventRecord(b->ev[2], st), "event");
    b->timed = true;
    return DCSB_OK;
}

extern "C" float dcsb_batch_last_kernel_ms(dcsb_batch *b, int which)
{
    if (!b || !b->timed || which < 0 || which > 1) return -1.f;
    float ms = -1.f;
    if (cudaEventSynchronize(b->ev[which + 1]) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, b->ev[which], b->ev[which + 1]) != cudaSuccess) return -1.f;
    return ms;
}

extern "C" int dcsb_batch_results(dcsb_batch *b, void *cuda_stream, dcsb_result *results)
{
    if (!b) return DCSB_E_ARG;
    dcsb_ctx *ctx = b->ctx;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(cudaStreamSynchronize((cudaStream_t)cuda_stream), "stream sync (kernel failure?)");
    if (!results || b->n == 0) return DCSB_OK;
    std::vector<int32_t> st(b->n);
    std::vector<uint32_t> np(b->n), eb(b->n);
    std::vector<unsigned long long> cs(b->n);
    CK(cudaMemcpy(st.data(), b->scan.status, b->n * 4, cudaMemcpyDeviceToHost), "D2H status");
    CK(cudaMemcpy(np.data(), b->scan.nplay, b->n * 4, cudaMemcpyDeviceToHost), "D2H nplay");
    CK(cudaMemcpy(eb.data(), b->scan.endbits, b->n * 4, cudaMemcpyDeviceToHost), "D2H endbits");
    CK(cudaMemcpy(cs.data(), b->d_checksums, b->n * 8, cudaMemcpyDeviceToHost), "D2H checksums");
    for (size_t i = 0; i < b->n; ++i) {
        results[i].status = b->host_status[i] ? b->host_status[i] : st[i];
        results[i].frames = b->recs[i].out_frames;
        results[i].frames_decoded = np[i];
        results[i].stream_bytes = b->host_status[i] ? 0 : 2 + b->recs[i].hdr_len + (eb[i] + 7) / 8;
        results[i].checksum = cs[i];
    }
    return DCSB_OK;
}

extern "C" int dcsb_batch_read_pcm(dcsb_batch *b, size_t i, int16_t *pcm, size_t max_samples)
{
    if (!b || i >= b->n || !pcm) return DCSB_E_ARG;
    dcsb_ctx *ctx = b->ctx;
    if (!b->d_pcm) return fail(ctx, DCSB_E_ARG, "dcsb_batch_read_pcm: nothing decoded into the internal buffer");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const size_t ns = std::min<size_t>(max_samples, (size_t)b->recs[i].out_frames * 240);
    CK(cudaMemcpy(pcm, b->d_pcm + b->recs[i].pcm_off, ns * 2, cudaMemcpyDeviceToHost), "D2H pcm");
    return (int)std::min<size_t>(ns, 0x7FFFFFFF);
}

extern "C" int dcsb_batch_read_scan(dcsb_batch *b, size_t i, uint32_t *bitpos, uint8_t *bandtypes, size_t max_frames)
{
    if (!b || i >= b->n) return DCSB_E_ARG;
    dcsb_ctx *ctx = b->ctx;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const size_t nf = std::min<size_t>(max_frames, b->recs[i].nframes);
    if (nf == 0) return 0;
    if (bitpos) CK(cudaMemcpy(bitpos, b->scan.bitpos + b->recs[i].frame_base, nf * 4, cudaMemcpyDeviceToHost), "D2H bitpos");
    if (bandtypes) {
        std::vector<uint2> bt(nf);
        CK(cudaMemcpy(bt.data(), b->scan.bt + b->recs[i].frame_base, nf * 8, cudaMemcpyDeviceToHost), "D2H bt");
        for (size_t f = 0; f < nf; ++f) {
            const uint64_t v = ((uint64_t)bt[f].y << 32) | bt[f].x;
            for (int k = 0; k < 16; ++k) bandtypes[f * 16 + k] = (uint8_t)((v >> (4 * k)) & 15);
        }
    }
    return (int)nf;
}

extern "C" int dcsb_decode_streams(dcsb_ctx *ctx, const dcsb_stream_desc *descs, size_t n,
                                   int16_t *pcm_out, const uint64_t *pcm_offsets, dcsb_result *results)
{
    if (!ctx || (!descs && n) || (!pcm_out && n)) return fail(ctx, DCSB_E_ARG, "dcsb_decode_streams: bad argument");
    dcsb_batch *b = nullptr;
    int rc = dcsb_batch_create(ctx, descs, n, &b);
    if (rc != DCSB_OK) return rc;
    rc = dcsb_batch_decode(b, nullptr, nullptr);
    if (rc == DCSB_OK) rc = dcsb_batch_results(b, nullptr, results);
    if (rc == DCSB_OK && n) {
        bool packed = true;
        if (pcm_offsets)
            for (size_t i = 0; i < n && packed; ++i) packed = pcm_offsets[i] == b->recs[i].pcm_off;
        cudaError_t e = cudaSuccess;
        if (packed) e = cudaMemcpy(pcm_out, b->d_pcm, b->total_out_frames * 480, cudaMemcpyDeviceToHost);
        else
            for (size_t i = 0; i < n && e == cudaSuccess; ++i)
                e = cudaMemcpyAsync(pcm_out + pcm_offsets[i], b->d_pcm + b->recs[i].pcm_off,
                                    (size_t)b->recs[i].out_frames * 480, cudaMemcpyDeviceToHost, nullptr);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) rc = fail(ctx, DCSB_E_CUDA, "D2H pcm", e);
    }
    dcsb_batch_destroy(b);
    return rc;
}
