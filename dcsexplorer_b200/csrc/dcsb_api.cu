// dcsb200 C-ABI implementation: host-side control plane (stream validation, gain staging,
// slab packing, tiling) and kernel orchestration.  See include/dcsb200.h.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <thread>
#include <vector>
#include <algorithm>
#include "../../include/dcsb200.h"
#include "dcsb_internal.h"

#include "dcsb_ctx.h"

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
    {
        int sms = 0, major = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess) dcsb_set_num_sms(sms);
        // the kernels are built for sm_100a only: anything else has no code to run (and there is no CPU fallback)
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess || major != 10) { delete ctx; return DCSB_E_CUDA; }
    }
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
    for (DcsbLane &l : ctx->lanes) {
        if (l.st) { cudaStreamSynchronize(l.st); cudaStreamDestroy(l.st); }
        if (l.aux) { cudaStreamSynchronize(l.aux); cudaStreamDestroy(l.aux); }
        if (l.ev_go) cudaEventDestroy(l.ev_go);
        if (l.ev_scan) cudaEventDestroy(l.ev_scan);
        for (cudaEvent_t e : l.ev_slices) cudaEventDestroy(e);
        l.d_progress.release(false);
        l.h_slab.release(true); l.h_res.release(true); l.h_meta.release(true);
        for (DcsbBuf *b : { &l.d_slab, &l.d_recs, &l.d_tiles, &l.d_bitpos, &l.d_bt, &l.d_hdrbits, &l.d_status, &l.d_nplay,
                            &l.d_endbits, &l.d_stopband, &l.d_csum, &l.d_pcm, &l.d_queue, &l.d_order }) b->release(false);
    }
    if (ctx->timeline_cache && ctx->timeline_cache_free) ctx->timeline_cache_free(ctx->timeline_cache);
    if (ctx->encode_cache && ctx->encode_cache_free) ctx->encode_cache_free(ctx->encode_cache);
    if (ctx->aux) { cudaStreamSynchronize(ctx->aux); cudaStreamDestroy(ctx->aux); }
    if (ctx->up) { cudaStreamSynchronize(ctx->up); cudaStreamDestroy(ctx->up); }
    if (ctx->down) { cudaStreamSynchronize(ctx->down); cudaStreamDestroy(ctx->down); }
    cudaFree(ctx->d_tables);
    delete ctx;
}

extern "C" int dcsb_set_overlap(dcsb_ctx *ctx, int on)
{
    if (!ctx) return DCSB_E_ARG;
    ctx->overlap = on != 0;
    return DCSB_OK;
}

extern "C" int dcsb_set_pipeline(dcsb_ctx *ctx, int max_chunks, int slice_frames)
{
    if (!ctx || max_chunks < 0 || max_chunks > DCSB_MAX_LANES) return DCSB_E_ARG;
    ctx->max_chunks = max_chunks;
    ctx->slice_frames = slice_frames;
    return DCSB_OK;
}

extern "C" const char *dcsb_last_error(const dcsb_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }

extern "C" void dcsb_batch_destroy(dcsb_batch *b)
{
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaFree(b->d_slab); cudaFree(b->d_recs); cudaFree(b->d_tiles); cudaFree(b->d_order);
    cudaFree(b->scan.bitpos); cudaFree(b->scan.bt); cudaFree(b->scan.hdrbits); cudaFree(b->scan.status); cudaFree(b->scan.nplay);
    cudaFree(b->scan.endbits); cudaFree(b->scan.stopband); cudaFree(b->scan.dbg);
    cudaFree(b->d_pcm); cudaFree(b->d_checksums); cudaFree(b->d_progress); cudaFree(b->d_queue);
    for (auto &e : b->ev) if (e) cudaEventDestroy(e);
    delete b;
}

extern "C" int dcsb_batch_create(dcsb_ctx *ctx, const dcsb_stream_desc *descs, size_t n, dcsb_batch **out)
{
    return dcsb_batch_create_impl(ctx, descs, n, nullptr, 0, out);
}

// in_place_base != NULL: the streams all lie inside host bytes [in_place_base, +in_place_span),
// which become the device slab verbatim (ROM images: streams keep their chip positions)
int dcsb_batch_create_impl(dcsb_ctx *ctx, const dcsb_stream_desc *descs, size_t n, const uint8_t *in_place_base,
                           size_t in_place_span, dcsb_batch **out)
{
    if (!ctx || !out || (!descs && n)) return fail(ctx, DCSB_E_ARG, "dcsb_batch_create: bad argument");
    *out = nullptr;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    dcsb_batch *b = new dcsb_batch();
    b->ctx = ctx;
    b->n = n;

    DcsbPrepared prep;
    {
        int rc = dcsb_prepare(descs, n, &prep, in_place_base, in_place_span);
        if (rc != DCSB_OK) { delete b; return fail(ctx, rc, "dcsb_batch_create: unknown os_version or batch too large"); }
    }
    b->recs = prep.recs;
    b->host_status = prep.host_status;
    b->tiles = prep.tiles;
    b->ntiles94 = prep.ntiles94;
    b->ntiles93 = prep.ntiles93;
    b->nqueue94 = prep.nqueue94;
    b->total_frames_in = prep.total_frames_in;
    b->total_out_frames = prep.total_out_frames;
    b->compressed_bytes = prep.compressed_bytes;
    b->slab_bytes = prep.slab_bytes;
    const uint64_t frames = prep.total_checkpoints;

    // pack the compressed slab through pinned staging (multi-threaded) and upload it; a slab larger than the
    // staging buffer goes piece by piece (whole streams per piece), two buffers so that packing overlaps the copy
    uint8_t *h_slab = nullptr;
    const size_t STAGE = (size_t)256 << 20;
    const bool whole = in_place_base || b->slab_bytes <= STAGE;
    const size_t stage_bytes = whole ? b->slab_bytes : STAGE;
    cudaError_t e = cudaMallocHost(&h_slab, whole ? stage_bytes : 2 * stage_bytes);
    if (e != cudaSuccess) { delete b; return fail(ctx, DCSB_E_NOMEM, "cudaMallocHost(slab)", e); }
#define CKB(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cudaFreeHost(h_slab); dcsb_batch_destroy(b); return fail(ctx, DCSB_E_CUDA, what, e_); } } while (0)
    CKB(cudaMalloc(&b->d_slab, b->slab_bytes), "cudaMalloc(slab)");
    if (whole) {
        if (in_place_base) {
            memcpy(h_slab, in_place_base, in_place_span);
            memset(h_slab + in_place_span, 0, b->slab_bytes - in_place_span);
        } else dcsb_pack_slab(descs, n, &prep, h_slab);
        CKB(cudaMemcpy(b->d_slab, h_slab, b->slab_bytes, cudaMemcpyHostToDevice), "H2D slab");
    } else {
        cudaStream_t cs;
        CKB(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking), "cudaStreamCreate");
        cudaEvent_t done[2];
        for (auto &ev : done) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        size_t i0 = 0;
        int k = 0;
        cudaError_t ce = cudaSuccess;
        while (i0 < n && ce == cudaSuccess) {
            const uint64_t base = prep.recs[i0].data_off;
            size_t i1 = i0 + 1;
            auto end_of = [&](size_t i) { return i < n ? prep.recs[i].data_off : (uint64_t)b->slab_bytes; };
            while (i1 < n && end_of(i1 + 1) - base <= stage_bytes) ++i1;
            if (end_of(i1) - base > stage_bytes) { ce = cudaErrorInvalidValue; break; }     // (one stream larger than the staging buffer: cannot happen, a stream is < 36 MB)
            uint8_t *dst = h_slab + (size_t)(k & 1) * stage_bytes;
            if (k >= 2) ce = cudaEventSynchronize(done[k & 1]);      // the copy that last used this buffer
            if (ce != cudaSuccess) break;
            dcsb_pack_slab_range(descs, n, &prep, i0, i1, dst);
            ce = cudaMemcpyAsync(b->d_slab + base, dst, end_of(i1) - base, cudaMemcpyHostToDevice, cs);
            if (ce == cudaSuccess) ce = cudaEventRecord(done[k & 1], cs);
            i0 = i1;
            ++k;
        }
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(cs);
        for (auto &ev : done) cudaEventDestroy(ev);
        cudaStreamDestroy(cs);
        if (n && prep.recs[0].data_off) cudaMemset(b->d_slab, 0, prep.recs[0].data_off);
        CKB(ce, "H2D slab (staged)");
    }
    cudaFreeHost(h_slab);
    h_slab = nullptr;
    CKB(cudaMalloc(&b->d_recs, std::max<size_t>(1, n) * sizeof(DcsbStreamRec)), "cudaMalloc(recs)");
    CKB(cudaMemcpy(b->d_recs, b->recs.data(), n * sizeof(DcsbStreamRec), cudaMemcpyHostToDevice), "H2D recs");
    CKB(cudaMalloc(&b->d_tiles, std::max<size_t>(1, b->tiles.size()) * sizeof(DcsbTile)), "cudaMalloc(tiles)");
    CKB(cudaMemcpy(b->d_tiles, b->tiles.data(), b->tiles.size() * sizeof(DcsbTile), cudaMemcpyHostToDevice), "H2D tiles");
    CKB(cudaMalloc(&b->d_order, std::max<size_t>(1, n) * sizeof(uint32_t)), "cudaMalloc(order)");
    CKB(cudaMemcpy(b->d_order, prep.scan_order.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice), "H2D order");
    b->n94 = prep.n_scan94;
    CKB(cudaMalloc(&b->scan.bitpos, std::max<uint64_t>(1, frames) * sizeof(uint32_t)), "cudaMalloc(bitpos)");
    CKB(cudaMalloc(&b->scan.bt, std::max<uint64_t>(1, frames) * sizeof(uint2)), "cudaMalloc(bt)");
    CKB(cudaMalloc(&b->scan.hdrbits, std::max<uint64_t>(1, frames) * sizeof(uint16_t)), "cudaMalloc(hdrbits)");
    CKB(cudaMalloc(&b->scan.status, std::max<size_t>(1, n) * sizeof(int32_t)), "cudaMalloc(status)");
    CKB(cudaMalloc(&b->scan.nplay, std::max<size_t>(1, n) * sizeof(uint32_t)), "cudaMalloc(nplay)");
    CKB(cudaMalloc(&b->scan.endbits, std::max<size_t>(1, n) * sizeof(uint32_t)), "cudaMalloc(endbits)");
    CKB(cudaMalloc(&b->scan.stopband, std::max<size_t>(1, n)), "cudaMalloc(stopband)");
#ifdef DCSB_SCAN_DEBUG
    CKB(cudaMalloc(&b->scan.dbg, std::max<size_t>(2048, n) * 32), "cudaMalloc(dbg)");
    CKB(cudaMemset(b->scan.dbg, 0, std::max<size_t>(2048, n) * 32), "memset(dbg)");
#endif
    CKB(cudaMalloc(&b->d_checksums, std::max<size_t>(1, n) * sizeof(unsigned long long)), "cudaMalloc(checksums)");
    CKB(cudaMalloc(&b->d_progress, (n + 4) * sizeof(uint32_t)), "cudaMalloc(progress)");
    CKB(cudaMemset(b->d_progress, 0, (n + 4) * sizeof(uint32_t)), "memset(progress)");
    CKB(cudaMalloc(&b->d_queue, std::max<size_t>(1, (size_t)b->nqueue94) * sizeof(unsigned long long)), "cudaMalloc(queue)");
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
    // scan (+ the one-thread gate when scan and decode overlap) + one decode launch per transform family
    // (the scan is one kernel per layout family: lock-step warps for the 1994 layout, a lane per stream for the 1993 ones)
    return (b->n94 ? 1 : 0) + (b->n > b->n94 ? 1 : 0) + (b->n && b->ctx->overlap ? 1 : 0) + (b->ntiles94 ? 1 : 0) + (b->ntiles93 ? 1 : 0);
}

extern "C" int dcsb_batch_launch_shape(const dcsb_batch *b, int which, int *grid, int *block)
{
    if (!b || !grid || !block || which < 0 || which > 2) return DCSB_E_ARG;
    if (which == 0) {
        int warps, g;
        dcsb_scan_shape((int)b->n94, 0, &warps, &g);
        *grid = b->n94 ? g : 0;
        *block = warps * 32;
        return DCSB_OK;
    }
    int gi, gq, blk;
    dcsb_decode_shapes(which == 1 ? b->ntiles94 : b->nqueue94, &gi, &gq, &blk);
    *grid = which == 1 ? gi : gq;
    *block = blk;
    return DCSB_OK;
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
    if (!ctx->overlap) {
        // one kernel after the other on the caller's stream
        DcsbScanOut so = b->scan;
        so.progress = so.started = so.qctl = nullptr;
        so.queue = nullptr;
        CK(cudaEventRecord(b->ev[0], st), "event");
        CK(dcsb_launch_scan(b->d_slab, b->d_recs, b->d_order, (int)b->n, (int)b->n94, 0, ctx->d_tables, so, st), "scan kernel launch");
        CK(cudaEventRecord(b->ev[1], st), "event");
        CK(cudaEventRecord(b->ev[3], st), "event");
        CK(dcsb_launch_decode(b->d_slab, b->d_recs, b->d_tiles, b->ntiles94, b->ntiles93, ctx->d_tables, so,
                              pcm, b->d_checksums, st), "decode kernel launch");
        CK(cudaEventRecord(b->ev[2], st), "event");
    } else {
        // the scan (a few latency-bound warps per SM) runs on its own stream BESIDE the decode kernel,
        // whose warps wait per work item for the checkpoints they need (dcsb_await)
        if (!ctx->aux) CK(cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking), "cudaStreamCreate");
        DcsbScanOut so = b->scan;
        so.progress = b->d_progress;
        so.started = b->d_progress + b->n;
        so.qctl = b->d_progress + b->n + 1;
        so.queue = b->d_queue;
        CK(cudaMemsetAsync(b->d_progress, 0, (b->n + 4) * sizeof(uint32_t), st), "memset progress");
        if (b->nqueue94) CK(cudaMemsetAsync(b->d_queue, 0, (size_t)b->nqueue94 * sizeof(unsigned long long), st), "memset queue");
        CK(cudaEventRecord(b->ev[0], st), "event");
        CK(cudaStreamWaitEvent(ctx->aux, b->ev[0], 0), "stream wait");
        CK(dcsb_launch_scan(b->d_slab, b->d_recs, b->d_order, (int)b->n, (int)b->n94, 0, ctx->d_tables, so, ctx->aux), "scan kernel launch");
        CK(cudaEventRecord(b->ev[1], ctx->aux), "event");
        CK(dcsb_launch_gate(so, dcsb_scan_grid((int)b->n, (int)b->n94, 0), st), "gate kernel launch");
        CK(cudaEventRecord(b->ev[3], st), "event");
        CK(dcsb_launch_decode_queue(b->d_slab, b->d_recs, (int)b->n, b->nqueue94, ctx->d_tables, so, pcm, b->d_checksums, st), "decode kernel launch");
        CK(dcsb_launch_decode(b->d_slab, b->d_recs, b->d_tiles + b->ntiles94, 0, b->ntiles93, ctx->d_tables, so,
                              pcm, b->d_checksums, st), "decode kernel launch");
        CK(cudaStreamWaitEvent(st, b->ev[1], 0), "stream wait");
        CK(cudaEventRecord(b->ev[2], st), "event");
    }
    b->timed = true;
    return DCSB_OK;
}

extern "C" float dcsb_batch_last_kernel_ms(dcsb_batch *b, int which)
{
    // 0 = scan span, 1 = decode span (from its launch; overlaps the scan unless dcsb_set_overlap(ctx, 0)), 2 = whole step
    if (!b || !b->timed || which < 0 || which > 2) return -1.f;
    float ms = -1.f;
    cudaEvent_t a = which == 1 ? b->ev[3] : b->ev[0], z = which == 0 ? b->ev[1] : b->ev[2];
    if (cudaEventSynchronize(b->ev[2]) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, a, z) != cudaSuccess) return -1.f;
    return ms;
}

extern "C" int dcsb_batch_results(dcsb_batch *b, void *cuda_stream, dcsb_result *results)
{
    if (!b) return DCSB_E_ARG;
    dcsb_ctx *ctx = b->ctx;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(cudaStreamSynchronize((cudaStream_t)cuda_stream), "stream sync (kernel failure?)");
    if (b->n == 0) return DCSB_OK;
    {
        // a decode warp (or the gate) that gave up waiting for the scan left its mark here: the PCM is incomplete
        uint32_t errw = 0;
        CK(cudaMemcpy(&errw, b->d_progress + b->n + 3, 4, cudaMemcpyDeviceToHost), "D2H error word");
        if (errw) return fail(ctx, DCSB_E_CUDA, "dcsb_batch_results: the decode kernel timed out waiting for the scan (PCM incomplete)");
    }
    if (!results) return DCSB_OK;
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

#ifdef DCSB_SCAN_DEBUG
// tuning builds only: per-stream {cycles lo, cycles hi, table steps, header steps} of the last scan
extern "C" int dcsb_batch_scan_debug(dcsb_batch *b, uint32_t *out4)
{
    if (!b || !out4) return DCSB_E_ARG;
    return cudaMemcpy(out4, b->scan.dbg, std::max<size_t>(2048, b->n) * 32, cudaMemcpyDeviceToHost) == cudaSuccess ? DCSB_OK : DCSB_E_CUDA;
}
#endif

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

// ---- one-shot decode with host buffers: a pipeline of stream chunks -------------------
// The batch is cut into chunks of streams (one lane each) and, where a chunk's streams all render
// the same number of frames into a packed pinned buffer, every chunk also into time slices:
//
//   upload stream   H2D chunk 0 | H2D chunk 1 | ...                       (in place from pinned memory,
//                                                                          or from the lane's packed staging slab)
//   lane c stream   scan(c, slice 0) decode(c, 0) scan(c, 1) decode(c, 1) ...   (resumed scans)
//   copy stream     PCM(0, 0) PCM(1, 0) ... PCM(0, 1) PCM(1, 1) ...       (one strided copy per chunk and slice)
//
// The scan of a 10 s stream is a 10+ ms dependent chain whatever the chunk size; sliced, the first
// PCM leaves after 1/8 of it and the copy engine -- the bottleneck of the whole call: PCM is 4.7x
// the compressed bytes -- stays busy from then on.  Everything is submitted slice-major, the order
// in which the work becomes ready: the copy engine takes its copies in submission order, so a copy
// queued behind one that still waits for its kernels would wait with it.  Streams are few (lanes +
// 2) so that they do not alias on the device's hardware queues.
static bool is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

#define ENS(buf, bytes, host, what) do { cudaError_t e_ = (buf).ensure((bytes), (host)); if (e_ != cudaSuccess) return fail(ctx, DCSB_E_NOMEM, what, e_); } while (0)

// phase 1 of a lane: lay the chunk out, plan its time slices, upload it (on the context's upload stream)
// pcm_pinned: the caller's PCM buffer is page-locked (PCM is copied straight into it); pcm_packed: stream
// i's PCM starts where stream i - 1's ends; dst_off[i] = sample offset of stream i in the caller's buffer
static int lane_upload(dcsb_ctx *ctx, DcsbLane &l, const dcsb_stream_desc *descs, bool pcm_pinned, bool pcm_packed,
                       const uint64_t *dst_off, size_t total_streams)
{
    const size_t n = l.count;
    const dcsb_stream_desc *d = descs + l.first;
    const int lane_id = (int)(&l - ctx->lanes);
    if (!l.st) {
        CK(cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking), "cudaStreamCreate");
        CK(cudaEventCreateWithFlags(&l.ev_go, cudaEventDisableTiming), "cudaEventCreate");
        CK(cudaEventCreateWithFlags(&l.ev_scan, cudaEventDisableTiming), "cudaEventCreate");
    }
    // in-place upload? (all streams close together inside one pinned host allocation)
    const uint8_t *lo = nullptr, *hi = nullptr;
    uint64_t sum = 0;
    bool have_all = true;
    for (size_t i = 0; i < n; ++i) {
        if (!d[i].data || !d[i].nbytes) { have_all = false; break; }
        if (!lo || d[i].data < lo) lo = d[i].data;
        if (!hi || d[i].data + d[i].nbytes > hi) hi = d[i].data + d[i].nbytes;
        sum += d[i].nbytes;
    }
    const bool in_place = have_all && n && (uint64_t)(hi - lo) <= sum + sum / 4 + 4096 && is_pinned(lo) && is_pinned(hi - 1);
    l.prep.concurrent_streams = total_streams;
    int rc = dcsb_prepare(d, n, &l.prep, in_place ? lo : nullptr, in_place ? (size_t)(hi - lo) : 0);
    if (rc != DCSB_OK) return fail(ctx, rc, "dcsb_decode_streams: unknown os_version or batch too large");
    const DcsbPrepared &p = l.prep;
    const uint64_t ck = std::max<uint64_t>(1, p.total_checkpoints), nn = std::max<size_t>(1, n);
    // Time slices.  Frames [a, b) of every stream of the chunk are scanned, decoded and copied out together:
    // when all streams render the same number of frames into a packed buffer that is ONE strided copy
    // (pitch = stream length), otherwise one batched copy (cudaMemcpyBatchAsync) of a piece per stream.
    // The scan of a 10 s stream is a 10+ ms dependent chain whatever the chunk size; sliced, the first PCM
    // leaves after a fraction of it and the copy engine stays busy from then on.
    l.slice = 0;
    l.nslices = 1;
    l.direct_pcm = pcm_pinned && pcm_packed;
    l.copy2d = false;
    l.dst_off.clear();
    uint32_t U = 0;                               // longest stream of the chunk, in output frames
    for (size_t i = 0; i < n; ++i) U = std::max(U, p.recs[i].out_frames);
    if (pcm_pinned && n && ctx->slice_frames >= 0 && U > 1) {
        bool uniform = pcm_packed && (uint64_t)U * 480 <= 0x7FFFFFFFull;
        for (size_t i = 0; i < n && uniform; ++i) uniform = p.recs[i].out_frames == U;
        if (ctx->slice_frames > 0) l.slice = (uint32_t)ctx->slice_frames;
        else if (U >= 256) l.slice = (((U + 7) / 8 + p.item_len - 1) / p.item_len) * p.item_len;     // ~8 slices of whole work items
        if (l.slice >= U) l.slice = 0;
        if (l.slice && (U + l.slice - 1) / l.slice > 64) l.slice = (U + 63) / 64;
        if (l.slice) {
            l.copy2d = uniform;
            l.direct_pcm = true;
            if (!uniform) l.dst_off.assign(dst_off + l.first, dst_off + l.first + n);
        }
    }
    // slice boundaries.  Automatic slicing starts with short slices (32, 32, 64, 64, 128 frames): the rate at
    // which PCM is produced is set by how many streams are being scanned, not by the slice length, so short
    // first slices only bring the first copies forward while the later chunks are still being uploaded
    l.sl_bound.clear();
    l.sl_bound.push_back(0);
    if (l.slice) {
        uint32_t f = 0;
        if (ctx->slice_frames == 0)
            for (uint32_t len : { 32u, 32u, 64u, 64u, 128u })
                if (f + len + l.slice <= U) { f += len; l.sl_bound.push_back(f); }
        while (f + l.slice < U) { f += l.slice; l.sl_bound.push_back(f); }
        l.sl_bound.push_back(U);
        l.nslices = (uint32_t)l.sl_bound.size() - 1;
    }
    l.sl_off.clear();
    if (l.slice) {
        l.slice_tiles.clear();
        std::vector<DcsbTile> t93;
        for (uint32_t k = 0; k < l.nslices; ++k) {
            t93.clear();
            l.sl_off.push_back(l.slice_tiles.size());
            dcsb_build_tiles(&p, l.sl_bound[k], k + 1 == l.nslices ? 0xFFFFFFFFu : l.sl_bound[k + 1], &l.slice_tiles, &t93);
            l.sl_off.push_back(l.slice_tiles.size());
            l.slice_tiles.insert(l.slice_tiles.end(), t93.begin(), t93.end());
        }
        l.sl_off.push_back(l.slice_tiles.size());
    }
    while (l.ev_slices.size() < l.nslices) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
        l.ev_slices.push_back(e);
    }
    const std::vector<DcsbTile> &tiles = l.slice ? l.slice_tiles : p.tiles;
    ENS(l.d_slab, p.slab_bytes, false, "cudaMalloc(slab)");
    ENS(l.d_recs, nn * sizeof(DcsbStreamRec), false, "cudaMalloc(recs)");
    ENS(l.d_tiles, std::max<size_t>(1, tiles.size()) * sizeof(DcsbTile), false, "cudaMalloc(tiles)");
    ENS(l.d_order, nn * sizeof(uint32_t), false, "cudaMalloc(order)");
    ENS(l.d_bitpos, ck * 4, false, "cudaMalloc(bitpos)");
    ENS(l.d_bt, ck * 8, false, "cudaMalloc(bt)");
    ENS(l.d_hdrbits, ck * 2, false, "cudaMalloc(hdrbits)");
    ENS(l.d_status, nn * 4, false, "cudaMalloc(status)");
    ENS(l.d_nplay, nn * 4, false, "cudaMalloc(nplay)");
    ENS(l.d_endbits, nn * 4, false, "cudaMalloc(endbits)");
    ENS(l.d_stopband, nn, false, "cudaMalloc(stopband)");
    ENS(l.d_csum, nn * 8, false, "cudaMalloc(checksums)");
    ENS(l.d_progress, (nn + 4) * 4, false, "cudaMalloc(progress)");
    ENS(l.d_queue, std::max<size_t>(1, (size_t)p.nqueue94) * 8, false, "cudaMalloc(queue)");
    // (PCM goes to a device buffer and is copied by the DMA engine afterwards: letting the decode
    // warps store straight into the caller's pinned buffer over PCIe measured 86 ms against 70 ms)
    ENS(l.d_pcm, std::max<uint64_t>(2, p.total_out_frames * 480), false, "cudaMalloc(pcm)");
    ENS(l.h_res, nn * 20 + 8, true, "cudaMallocHost(results)");
    if (n == 0) return DCSB_OK;
    cudaStream_t up = ctx->up;
    if (in_place) {
        const size_t span = (size_t)(hi - lo);
        // (the 1 KB of slack behind the span is not cleared: nothing read from there can reach the output --
        // look-ahead bits only select among table entries that share the valid prefix, and a frame that
        // consumes bits behind its stream is a truncation whatever those bits are.  A memset here is a
        // kernel that would queue behind the other lanes' decode grids and hold the next upload back.)
        CK(cudaMemcpyAsync(l.d_slab.p, lo, span, cudaMemcpyHostToDevice, up), "H2D streams (in place)");
    } else {
        ENS(l.h_slab, p.slab_bytes, true, "cudaMallocHost(slab)");
        dcsb_pack_slab(d, n, &p, (uint8_t *)l.h_slab.p);
        CK(cudaMemcpyAsync(l.d_slab.p, l.h_slab.p, p.slab_bytes, cudaMemcpyHostToDevice, up), "H2D slab");
    }
    // descriptors, work items and scan order go through pinned staging: a copy from pageable memory
    // would make the host wait for everything queued on the upload stream before it
    const size_t b_recs = n * sizeof(DcsbStreamRec), b_tiles = tiles.size() * sizeof(DcsbTile), b_order = n * sizeof(uint32_t);
    ENS(l.h_meta, b_recs + b_tiles + b_order + 64, true, "cudaMallocHost(meta)");
    uint8_t *hm = (uint8_t *)l.h_meta.p;
    memcpy(hm, p.recs.data(), b_recs);
    memcpy(hm + b_recs, tiles.data(), b_tiles);
    memcpy(hm + b_recs + b_tiles, p.scan_order.data(), b_order);
    CK(cudaMemcpyAsync(l.d_recs.p, hm, b_recs, cudaMemcpyHostToDevice, up), "H2D recs");
    if (b_tiles) CK(cudaMemcpyAsync(l.d_tiles.p, hm + b_recs, b_tiles, cudaMemcpyHostToDevice, up), "H2D tiles");
    CK(cudaMemcpyAsync(l.d_order.p, hm + b_recs + b_tiles, b_order, cudaMemcpyHostToDevice, up), "H2D order");
    CK(cudaEventRecord(l.ev_go, up), "event");
    ctx->trace.mark(lane_id, -1, "h2d", up);
    // (memsets are kernels: they go on the lane's own stream, not between the uploads)
    CK(cudaMemsetAsync(l.d_csum.p, 0, n * 8, l.st), "memset checksums");
    if (!(!l.slice && ctx->overlap)) CK(cudaMemsetAsync((uint32_t *)l.d_progress.p + n, 0, 4 * 4, l.st), "memset error word");
    if (!l.slice && ctx->overlap) {
        CK(cudaMemsetAsync(l.d_progress.p, 0, (n + 4) * 4, l.st), "memset progress");
        if (p.nqueue94) CK(cudaMemsetAsync(l.d_queue.p, 0, (size_t)p.nqueue94 * 8, l.st), "memset queue");
        CK(cudaEventRecord(l.ev_scan, l.st), "event");       // (re-recorded after the scan; here: the memsets are done)
    }
    CK(cudaStreamWaitEvent(l.st, l.ev_go, 0), "stream wait");
    return DCSB_OK;
}

// phase 2 of a lane: kernels of slice k (the whole chunk when it is not sliced) and the copy of its PCM
static int lane_slice(dcsb_ctx *ctx, DcsbLane &l, uint32_t k, int16_t *pcm_out, int concurrent)
{
    const size_t n = l.count;
    if (n == 0 || k >= l.nslices) return DCSB_OK;
    const DcsbPrepared &p = l.prep;
    const int lane_id = (int)(&l - ctx->lanes);
    const uint8_t *slab = (const uint8_t *)l.d_slab.p;
    const DcsbStreamRec *recs = (const DcsbStreamRec *)l.d_recs.p;
    const DcsbTile *tiles = (const DcsbTile *)l.d_tiles.p;
    int16_t *d_pcm = (int16_t *)l.d_pcm.p;
    unsigned long long *d_csum = (unsigned long long *)l.d_csum.p;
    DcsbScanOut so{ (uint32_t *)l.d_bitpos.p, (uint2 *)l.d_bt.p, (uint16_t *)l.d_hdrbits.p, (int32_t *)l.d_status.p,
                    (uint32_t *)l.d_nplay.p, (uint32_t *)l.d_endbits.p, (uint8_t *)l.d_stopband.p, nullptr, nullptr, nullptr, nullptr, nullptr };
    if (l.slice) {
        const uint32_t U = l.sl_bound.back();
        const uint32_t fa = l.sl_bound[k], fb = l.sl_bound[k + 1];
        CK(dcsb_launch_scan(slab, recs, (const uint32_t *)l.d_order.p, (int)n, (int)p.n_scan94, concurrent, ctx->d_tables, so, l.st, fa,
                            k + 1 == l.nslices ? 0xFFFFFFFFu : fb), "scan kernel launch");
        ctx->trace.mark(lane_id, (int)k, "scan", l.st);
        CK(dcsb_launch_decode(slab, recs, tiles + l.sl_off[2 * k], (int)(l.sl_off[2 * k + 1] - l.sl_off[2 * k]),
                              (int)(l.sl_off[2 * k + 2] - l.sl_off[2 * k + 1]), ctx->d_tables, so, d_pcm, d_csum, l.st), "decode kernel launch");
        ctx->trace.mark(lane_id, (int)k, "decode", l.st);
        CK(cudaEventRecord(l.ev_slices[k], l.st), "event");
        CK(cudaStreamWaitEvent(ctx->down, l.ev_slices[k], 0), "stream wait");
        if (l.copy2d) {
            CK(cudaMemcpy2DAsync(pcm_out + l.pcm_base + (uint64_t)fa * 240, (size_t)U * 480, d_pcm + (uint64_t)fa * 240, (size_t)U * 480,
                                 (size_t)(fb - fa) * 480, n, cudaMemcpyDeviceToHost, ctx->down), "D2H pcm slice");
        } else {
            // one piece per stream that reaches into the slice
            l.cp_dst.clear(); l.cp_src.clear(); l.cp_size.clear();
            for (size_t i = 0; i < n; ++i) {
                const uint32_t of = p.recs[i].out_frames;
                if (of <= fa) continue;
                l.cp_dst.push_back(pcm_out + l.dst_off[i] + (uint64_t)fa * 240);
                l.cp_src.push_back(d_pcm + p.recs[i].pcm_off + (uint64_t)fa * 240);
                l.cp_size.push_back((size_t)(std::min(of, fb) - fa) * 480);
            }
            if (!l.cp_dst.empty()) {
                cudaMemcpyAttributes at;
                memset(&at, 0, sizeof(at));
                at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
                size_t at_idx = 0, fail_idx = 0;
                CK(cudaMemcpyBatchAsync(l.cp_dst.data(), l.cp_src.data(), l.cp_size.data(), l.cp_dst.size(), &at, &at_idx, 1, &fail_idx,
                                        ctx->down), "D2H pcm slice (batched copy)");
            }
        }
        ctx->trace.mark(lane_id, (int)k, "d2h", ctx->down);
        return DCSB_OK;
    }
    if (ctx->overlap) {
        // the scan (a few latency-bound warps per SM) runs on a second stream BESIDE the persistent decode kernel
        if (!l.aux) CK(cudaStreamCreateWithFlags(&l.aux, cudaStreamNonBlocking), "cudaStreamCreate");
        so.progress = (uint32_t *)l.d_progress.p;
        so.started = so.progress + n;
        so.qctl = so.progress + n + 1;
        so.queue = (unsigned long long *)l.d_queue.p;
        CK(cudaStreamWaitEvent(l.aux, l.ev_go, 0), "stream wait");
        CK(cudaStreamWaitEvent(l.aux, l.ev_scan, 0), "stream wait");
        CK(dcsb_launch_scan(slab, recs, (const uint32_t *)l.d_order.p, (int)n, (int)p.n_scan94, concurrent, ctx->d_tables, so, l.aux), "scan kernel launch");
        CK(cudaEventRecord(l.ev_scan, l.aux), "event");
        CK(dcsb_launch_gate(so, dcsb_scan_grid((int)n, (int)p.n_scan94, concurrent), l.st), "gate kernel launch");
        CK(dcsb_launch_decode_queue(slab, recs, (int)n, p.nqueue94, ctx->d_tables, so, d_pcm, d_csum, l.st), "decode kernel launch");
        CK(dcsb_launch_decode(slab, recs, tiles + p.ntiles94, 0, p.ntiles93, ctx->d_tables, so, d_pcm, d_csum, l.st), "decode kernel launch");
        CK(cudaStreamWaitEvent(l.st, l.ev_scan, 0), "stream wait");
    } else {
        CK(dcsb_launch_scan(slab, recs, (const uint32_t *)l.d_order.p, (int)n, (int)p.n_scan94, concurrent, ctx->d_tables, so, l.st), "scan kernel launch");
        CK(dcsb_launch_decode(slab, recs, tiles, p.ntiles94, p.ntiles93, ctx->d_tables, so, d_pcm, d_csum, l.st), "decode kernel launch");
    }
    ctx->trace.mark(lane_id, -1, "kernels", l.st);
    if (l.direct_pcm) {
        CK(cudaEventRecord(l.ev_slices[0], l.st), "event");
        CK(cudaStreamWaitEvent(ctx->down, l.ev_slices[0], 0), "stream wait");
        CK(cudaMemcpyAsync(pcm_out + l.pcm_base, d_pcm, p.total_out_frames * 480, cudaMemcpyDeviceToHost, ctx->down), "D2H pcm");
        ctx->trace.mark(lane_id, -1, "d2h", ctx->down);
    }
    return DCSB_OK;
}

// phase 3 of a lane: per-stream results into the lane's pinned block
static int lane_results(dcsb_ctx *ctx, DcsbLane &l)
{
    const size_t n = l.count, nn = std::max<size_t>(1, n);
    if (n == 0) return DCSB_OK;
    uint8_t *hr = (uint8_t *)l.h_res.p;
    CK(cudaMemcpyAsync(hr, l.d_status.p, n * 4, cudaMemcpyDeviceToHost, l.st), "D2H status");
    CK(cudaMemcpyAsync(hr + nn * 4, l.d_nplay.p, n * 4, cudaMemcpyDeviceToHost, l.st), "D2H nplay");
    CK(cudaMemcpyAsync(hr + nn * 8, l.d_endbits.p, n * 4, cudaMemcpyDeviceToHost, l.st), "D2H endbits");
    CK(cudaMemcpyAsync(hr + nn * 12, l.d_csum.p, n * 8, cudaMemcpyDeviceToHost, l.st), "D2H checksums");
    CK(cudaMemcpyAsync(hr + nn * 20, (uint32_t *)l.d_progress.p + n + 3, 4, cudaMemcpyDeviceToHost, l.st), "D2H error word");
    return DCSB_OK;
}
#undef ENS

extern "C" int dcsb_decode_streams(dcsb_ctx *ctx, const dcsb_stream_desc *descs, size_t n,
                                   int16_t *pcm_out, const uint64_t *pcm_offsets, dcsb_result *results)
{
    if (!ctx || (!descs && n) || (!pcm_out && n)) return fail(ctx, DCSB_E_ARG, "dcsb_decode_streams: bad argument");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    if (n == 0) return DCSB_OK;
    // output layout: tightly packed unless the caller's offsets say otherwise
    std::vector<uint64_t> off(n + 1, 0);
    bool packed = true;
    for (size_t i = 0; i < n; ++i) {
        if (descs[i].os_version != DCSB_OS94 && descs[i].os_version != DCSB_OS95 && descs[i].os_version != DCSB_OS93A &&
            descs[i].os_version != DCSB_OS93B)
            return fail(ctx, DCSB_E_ARG, "dcsb_decode_streams: unknown os_version");
        // DCSB_STREAM_WRAP_EMPTY (a zero frame count plays 65,536 frames) belongs to ROM playback, which sizes its
        // own buffers; here the caller's buffer is sized from the count as written, so the flag is refused
        if (descs[i].reserved != 0) return fail(ctx, DCSB_E_ARG, "dcsb_decode_streams: dcsb_stream_desc.reserved must be 0");
        const uint32_t nf = (descs[i].data && descs[i].nbytes >= 2) ? (((uint32_t)descs[i].data[0] << 8) | descs[i].data[1]) : 0;
        off[i + 1] = off[i] + (uint64_t)(nf + descs[i].tail_frames) * 240;
        if (pcm_offsets && pcm_offsets[i] != off[i]) packed = false;
    }
    uint64_t extent = off[n];                      // samples of pcm_out the call may write
    if (!packed) { extent = 0; for (size_t i = 0; i < n; ++i) extent = std::max(extent, pcm_offsets[i] + (off[i + 1] - off[i])); }
    const bool pinned = extent && is_pinned(pcm_out) && is_pinned(pcm_out + extent - 1);
    if (!ctx->up) {
        CK(cudaStreamCreateWithFlags(&ctx->up, cudaStreamNonBlocking), "cudaStreamCreate");
        CK(cudaStreamCreateWithFlags(&ctx->down, cudaStreamNonBlocking), "cudaStreamCreate");
    }
    ctx->trace.on = getenv("DCSB_TRACE") != nullptr;
    if (ctx->trace.on) { cudaEventCreate(&ctx->trace.t0); cudaEventRecord(ctx->trace.t0, ctx->up); }
    // chunks of about equal PCM size; few enough that every chunk still fills the GPU
    const uint64_t total = off[n];
    // (a batch that cannot be cut in time -- a pageable output -- gets more, smaller chunks instead: the
    // first PCM copy can only start when a whole chunk is decoded)
    const bool sliceable = pinned && ctx->slice_frames >= 0;
    int nchunks = (int)std::min<uint64_t>(sliceable ? DCSB_DEFAULT_LANES : DCSB_MAX_LANES, std::max<uint64_t>(1, total / (48ull << 20)));
    nchunks = (int)std::min<size_t>((size_t)nchunks, std::max<size_t>(1, n / 64));
    if (ctx->max_chunks > 0) nchunks = (int)std::min<size_t>((size_t)ctx->max_chunks, n);
    int used = 0, rc = DCSB_OK;
    size_t i0 = 0;
    // the first chunks are smaller (weights 1, 2, 3, 3, ...): the first PCM can only leave once a chunk is
    // uploaded and its first slice scanned, and until then the copy engine idles
    uint64_t wsum = 0, wacc = 0;
    for (int c = 0; c < nchunks; ++c) wsum += (uint64_t)std::min(c + 1, 3);
    for (int c = 0; c < nchunks && i0 < n && rc == DCSB_OK; ++c) {
        wacc += (uint64_t)std::min(c + 1, 3);
        const uint64_t goal = total / wsum * wacc;
        size_t i1 = i0 + 1;
        while (i1 < n && (c == nchunks - 1 || off[i1] < goal)) ++i1;
        DcsbLane &l = ctx->lanes[used++];
        l.first = i0;
        l.count = i1 - i0;
        l.pcm_base = off[i0];
        rc = lane_upload(ctx, l, descs, pinned, packed, packed ? off.data() : pcm_offsets, n);
        i0 = i1;
    }
    // submit the (chunk, slice) work in the order it is expected to become ready: chunk c is uploaded
    // after the chunks before it (~45 GB/s), its slices then follow each other at the pace of the scan
    // chain (~14 us per frame with the chip full)
    struct Job { double ready, start; int c; uint32_t k; int concurrent; };
    std::vector<Job> jobs;
    std::vector<double> up_done((size_t)used, 0.0);
    double up_ms = 0.2;
    for (int c = 0; c < used; ++c) {
        const DcsbLane &l = ctx->lanes[c];
        up_ms += (double)l.prep.slab_bytes / 45e6;
        up_done[c] = up_ms;
        double t = up_ms;
        for (uint32_t k = 0; k < l.nslices; ++k) {
            const double frames = l.slice ? (double)(l.sl_bound[k + 1] - l.sl_bound[k]) : (l.count ? (double)l.prep.total_frames_in / (double)l.count : 0.0);
            const double t_slice = frames * 0.012 + 0.15;
            jobs.push_back(Job{ t + t_slice, t, c, k, 0 });
            t += t_slice;
        }
    }
    std::stable_sort(jobs.begin(), jobs.end(), [](const Job &a, const Job &b) { return a.ready < b.ready; });
    // a scan launch shares the SMs (one scan CTA each) with the scans of the chunks that are on the device by
    // then: the first chunks' first slices spread over the whole chip (short chains: the first PCM leaves early),
    // later launches pack their streams so that all chunks fit side by side
    for (Job &j : jobs) {
        size_t live = 0;
        for (int c = 0; c < used; ++c) if (up_done[c] <= j.start + 1e-9) live += ctx->lanes[c].count;
        j.concurrent = (int)std::min<size_t>(std::max(live, ctx->lanes[j.c].count), 0x7FFFFFFF);
    }
    for (const Job &j : jobs) {
        if (rc != DCSB_OK) break;
        rc = lane_slice(ctx, ctx->lanes[j.c], j.k, pcm_out, j.concurrent);
    }
    for (int c = 0; c < used && rc == DCSB_OK; ++c) rc = lane_results(ctx, ctx->lanes[c]);
    // drain
    {
        cudaError_t e = cudaStreamSynchronize(ctx->down);
        if (e != cudaSuccess && rc == DCSB_OK) rc = fail(ctx, DCSB_E_CUDA, "dcsb_decode_streams: stream sync (copy failure?)", e);
        e = cudaStreamSynchronize(ctx->up);
        if (e != cudaSuccess && rc == DCSB_OK) rc = fail(ctx, DCSB_E_CUDA, "dcsb_decode_streams: stream sync (copy failure?)", e);
    }
    for (int c = 0; c < used; ++c) {
        DcsbLane &l = ctx->lanes[c];
        if (!l.st) continue;
        cudaError_t e = cudaStreamSynchronize(l.st);
        if (e != cudaSuccess && rc == DCSB_OK) rc = fail(ctx, DCSB_E_CUDA, "dcsb_decode_streams: stream sync (kernel failure?)", e);
        if (rc != DCSB_OK) continue;
        const size_t nn = std::max<size_t>(1, l.count);
        if (l.count) {
            uint32_t errw;
            memcpy(&errw, (const uint8_t *)l.h_res.p + nn * 20, 4);
            if (errw) { rc = fail(ctx, DCSB_E_CUDA, "dcsb_decode_streams: the decode kernel timed out waiting for the scan (PCM incomplete)"); continue; }
        }
        if (!l.direct_pcm) {
            if (packed) e = cudaMemcpy(pcm_out + l.pcm_base, l.d_pcm.p, l.prep.total_out_frames * 480, cudaMemcpyDeviceToHost);
            else
                for (size_t i = 0; i < l.count && e == cudaSuccess; ++i)
                    e = cudaMemcpy(pcm_out + pcm_offsets[l.first + i], (int16_t *)l.d_pcm.p + l.prep.recs[i].pcm_off,
                                   (size_t)l.prep.recs[i].out_frames * 480, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { rc = fail(ctx, DCSB_E_CUDA, "D2H pcm", e); continue; }
        }
        if (results) {
            const uint8_t *hr = (const uint8_t *)l.h_res.p;
            const int32_t *st = (const int32_t *)hr;
            const uint32_t *np = (const uint32_t *)(hr + nn * 4), *eb = (const uint32_t *)(hr + nn * 8);
            const unsigned long long *cs = (const unsigned long long *)(hr + nn * 12);
            for (size_t i = 0; i < l.count; ++i) {
                dcsb_result &r = results[l.first + i];
                const int32_t hs = l.prep.host_status[i];
                r.status = hs ? hs : st[i];
                r.frames = l.prep.recs[i].out_frames;
                r.frames_decoded = np[i];
                r.stream_bytes = hs ? 0 : 2 + l.prep.recs[i].hdr_len + (eb[i] + 7) / 8;
                unsigned long long c8;
                memcpy(&c8, cs + i, 8);
                r.checksum = c8;
            }
        }
    }
    ctx->trace.dump();
    return rc;
}
