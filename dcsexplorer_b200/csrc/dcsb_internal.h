// dcsb200 internal definitions shared by the host API (dcsb_api.cu) and the kernels.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

// frame layout family of a stream
enum : uint8_t { DCSB_FMT_94 = 0, DCSB_FMT_93 = 1, DCSB_FMT_93A1 = 2 };

// One stream, as the kernels see it (64 bytes).  Read-only on the device.
struct DcsbStreamRec {
    uint64_t data_off;      // byte offset of the stream's first byte in the compressed slab (16-byte aligned)
    uint64_t pcm_off;       // sample offset of the stream's first PCM sample in the output
    uint32_t nbytes;        // stream bytes
    uint32_t frame_base;    // index of frame 0 in the frame-checkpoint arrays
    uint32_t out_frames;    // frames rendered (nFrames + tail)
    uint16_t pad0;
    uint8_t  fmt;           // DCSB_FMT_*
    uint8_t  hdr_len;       // 16, or 1 for OS93a type 1
    uint16_t mult0, mult1;  // effective channel multiplier for frame 0 / frames >= 1 (host gain staging)
    uint8_t  vs0, vs1;      // volShift for frame 0 / frames >= 1
    uint8_t  vs_idle;       // volShift once no stream is active (8)
    uint8_t  pad[1];
    uint8_t  hdr[16];       // stream header copy
    uint32_t nframes;       // frames to walk: the preamble's count (0 when rejected on the host; 65536 for a
                            // zero count under DCSB_STREAM_WRAP_EMPTY)
    uint32_t pad2;
};
static_assert(sizeof(DcsbStreamRec) == 64, "DcsbStreamRec layout");

// A work item = `count` consecutive output frames of one stream starting at `first`, handled by
// one warp.  1993-family items are single tiles of <= DCSB_TILE_OUT frames (lane 0 re-decodes
// the frame before the tile so the 16-sample overlap is available); 1994-family items span
// several 32-frame tiles and carry the overlap from tile to tile (one warm-up frame per item).
#define DCSB_TILE_OUT 31
struct DcsbTile { uint32_t stream; uint32_t first; uint32_t count; };

// Peek-LUT block (uint16 entries: len<<8 | value), copied to shared memory by each CTA.
#define DCSB_LUT_HDR94   0      // 256: 1994 frame-header delta code, 8-bit peek (0 = longer code)
#define DCSB_LUT_CB      256    // 940: 1994 sample codebooks 1..6 (4+8+32+128+256+512): len<<12 | two-zeros<<11 | signed value
#define DCSB_LUT_HDR93   1196   // 256: 1993 type-1 band-type delta code, 8-bit peek (0 = longer code)
#define DCSB_LUT_BB93A   1452   // 64:  OS93a band-bits codes, 4 groups x 4-bit peek
#define DCSB_LUT_SC93A   1516   // 256: OS93a scale-delta code, 8-bit peek
#define DCSB_LUT_XLAT    1772   // 48:  1994 type-1 band translation, 3 groups x 16: (codebook/width << 8) | scale adjust
#define DCSB_LUT_HDR94B 1820   // 256: 1994 frame-header delta codes of 9..16 bits, by the 8 bits behind their common 8-bit prefix
#define DCSB_LUT_WORDS   2076

// scan: length table of the 1994 sample codebooks (dcsb_scan94.cuh); entry = y1 << 16 | y8, each slots << 12 | bits
#define DCSB_T8_PEEK 12
#define DCSB_T8_CB   (1 << DCSB_T8_PEEK)            // entries per codebook (16 KB)
#define DCSB_T8_CAP  15                             // most output slots one multi-codeword step covers
#define DCSB_T1_PEEK 9                              // the single-codeword part depends on the first 9 bits only

struct DcsbLongCode { uint32_t code; uint8_t len; uint8_t val; uint16_t pad; };

struct alignas(8) DcsbTw2 { int c2, s2; };      // a butterfly twiddle pre-doubled: 2 cos, 2 sin
struct DcsbTables {
    uint16_t lut[DCSB_LUT_WORDS];
    uint16_t overlap[16];
    uint32_t twiddle[128];     // (cos << 16) | (sin & 0xffff), reference table order (bit-reversed partitions)
    uint32_t pretw[64];        // twiddle pass of the 1994 transform in natural order i: (c0 << 16) | c1
    uint16_t pairs93a[2048];
    DcsbLongCode long94[32];   // 1994 header codes longer than 8 bits
    DcsbLongCode long93[64];   // 1993 header codes longer than 8 bits
    int n_long94, n_long93;
    // 1994 fast path
    int tw_c2[64], tw_s2[64];      // butterfly twiddles pre-doubled (2cos, 2sin), partition order
    int pre_c0[64], pre_c1[64];    // pre-pass coefficients pre-doubled, natural order
    DcsbTw2 tw93[128];             // 1993 lane transform: the twiddles of `twiddle` pre-doubled
    // scan: tx[codebook][next 12 bits] = y1 << 16 | y8; y8 = as many whole codewords as fit (at most 15
    // output slots), y1 = exactly one codeword; y = output slots covered << 12 | bits consumed
    uint32_t tx[6 * DCSB_T8_CB];
};

#define DCSB_DEV_E_QUEUE 1u                // decode warp: no work item arrived (scan CTA not resident / lost)
#define DCSB_DEV_E_AWAIT 2u                // decode warp: the checkpoints of its item never arrived
#define DCSB_DEV_E_GATE  4u                // gate kernel: the scan CTAs did not become resident
#define DCSB_SCAN_DONE 0x80000000u
#define DCSB_SCAN_RUNNING 1                 // DcsbScanOut::status between two time slices of a stream (never seen by callers)
#define DCSB_QITEM 63u                      // frames per queued work item: with the warm-up frame, two full 32-lane tiles
#define DCSB_Q_VALID (1ull << 63)           // queue entry: VALID | FINAL? | stream << 24 | first frame
#define DCSB_Q_FINAL (1ull << 62)           // the stream's scan is finished: nplay / stopband are final
struct DcsbScanOut {
    // checkpoints: one per stream frame plus one end entry per stream (frame_base counts both)
    uint32_t *bitpos;          // frame start, bits from the first byte after the stream header
    uint2    *bt;              // band-type state carried into the frame, 16 x 4 bits
    uint16_t *hdrbits;         // 1994 layout: length of the frame header in bits
    int32_t  *status;          // [nstreams]
    uint32_t *nplay;           // [nstreams] frames that decode before the channel goes silent
    uint32_t *endbits;         // [nstreams] bit position after the last decoded frame
    uint8_t  *stopband;        // [nstreams] band at which the reference's error path fired (else 0xFF)
    // scan -> decode hand-off while both kernels run (NULL = the scan has finished before the decode
    // starts): progress[s] = checkpoints of stream s that are valid so far, DCSB_SCAN_DONE once the
    // stream is finished and status / nplay / stopband are final; started[0] counts resident scan CTAs
    uint32_t *progress;
    uint32_t *started;
    // 1994 layout, overlapped mode: the scan appends a work item to `queue` every DCSB_QITEM frames
    // of a stream (and the rest of the stream when it finishes it); the persistent decode kernel
    // takes the items in that order, so it always works on frames whose checkpoints exist.
    // qctl[0] = tail (scan side), qctl[1] = head (decode side), qctl[2] = error word: set when a consumer gave
    // up waiting for its producer (DCSB_DEV_E_*); the host turns it into DCSB_E_CUDA -- a timeout is never a
    // silent success.
    unsigned long long *queue;
    uint32_t *qctl;
    uint32_t *dbg;             // tuning builds (-DDCSB_SCAN_DEBUG): [nstreams][4] cycles lo/hi, table steps, header steps; else NULL
};

void dcsb_build_tables(DcsbTables *t);   // host
#define DCSB_SCAN_MAXWARPS 2              // most lock-step warps (32 streams, 1 KB ring each) a scan CTA holds
#define DCSB_SCAN_MAXWARPS_DIRECT 8       // ... when the stream bytes are read from global memory instead (no rings)
// multi-wave batches (more 32-stream groups than 2 per SM) read their streams straight from global memory
bool dcsb_scan_direct(int nstreams, int concurrent);
// SM count of the device the context runs on (cudaDevAttrMultiProcessorCount; 148 on a B200): sizes the
// scan grid and the persistent decode grid
int dcsb_num_sms();
void dcsb_set_num_sms(int n);

// host-side batch layout (dcsb_host.cpp)
#include <vector>
#include "../../include/dcsb200.h"
struct DcsbPrepared {
    std::vector<DcsbStreamRec> recs;
    std::vector<int32_t> host_status;     // host-side rejections (0 = let the scan decide)
    std::vector<DcsbTile> tiles;          // 1994-family items first, then 1993-family tiles
    size_t concurrent_streams = 0;        // in: streams of all scans running side by side (0 = this batch alone)
    std::vector<uint32_t> scan_order;     // streams in the order the scan assigns them to lanes: alike streams side by side;
                                          // the n_scan94 streams of the 1994 layout first, the 1993 layouts behind them
    size_t n_scan94 = 0;
    int ntiles94 = 0, ntiles93 = 0;
    uint32_t item_len = 31;               // output frames per 1994-family work item
    int nqueue94 = 0;                     // work items the scan queues for the 1994-layout streams (overlapped mode)
    uint64_t total_frames_in = 0;         // stream frames
    uint64_t total_checkpoints = 0;       // stream frames + one end entry per stream
    uint64_t total_out_frames = 0, compressed_bytes = 0;
    size_t slab_bytes = 0;
};
// validate + lay out a batch (no CUDA calls); DCSB_OK or DCSB_E_ARG.  With in_place_base the
// device slab is a verbatim copy of host bytes [in_place_base, in_place_base + in_place_span)
// and each stream keeps its position inside it (any byte alignment); otherwise streams are laid
// out back to back, 16-byte aligned and zero padded, for dcsb_pack_slab.
int dcsb_prepare(const dcsb_stream_desc *descs, size_t n, DcsbPrepared *p, const uint8_t *in_place_base, size_t in_place_span);
// work items covering output frames [fa, fb) of every stream (appended to t94 / t93), frame-major
void dcsb_build_tiles(const DcsbPrepared *p, uint32_t fa, uint32_t fb, std::vector<DcsbTile> *t94, std::vector<DcsbTile> *t93);
// scan launch shape (lock-step warps per CTA, CTAs) and stream -> lane assignment (dcsb_host.cpp)
void dcsb_scan_shape(int nstreams, int concurrent, int *warps, int *grid);
void dcsb_scan_order(DcsbPrepared *p);
// copy the streams into `slab` (p->slab_bytes bytes) at their 16-byte aligned offsets, zero padded
void dcsb_pack_slab(const dcsb_stream_desc *descs, size_t n, const DcsbPrepared *p, uint8_t *slab);
// the same for streams [i0, i1) only: `dst` stands for slab offset recs[i0].data_off; fills up to the offset of
// stream i1 (or the slab's end)
void dcsb_pack_slab_range(const dcsb_stream_desc *descs, size_t n, const DcsbPrepared *p, size_t i0, size_t i1, uint8_t *dst);

// concurrent: streams of all the scans launched side by side (0 = this launch alone), see dcsb_scan_shape
// order: device array of nstreams stream indices (NULL = identity): which stream each scan lane takes
// [f0, f1): frames of every stream this launch walks; f0 > 0 resumes from the checkpoints the launch
// for [.., f0) left (time-sliced chunks of dcsb_decode_streams)
// n94: order[0..n94) are the streams of the 1994 layout (lock-step kernel), order[n94..nstreams) those of the 1993
// layouts (dcsb_scan93_kernel, launched behind it on the same stream)
cudaError_t dcsb_launch_scan(const uint8_t *slab, const DcsbStreamRec *streams, const uint32_t *order, int nstreams, int n94, int concurrent,
                             const DcsbTables *tables, DcsbScanOut out, cudaStream_t st, uint32_t f0 = 0, uint32_t f1 = 0xFFFFFFFFu);
// tiles[0..ntiles94) use the 1994 transform, tiles[ntiles94..ntiles94+ntiles93) the 1993 one
// enqueue a one-thread kernel that returns once `ctas` scan CTAs are resident (scan.started)
cudaError_t dcsb_launch_gate(DcsbScanOut scan, int ctas, cudaStream_t st);
int dcsb_scan_grid(int nstreams, int n94, int concurrent);   // CTAs dcsb_launch_scan uses (both kernels)
void dcsb_decode_shapes(int nitems, int *grid_items, int *grid_queue, int *block);   // 1994-layout decode kernels: CTAs for nitems work items / of the persistent kernel
// persistent decode over the scan's ready queue (1994-layout streams, overlapped mode)
cudaError_t dcsb_launch_decode_queue(const uint8_t *slab, const DcsbStreamRec *streams, int nstreams, int nitems,
                                     const DcsbTables *tables, DcsbScanOut scan, int16_t *pcm,
                                     unsigned long long *checksums, cudaStream_t st);
cudaError_t dcsb_launch_decode(const uint8_t *slab, const DcsbStreamRec *streams, const DcsbTile *tiles,
                               int ntiles94, int ntiles93, const DcsbTables *tables, DcsbScanOut scan,
                               int16_t *pcm, unsigned long long *checksums, cudaStream_t st);

// K5: the track interpreter on the device (dcsb_seq.cuh): one thread per decoder instance (timeline) writes the mix
// schedule -- frames[first_frame + f], entries[8 * (first_frame + f) ...] -- K4 renders.  tls / writes / out are
// device pointers (DcsbSeqTimeline[n], dcsb_port_write[], uint32_t[2 n] = {fatal, host bytes} per timeline).
struct DcsbSeqTimeline { uint32_t first_write, n_writes, n_frames, first_frame; uint32_t master_volume; };
struct DcsbRomView;
cudaError_t dcsb_launch_seq(const DcsbRomView *view, const DcsbSeqTimeline *tls, int n, const void *writes,
                            void *frames, void *entries, uint32_t *out, cudaStream_t st);

// K4 launch (dcsb_mix.cuh): items = DcsbMixItem[], frames = DcsbSchedFrame[], entries = DcsbSchedEntry[]
// (device pointers; the types live in dcsb_rom.h / dcsb_mix.cuh).  family93 selects the 1993 transform.
cudaError_t dcsb_launch_mix(bool family93, const uint8_t *slab, const DcsbStreamRec *streams, const void *items, int nitems,
                            const void *frames, const void *entries, const DcsbTables *tables, DcsbScanOut scan,
                            int16_t *pcm, unsigned long long *checksums, cudaStream_t st);
