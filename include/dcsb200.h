/* dcsb200 -- B200-native batch decoder for DCS compressed audio streams.
 *
 * C-ABI boundary of the product (libdcsb200.so).  The reference (mjrgh/DCSExplorer) has
 * no C ABI: its decode path is the C++ class DCSDecoderNative behind the abstract
 * DCSDecoder interface (DCSDecoder/DCSDecoder.h:118-1363).  Each entry point below names
 * the reference interface it replaces; INTEGRATION.md shows the reference-side binding.
 *
 * Conventions: plain pointers and sizes, no exceptions, integer status codes, inputs are
 * borrowed for the duration of a call (dcsb_batch_create copies what it needs), outputs
 * are caller-allocated, one in-flight call per context (the reference decoder is
 * single-threaded and non-re-entrant as well, DCSDecoder.h:90-105).
 *
 * There is no CPU fallback: every decode entry point fails with DCSB_E_CUDA when no
 * usable sm_100 device is present.
 */
#ifndef DCSB200_H
#define DCSB200_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes --------------------------------------------------------------- */
#define DCSB_OK             0
#define DCSB_E_EMPTY       -1   /* stream has nFrames == 0 (reference: frame counter wraps, DCSDecoderNative.cpp:1411-1415) */
#define DCSB_E_TRUNCATED   -2   /* frame data runs past the end of the stream bytes */
#define DCSB_E_BANDTYPE    -3   /* band type code leaves 0..15 (reference behaviour undefined, :1932/:2183) */
#define DCSB_E_SHORT       -4   /* fewer bytes than the stream preamble needs */
#define DCSB_E_STOPPED     -5   /* the reference's channel.stop error path fired (:2213-2218); PCM up to and
                                   including that frame is produced, silence after, exactly as the reference */
#define DCSB_E_ARG        -16
#define DCSB_E_CUDA       -17
#define DCSB_E_NOMEM      -18

/* ---- format selectors (DCSEncoder.h CompressionParams::formatVersion; OSVersion in DCSDecoder.h:846) */
#define DCSB_OS93A 0x9301
#define DCSB_OS93B 0x9302
#define DCSB_OS94  0x9400
#define DCSB_OS95  0x9500   /* same stream layout as DCSB_OS94 */

typedef struct dcsb_ctx   dcsb_ctx;     /* one per (process, device) */
typedef struct dcsb_batch dcsb_batch;   /* a set of streams resident in HBM */

/* One stream to decode with the reference's canonical "one stream -> PCM" protocol
 * (DCSExplorer/DCSExplorer.cpp:1655-1721, DCSEncoder/DCSEncoder.cpp:547-571): a fresh
 * DCSDecoderNative, InitStandalone(os_version), SoftBoot(), SetMasterVolume(master_volume),
 * LoadAudioStream(0, data, mixing_level), then (nFrames + tail_frames) * 240 samples. */
typedef struct {
    const uint8_t *data;        /* U16BE nFrames, 16-byte header (1 byte for OS93a type 1), frame bits */
    uint32_t nbytes;
    uint16_t os_version;        /* DCSB_OS93A / OS93B / OS94 / OS95 */
    uint8_t  master_volume;     /* DCSDecoder::SetMasterVolume, DCSDecoder.h:546 */
    uint8_t  mixing_level;      /* DCSDecoderNative::LoadAudioStream mixing level, DCSDecoderNative.h:98 */
    uint16_t tail_frames;       /* frames rendered after the last stream frame (DCSExplorer uses 2) */
    uint16_t reserved;
} dcsb_stream_desc;

typedef struct {
    int32_t  status;            /* DCSB_OK or the DCSB_E_* that ended the stream early */
    uint32_t frames;            /* frames rendered = nFrames + tail_frames */
    uint32_t frames_decoded;    /* stream frames actually decoded before silence */
    uint32_t stream_bytes;      /* bytes the stream really occupies = StreamInfo::nBytes (DCSDecoderNative.h:106-123) */
    uint64_t checksum;          /* sum over samples i of (uint16)s[i] * (2*i+1) mod 2^64 (order-sensitive, parallel) */
} dcsb_result;

/* ---- context ---------------------------------------------------------------------- */
/* Replaces: constructing a DCSDecoderNative (DCSDecoderNative.cpp:22) + Initialize (:3143). */
int  dcsb_create(int cuda_device, dcsb_ctx **out);
void dcsb_destroy(dcsb_ctx *ctx);
const char *dcsb_last_error(const dcsb_ctx *ctx);    /* DCSDecoder::GetErrorMessage, DCSDecoder.h:213-222 */
const char *dcsb_version(void);

/* ---- one-shot batch decode with HOST buffers ------------------------------------- */
/* Replaces: the per-stream loop in DCSExplorer ExtractTracksOrStreams (DCSExplorer.cpp:1628-1939).
 * pcm_offsets[i] is the sample offset of stream i inside pcm_out (NULL = tightly packed in
 * order; stream i always occupies (U16BE(data) + tail_frames) * 240 samples, silence where
 * nothing decodes).  results may be NULL.  Host->device and device->host copies happen inside. */
int dcsb_decode_streams(dcsb_ctx *ctx, const dcsb_stream_desc *descs, size_t n,
                        int16_t *pcm_out, const uint64_t *pcm_offsets, dcsb_result *results);

/* ---- resident batches (upload once, decode many; device-pointer output) ---------- */
/* Replaces: LoadAudioStream on many decoder instances (DCSDecoderNative.cpp:1387-1463). */
int  dcsb_batch_create(dcsb_ctx *ctx, const dcsb_stream_desc *descs, size_t n, dcsb_batch **out);
void dcsb_batch_destroy(dcsb_batch *b);
uint64_t dcsb_batch_total_samples(const dcsb_batch *b);        /* PCM samples the batch renders */
uint64_t dcsb_batch_total_frames(const dcsb_batch *b);
uint64_t dcsb_batch_compressed_bytes(const dcsb_batch *b);     /* sum of nbytes */
uint64_t dcsb_batch_pcm_offset(const dcsb_batch *b, size_t i); /* sample offset of stream i (tightly packed) */

/* Run the hot path (frame-boundary scan + decode/transform/writeback kernels) on
 * cuda_stream (a cudaStream_t cast to void*, NULL = default stream).  d_pcm is a DEVICE
 * pointer to dcsb_batch_total_samples() int16 samples, or NULL to use an internal buffer.
 * Asynchronous: returns after enqueueing.  Replaces: DCSDecoderNative::MainLoop x frames
 * (DCSDecoderNative.cpp:89-306) drained by GetNextSample (DCSDecoder.cpp:1579-1690). */
int dcsb_batch_decode(dcsb_batch *b, void *d_pcm, void *cuda_stream);
/* number of kernels one dcsb_batch_decode enqueues */
int dcsb_batch_launches(const dcsb_batch *b);
/* wait for cuda_stream, then fetch per-stream results (status/checksum) */
int dcsb_batch_results(dcsb_batch *b, void *cuda_stream, dcsb_result *results);
/* copy stream i's PCM (from the internal buffer) to host */
int dcsb_batch_read_pcm(dcsb_batch *b, size_t i, int16_t *pcm, size_t max_samples);
/* device pointer of the internal PCM buffer (NULL until first decode into it) */
void *dcsb_batch_device_pcm(dcsb_batch *b);
/* time of the most recent dcsb_batch_decode per kernel, measured with CUDA events on the
 * launching stream (ms); which: 0 = scan, 1 = decode+transform */
float dcsb_batch_last_kernel_ms(dcsb_batch *b, int which);

/* Frame checkpoints produced by the scan kernel, for stage-level parity tests against
 * GetStreamInfo / the per-frame bit pointer (DCSDecoderNative.cpp:1486-1537):
 * bitpos[f] for f < nFrames, band types (16 bytes per frame) carried INTO each frame. */
int dcsb_batch_read_scan(dcsb_batch *b, size_t i, uint32_t *bitpos, uint8_t *bandtypes, size_t max_frames);

/* ---- gain helpers (host side) ------------------------------------ */
uint16_t dcsb_master_multiplier(int vol);                                  /* SetMasterVolume, :3250-3282 */
uint16_t dcsb_level_multiplier(int level_sum, int os_version, int channel_volume, int max_override); /* :3071-3121 */
int dcsb_gain_stage(const uint16_t mix_mult[8], unsigned active_mask, unsigned max_override_mask,
                    uint16_t vol_mult, uint16_t eff_mult[8]);                /* MainLoop :227-269 */

#ifdef __cplusplus
}
#endif
#endif
