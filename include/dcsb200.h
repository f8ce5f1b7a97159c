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
    uint16_t reserved;          /* flags: DCSB_STREAM_WRAP_EMPTY, else 0 */
} dcsb_stream_desc;
/* a zero frame count plays 65,536 frames, as the reference's wrapped 16-bit frame counter makes it
 * (DCSDecoderNative.cpp:1411-1415, :1565) -- what track playback from a ROM does; without the flag
 * such a stream is rejected with DCSB_E_EMPTY.  Honoured by dcsb_batch_create (whose PCM layout comes
 * from dcsb_batch_pcm_offset / dcsb_batch_total_samples) and used internally by ROM playback;
 * dcsb_decode_streams, whose caller sizes pcm_out from the count as written, refuses it (DCSB_E_ARG). */
#define DCSB_STREAM_WRAP_EMPTY 1

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
/* on (default): the frame-boundary scan runs on an internal stream beside the decode kernel, whose
 * warps wait for the checkpoints they need; off: one kernel after the other on the caller's stream
 * (what a profiler sees anyway, and what per-kernel timings should be read from) */
int dcsb_set_overlap(dcsb_ctx *ctx, int on);

/* Tuning of dcsb_decode_streams' pipeline (defaults: 0, 0).  max_chunks: how many chunks of streams
 * the batch is cut into (1..8; each chunk is uploaded, decoded and downloaded on its own CUDA
 * stream).  slice_frames: a chunk whose streams all render the same number of frames into a packed,
 * pinned output is also cut in TIME -- frames [k*slice, (k+1)*slice) of every stream are scanned
 * (resuming from the previous slice's checkpoints), decoded and copied out as one strided copy while
 * the next slice is being scanned, so the PCM starts to drain over PCIe after a fraction of the
 * per-stream scan chain; > 0 = that many frames per slice, 0 = choose, < 0 = never slice. */
int dcsb_set_pipeline(dcsb_ctx *ctx, int max_chunks, int slice_frames);

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
/* launch shape of the batch's kernels, for matching a profiler capture with what runs (which: 0 = the
 * frame-boundary scan, 1 = the 1994-layout decode kernel of the one-after-the-other mode, 2 = the persistent
 * 1994-layout decode kernel that runs beside the scan): grid and block size in threads; DCSB_E_ARG if unknown */
int dcsb_batch_launch_shape(const dcsb_batch *b, int which, int *grid, int *block);
/* wait for cuda_stream, then fetch per-stream results (status/checksum) */
int dcsb_batch_results(dcsb_batch *b, void *cuda_stream, dcsb_result *results);
/* copy stream i's PCM (from the internal buffer) to host */
int dcsb_batch_read_pcm(dcsb_batch *b, size_t i, int16_t *pcm, size_t max_samples);
/* device pointer of the internal PCM buffer (NULL until first decode into it) */
void *dcsb_batch_device_pcm(dcsb_batch *b);
/* time of the most recent dcsb_batch_decode, measured with CUDA events on the launching streams
 * (ms); which: 0 = scan span, 1 = decode+transform span (overlaps the scan unless
 * dcsb_set_overlap(ctx, 0)), 2 = the whole step */
float dcsb_batch_last_kernel_ms(dcsb_batch *b, int which);

/* Frame checkpoints produced by the scan kernel, for stage-level parity tests against
 * GetStreamInfo / the per-frame bit pointer (DCSDecoderNative.cpp:1486-1537):
 * bitpos[f] for f < nFrames, band types (16 bytes per frame) carried INTO each frame. */
int dcsb_batch_read_scan(dcsb_batch *b, size_t i, uint32_t *bitpos, uint8_t *bandtypes, size_t max_frames);

/* ---- ROM sets: load, look up tracks and streams (host side) ------------------------ */
/* Replaces the ROM half of the DCSDecoder API: AddROM (DCSDecoder.h:324), LoadROMFromZipFile
 * (:285, DCSDecoderZipLoader.cpp), CheckROMs (:347), GetVersionInfo / GetVersionNumber /
 * GetNumChannels (:1081-1110), GetMaxTrackNumber (:360), GetTrackInfo (:416), ListStreams
 * (:486), MakeROMPointer (:792).  A dcsb_rom copies the images it is given. */
typedef struct dcsb_rom dcsb_rom;

#define DCSB_ZIP_OK        0    /* ZipLoadStatus::Success (DCSDecoder.h:278-284) */
#define DCSB_ZIP_E_OPEN    1    /* OpenFileError */
#define DCSB_ZIP_E_EXTRACT 2    /* ExtractError */
#define DCSB_ZIP_E_NOU2    3    /* NoU2 */

typedef struct {
    uint16_t os_version;        /* DCSB_OS93A / OS93B / OS94 / OS95; 0 = not detected */
    uint8_t  hw_version;        /* 2 = DCS audio board (1993), 3 = DCS-95 A/V board, 1 = not detected, 0 = not checked */
    uint8_t  n_channels;        /* GetNumChannels(): 0 when the code pattern is absent */
    uint16_t version_number;    /* GetVersionNumber(): 0x0100, 0x0101, 0x0103..0x0105, 0 */
    uint16_t n_tracks;          /* entries in the track index = GetMaxTrackNumber() + 1 */
    uint32_t catalog_offset;    /* offset of the catalog inside U2 (0 = not found) */
    int32_t  post_code;         /* CheckROMs(): 1 = all chips verified, 2..9 = first failing chip U2..U9 */
    char     signature[128];    /* U2 signature string */
} dcsb_rom_info;

typedef struct {                /* DCSDecoder::TrackInfo (DCSDecoder.h:384-414) */
    uint32_t address;           /* 24-bit linear ROM address of the track */
    int32_t  channel;
    int32_t  type;              /* 1 = byte-code program, 2 = deferred, 3 = deferred indirect */
    uint16_t defer_code;        /* types 2/3; 0xFFFF otherwise */
    uint8_t  looping;
    uint8_t  reserved;
    uint32_t time;              /* playing time in frames (7.68 ms) */
} dcsb_track_info;

int  dcsb_rom_create(dcsb_rom **out);
void dcsb_rom_destroy(dcsb_rom *rom);
int  dcsb_rom_add(dcsb_rom *rom, int chip_number /* 2..9 */, const uint8_t *image, size_t nbytes);
/* returns DCSB_ZIP_*; explicit_u2 may be NULL; details in dcsb_rom_last_error */
int  dcsb_rom_load_zip(dcsb_rom *rom, const char *zip_path, const char *explicit_u2);
/* the files of the zip most recently loaded, as LoadROMFromZipFile hands them to its caller in its
 * std::list<ZipFileData> (DCSDecoder.h:225-235, :285-288): name, inflated bytes (owned by the rom object, valid
 * until the next load / destroy) and the chip the loader took the file for (2..9, or -1: not a ROM image).
 * Writes up to max records, returns how many files the zip held. */
typedef struct { const char *name; const uint8_t *data; size_t size; int32_t chip_number; } dcsb_zip_file;
size_t dcsb_rom_zip_files(const dcsb_rom *rom, dcsb_zip_file *files, size_t max);
int  dcsb_rom_check(dcsb_rom *rom);                                    /* CheckROMs(): POST code */
int  dcsb_rom_get_info(const dcsb_rom *rom, dcsb_rom_info *info);
int  dcsb_rom_track_info(const dcsb_rom *rom, uint16_t track, dcsb_track_info *info);   /* 1 = valid track, 0 = not */
/* DecompileTrackProgram (DCSDecoder.h:432-481, DCSDecoder.cpp:885-1135): the steps of a type-1 track program.
 * Writes up to max steps, returns how many the program has (0: no such track / not a type-1 track). */
typedef struct dcsb_opcode {
    int32_t  offset;            /* byte offset of the step's delay count inside the track */
    int32_t  nesting_level;     /* loops around the step */
    int32_t  loop_parent;       /* as the reference numbers it: 1 + index of the enclosing loop's step, -1 = top level */
    uint16_t delay_count;       /* frames to wait before the step; 0xFFFF = forever */
    uint8_t  opcode;
    uint8_t  n_operand_bytes;
    uint8_t  operand_bytes[8];
    char     desc[64];          /* mnemonic form, the reference's wording */
    char     hex_desc[40];      /* count, opcode and operands as grouped hex numbers */
} dcsb_opcode;
size_t dcsb_rom_decompile_track(const dcsb_rom *rom, uint16_t track, dcsb_opcode *steps, size_t max);
/* distinct stream addresses referenced by Play opcodes, ascending; returns how many exist */
size_t dcsb_rom_list_streams(const dcsb_rom *rom, uint32_t *addresses, size_t max);
/* MakeROMPointer: pointer into the rom's own copy of the chip + bytes left in that chip */
const uint8_t *dcsb_rom_pointer(const dcsb_rom *rom, uint32_t linear_address, uint32_t *bytes_left);
const char *dcsb_rom_last_error(const dcsb_rom *rom);

/* Size of a stream whose caller does not know it (the reference's clients hand the decoder an unsized
 * ROMPointer and let it read as far as the bits go: DCSEncoder.cpp:547-571; its own GetStreamInfo,
 * DCSDecoderNative.cpp:1486-1537, finds the size by walking every frame).  HOST side, lengths only (no PCM is
 * produced: this is stream validation, like reading the frame count): walks the frames the preamble counts
 * and returns the bytes the stream occupies, 0 if it is not decodable to its end.  Reads at most
 * DCSB_EXTENT_SLACK bytes past the stream's last byte (the reference reads 1). */
#define DCSB_EXTENT_SLACK 16
size_t dcsb_stream_extent(const uint8_t *data, int os_version);

/* ---- track playback: one decoder instance = one player ------------------------------ */
/* A player is the control state of one DCSDecoderNative (channels, track programs, command
 * queue, mixer levels and fades, master volume) plus the 16-sample overlap carried from frame to
 * frame.  The host runs the byte-code and gain staging; the GPU decodes, mixes (channels 0..7 in
 * order, in the frequency domain, DCSDecoderNative.cpp:272-278), transforms and writes PCM.
 * dcsb_player_create = construct + SoftBoot() (DCSDecoder.h:594); the ROM's streams are uploaded
 * and scanned on first use of the rom with a context. */
typedef struct dcsb_player dcsb_player;
int  dcsb_player_create(dcsb_ctx *ctx, dcsb_rom *rom, dcsb_player **out);
void dcsb_player_destroy(dcsb_player *p);
void dcsb_player_set_master_volume(dcsb_player *p, int vol);          /* SetMasterVolume, DCSDecoder.h:546 */
void dcsb_player_write_data_port(dcsb_player *p, uint8_t byte);       /* WriteDataPort, DCSDecoder.h:663 */
void dcsb_player_add_track_command(dcsb_player *p, uint16_t track);   /* AddTrackCommand, DCSDecoderNative.h:129 */
int  dcsb_player_load_audio_stream(dcsb_player *p, int channel, uint32_t stream_address, int mixing_level); /* :98 */
void dcsb_player_clear_tracks(dcsb_player *p);                        /* ClearTracks, DCSDecoderNative.h:126 */
int  dcsb_player_is_stream_playing(const dcsb_player *p, int channel);/* IsStreamPlaying, DCSDecoderNative.h:101 */
/* GetStreamInfo (DCSDecoderNative.h:106-123) from the GPU scan of the ROM's streams.  n_bytes is the
 * exact size of the stream (count + header + frame bits rounded up to a byte); the reference reports
 * up to 2 bytes more, depending on how far its bit reader had prefetched at the last sample. */
typedef struct {
    int32_t n_frames, n_bytes, stream_type, stream_subtype;
    int32_t status;                     /* DCSB_OK or the DCSB_E_* the scan found */
    uint8_t header[16];
} dcsb_stream_info;
int  dcsb_player_stream_info(const dcsb_player *p, uint32_t stream_address, dcsb_stream_info *info);
/* render the next n_frames * 240 samples into HOST memory (the GetNextSample pump,
 * DCSDecoder.cpp:1579-1690, n_frames main-loop passes at once) */
int  dcsb_player_render(dcsb_player *p, uint32_t n_frames, int16_t *pcm_out);
/* Render ahead: dcsb_player_render then renders n_frames at a time (one launch and one copy per block) and
 * hands frames out of the block; an input that arrives in the middle of a block (data port, volume, track
 * command, LoadAudioStream, ClearTracks) still takes effect at the frame it arrives at, as in the reference
 * (DCSDecoder.cpp:1625-1631: the port is drained before every main-loop pass) -- the unconsumed frames are
 * dropped and rendered anew.  0 (the default) = render exactly what each call asks for. */
int  dcsb_player_set_lookahead(dcsb_player *p, uint32_t n_frames);
/* bytes the decoder sent to the host since the last call (Host::ReceiveDataPort); returns count */
size_t dcsb_player_host_bytes(dcsb_player *p, uint8_t *out, size_t max);

/* One timeline = a fresh player fed data-port bytes at given frame numbers.  Many timelines are
 * rendered in one launch; timeline t's PCM starts at pcm_offsets[t] samples (NULL = packed). */
typedef struct { uint32_t frame; uint8_t byte; uint8_t pad[3]; } dcsb_port_write;   /* byte is written before frame `frame` renders */
typedef struct {
    const dcsb_port_write *writes;      /* sorted by frame */
    uint32_t n_writes;
    uint32_t n_frames;                  /* frames to render */
    uint8_t  master_volume;             /* SetMasterVolume after SoftBoot */
    uint8_t  pad[3];
} dcsb_timeline;
typedef struct {
    int32_t  status;                    /* DCSB_OK, or DCSB_E_STOPPED-class: the decoder hit its fatal-error state */
    uint32_t frames;
    uint64_t checksum;                  /* same definition as dcsb_result::checksum */
    uint32_t n_host_bytes;              /* bytes sent to the host during the timeline */
    uint32_t reserved;
} dcsb_timeline_result;
int dcsb_render_timelines(dcsb_ctx *ctx, dcsb_rom *rom, const dcsb_timeline *timelines, size_t n,
                          int16_t *pcm_out, const uint64_t *pcm_offsets, dcsb_timeline_result *results);

/* ---- output containers (host side): the step after the decode path ------------------ */
/* 44-byte RIFF/WAVE header + mono 16-bit PCM at 31,250 Hz, as ExtractToWAV writes it
 * (DCSExplorer.cpp:1686-1712).  Returns DCSB_OK or DCSB_E_ARG (cannot create / write). */
int dcsb_write_wav(const char *path, const int16_t *pcm, size_t n_samples);
/* Raw stream container: 36-byte "DCSa" header (format version $9301 / $9302 / $9400, 1 channel, 31,250 Hz,
 * 22 reserved bytes, U32BE data size) + the stream bytes (DCSExplorer.cpp:1831-1871; read back by
 * DCSEncoder::IsDCSFile / EncodeDCSFile, DCSEncoder.cpp:369-399). */
int dcsb_write_dcs_file(const char *path, uint16_t os_version, const uint8_t *stream, size_t nbytes);
/* Reads a "DCSa" file: *os_version = format version (DCSB_OS94 for $9400), stream bytes copied to out
 * (up to max); returns the stream size in bytes, or a negative DCSB_E_* (not a DCSa file / unreadable). */
long long dcsb_read_dcs_file(const char *path, uint16_t *os_version, uint8_t *out, size_t max);

/* ---- forward path: PCM -> DCS streams, in batch ------------------------------------------
 * The reference's DCSEncoder (DCSEncoder/DCSEncoder.cpp) minus its resampler: the clip must already be mono float
 * PCM at 31,250 Hz (full scale = 1.0), and is framed directly (16 samples of overlap + 240 new ones per frame, the
 * last frame zero padded).  Everything behind that -- window, transform, per-band statistics, power cut, scale codes,
 * band-type search, header and sample codes, bit packing -- runs on the GPU with the reference's float operations
 * in the reference's order, so the stream BYTES are the ones the reference's CloseStream() produces from the same
 * frames (DCSEncoder.cpp:717-858; checked by tests/test_gpu_encode.py against oracle/_ref).
 * The fields mirror DCSEncoder::CompressionParams (DCSEncoder.h:70-180), including the wildcard: -1 as stream type
 * and / or subtype encodes the clip with every matching format of {0.0, 0.3, 1.0, 1.3} and keeps the first of the
 * smallest streams, as CloseStream() does (DCSEncoder.cpp:779-836). */
typedef struct dcsb_encode_params {
    int32_t stream_type;            /* 0 | 1 | -1 (any) */
    int32_t stream_subtype;         /* 0 | 3 | -1 (any) */
    int32_t target_bit_rate;        /* bits per second, default 128000 */
    float power_band_cutoff;        /* default 0.97 */
    float max_quantization_error;   /* default 10 / 32768 */
    float min_dynamic_range;        /* default 10 / 32768 */
    int32_t format_version;         /* 0 or DCSB_OS94: the 1994 layout; DCSB_OS93A / DCSB_OS93B: the 1993 layout (CompressFrame93b,
                                     * DCSEncoder.cpp:2053-2473): subtype 0 (or -1), stream type 0, or 1 with DCSB_OS93B (the
                                     * reference has no encoder for OS93a type 1 either, :808-815) */
} dcsb_encode_params;
/* bytes a stream of n_samples samples can need at most (size `out` with the sum over the clips) */
uint64_t dcsb_encode_bound(uint64_t n_samples);
/* Encodes n clips.  pcm[i]: n_samples[i] floats (host memory); params[i]: that clip's parameters.  The streams are
 * written back to back into out (capacity out_capacity bytes); stream i occupies [out_offsets[i], out_offsets[i + 1])
 * (out_offsets has n + 1 entries) and starts with the U16BE frame count and the 16 header bytes, ready for
 * dcsb_decode_streams / the reference decoder.  frames_out (may be NULL): the transformed frames, 256 floats each
 * (DCSEncoder.h:322), all clips back to back -- for tests; explicit stream types only.
 * DCSB_OK, DCSB_E_ARG (empty clip, more than 65,535 frames, bad type), DCSB_E_NOMEM (out too small), DCSB_E_CUDA. */
int dcsb_encode_streams(dcsb_ctx *ctx, const float *const *pcm, const uint64_t *n_samples, size_t n,
                        const dcsb_encode_params *params, uint8_t *out, uint64_t out_capacity, uint64_t *out_offsets,
                        float *frames_out);

/* ---- multi-GPU work partitioning (host side) ---------------------------------------- */
/* Streams are independent (one DCSDecoderNative instance each in the reference), so a batch is
 * sharded by stream with no data-path collective: longest-processing-time greedy on the frame
 * counts, ties broken by index, so every rank computes the same assignment.  part_out[i] =
 * the part (0..n_parts-1) stream i goes to; frames_per_part (may be NULL) receives the load. */
int dcsb_partition_streams(const uint32_t *frames, size_t n, int n_parts, uint32_t *part_out, uint64_t *frames_per_part);

/* ---- gain helpers (host side) ------------------------------------ */
uint16_t dcsb_master_multiplier(int vol);                                  /* SetMasterVolume, :3250-3282 */
uint16_t dcsb_level_multiplier(int level_sum, int os_version, int channel_volume, int max_override); /* :3071-3121 */
int dcsb_gain_stage(const uint16_t mix_mult[8], unsigned active_mask, unsigned max_override_mask,
                    uint16_t vol_mult, uint16_t eff_mult[8]);                /* MainLoop :227-269 */

#ifdef __cplusplus
}
#endif
#endif
