/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the reference's
 * "compressed DCS stream bytes -> int16 PCM" path, written from the format as
 * implemented by /root/reference/DCSDecoder/DCSDecoderNative.cpp (each function
 * cites the lines it follows).  It is the checker for the CUDA path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * use it.  The product (dcsexplorer_b200/) never links or loads this file.
 *
 * Parity status: the reference ships no golden vectors / KATs for this path
 * (SURVEY.md section 8c), so this restatement is pinned against outputs of the
 * reference itself compiled here (oracle/_ref, see oracle/Makefile) and against the
 * fixtures under tests/golden/ that were generated from it (tests/golden/make_golden.py).
 */
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCSO_OS93A 0x9301
#define DCSO_OS93B 0x9302
#define DCSO_OS94  0x9400
#define DCSO_OS95  0x9500

/* per-frame checkpoint produced by the scan: where the frame starts and the
 * band-type state carried INTO it (DCSDecoderNative.h:364-454) */
typedef struct {
    uint32_t bitpos;     /* bits from the first byte after the stream header */
    uint8_t  bt[16];     /* band type codes before this frame's header deltas */
} dcso_frame_t;

/* status codes (shared numbering with include/dcsb200.h) */
#define DCSO_OK            0
#define DCSO_E_EMPTY      -1   /* nFrames == 0 */
#define DCSO_E_TRUNCATED  -2   /* frame data runs past nbytes */
#define DCSO_E_BANDTYPE   -3   /* band type code outside 0..15 (reference behaviour undefined) */
#define DCSO_E_SHORT      -4   /* fewer bytes than the preamble needs */

/* header length: 16, or 1 for OS93a type-1 streams (DCSDecoderNative.cpp:1457) */
int dcso_header_len(const uint8_t *stream, int os);

/* Frame-boundary scan.  frames[] must hold nFrames+1 entries; entry nFrames gets the
 * end position.  *stop_frame = index of the first frame that raises the reference's
 * channel.stop flag (DCSDecoderNative.cpp:2213-2218), or -1.  Returns nFrames (>0) or
 * a negative DCSO_E_* code; on DCSO_E_BANDTYPE/TRUNCATED *stop_frame holds the first
 * undecodable frame (frames before it are valid). */
int dcso_scan(const uint8_t *stream, size_t nbytes, int os, dcso_frame_t *frames, int *stop_frame);

/* Decode one frame from its checkpoint, ADDING its contribution (scaled by the
 * channel's effective mixing multiplier) into fb[512] exactly as DecompressFrame does.
 * *stop is set when the frame raises channel.stop. */
void dcso_decode_frame(const uint8_t *stream, int os, const dcso_frame_t *f, uint16_t mult,
                       uint16_t *fb, int *stop);

/* TransformFrame (1994: DCSDecoderNative.cpp:397-576; 1993: :614-813).
 * fb[512] in (destroyed), ovl[16] in/out, pcm[240] out. */
void dcso_transform(int os, uint16_t *fb, uint16_t *ovl, int vol_shift, int16_t *pcm);

/* gain helpers */
uint16_t dcso_master_multiplier(int vol);                                    /* :3250-3282 */
uint16_t dcso_level_multiplier(int level_sum, int os, int chan_vol, int max_override); /* :3071-3121 */
int      dcso_calc_exp32(uint32_t x);                                         /* :3447-3459 */
/* gain staging for one frame (:227-269): in = per-channel multipliers + active mask,
 * out = effective multipliers; returns volShift. */
int dcso_gain_stage(const uint16_t mix_mult[8], unsigned active_mask, unsigned max_override_mask,
                    uint16_t vol_mult, uint16_t eff_mult[8]);

/* Whole single-stream protocol: fresh decoder, master volume
 * vol, stream loaded on channel 0 at mixing level `level`, n_frames_out*240 samples.
 * Returns DCSO_OK or the DCSO_E_* code that ended the stream early (PCM is still
 * fully written: silence after the failure point, like the reference after stop). */
int dcso_decode_stream(const uint8_t *stream, size_t nbytes, int os, int vol, int level,
                       int n_frames_out, int16_t *pcm);

/* FNV-1a 64 over the little-endian bytes of pcm[] */
uint64_t dcso_fnv1a(const int16_t *pcm, size_t n);

#ifdef __cplusplus
}
#endif
