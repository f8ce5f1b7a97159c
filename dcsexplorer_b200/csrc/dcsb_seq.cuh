// dcsb200 sequencer core: everything DCSDecoderNative::MainLoop does EXCEPT touching audio bits --
// stop flags, the command queue, the track byte-code interpreter (all 19 opcodes incl. the OS93a /
// 1.05 variants), loops, deferred and indirect-deferred links, mixing-level ops and fades,
// UpdateMixingLevels, the data-port state machine (IRQ2Handler), self-reset / fatal-error handling,
// host event timers and the per-frame gain staging.  Its product per frame is a mix schedule:
// volume shift + up to 8 entries {stream, frame, effective multiplier} in channel order.
//
// Written once with fixed-size state and no exceptions, compiled twice:
//   * by the host compiler for the DCSDecoder-style player (one instance, dcsb_rom.cpp: DcsbSequencer),
//   * by nvcc for sm_100a as the body of dcsb_seq_kernel (dcsb_kernels.cu): one THREAD per decoder
//     instance (timeline), so that dcsb_render_timelines needs no host round trip for the schedule
//     (SURVEY 8(f) rank 3).
//
// Reference behaviour followed (restated, not copied): DCSDecoderNative.cpp:89-306 (MainLoop),
// :826-1228 (LoadTrack, ExecTrack), :1241-1371 (loops, mixing level ops), :1387-1463 (stream load),
// :1546-1589 (DecodeStream), :3042-3135 (UpdateMixingLevels), :3250-3282 (SetMasterVolume),
// :3297-3437 (IRQ2Handler); DCSDecoder.cpp:1631-1668 (self-reset retries).
#pragma once
#include <stdint.h>
#include "../../include/dcsb200.h"

#if defined(__CUDACC__)
#define DCSB_SEQ_HD __host__ __device__ __forceinline__
#else
#define DCSB_SEQ_HD inline
#endif

#ifndef DCSB_MAX_CHANNELS
#define DCSB_MAX_CHANNELS 8
#endif
#define DCSB_SEQ_MAX_LOOPS 16                  // nested loops per track program (deeper: treated like a bad opcode)
#define DCSB_SEQ_QCAP 256                      // queued track commands (more: treated like a runaway program)
#define DCSB_MAX_STEPS_PER_FRAME 65536u        // track-program steps in one frame before the program counts as runaway
#define DCSB_SEQ_FRAME_HOST_BYTES 32           // host bytes one frame can hand back with their values (more are only counted)

enum { DCSB_HW_UNKNOWN = 0, DCSB_HW_INVALID = 1, DCSB_HW_DCS93 = 2, DCSB_HW_DCS95 = 3 };

// cursor into one chip image; chip < 0 is the null pointer
struct DcsbRomPtr {
    int chip = -1;
    uint32_t ofs = 0;
    DCSB_SEQ_HD bool null() const { return chip < 0; }
    DCSB_SEQ_HD void clear() { chip = -1; ofs = 0; }
    DCSB_SEQ_HD bool operator==(const DcsbRomPtr &o) const { return chip == o.chip && ofs == o.ofs; }
};

// One output frame's worth of mixing work
struct DcsbSchedEntry { uint32_t stream; uint16_t frame; uint16_t mult; };
#define DCSB_FRAME_MUTE 1       // the decoder is in its fatal-error state: pure silence, no overlap tail
struct DcsbSchedFrame { uint32_t first_entry; uint8_t n_entries; uint8_t vs; uint8_t flags; uint8_t pad; };

// what the sequencer needs to know about a stream without decoding it (from the GPU scan of the ROM's streams)
struct DcsbSeqStream { uint32_t linear; uint32_t nplay; int32_t status; uint32_t id; };      // id: the stream's index in the ROM's resident batch (what a mix entry names)

// A ROM set as the sequencer reads it: plain pointers and numbers, valid on the host (dcsb_rom's own copy)
// and on the device (the slab that mirrors the chips back to back).
struct DcsbRomView {
    const uint8_t *image;               // all chips back to back
    uint32_t chip_ofs[8], chip_size[8], chip_mask[8];
    uint8_t present[8];
    int hw, os;
    uint16_t nominal_version, n_tracks;
    uint8_t totan;
    uint32_t track_index, indirect_index;       // offsets inside U2
    const DcsbSeqStream *streams;       // sorted by linear address
    uint32_t n_streams;
};

DCSB_SEQ_HD DcsbRomPtr dcsb_rv_ptr(const DcsbRomView &rv, uint32_t linear)
{
    // chip select in bits 21-23 on the DCS-95 board, 20-22 on the original one (DCSDecoder.cpp:67-76)
    DcsbRomPtr p;
    p.chip = (int)((linear >> (rv.hw == DCSB_HW_DCS95 ? 21 : 20)) & 7);
    p.ofs = linear & (rv.present[p.chip] ? rv.chip_mask[p.chip] : 0x1FFFu);     // absent chips read as 8 KB of $FF
    return p;
}
DCSB_SEQ_HD uint8_t dcsb_rv_u8(const DcsbRomView &rv, const DcsbRomPtr &p, uint32_t d = 0)
{
    if (p.chip < 0) return 0xFF;
    const uint64_t o = (uint64_t)p.ofs + d;
    return (rv.present[p.chip] && o < rv.chip_size[p.chip]) ? rv.image[rv.chip_ofs[p.chip] + o] : 0xFF;
}
DCSB_SEQ_HD uint32_t dcsb_rv_be(const DcsbRomView &rv, const DcsbRomPtr &p, int nbytes, uint32_t d = 0)
{
    uint32_t v = 0;
    for (int i = 0; i < nbytes; ++i) v = (v << 8) | dcsb_rv_u8(rv, p, d + (uint32_t)i);
    return v;
}
DCSB_SEQ_HD uint32_t dcsb_rv_u2_be(const DcsbRomView &rv, uint32_t ofs, int nbytes)
{
    uint32_t v = 0;
    for (int i = 0; i < nbytes; ++i) v = (v << 8) | ((rv.present[0] && ofs + i < rv.chip_size[0]) ? rv.image[rv.chip_ofs[0] + ofs + i] : 0xFFu);
    return v;
}
// index of the stream at a 24-bit ROM address, 0xFFFFFFFF if the ROM scan never saw one there
DCSB_SEQ_HD uint32_t dcsb_rv_stream(const DcsbRomView &rv, uint32_t linear)
{
    linear &= 0xFFFFFFu;
    uint32_t lo = 0, hi = rv.n_streams;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (rv.streams[mid].linear < linear) lo = mid + 1; else hi = mid;
    }
    return (lo < rv.n_streams && rv.streams[lo].linear == linear) ? lo : 0xFFFFFFFFu;
}

// ---- gain arithmetic -------------------------------------------------------
DCSB_SEQ_HD int dcsb_seq_clz(uint32_t v)        // leading zeros, 32 for 0
{
#if defined(__CUDA_ARCH__)
    return __clz((int)v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}
// ADSP-2105 EXP on a 32-bit mantissa (DCSDecoderNative.cpp:3447-3459): minus the number of redundant sign
// bits -- the reference shifts left until bit 30 differs from the sign (at most 31 times); here counted
DCSB_SEQ_HD int dcsb_seq_exp32(uint32_t x)
{
    const int n = dcsb_seq_clz(((x & 0x80000000u) ? ~x : x) << 1);      // sign copies behind bit 31
    return -(n > 31 ? 31 : n);
}
DCSB_SEQ_HD uint16_t dcsb_seq_master_multiplier(int vol)      // SetMasterVolume, :3250-3282
{
    if (vol > 255) vol = 255;
    if (vol <= 0) return 0;
    uint32_t x = 0x3fff, y = 0x7d98;       // 0.5 * 0.981201^(255-vol) in 1.15
    for (int i = 0; i < 8; ++i, vol >>= 1) {
        if (!(vol & 1)) x = ((x * y) >> 15) & 0xFFFFu;
        y = ((y * y) >> 15) & 0xFFFFu;
    }
    return (uint16_t)(x << 1);
}
DCSB_SEQ_HD uint16_t dcsb_seq_level_multiplier(int level_sum, int os_version, int channel_volume, int max_override)   // :3071-3121
{
    level_sum = level_sum < -8191 ? -8191 : (level_sum > 8191 ? 8191 : level_sum);
    const uint32_t e = (uint32_t)((level_sum >> 6) & 0x3FF) + 0x80;
    uint32_t m = os_version == DCSB_OS93A ? 0x7FFFu : ((uint32_t)channel_volume << 7) & 0xFFFFu;
    if (max_override) m = 0xFFu << 7;
    uint32_t p = 0x7C94;                     // 0.9733^(2^j) ladder
    for (int j = 0; j < 8; ++j) {
        if (!(e & (1u << j))) m = ((m * p) >> 15) & 0xFFFFu;
        p = ((p * p) >> 15) & 0xFFFFu;
    }
    return (uint16_t)(m << 1);
}
DCSB_SEQ_HD int dcsb_seq_gain_stage(const uint16_t mix_mult[8], unsigned active_mask, unsigned max_override_mask,
                                    uint16_t vol_mult, uint16_t eff_mult[8])      // MainLoop :227-269
{
    uint64_t sum = 0;
    for (int i = 0; i < 8; ++i) {
        if (max_override_mask & (1u << i)) sum += (uint64_t)mix_mult[i] * 0x7FFE;
        else if (active_mask & (1u << i)) sum += (uint64_t)mix_mult[i] * vol_mult;
    }
    int vs = -(dcsb_seq_exp32((uint32_t)(sum >> 2)) + 3);
    vs = vs < 0 ? 0 : (vs > 8 ? 8 : vs);
    for (int i = 0; i < 8; ++i) {
        const uint64_t v = (max_override_mask & (1u << i)) ? 0x7FFE : vol_mult;
        eff_mult[i] = (uint16_t)(((((uint64_t)mix_mult[i] * v) << 1) << vs) >> 16);
    }
    return vs;
}

// ---- decoder instance state ---------------------------------------------------------------
struct DcsbSeqMixer { int cur, target, delta, steps; };
struct DcsbSeqTimer { uint8_t data; uint16_t interval, counter; };
struct DcsbSeqLoop { uint16_t counter; DcsbRomPtr pos; };
struct DcsbSeqStreamState {
    bool active, at_start;
    uint32_t id;                                // index into rv.streams (0xFFFFFFFF: unknown address, plays silence)
    uint16_t nframes, counter, loops, pos;
};
struct DcsbSeqChannel {
    DcsbRomPtr track;
    uint16_t track_counter;
    uint8_t next_type;
    uint16_t next_link;
    bool stop;
    DcsbSeqStreamState st;
    int source;
    DcsbSeqMixer mixer[DCSB_MAX_CHANNELS];
    uint8_t fading;                             // mixers with a fade in progress (steps != 0), one bit each
    bool max_override;
    uint16_t mult;
    uint32_t level_key;                         // (level sum, volume, override) the cached level_mult belongs to
    uint16_t level_mult;
    DcsbSeqTimer timer;
    uint16_t volume;
    DcsbSeqLoop loops[DCSB_SEQ_MAX_LOOPS];
    uint8_t n_loops;
    bool levels_dirty;                          // a mixer level of this channel moved: level_sum is stale
    int level_sum;                              // sum of the channel's mixer levels (cached)
};
struct DcsbSeqState {
    DcsbSeqChannel chan[DCSB_MAX_CHANNELS];
    uint8_t vars[256];
    uint16_t cmdq[DCSB_SEQ_QCAP];
    uint32_t q_head, q_count;
    uint16_t port_word, port_ext;
    int port_bytes, port_timeout;
    uint16_t vol_mult;
    uint16_t reported_version;                  // DCSDecoderNative.h:168
    unsigned done_mask;
    uint32_t steps_this_frame;                  // track-program steps + queued commands taken in the current main-loop pass
    bool fatal;
    uint32_t frame_no;
    // bytes for the host (Host::ReceiveDataPort) since the sink was last emptied: the first few with their
    // values, all of them counted
    uint8_t host_buf[DCSB_SEQ_FRAME_HOST_BYTES];
    uint32_t host_n;
    uint32_t host_total;
    // gain staging of the last frame: inputs and results (most frames nothing that feeds it has moved)
    uint16_t gs_mix[DCSB_MAX_CHANNELS], gs_eff[DCSB_MAX_CHANNELS];
    uint32_t gs_key;                            // active mask | max-override mask << 8 | volume multiplier << 16
    uint8_t gs_vs;
    bool gs_valid;
    // main-loop passes ahead in which nothing but counters moves (dcsb_seq_horizon); any input zeroes it
    uint32_t quiet;
};

DCSB_SEQ_HD void dcsb_seq_to_host(DcsbSeqState &s, uint8_t b)
{
    if (s.host_n < DCSB_SEQ_FRAME_HOST_BYTES) s.host_buf[s.host_n] = b;
    ++s.host_n;
    ++s.host_total;
}
DCSB_SEQ_HD bool dcsb_seq_q_push(DcsbSeqState &s, uint16_t v)      // false: the queue is full (a runaway program)
{
    s.quiet = 0;
    if (s.q_count >= DCSB_SEQ_QCAP) return false;
    s.cmdq[(s.q_head + s.q_count) % DCSB_SEQ_QCAP] = v;
    ++s.q_count;
    return true;
}
DCSB_SEQ_HD void dcsb_seq_mixer_reset(DcsbSeqMixer &m) { m.cur = m.target = m.steps = 0; }
DCSB_SEQ_HD void dcsb_seq_timer_clear(DcsbSeqTimer &t) { t.interval = t.counter = 0; }

DCSB_SEQ_HD void dcsb_seq_set_master_volume(DcsbSeqState &s, int vol) { s.vol_mult = dcsb_seq_master_multiplier(vol); s.quiet = 0; }

DCSB_SEQ_HD void dcsb_seq_init(DcsbSeqState &s)        // a freshly constructed decoder + Initialize() (SoftBoot)
{
    for (int i = 0; i < DCSB_MAX_CHANNELS; ++i) {
        DcsbSeqChannel &c = s.chan[i];
        c.track.clear();
        c.track_counter = 0;
        c.next_type = 0;
        c.next_link = 0;
        c.stop = false;
        c.st.active = c.st.at_start = false;
        c.st.id = 0xFFFFFFFFu;
        c.st.nframes = c.st.counter = c.st.loops = c.st.pos = 0;
        c.source = -1;
        for (int k = 0; k < DCSB_MAX_CHANNELS; ++k) { c.mixer[k].cur = c.mixer[k].target = c.mixer[k].delta = c.mixer[k].steps = 0; }
        c.fading = 0;
        c.max_override = false;
        c.mult = 0x7FFF;                        // the constructor default frame 0 is mixed with (DCSDecoderNative.h:514)
        c.level_key = 0xFFFFFFFFu;
        c.level_mult = 0;
        c.timer.data = 0;
        c.timer.interval = c.timer.counter = 0;
        c.volume = 0xFF;
        c.n_loops = 0;
        c.levels_dirty = true;
        c.level_sum = 0;
    }
    s.gs_valid = false;
    s.quiet = 0;
    for (int i = 0; i < 256; ++i) s.vars[i] = 0;
    s.q_head = s.q_count = 0;
    s.port_word = s.port_ext = 0;
    s.port_bytes = 0;
    s.port_timeout = 0;
    s.reported_version = 0x0106;
    s.done_mask = 0;
    s.steps_this_frame = 0;
    s.fatal = false;
    s.frame_no = 0;
    s.host_n = s.host_total = 0;
    dcsb_seq_set_master_volume(s, 0x67);        // DCSDecoder's default volume until the host says otherwise
}

DCSB_SEQ_HD void dcsb_seq_reset_mix(DcsbSeqState &s, int ch)
{
    for (int i = 0; i < DCSB_MAX_CHANNELS; ++i) {
        dcsb_seq_mixer_reset(s.chan[i].mixer[ch]);
        s.chan[i].fading &= (uint8_t)~(1u << ch);
        s.chan[i].levels_dirty = true;
    }
}
DCSB_SEQ_HD void dcsb_seq_clear_tracks(DcsbSeqState &s)
{
    s.quiet = 0;
    for (int i = 0; i < DCSB_MAX_CHANNELS; ++i) { s.chan[i].track.clear(); s.chan[i].st.active = false; }
}

// WriteDataPort + IRQ2Handler: a byte from the host, taken before the next frame (:3297-3437)
DCSB_SEQ_HD void dcsb_seq_write_port(DcsbSeqState &s, const DcsbRomView &rv, uint8_t data)
{
    s.quiet = 0;
    if (s.port_timeout >= 13) s.port_bytes = 0;
    switch (s.port_bytes) {
    case 0:
        s.port_word = (uint16_t)(data << 8);
        s.port_bytes = 1;
        break;
    case 1:
        s.port_word |= data;
        if ((s.port_word >= 0x55AA && s.port_word <= 0x55B2) || (s.port_word >= 0x55BA && s.port_word <= 0x55C1)) {
            s.port_ext = s.port_word;
            s.port_bytes = 2;
        } else if (s.port_word > 0x55B2 && s.port_word < 0x55BA) s.port_bytes = 0;
        else if (s.port_word == 0x55C2 || s.port_word == 0x55C3) {
            const uint16_t rep = s.reported_version;
            dcsb_seq_to_host(s, (uint8_t)((s.port_word == 0x55C2 ? rep >> 8 : rep) & 0xFF));
            s.port_bytes = 0;
        } else if (s.port_word & 0x8000) s.port_bytes = 0;
        else if (s.port_word == 0x03E7 && rv.totan) { dcsb_seq_to_host(s, 0x11); s.port_bytes = 0; }
        else { dcsb_seq_q_push(s, s.port_word); s.port_bytes = 0; }      // (a full queue drops the command: the program is runaway and resets anyway)
        break;
    case 2:
        s.port_word = data;
        s.port_bytes = 3;
        break;
    default:
        if (s.port_word == (uint16_t)(data ^ 0xFF)) {
            if (s.port_ext == 0x55AA) dcsb_seq_set_master_volume(s, (uint8_t)s.port_word);
            else if (s.port_ext <= 0x55B2) { const int ch = s.port_ext - 0x55AB; if (ch >= 0 && ch < DCSB_MAX_CHANNELS) s.chan[ch].volume = (uint8_t)s.port_word; }
            // 55BA..55C1 only touch state no audio path reads
        }
        s.port_bytes = 0;
        break;
    }
    s.port_timeout = 0;
}

DCSB_SEQ_HD void dcsb_seq_load_track(DcsbSeqState &s, int ch, DcsbRomPtr p)
{
    DcsbSeqChannel &c = s.chan[ch];
    c.track = p;
    c.st.active = false;
    c.track_counter = 0;
    dcsb_seq_timer_clear(c.timer);
    c.n_loops = 0;
    s.done_mask &= ~(1u << ch);
    dcsb_seq_reset_mix(s, ch);
}

DCSB_SEQ_HD void dcsb_seq_start_stream(DcsbSeqState &s, const DcsbRomView &rv, int sch, int source, int loops, uint32_t linear)
{
    DcsbSeqChannel &c = s.chan[sch];
    const DcsbRomPtr sp = dcsb_rv_ptr(rv, linear);
    DcsbSeqStreamState &st = c.st;
    st.nframes = st.counter = (uint16_t)dcsb_rv_be(rv, sp, 2);
    st.active = true;
    st.at_start = true;
    st.pos = 0;
    st.id = dcsb_rv_stream(rv, linear);
    if (st.nframes == 0) return;            // the reference leaves such a stream playing: its uint16 counter wraps to 65,536 frames
    st.loops = (uint16_t)loops;
    if (c.source >= 0 && c.source != source) { dcsb_seq_mixer_reset(c.mixer[c.source]); c.levels_dirty = true; }
    c.source = source;
}

// LoadAudioStream(ch, ptr, level)
DCSB_SEQ_HD void dcsb_seq_load_stream(DcsbSeqState &s, const DcsbRomView &rv, int ch, uint32_t linear, int level)
{
    if (ch < 0 || ch >= DCSB_MAX_CHANNELS) return;
    s.quiet = 0;
    s.chan[ch].track.clear();
    dcsb_seq_start_stream(s, rv, ch, ch, 1, linear);
    DcsbSeqMixer &m = s.chan[ch].mixer[ch];
    dcsb_seq_mixer_reset(m);
    m.cur = m.target = level << 6;
    s.chan[ch].levels_dirty = true;
}

DCSB_SEQ_HD void dcsb_seq_mix_op(DcsbSeqState &s, const DcsbRomView &rv, int cur, DcsbRomPtr &p, int mode, bool fade)
{
    const int target_ch = dcsb_rv_u8(rv, p) & 7;       // (the reference indexes with the raw byte)
    const int param = (int)(int8_t)dcsb_rv_u8(rv, p, 1) * 64;
    p.ofs += 2;
    int steps = 0;
    if (fade) { steps = (int)dcsb_rv_be(rv, p, 2); p.ofs += 2; }
    DcsbSeqMixer &m = s.chan[target_ch].mixer[cur];
    m.steps = steps;
    if (steps) s.chan[target_ch].fading |= (uint8_t)(1u << cur); else s.chan[target_ch].fading &= (uint8_t)~(1u << cur);
    const int old = m.cur;
    int lvl = mode == 0 ? param : (mode == 1 ? old + param : old - param);
    const int delta = lvl - old;                // taken before the range limit, as the original does
    lvl = lvl < -8191 ? -8191 : (lvl > 8191 ? 8191 : lvl);
    m.target = lvl;
    if (steps != 0) m.delta = delta / steps;
    else m.cur = lvl;
    s.chan[target_ch].levels_dirty = true;
}

// one channel's track program up to its next wait.  Returns false when the decoder must reset itself
// (bad opcode, runaway program): the reference throws ResetException there (:1225)
DCSB_SEQ_HD bool dcsb_seq_exec_track(DcsbSeqState &s, const DcsbRomView &rv, int cur)
{
    DcsbSeqChannel &me = s.chan[cur];
    DcsbRomPtr p = me.track;
    if (p.null()) return true;
    for (;;) {
        // A track program that loops without ever waiting (or queues itself over and over) would keep the
        // reference's MainLoop -- and a whole dcsb_render_timelines batch with it -- busy for ever.  No
        // well-formed program comes near this many steps in one 7.68 ms frame: treat it like the other
        // malformed-program cases (bad opcode / track type): self-reset, fatal after four in a row.
        if (++s.steps_this_frame > DCSB_MAX_STEPS_PER_FRAME) return false;
        const uint32_t wait = dcsb_rv_be(rv, p, 2);
        if (wait == 0xFFFF || me.track_counter != wait) { me.track = p; return true; }
        p.ofs += 2;
        me.track_counter = 0;
        const int op = dcsb_rv_u8(rv, p);
        p.ofs += 1;
        switch (op) {
        case 0x00:
            me.track.clear();
            me.st.active = false;
            me.n_loops = 0;
            dcsb_seq_timer_clear(me.timer);
            dcsb_seq_reset_mix(s, cur);
            return true;
        case 0x01: {
            const int sch = dcsb_rv_u8(rv, p) & 7;
            if (sch == 5) s.chan[5].max_override = false;
            const uint32_t addr = dcsb_rv_be(rv, p, 3, 1);
            const int loops = dcsb_rv_u8(rv, p, 4);
            p.ofs += 5;
            dcsb_seq_start_stream(s, rv, sch, cur, loops, addr);
            break;
        }
        case 0x02: {
            const int t = dcsb_rv_u8(rv, p) & 7;
            p.ofs += 1;
            if (s.chan[t].st.active) { s.chan[t].st.active = false; dcsb_seq_reset_mix(s, t); }
            s.chan[t].track.clear();
            dcsb_seq_timer_clear(s.chan[t].timer);
            if (me.track.null()) return true;
            break;
        }
        case 0x03:
            if (!dcsb_seq_q_push(s, (uint16_t)dcsb_rv_be(rv, p, 2))) return false;
            p.ofs += 2;
            break;
        case 0x04:
            if (rv.os == DCSB_OS93A) {
                const uint8_t b = dcsb_rv_u8(rv, p);
                const uint16_t counter = (uint16_t)dcsb_rv_be(rv, p, 2, 1);
                p.ofs += 3;
                if (b == 0) dcsb_seq_timer_clear(me.timer);
                else {
                    dcsb_seq_to_host(s, b);
                    if (counter) { me.timer.data = b; me.timer.interval = me.timer.counter = counter; }
                    else dcsb_seq_timer_clear(me.timer);
                }
            } else {
                const uint8_t b = dcsb_rv_u8(rv, p);
                p.ofs += 1;
                dcsb_seq_to_host(s, b);
                if (rv.nominal_version == 0x0105) {
                    if (b == 0x69) s.chan[5].max_override = true;
                    else if (b == 0x6A) s.chan[5].max_override = false;
                }
            }
            break;
        case 0x05: {
            const int t = dcsb_rv_u8(rv, p) & 7;
            p.ofs += 1;
            const int type = s.chan[t].next_type;
            if (type == 0) break;
            s.chan[t].next_type = 0;
            if (type == 2) { if (!dcsb_seq_q_push(s, s.chan[t].next_link)) return false; }
            else if (type == 3) {
                // Catalog[$43][low byte][variables[high byte]] -> track number
                const uint16_t link = s.chan[t].next_link;
                const uint32_t table = dcsb_rv_u2_be(rv, rv.indirect_index + 3u * (link & 0xFF), 3);
                const DcsbRomPtr tp = dcsb_rv_ptr(rv, table);
                if (!dcsb_seq_q_push(s, (uint16_t)dcsb_rv_be(rv, tp, 2, 2u * s.vars[(link >> 8) & 0xFF]))) return false;
            }
            break;
        }
        case 0x06:
            if (rv.os != DCSB_OS93A && rv.os != DCSB_OS93B) {
                s.vars[dcsb_rv_u8(rv, p)] = dcsb_rv_u8(rv, p, 1);
                p.ofs += 2;
            }
            break;
        case 0x07: case 0x08: case 0x09: dcsb_seq_mix_op(s, rv, cur, p, op - 0x07, false); break;
        case 0x0A: case 0x0B: case 0x0C: dcsb_seq_mix_op(s, rv, cur, p, op - 0x0A, true); break;
        case 0x0D: break;
        case 0x0E: {
            const uint16_t n = dcsb_rv_u8(rv, p);
            p.ofs += 1;
            if (me.n_loops >= DCSB_SEQ_MAX_LOOPS) return false;
            me.loops[me.n_loops].counter = n;
            me.loops[me.n_loops].pos = p;
            ++me.n_loops;
            break;
        }
        case 0x0F:
            if (me.n_loops) {
                DcsbSeqLoop &l = me.loops[me.n_loops - 1];
                if (l.counter == 0) p = l.pos;
                else if (l.counter == 1) --me.n_loops;
                else { --l.counter; p = l.pos; }
            }
            break;
        case 0x10: p.ofs += 2; break;           // 0x10-0x12 set parameters nothing audible reads
        case 0x11: case 0x12: p.ofs += 4; break;
        default: return false;
        }
    }
}

DCSB_SEQ_HD void dcsb_seq_update_levels(DcsbSeqState &s, const DcsbRomView &rv)
{
    for (int i = 0; i < DCSB_MAX_CHANNELS; ++i) {
        DcsbSeqChannel &c = s.chan[i];
        for (unsigned f = c.fading; f; f &= f - 1) {        // only the mixers with a fade in progress
            int k = 0;
            while (!(f & (1u << k))) ++k;
            DcsbSeqMixer &m = c.mixer[k];
            if (m.steps == 1) { m.steps = 0; m.cur = m.target; }
            else if (m.steps > 1) {
                --m.steps;
                const int v = m.cur + m.delta;
                m.cur = v < -8191 ? -8191 : (v > 8191 ? 8191 : v);
            }
            if (m.steps == 0) c.fading &= (uint8_t)~(1u << k);
            c.levels_dirty = true;
        }
    }
    for (int i = 0; i < DCSB_MAX_CHANNELS; ++i) {
        DcsbSeqChannel &c = s.chan[i];
        if (c.levels_dirty) {                   // (most frames no level moves: the sum is kept)
            int sum = 0;
            for (int k = 0; k < DCSB_MAX_CHANNELS; ++k) sum += c.mixer[k].cur;
            c.level_sum = sum < -8191 ? -8191 : (sum > 8191 ? 8191 : sum);
            c.levels_dirty = false;
        }
        // the multiplier ladder (16 dependent 1.15 multiplies) only when its inputs moved
        const uint32_t key = (uint32_t)(c.level_sum + 8192) | ((uint32_t)c.volume << 14) | (c.max_override ? 1u << 30 : 0u);
        if (key != c.level_key) {
            c.level_key = key;
            c.level_mult = dcsb_seq_level_multiplier(c.level_sum, rv.os, c.volume, c.max_override ? 1 : 0);
        }
        c.mult = c.level_mult;
    }
    for (int i = 0; i < DCSB_MAX_CHANNELS; ++i) {
        DcsbSeqChannel &c = s.chan[i];
        c.track_counter += 1;
        if (c.timer.interval != 0 && --c.timer.counter == 0) { c.timer.counter = c.timer.interval; dcsb_seq_to_host(s, c.timer.data); }
    }
}

// one main-loop pass (MainLoop :89-306).  entries: room for DCSB_MAX_CHANNELS; *n_entries / *vs out.
// Returns false when the decoder must reset itself.
DCSB_SEQ_HD bool dcsb_seq_main_loop(DcsbSeqState &s, const DcsbRomView &rv, DcsbSchedEntry *entries, int *n_entries, uint8_t *vs)
{
    s.steps_this_frame = 0;
    *n_entries = 0;
    // channels the decoder's error path flagged last frame
    for (int ch = 0; ch < DCSB_MAX_CHANNELS; ++ch) {
        DcsbSeqChannel &c = s.chan[ch];
        if (!c.stop) continue;
        c.stop = false;
        if (c.st.active) { c.st.active = false; dcsb_seq_reset_mix(s, ch); }
        dcsb_seq_timer_clear(c.timer);
        c.track.clear();
    }
    // pending commands = indices into the track index
    while (s.q_count) {
        if (++s.steps_this_frame > DCSB_MAX_STEPS_PER_FRAME) { s.q_count = 0; return false; }
        const uint16_t cmd = s.cmdq[s.q_head];
        s.q_head = (s.q_head + 1) % DCSB_SEQ_QCAP;
        --s.q_count;
        if (cmd >= rv.n_tracks) continue;
        const uint32_t ofs = dcsb_rv_u2_be(rv, rv.track_index + 3u * cmd, 3);
        if ((ofs & 0xFF0000u) == 0xFF0000u) continue;
        DcsbRomPtr tp = dcsb_rv_ptr(rv, ofs);
        const int type = dcsb_rv_u8(rv, tp), ch = dcsb_rv_u8(rv, tp, 1) & 7;
        tp.ofs += 2;
        if (type == 1) dcsb_seq_load_track(s, ch, tp);
        else if (type <= 3) { s.chan[ch].next_type = (uint8_t)type; s.chan[ch].next_link = (uint16_t)dcsb_rv_be(rv, tp, 2); }
        else return false;
    }
    // run the track programs until every channel has had its turn
    s.done_mask = 0;
    for (int ch = 0; s.done_mask != 0xFFu; ch = (ch + 1) % DCSB_MAX_CHANNELS)
        if (!(s.done_mask & (1u << ch))) {
            if (!dcsb_seq_exec_track(s, rv, ch)) return false;
            s.done_mask |= 1u << ch;
        }
    // gain staging (MainLoop :227-269)
    uint16_t mix[8], eff[8];
    unsigned active = 0, maxo = 0;
    for (int i = 0; i < DCSB_MAX_CHANNELS; ++i) {
        mix[i] = s.chan[i].mult;
        if (s.chan[i].st.active) active |= 1u << i;
        if (s.chan[i].max_override) maxo |= 1u << i;
    }
    {
        // (the same inputs as last frame -- nothing started, stopped or faded -- give the same shift and multipliers)
        const uint32_t key = active | (maxo << 8) | ((uint32_t)s.vol_mult << 16);
        bool same = s.gs_valid && key == s.gs_key;
        for (int i = 0; i < DCSB_MAX_CHANNELS; ++i) same = same && mix[i] == s.gs_mix[i];
        if (!same) {
            s.gs_vs = (uint8_t)dcsb_seq_gain_stage(mix, active, maxo, s.vol_mult, eff);
            for (int i = 0; i < DCSB_MAX_CHANNELS; ++i) { s.gs_mix[i] = mix[i]; s.gs_eff[i] = eff[i]; }
            s.gs_key = key;
            s.gs_valid = true;
        }
        *vs = s.gs_vs;
        for (int i = 0; i < DCSB_MAX_CHANNELS; ++i) s.chan[i].mult = s.gs_eff[i];
    }
    // one frame from each active stream, channel order (DecodeStream :1546-1589)
    for (int ch = 0; ch < DCSB_MAX_CHANNELS; ++ch) {
        DcsbSeqChannel &c = s.chan[ch];
        DcsbSeqStreamState &st = c.st;
        if (!st.active) continue;
        st.at_start = false;
        bool decodes = false;
        uint32_t batch_id = 0;
        if (st.id != 0xFFFFFFFFu) {
            const DcsbSeqStream &sf = rv.streams[st.id];
            batch_id = sf.id;
            if (st.pos < sf.nplay) {
                decodes = true;
                if (sf.status == DCSB_E_STOPPED && st.pos + 1u == sf.nplay) c.stop = true;      // decoder error path: partial frame, channel stops
            } else c.stop = true;               // frame cannot be decoded (truncated / invalid band type): silence, channel stops
        } else c.stop = true;                   // a stream the ROM scan never saw: nothing to decode
        if (decodes) {
            DcsbSchedEntry &e = entries[(*n_entries)++];
            e.stream = batch_id; e.frame = st.pos; e.mult = c.mult;
        }
        ++st.pos;
        if (--st.counter != 0) continue;
        st.counter = st.nframes;
        st.pos = 0;
        st.at_start = true;
        if (st.loops == 0) continue;
        if (--st.loops != 0) continue;
        st.active = false;
        c.source = -1;
    }
    dcsb_seq_update_levels(s, rv);
    if (++s.port_timeout > 13) s.port_timeout = 13;
    return true;
}

// How many of the next main-loop passes are QUIET: no stop flag, no queued command, no track program whose wait
// runs out, no stream that ends, wraps or hits a damaged frame, no fade, no timer firing, gain staging unchanged.
// Such a pass moves counters and emits the same channels one frame further (dcsb_seq_quiet_frame); what it would
// compute is already in the state.  Conservative: 0 whenever in doubt.
DCSB_SEQ_HD uint32_t dcsb_seq_horizon(const DcsbSeqState &s, const DcsbRomView &rv)
{
    if (s.fatal || s.q_count || !s.gs_valid) return 0;
    uint32_t h = 0xFFFFu;
    unsigned active = 0, maxo = 0;
    for (int i = 0; i < DCSB_MAX_CHANNELS; ++i) {
        const DcsbSeqChannel &c = s.chan[i];
        if (c.stop || c.fading || c.levels_dirty) return 0;
        if (c.mult != c.level_mult || c.level_mult != s.gs_mix[i]) return 0;
        if (!c.track.null()) {
            const uint32_t wait = dcsb_rv_be(rv, c.track, 2);
            if (wait != 0xFFFFu) {                  // acts in the pass that finds the counter at `wait`
                const uint32_t d = (wait - c.track_counter) & 0xFFFFu;
                if (d < h) h = d;
            }
        }
        if (c.st.active) {
            active |= 1u << i;
            if (c.st.id == 0xFFFFFFFFu) return 0;
            const uint32_t np = rv.streams[c.st.id].nplay;
            const uint32_t left = np > (uint32_t)c.st.pos + 1u ? np - 1u - c.st.pos : 0u;      // whole frames before the last one
            if (left < h) h = left;
            const uint32_t cnt = c.st.counter ? c.st.counter - 1u : 0u;                         // passes before the counter runs out
            if (cnt < h) h = cnt;
        }
        if (c.max_override) maxo |= 1u << i;
        if (c.timer.interval != 0) {
            const uint32_t t = c.timer.counter ? c.timer.counter - 1u : 0u;
            if (t < h) h = t;
        }
    }
    if (s.gs_key != (active | (maxo << 8) | ((uint32_t)s.vol_mult << 16))) return 0;
    return h;
}
DCSB_SEQ_HD void dcsb_seq_quiet_frame(DcsbSeqState &s, const DcsbRomView &rv, DcsbSchedFrame *fr, DcsbSchedEntry *entries)
{
    int n = 0;
    for (int ch = 0; ch < DCSB_MAX_CHANNELS; ++ch) {
        DcsbSeqChannel &c = s.chan[ch];
        if (c.st.active) {
            DcsbSchedEntry &e = entries[n++];
            e.stream = rv.streams[c.st.id].id; e.frame = c.st.pos; e.mult = s.gs_eff[ch];
            c.st.at_start = false;
            ++c.st.pos;
            --c.st.counter;
        }
        c.track_counter += 1;
        if (c.timer.interval != 0) --c.timer.counter;
    }
    if (++s.port_timeout > 13) s.port_timeout = 13;
    fr->flags = 0;
    fr->pad = 0;
    fr->n_entries = (uint8_t)n;
    fr->vs = s.gs_vs;
    ++s.frame_no;
}

// One output frame: the main loop with the reference's self-reset retries (a reset is retried; four in a row are
// fatal, DCSDecoder.cpp:1631-1668).  entries: room for DCSB_MAX_CHANNELS.  fr->first_entry is left to the caller.
// Returns false once the decoder is in its fatal-error state (the frame is then silent).
DCSB_SEQ_HD bool dcsb_seq_frame(DcsbSeqState &s, const DcsbRomView &rv, DcsbSchedFrame *fr, DcsbSchedEntry *entries)
{
    if (s.quiet) {
        --s.quiet;
        dcsb_seq_quiet_frame(s, rv, fr, entries);
        return true;
    }
    int n = 0;
    uint8_t vs = 8;
    fr->flags = 0;
    fr->pad = 0;
    if (!s.fatal) {
        for (int tries = 0;; ++tries) {
            if (dcsb_seq_main_loop(s, rv, entries, &n, &vs)) break;
            n = 0;
            vs = 8;
            if (tries >= 3) { s.fatal = true; break; }
        }
    }
    if (s.fatal) { n = 0; vs = 8; fr->flags = DCSB_FRAME_MUTE; }
    fr->n_entries = (uint8_t)n;
    fr->vs = vs;
    ++s.frame_no;
    s.quiet = dcsb_seq_horizon(s, rv);
    return !s.fatal;
}
