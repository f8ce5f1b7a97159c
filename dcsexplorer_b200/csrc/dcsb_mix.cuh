// dcsb200 K4: channel mix / volume / PCM writeback for track playback.
//
// An output frame of a timeline is the sum of up to 8 channels, each contributing one frame
// of one stream scaled by its effective mixing multiplier; the channels are accumulated IN
// ORDER 0..7 into the same 16-bit frequency bins (DCSDecoderNative.cpp:272-273: the bin-0
// fix-up saturates per channel, :2255-2257, and the OS93a type-1 rounding depends on the
// accumulator, :3010-3015), then one inverse transform with the frame's volume shift, the
// 16-sample overlap-add against the previous OUTPUT frame and the PCM store (:278, :532-575).
// The host sequencer (dcsb_rom.cpp) supplies the schedule; everything audible happens here.
// The single-stream batch path (dcsb_fast94.cuh / dcsb_core.cuh) is the 1-channel special case
// with the schedule implied by the work item.
//
// Compiled by nvcc for sm_100a (the product) and by g++ for the CPU-side kernel simulator.
#pragma once
#include "dcsb_core.cuh"
#include "dcsb_fast94.cuh"
#include "dcsb_rom.h"

// `count` consecutive output frames of one timeline, frames[] index space
struct DcsbMixItem { uint32_t frame0, count, tl_first, timeline; };

struct DcsbMixSched {
    const DcsbSchedFrame *frames;
    const DcsbSchedEntry *entries;
};

// ---- 1994 layout: one warp per item, one lane per output frame ----------------------------
// rows = 32 x 129 words + 8 words of carried tail; hdrs = 32 x 16 bytes (one stream header per lane)
DCSB_HD unsigned long long dcsb_mix94_item(const uint8_t *slab, const DcsbStreamRec *streams, DcsbMixItem it, DcsbMixSched sc,
                                           const DcsbTables *tab, const uint16_t *lut, const DcsbTw94 *tw,
                                           const DcsbScanOut &scan, int16_t *pcm, uint32_t *rows, uint8_t *hdrs)
{
    int16_t *tail = reinterpret_cast<int16_t *>(rows + 32 * DCSB_ROW94_WORDS);
    uint32_t *pcm32 = reinterpret_cast<uint32_t *>(pcm + (size_t)it.tl_first * 240);
    const uint32_t fend = it.frame0 + it.count;
    unsigned long long csum = 0;
    uint32_t cur = it.frame0;
    bool have_tail = cur == it.tl_first;           // a fresh decoder's overlap buffer is zero
    if (have_tail) {
        DCSB_FOR_LANES(i, 8) rows[32 * DCSB_ROW94_WORDS + i] = 0;
        DCSB_SYNCWARP();
    }
    while (cur < fend) {
        const uint32_t tb = have_tail ? cur : cur - 1;          // frame of lane 0 (warm-up frame if no tail yet)
        const int out_from = have_tail ? 0 : 1;
        const int nfr = (int)(fend - tb < 32u ? fend - tb : 32u);
        // ---- decode every channel of the lane's frame into its row, then transform
        DCSB_FOR_LANES(l, 32) {
            uint32_t *row = rows + l * DCSB_ROW94_WORDS;
            if (l < nfr) {
                const DcsbSchedFrame fr = sc.frames[tb + (uint32_t)l];
                for (int i = 0; i < 128; ++i) row[i] = 0;
                int16_t *r16 = reinterpret_cast<int16_t *>(row);
                uint8_t *hdr = hdrs + l * 16;
                bool any = false;
                for (uint32_t e = 0; e < fr.n_entries; ++e) {
                    const DcsbSchedEntry en = sc.entries[fr.first_entry + e];
                    const DcsbStreamRec *sp = streams + en.stream;
                    const uint32_t nplay = scan.nplay[en.stream];
                    if (sp->fmt != DCSB_FMT_94 || en.frame >= nplay) continue;
                    const uint32_t fb = sp->frame_base, f = en.frame;
                    for (int i = 0; i < 16; ++i) hdr[i] = sp->hdr[i];
                    const uint2 bp = scan.bt[fb + f], bc = scan.bt[fb + f + 1];
                    DcsbWin win = dcsb_make_window(slab, *sp, scan.bitpos[fb + f] + scan.hdrbits[fb + f]);
                    const int stopband = scan.stopband[en.stream];
                    const int zero_from = (f == nplay - 1 && stopband != 0xFF) ? stopband : 16;
                    dcsb_lane_decode94<true>(hdr, lut, win, ((uint64_t)bp.y << 32) | bp.x, ((uint64_t)bc.y << 32) | bc.x,
                                             en.mult, zero_from, r16);
                    any = true;
                }
                if (any) dcsb_lane_transform94(r16, tw, fr.vs);
            }
        }
        DCSB_SYNCWARP();
        // ---- overlap-add, lane-private (reads the neighbour row's tail, writes its own head)
        DCSB_FOR_LANES(l, 32) {
            if (l >= out_from && l < nfr && !(sc.frames[tb + (uint32_t)l].flags & DCSB_FRAME_MUTE)) {
                int16_t *r16 = reinterpret_cast<int16_t *>(rows + l * DCSB_ROW94_WORDS);
                const int16_t *pr = l ? r16 - 2 * DCSB_ROW94_WORDS + 120 : tail;
                const int16_t *pi = l ? r16 - 2 * DCSB_ROW94_WORDS + 248 : tail + 8;
                dcsb_lane_overlap94(r16, pr, pi, tab);
            }
        }
        DCSB_SYNCWARP();
        for (int k = out_from; k < nfr; ++k)
            dcsb_frame_output94(reinterpret_cast<const int16_t *>(rows + k * DCSB_ROW94_WORDS), pcm32, tb + (uint32_t)k - it.tl_first, csum);
        DCSB_SYNCWARP();
        {
            const int16_t *last = reinterpret_cast<const int16_t *>(rows + (nfr - 1) * DCSB_ROW94_WORDS);
            DCSB_FOR_LANES(i, 16) tail[i] = i < 8 ? last[120 + i] : last[248 + i - 8];
            DCSB_SYNCWARP();
        }
        cur = tb + (uint32_t)nfr;
        have_tail = true;
    }
    return csum;
}

// ---- 1993 layouts: one warp per tile of <= 31 output frames (lane 0 = warm-up frame) -------
DCSB_HD unsigned long long dcsb_mix93_tile(const uint8_t *slab, const DcsbStreamRec *streams, DcsbMixItem it, DcsbMixSched sc,
                                           const DcsbTables *tab, const uint16_t *lut, const DcsbScanOut &scan,
                                           int16_t *pcm, uint32_t *rows)
{
    constexpr int ROWW = DcsbRow<true>::WORDS;
    uint32_t *tails = rows + 32 * ROWW;
    DCSB_FOR_LANES(i, DcsbWarpSmem<true>::WORDS) rows[i] = 0;
    DCSB_SYNCWARP();
    const long long first = (long long)it.frame0, lo = (long long)it.tl_first, end = first + it.count;
    // ---- phase A: lane l accumulates the channels of frame first-1+l
    DCSB_FOR_LANES(l, 32) {
        const long long g = first - 1 + l;
        if (g >= lo && g < end) {
            const DcsbSchedFrame fr = sc.frames[g];
            int16_t *row = reinterpret_cast<int16_t *>(rows + l * ROWW);
            for (uint32_t e = 0; e < fr.n_entries; ++e) {
                const DcsbSchedEntry en = sc.entries[fr.first_entry + e];
                const DcsbStreamRec *sp = streams + en.stream;
                if (sp->fmt == DCSB_FMT_94 || en.frame >= scan.nplay[en.stream]) continue;
                DcsbWalkCtx cx;
                cx.rd = dcsb_make_reader(slab, *sp);
                cx.hdr = sp->hdr;
                cx.lut = lut;
                cx.tab = tab;
                cx.mult = en.mult;
                cx.zero_from = 16;
                uint32_t pos = scan.bitpos[sp->frame_base + en.frame];
                const uint2 b2 = scan.bt[sp->frame_base + en.frame];
                uint64_t bt = ((uint64_t)b2.y << 32) | b2.x;
                int sb = 99;
                dcsb_walk<true>(sp->fmt, cx, pos, bt, row, sb);
            }
        }
    }
    DCSB_SYNCWARP();
    // ---- phase B: transform the frames in order, carrying the 16-sample tail
    unsigned long long csum = 0;
    uint32_t *pcm32 = reinterpret_cast<uint32_t *>(pcm + (size_t)it.tl_first * 240);
    for (int k = 0; k < 32; ++k) {
        const long long g = first - 1 + k;
        if (g < lo) continue;
        if (g >= end) break;
        const DcsbSchedFrame fr = sc.frames[g];
        uint32_t *c = rows + k * ROWW;
        if (fr.n_entries) dcsb_transform93_warp(c, tab);
        const int vs = fr.vs;
        const bool mute = (fr.flags & DCSB_FRAME_MUTE) != 0;
        const uint32_t *tin = tails + (k & 1) * 8;
        uint32_t *tout = tails + ((k + 1) & 1) * 8;
        const long long f = g - lo;
        DCSB_FOR_LANES(m, 128) {
            int s0 = dcsb_re(c[dcsb_rev8(2 * m)]) >> vs;
            int s1 = dcsb_re(c[dcsb_rev8(2 * m + 1)]) >> vs;
            if (m < 8 && !mute) {
                const uint32_t t = tin[m];
                s0 = dcsb_overlap_mix(s0, dcsb_re(t), tab->overlap[2 * m], tab->overlap[15 - 2 * m]);
                s1 = dcsb_overlap_mix(s1, dcsb_im(t), tab->overlap[2 * m + 1], tab->overlap[14 - 2 * m]);
            }
            const uint32_t w = dcsb_pack(s0, s1);
            if (m >= 120) tout[m - 120] = w;
            else if (k > 0) {
                pcm32[(size_t)f * 120 + m] = w;
                const unsigned long long i0 = (unsigned long long)f * 240 + 2 * m;
                csum += (unsigned long long)(w & 0xFFFFu) * (2 * i0 + 1) + (unsigned long long)(w >> 16) * (2 * i0 + 3);
            }
        }
        DCSB_SYNCWARP();
    }
    return csum;
}
