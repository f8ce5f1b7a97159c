// dcsb200 host-side control plane that needs no CUDA call: format tables, the gain
// staging arithmetic, stream validation, slab layout and tiling.  Linked into
// libdcsb200.so; also compiled on its own by the CPU-side kernel simulator (tests/hostsim).
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <thread>
#include <vector>
#include "../../include/dcsb200.h"
#include "dcsb_internal.h"
#include "dcs_tables.h"
#include "dcsb_seq.cuh"

// ======================================================================================
// tables: prefix-code lists (dcs_tables.h) -> peek LUTs
// 1994 sample codebook k: entry = code length << 12 | 'two zeros' flag << 11 | signed value (raw - 2^(k-1))
static void fill_cb94(uint16_t *lut, int peek_bits, const dcs_code_t *codes, int n, int k)
{
    for (int i = 0; i < n; ++i) {
        const int rep = 1 << (peek_bits - codes[i].len);
        const uint32_t base = codes[i].code << (peek_bits - codes[i].len);
        const bool dz = (codes[i].val & 0x80) != 0;
        const int v = dz ? 0 : (int)codes[i].val - (1 << (k - 1));
        for (int r = 0; r < rep; ++r) lut[base + r] = (uint16_t)((codes[i].len << 12) | (dz ? 0x800 : 0) | (v & 0xFF));
    }
}

static void fill_lut(uint16_t *lut, int peek_bits, const dcs_code_t *codes, int n)
{
    for (int i = 0; i < n; ++i) {
        if (codes[i].len > peek_bits) continue;
        const int rep = 1 << (peek_bits - codes[i].len);
        const uint32_t base = codes[i].code << (peek_bits - codes[i].len);
        for (int r = 0; r < rep; ++r) lut[base + r] = (uint16_t)((codes[i].len << 8) | codes[i].val);
    }
}
#define NCODES(t) ((int)(sizeof(t) / sizeof((t)[0])))

static int rev_bits(int x, int n) { int r = 0; for (int i = 0; i < n; ++i) r |= ((x >> i) & 1) << (n - 1 - i); return r; }

void dcsb_build_tables(DcsbTables *t)
{
    memset(t, 0, sizeof(*t));
    fill_lut(t->lut + DCSB_LUT_HDR94, 8, dcs94_hdr, NCODES(dcs94_hdr));
    fill_cb94(t->lut + DCSB_LUT_CB + 0, 2, dcs94_cb1, NCODES(dcs94_cb1), 1);
    fill_cb94(t->lut + DCSB_LUT_CB + 4, 3, dcs94_cb2, NCODES(dcs94_cb2), 2);
    fill_cb94(t->lut + DCSB_LUT_CB + 12, 5, dcs94_cb3, NCODES(dcs94_cb3), 3);
    fill_cb94(t->lut + DCSB_LUT_CB + 44, 7, dcs94_cb4, NCODES(dcs94_cb4), 4);
    fill_cb94(t->lut + DCSB_LUT_CB + 172, 8, dcs94_cb5, NCODES(dcs94_cb5), 5);
    fill_cb94(t->lut + DCSB_LUT_CB + 428, 9, dcs94_cb6, NCODES(dcs94_cb6), 6);
    fill_lut(t->lut + DCSB_LUT_HDR93, 8, dcs93_hdr, NCODES(dcs93_hdr));
    fill_lut(t->lut + DCSB_LUT_BB93A + 0, 4, dcs93a_bandbits0, NCODES(dcs93a_bandbits0));
    fill_lut(t->lut + DCSB_LUT_BB93A + 16, 4, dcs93a_bandbits1, NCODES(dcs93a_bandbits1));
    fill_lut(t->lut + DCSB_LUT_BB93A + 32, 4, dcs93a_bandbits2, NCODES(dcs93a_bandbits2));
    fill_lut(t->lut + DCSB_LUT_BB93A + 48, 4, dcs93a_bandbits3, NCODES(dcs93a_bandbits3));
    fill_lut(t->lut + DCSB_LUT_SC93A, 8, dcs93a_scale, NCODES(dcs93a_scale));
    for (int i = 0; i < 16; ++i) {
        t->lut[DCSB_LUT_XLAT + i] = (uint16_t)((dcs94_xlat_lo[i][0] << 8) | dcs94_xlat_lo[i][1]);
        t->lut[DCSB_LUT_XLAT + 16 + i] = (uint16_t)((dcs94_xlat_mid[i][0] << 8) | dcs94_xlat_mid[i][1]);
        t->lut[DCSB_LUT_XLAT + 32 + i] = (uint16_t)((dcs94_xlat_hi[i][0] << 8) | dcs94_xlat_hi[i][1]);
    }
    for (int i = 0; i < NCODES(dcs94_hdr); ++i)
        if (dcs94_hdr[i].len > 8) t->long94[t->n_long94++] = DcsbLongCode{ dcs94_hdr[i].code, dcs94_hdr[i].len, dcs94_hdr[i].val, 0 };
    {
        // second level of the 1994 header LUT: every code longer than 8 bits starts with the one 8-bit
        // pattern the first level has no entry for; codes of 9..16 bits are resolved by the 8 bits behind it
        uint32_t prefix = 0xFFFFFFFFu;
        for (int i = 0; i < NCODES(dcs94_hdr); ++i) {
            const dcs_code_t &c = dcs94_hdr[i];
            if (c.len <= 8) continue;
            const uint32_t pre = c.code >> (c.len - 8);
            if (prefix == 0xFFFFFFFFu) prefix = pre;
            if (pre != prefix) abort();         // (a property of the format's table, checked once)
            if (c.len > 16) continue;
            const int rest = c.len - 8, rep = 1 << (8 - rest);
            const uint32_t base = (c.code & ((1u << rest) - 1u)) << (8 - rest);
            for (int r = 0; r < rep; ++r) t->lut[DCSB_LUT_HDR94B + base + r] = (uint16_t)((c.len << 8) | c.val);
        }
        for (int x = 0; x < 256; ++x)
            if ((uint32_t)x != prefix && t->lut[DCSB_LUT_HDR94 + x] == 0) abort();
    }
    for (int i = 0; i < NCODES(dcs93_hdr); ++i)
        if (dcs93_hdr[i].len > 8) t->long93[t->n_long93++] = DcsbLongCode{ dcs93_hdr[i].code, dcs93_hdr[i].len, dcs93_hdr[i].val, 0 };
    memcpy(t->overlap, dcs_overlap_win, sizeof(t->overlap));
    for (int p = 0; p < 128; ++p) {
        t->twiddle[p] = ((uint32_t)dcs_twiddle[128 + p] << 16) | dcs_twiddle[p];
        t->tw93[p] = DcsbTw2{ 2 * (int)(int16_t)dcs_twiddle[128 + p], 2 * (int)(int16_t)dcs_twiddle[p] };
    }
    for (int i = 0; i < 64; ++i) {
        // twiddle pass coefficients c0 = table[bitrev9(2+4i)], c1 = table[bitrev9(4i)] (DCSDecoderNative.cpp:428-429)
        const uint16_t c0 = dcs_twiddle[rev_bits(2 + 4 * i, 9)], c1 = dcs_twiddle[rev_bits(4 * i, 9)];
        t->pretw[i] = ((uint32_t)c0 << 16) | c1;
    }
    memcpy(t->pairs93a, dcs93a_pairs, sizeof(t->pairs93a));
    // 1994 fast path: pre-doubled coefficients (a multiply-accumulate then yields the ADSP's
    // left-shifted MR directly)
    for (int p = 0; p < 64; ++p) {
        t->tw_c2[p] = 2 * (int)(int16_t)dcs_twiddle[128 + p];
        t->tw_s2[p] = 2 * (int)(int16_t)dcs_twiddle[p];
        t->pre_c0[p] = 2 * (int)(int16_t)(t->pretw[p] >> 16);
        t->pre_c1[p] = 2 * (int)(int16_t)(t->pretw[p] & 0xFFFFu);
    }
    // length tables for the scan: greedily chain whole codewords inside the peek without covering
    // more than `cap` output slots (the first codeword is always taken)
    const dcs_code_t *cbs[6] = { dcs94_cb1, dcs94_cb2, dcs94_cb3, dcs94_cb4, dcs94_cb5, dcs94_cb6 };
    const int ncb[6] = { NCODES(dcs94_cb1), NCODES(dcs94_cb2), NCODES(dcs94_cb3), NCODES(dcs94_cb4), NCODES(dcs94_cb5), NCODES(dcs94_cb6) };
    for (int k = 0; k < 6; ++k)
        for (int x = 0; x < DCSB_T8_CB; ++x) {
            uint32_t m[2];
            for (int which = 0; which < 2; ++which) {
                const int P = which ? DCSB_T1_PEEK : DCSB_T8_PEEK, cap = which ? 1 : DCSB_T8_CAP;
                const int xp = x >> (DCSB_T8_PEEK - P);
                int used = 0, slots = 0;
                for (;;) {
                    int hit = -1;
                    for (int i = 0; i < ncb[k] && hit < 0; ++i) {
                        const int len = cbs[k][i].len;
                        if (used + len <= P && (uint32_t)((xp >> (P - used - len)) & ((1 << len) - 1)) == cbs[k][i].code) hit = i;
                    }
                    if (hit < 0) break;
                    const int add = (cbs[k][hit].val & 0x80) ? 2 : 1;
                    if (slots && slots + add > cap) break;
                    used += cbs[k][hit].len;
                    slots += add;
                    if (slots >= cap) break;
                }
                m[which] = (uint32_t)((slots << 12) | used);
            }
            t->tx[(k << DCSB_T8_PEEK) + x] = (m[1] << 16) | m[0];
        }
}

// ======================================================================================
// gain helpers (host side)
// (the arithmetic lives in dcsb_seq.cuh, shared with the device-side sequencer)
extern "C" uint16_t dcsb_master_multiplier(int vol) { return dcsb_seq_master_multiplier(vol); }

extern "C" uint16_t dcsb_level_multiplier(int level_sum, int os_version, int channel_volume, int max_override)
{
    return dcsb_seq_level_multiplier(level_sum, os_version, channel_volume, max_override);
}

extern "C" int dcsb_gain_stage(const uint16_t mix_mult[8], unsigned active_mask, unsigned max_override_mask,
                               uint16_t vol_mult, uint16_t eff_mult[8])
{
    return dcsb_seq_gain_stage(mix_mult, active_mask, max_override_mask, vol_mult, eff_mult);
}

// Per-stream gain schedule of the single-stream protocol: frame 0 still carries the
// constructor multiplier 0x7FFF (DCSDecoderNative.h:514); the mixing level only takes
// effect in UpdateMixingLevels at the end of the first main loop (:3042-3121).
static void stream_gain(const dcsb_stream_desc &d, DcsbStreamRec &r)
{
    const uint16_t vol = dcsb_master_multiplier(d.master_volume);
    uint16_t mix[8], eff[8];
    for (int i = 0; i < 8; ++i) mix[i] = 0x7FFF;
    r.vs0 = (uint8_t)dcsb_gain_stage(mix, 1u, 0u, vol, eff);
    r.mult0 = eff[0];
    mix[0] = dcsb_level_multiplier((int)d.mixing_level << 6, d.os_version, 0xFF, 0);
    for (int i = 1; i < 8; ++i) mix[i] = dcsb_level_multiplier(0, d.os_version, 0xFF, 0);
    r.vs1 = (uint8_t)dcsb_gain_stage(mix, 1u, 0u, vol, eff);
    r.mult1 = eff[0];
    r.vs_idle = (uint8_t)dcsb_gain_stage(mix, 0u, 0u, vol, eff);
}


// ---- scan launch shape (shared by the launcher in dcsb_kernels.cu and the stream -> lane assignment below)
// spread the streams over the SMs first (one CTA per SM), then fill the CTAs up.  `concurrent` = the
// streams of all the scans that run side by side (a scan CTA's tables fill most of an SM's shared
// memory: there is room for one per SM, so the launches of a pipelined call share the 148 between them)
static int g_num_sms = 148;
int dcsb_num_sms() { return g_num_sms; }
void dcsb_set_num_sms(int n) { if (n > 0) g_num_sms = n; }

// Scan launch shape.  A lock-step warp walks 32 streams; the scan is as long as its slowest warp's
// dependent chain, so the warps are spread as thin as the SMs allow: one warp per CTA (150 KB of
// shared memory, leaving room for a decode CTA beside it) while the groups fit one wave, more
// warps per CTA (sharing the 96 KB length table) when there are more groups than that.
// concurrent = streams of all the scans launched side by side (0 = this launch alone).
// More groups than two waves' worth: the rings are given up (DcsbWinT<false>: stream words through L1) and a
// CTA holds up to eight warps -- in that regime an SM's throughput counts, not a warp's latency.
bool dcsb_scan_direct(int nstreams, int concurrent)
{
    if (concurrent < nstreams) concurrent = nstreams;
    if (const char *e = getenv("DCSB_SCAN_DIRECT")) return atoi(e) != 0;      // tuning override
    return (concurrent + 31) / 32 > DCSB_SCAN_MAXWARPS * dcsb_num_sms();
}

void dcsb_scan_shape(int nstreams, int concurrent, int *warps, int *grid)
{
    if (concurrent < nstreams) concurrent = nstreams;
    const int sms = dcsb_num_sms();
    const int groups = (nstreams + 31) / 32, cgroups = (concurrent + 31) / 32;
    const bool direct = dcsb_scan_direct(nstreams, concurrent);
    const int maxw = direct ? DCSB_SCAN_MAXWARPS_DIRECT : DCSB_SCAN_MAXWARPS;
    // (measured on 131,072 one-second streams, scan beside the decode kernel: rings, 2 warps per CTA 13.4 ms; direct,
    // 4 warps 10.9 ms, 6 warps 8.7 ms, 7 warps 7.7 ms, 8 warps 20.9 ms -- with eight warps' band entries in shared
    // memory L1 is too small for the lanes' lines; profiles/r03a_scan_warps.txt, r03b_scan_warps.txt)
    const int defw = direct ? 7 : DCSB_SCAN_MAXWARPS;
    int w = (cgroups + sms - 1) / sms;
    w = w > defw ? defw : (w < 1 ? 1 : w);
    // Single wave with rings: two warps per CTA even when every group could have an SM of its own.  A one-warp scan
    // CTA (153.5 KB) leaves room for ONE decode CTA beside it, so the 128 SMs under config 2's scan run the decode
    // kernel at a third of their rate and the step waits for the DECODE kernel (14.1 ms); two-warp CTAs (195 KB, no
    // decode CTA beside them) take half as many SMs and leave the others to three decode CTAs each: 13.0 ms, the
    // scan's own 12.8 ms chain plus the tail (profiles/r03h_scan_shape.txt).
    if (!direct && cgroups >= 2 && w < 2) w = 2;
    if (const char *e = getenv("DCSB_SCAN_WARPS")) {         // tuning override
        const int v = atoi(e);
        if (v >= 1 && v <= maxw) w = v;
    }
    int g = (groups + w - 1) / w;
    *warps = w;
    *grid = g > sms ? sms : (g < 1 ? 1 : g);
}

// Which stream each scan lane takes: lane l of group g walks order[32 g + l]; warp w of CTA c takes
// groups c * warps + w, then grid-stride.  A warp is as slow as its slowest lane in every frame and
// as long as its longest stream, so a group holds alike streams -- same layout, about the same
// frame count (buckets of a quarter octave), then ranked by bits per frame -- and the groups run
// from the most expensive one down, so that the longest chains start first.
void dcsb_scan_order(DcsbPrepared *p)
{
    const size_t n = p->recs.size();
    p->scan_order.resize(n);
    p->n_scan94 = 0;
    if (!n) return;
    // 1994-layout streams first (lock-step warps), the 1993 layouts behind them (one lane each, dcsb_scan93_kernel)
    std::vector<uint32_t> rank, rank93;
    std::vector<uint64_t> key(n);
    for (size_t i = 0; i < n; ++i) {
        const DcsbStreamRec &x = p->recs[i];
        if (x.fmt != DCSB_FMT_94) {
            // alike streams side by side here too (the lanes of a warp serialise where their walks part): layout and
            // stream type first, then size
            rank93.push_back((uint32_t)i);
            key[i] = ((uint64_t)x.fmt << 41) | ((uint64_t)(x.hdr[0] >> 7) << 40) | std::min<uint64_t>(x.nbytes, 0xFFFFFFFFFFull);
            continue;
        }
        rank.push_back((uint32_t)i);
        uint32_t bucket = 0;                          // quarter-octave bucket of the frame count
        if (x.nframes) {
            int lg = 31;
            while (!(x.nframes >> lg)) --lg;
            bucket = 1u + (uint32_t)lg * 4u + ((lg >= 2 ? x.nframes >> (lg - 2) : x.nframes << (2 - lg)) & 3u);
        }
        const uint64_t bpf = x.nframes ? std::min<uint64_t>((uint64_t)x.nbytes * 8 / x.nframes, 0xFFFFFull) : 0;
        key[i] = ((uint64_t)bucket << 20) | bpf;
    }
    const size_t n94 = rank.size();
    p->n_scan94 = n94;
    std::stable_sort(rank.begin(), rank.end(), [&](uint32_t a, uint32_t b) { return key[a] > key[b]; });
    // groups of 32 by estimated cost (frames of the longest stream x bits per frame of the densest), largest first
    const size_t ng = (n94 + 31) / 32;
    std::vector<uint32_t> gidx(ng);
    std::vector<uint64_t> gcost(ng, 0);
    for (size_t g = 0; g < ng; ++g) {
        gidx[g] = (uint32_t)g;
        uint64_t mf = 0, mb = 0;
        for (size_t k = g * 32; k < std::min(n94, g * 32 + 32); ++k) {
            const DcsbStreamRec &x = p->recs[rank[k]];
            mf = std::max<uint64_t>(mf, x.nframes);
            mb = std::max<uint64_t>(mb, x.nframes ? (uint64_t)x.nbytes * 8 / x.nframes : 0);
        }
        gcost[g] = mf * (mb + 64);
    }
    std::stable_sort(gidx.begin(), gidx.end(), [&](uint32_t a, uint32_t b) { return gcost[a] > gcost[b]; });
    // a short last group must stay the last one (lanes are assigned by position in the order)
    size_t o = 0;
    const size_t tail_g = (n94 % 32) ? ng - 1 : ng;   // index of the short group, if any
    for (size_t q = 0; q < ng; ++q) {
        const size_t g = gidx[q];
        if (g == tail_g) continue;
        for (size_t k = g * 32; k < g * 32 + 32; ++k) p->scan_order[o++] = rank[k];
    }
    if (tail_g < ng) for (size_t k = tail_g * 32; k < n94; ++k) p->scan_order[o++] = rank[k];
    // 1993 layouts: the longest chains first
    std::stable_sort(rank93.begin(), rank93.end(), [&](uint32_t a, uint32_t b) { return key[a] > key[b]; });
    for (uint32_t i : rank93) p->scan_order[o++] = i;
}

// Work items covering output frames [fa, fb) of every stream, in frame-major order (item k of every
// stream before item k + 1 of any): CTAs are dispatched in index order, so the decode kernel works
// its way through the streams at the pace the scan running beside it delivers their checkpoints.
void dcsb_build_tiles(const DcsbPrepared *p, uint32_t fa, uint32_t fb, std::vector<DcsbTile> *t94, std::vector<DcsbTile> *t93)
{
    const size_t n = p->recs.size();
    for (int fam = 0; fam < 2; ++fam) {
        const uint32_t len = fam ? (uint32_t)DCSB_TILE_OUT : p->item_len;
        std::vector<DcsbTile> &dst = fam ? *t93 : *t94;
        std::vector<uint32_t> live;
        for (size_t i = 0; i < n; ++i)
            if ((p->recs[i].fmt != DCSB_FMT_94) == (fam == 1) && p->recs[i].out_frames > fa) live.push_back((uint32_t)i);
        for (uint32_t f = fa; !live.empty(); f += len) {
            size_t keep = 0;
            for (uint32_t i : live) {
                const uint32_t of = std::min(p->recs[i].out_frames, fb);
                dst.push_back(DcsbTile{ i, f, std::min<uint32_t>(len, of - f) });
                if (f + len < of) live[keep++] = i;
            }
            live.resize(keep);
        }
    }
}

int dcsb_prepare(const dcsb_stream_desc *descs, size_t n, DcsbPrepared *p, const uint8_t *in_place_base, size_t in_place_span)
{
    if (p->concurrent_streams < n) p->concurrent_streams = n;
    p->recs.resize(n);
    p->host_status.assign(n, 0);
    p->tiles.clear();
    p->compressed_bytes = 0;
    uint64_t off = 0, frames = 0, ckpt = 0, pcm = 0;
    std::vector<DcsbTile> t94, t93;
    // 1994 work items: long enough that the one warm-up frame per item is noise, short enough
    // that the batch spreads over every resident warp of the chip several times
    uint64_t est_frames = 0;
    for (size_t i = 0; i < n; ++i)
        if (descs[i].data && descs[i].nbytes >= 2) est_frames += (((uint32_t)descs[i].data[0] << 8) | descs[i].data[1]) + descs[i].tail_frames;
    // (and short enough that the decode kernel follows the scan running beside it closely: an item
    // can only start once the scan has passed its last frame)
    uint32_t item_len = (uint32_t)std::min<uint64_t>(63, std::max<uint64_t>(31, est_frames / 8192));
    item_len = item_len < 63 ? 31 : 31 + ((item_len - 31) / 32) * 32;       // 31 + whole 32-frame tiles
    size_t items94 = 0, items93 = 0;
    uint64_t nq94 = 0;
    for (size_t i = 0; i < n; ++i) {
        const dcsb_stream_desc &d = descs[i];
        DcsbStreamRec &r = p->recs[i];
        memset(&r, 0, sizeof(r));
        int fmt;
        switch (d.os_version) {
        case DCSB_OS94: case DCSB_OS95: fmt = DCSB_FMT_94; break;
        case DCSB_OS93B: fmt = DCSB_FMT_93; break;
        case DCSB_OS93A: fmt = (d.nbytes >= 3 && d.data && (d.data[2] & 0x80)) ? DCSB_FMT_93A1 : DCSB_FMT_93; break;
        default: return DCSB_E_ARG;
        }
        r.fmt = (uint8_t)fmt;
        r.hdr_len = fmt == DCSB_FMT_93A1 ? 1 : 16;
        r.data_off = in_place_base ? (uint64_t)(d.data - in_place_base) : off;
        r.nbytes = d.nbytes;
        r.frame_base = (uint32_t)ckpt;
        r.pcm_off = pcm;
        // frames rendered = U16 frame count + tail, whatever happens later (rejected and
        // failing streams render silence), so callers can lay out PCM from the first 2 bytes
        uint32_t nf = (d.data && d.nbytes >= 2) ? (((uint32_t)d.data[0] << 8) | d.data[1]) : 0;
        const bool wrap = (d.reserved & DCSB_STREAM_WRAP_EMPTY) != 0;
        r.out_frames = nf + d.tail_frames;
        if (!d.data || d.nbytes < 3 || d.nbytes < 2u + r.hdr_len) { p->host_status[i] = DCSB_E_SHORT; nf = 0; }
        else {
            if (nf == 0 && wrap) { nf = 65536; r.out_frames = nf + d.tail_frames; }      // the reference's uint16 counter wraps (:1411-1415, :1565)
            else if (nf == 0) p->host_status[i] = DCSB_E_EMPTY;
            memcpy(r.hdr, d.data + 2, r.hdr_len);
        }
        r.nframes = nf;
        stream_gain(d, r);
        if (fmt == DCSB_FMT_94) { items94 += (r.out_frames + item_len - 1) / item_len; nq94 += (r.out_frames + DCSB_QITEM - 1) / DCSB_QITEM; }
        else items93 += (r.out_frames + DCSB_TILE_OUT - 1) / DCSB_TILE_OUT;
        off += ((uint64_t)d.nbytes + 15 + 16) & ~15ull;     // 16-byte aligned, >= 16 bytes of zero padding
        frames += nf;
        ckpt += (uint64_t)nf + 1;
        pcm += (uint64_t)r.out_frames * 240;
        p->compressed_bytes += d.nbytes;
    }
    p->item_len = item_len;
    t94.reserve(items94);
    t93.reserve(items93);
    dcsb_build_tiles(p, 0, 0xFFFFFFFFu, &t94, &t93);
    dcsb_scan_order(p);
    if (ckpt > 0xFFFFFFF0ull || t94.size() + t93.size() > 0x7FFFFFF0ull || nq94 > 0x7FFFFFF0ull || n > 0x3FFFFFF0ull) return DCSB_E_ARG;
    p->nqueue94 = (int)nq94;
    p->total_frames_in = frames;
    p->total_checkpoints = ckpt;
    p->total_out_frames = pcm / 240;
    p->slab_bytes = (in_place_base ? in_place_span : (size_t)off) + 1024;     // slack: window prefetch + L2 prefetch run ahead of the data
    p->ntiles94 = (int)t94.size();
    p->ntiles93 = (int)t93.size();
    p->tiles = std::move(t94);
    p->tiles.insert(p->tiles.end(), t93.begin(), t93.end());
    return DCSB_OK;
}

void dcsb_pack_slab_range(const dcsb_stream_desc *descs, size_t n, const DcsbPrepared *p, size_t i0, size_t i1, uint8_t *dst)
{
    if (i0 >= i1) return;
    const uint64_t base = p->recs[i0].data_off;
    const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    auto work = [&](unsigned t) {
        for (size_t i = i0 + t; i < i1; i += nt) {
            const DcsbStreamRec &r = p->recs[i];
            const uint64_t span = (i + 1 < n ? p->recs[i + 1].data_off : (uint64_t)p->slab_bytes) - r.data_off;
            if (descs[i].data && descs[i].nbytes) memcpy(dst + (r.data_off - base), descs[i].data, descs[i].nbytes);
            memset(dst + (r.data_off - base) + descs[i].nbytes, 0, span - descs[i].nbytes);
        }
    };
    if (i1 - i0 < 64 || nt == 1) { for (unsigned t = 0; t < nt; ++t) work(t); return; }
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
}

void dcsb_pack_slab(const dcsb_stream_desc *descs, size_t n, const DcsbPrepared *p, uint8_t *slab)
{
    if (n) dcsb_pack_slab_range(descs, n, p, 0, n, slab + p->recs[0].data_off);
}

// ======================================================================================
// multi-GPU sharding: longest-processing-time greedy on frame counts
extern "C" int dcsb_partition_streams(const uint32_t *frames, size_t n, int n_parts, uint32_t *part_out, uint64_t *frames_per_part)
{
    if ((!frames && n) || (!part_out && n) || n_parts < 1) return DCSB_E_ARG;
    std::vector<uint32_t> order(n);
    for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return frames[a] > frames[b]; });
    std::vector<uint64_t> load((size_t)n_parts, 0);
    for (uint32_t i : order) {
        int best = 0;
        for (int p = 1; p < n_parts; ++p) if (load[p] < load[best]) best = p;
        part_out[i] = (uint32_t)best;
        load[best] += (uint64_t)frames[i] + 1;          // + 1: per-stream fixed cost (an empty stream is not free)
    }
    if (frames_per_part) for (int p = 0; p < n_parts; ++p) frames_per_part[p] = load[p];
    return DCSB_OK;
}

// ======================================================================================
// output containers
static void put_le32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static void put_le16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }

extern "C" int dcsb_write_wav(const char *path, const int16_t *pcm, size_t n_samples)
{
    if (!path || (!pcm && n_samples) || n_samples > 0x7FFFFFF0u / 2) return DCSB_E_ARG;
    FILE *f = fopen(path, "wb");
    if (!f) return DCSB_E_ARG;
    uint8_t h[44];
    memset(h, 0, sizeof(h));
    const uint32_t bytes = (uint32_t)(n_samples * 2);
    memcpy(h, "RIFF", 4);
    put_le32(h + 4, bytes + 44 - 8);
    memcpy(h + 8, "WAVEfmt ", 8);
    put_le32(h + 16, 16);            // fmt chunk length
    put_le16(h + 20, 1);             // PCM
    put_le16(h + 22, 1);             // mono
    put_le32(h + 24, 31250);
    put_le32(h + 28, 31250 * 2);     // bytes per second
    put_le16(h + 32, 2);             // block align
    put_le16(h + 34, 16);            // bits per sample
    memcpy(h + 36, "data", 4);
    put_le32(h + 40, bytes);
    bool ok = fwrite(h, 1, 44, f) == 44;
    // samples are little-endian on disk; the hosts this library runs on are little-endian
    if (ok && n_samples) ok = fwrite(pcm, 2, n_samples, f) == n_samples;
    if (fclose(f) != 0) ok = false;
    return ok ? DCSB_OK : DCSB_E_ARG;
}

extern "C" int dcsb_write_dcs_file(const char *path, uint16_t os_version, const uint8_t *stream, size_t nbytes)
{
    if (!path || !stream || nbytes == 0 || nbytes > 0xFFFFFFFFull) return DCSB_E_ARG;
    uint8_t h[36];
    memset(h, 0, sizeof(h));
    memcpy(h, "DCSa", 4);
    const bool os93 = os_version == DCSB_OS93A || os_version == DCSB_OS93B;
    h[4] = os93 ? 0x93 : 0x94;
    h[5] = os_version == DCSB_OS93A ? 0x01 : os_version == DCSB_OS93B ? 0x02 : 0x00;
    h[7] = 0x01;                     // channels
    h[8] = 0x7A; h[9] = 0x12;        // 31,250 Hz
    h[32] = (uint8_t)(nbytes >> 24); h[33] = (uint8_t)(nbytes >> 16); h[34] = (uint8_t)(nbytes >> 8); h[35] = (uint8_t)nbytes;
    FILE *f = fopen(path, "wb");
    if (!f) return DCSB_E_ARG;
    bool ok = fwrite(h, 1, 36, f) == 36 && fwrite(stream, 1, nbytes, f) == nbytes;
    if (fclose(f) != 0) ok = false;
    return ok ? DCSB_OK : DCSB_E_ARG;
}

extern "C" long long dcsb_read_dcs_file(const char *path, uint16_t *os_version, uint8_t *out, size_t max)
{
    if (!path) return DCSB_E_ARG;
    FILE *f = fopen(path, "rb");
    if (!f) return DCSB_E_ARG;
    uint8_t h[36];
    long long rc = DCSB_E_ARG;
    if (fread(h, 1, 36, f) == 36 && memcmp(h, "DCSa", 4) == 0 && (h[4] == 0x93 || h[4] == 0x94) && h[6] == 0 && h[7] == 1 &&
        h[8] == 0x7A && h[9] == 0x12) {
        const uint32_t n = ((uint32_t)h[32] << 24) | (h[33] << 16) | (h[34] << 8) | h[35];
        if (os_version) *os_version = (uint16_t)((h[4] << 8) | h[5]);
        rc = n;
        if (out && max) {
            const size_t want = n < max ? n : max;
            if (fread(out, 1, want, f) != want) rc = DCSB_E_TRUNCATED;
        }
    }
    fclose(f);
    return rc;
}
