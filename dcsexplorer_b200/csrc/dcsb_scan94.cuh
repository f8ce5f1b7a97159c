// dcsb200 K1 body for the 1994+ frame layout: the frame-boundary scan.
//
// One thread walks one stream (lengths only) and writes one checkpoint per frame plus the end
// checkpoint:
//   bitpos[f]  bit position of the frame start,   bt[f] band types carried INTO frame f,
//   hdrbits[f] length of the frame header (the decode lanes start at the first band and take
//              the frame's own band types from bt[f + 1]).
// It replaces the frame walk implicit in DecodeStream / GetStreamInfo (DCSDecoderNative.cpp:
// 1486-1589) and the header part of DecompressFrame (:1780-1834).
//
// The scan is one dependent chain per stream (position -> bits at that position -> next
// position); what counts is the length of that chain and how many instructions hang off it:
//  * the stream bytes are staged in a per-stream 1 KB shared-memory ring that cp.async fills
//    a whole frame ahead, so the register bit window refills with LDS (no global latency on the
//    chain) and a long skip (fixed-width bands) re-seeks inside the ring;
//  * Huffman bands advance with a multi-symbol length table tx[codebook][next 13 bits] =
//    {m8, m1}: m8 = {bits consumed, output slots covered} of as many whole codewords as fit in
//    the peek (at most 8 slots), m1 = the same for the first codeword alone.  When the m8 step
//    would cover more slots than the band has left, m1 is taken instead, so a step never
//    overruns the band and the table address depends on the bit position only, not on the slot
//    count; both come with ONE load (two tables and a predicated second load cost a second
//    shared-memory latency on the chain).  (A 'two zeros' codeword with one slot left leaves
//    rem < 0: the reference's error case, :2213-2218);
//  * band descriptors (kind, codebook, slot count) are kept per band and only looked up again
//    (one load from a 4 KB table) when a frame header changes the band's type; the band loop
//    walks the non-empty bands only;
//  * the frame header's 1-bit "unchanged" codes are skipped as a run (count leading ones);
//  * fixed-width bands advance in closed form.
//
// Compiled by nvcc for sm_100a (the product) and by g++ for the CPU-side kernel simulator
// (tests/hostsim, test infrastructure only; cp.async becomes an immediate copy there).
#pragma once
#include <string.h>
#include "dcsb_core.cuh"
#include "dcsb_fast94.cuh"

#define DCSB_RING_BYTES  1024u
#define DCSB_RING_CHUNKS (DCSB_RING_BYTES / 16u)
// the farthest a frame can reach past its first chunk: 16 header codes of up to 23 bits plus
// 255 samples of up to 15 bits (543 bytes), plus the window's three words and chunk rounding
#define DCSB_RING_FRAME_CHUNKS 36u

struct DcsbRingWin {
    uint32_t w0, w1, nx;    // current / next word (big-endian order), prefetched raw word
    uint32_t s;             // bit offset inside w0
    uint32_t wa;            // byte offset (from chunk 0) of the next word the ring hands out
    uint32_t bias;          // bit offset of the stream's first data bit from chunk 0
    uint32_t fill;          // chunks issued so far
    uint32_t limit;         // chunks that exist (stream bytes + slack)
    const uint8_t *g;       // global address of chunk 0 (16-byte aligned)
#if DCSB_DEVICE_PASS
    uint32_t ring;          // shared-window address of this stream's ring (16-byte aligned)
#else
    uint8_t *ring;
#endif

    DCSB_HD uint32_t ring_word(uint32_t off) const
    {
#if DCSB_DEVICE_PASS
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ring + (off & (DCSB_RING_BYTES - 1))));
        return v;
#else
        uint32_t v;
        memcpy(&v, ring + (off & (DCSB_RING_BYTES - 1)), 4);
        return v;
#endif
    }
    // issue the chunks up to DCSB_RING_CHUNKS - 1 ahead of the window, then make sure everything
    // the next frame can touch has landed.  Call at a frame start only.
    DCSB_HD void topup()
    {
        const uint32_t cc = (wa - 12u) >> 4;                       // chunk holding w0
        uint32_t target = cc + DCSB_RING_CHUNKS - 1u;
        if (target > limit) target = limit;
        const bool behind = fill < cc + DCSB_RING_FRAME_CHUNKS && fill < limit;
        while (fill < target) {
#if DCSB_DEVICE_PASS
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(ring + ((fill * 16u) & (DCSB_RING_BYTES - 1))),
                         "l"(g + (size_t)fill * 16u) : "memory");
#else
            memcpy(ring + ((fill * 16u) & (DCSB_RING_BYTES - 1)), g + (size_t)fill * 16u, 16);
#endif
            ++fill;
        }
#if DCSB_DEVICE_PASS
        asm volatile("cp.async.commit_group;" ::: "memory");
        // usual case: this frame's bytes were issued at least one frame ago
        if (behind) asm volatile("cp.async.wait_group 0;" ::: "memory");
        else asm volatile("cp.async.wait_group 1;" ::: "memory");
#else
        (void)behind;
#endif
    }
    DCSB_HD void seek(uint32_t pos)
    {
        const uint32_t a = pos + bias, off = (a >> 5) * 4u;
        s = a & 31u;
        w0 = DcsbBits::be(ring_word(off));
        w1 = DcsbBits::be(ring_word(off + 4u));
        nx = ring_word(off + 8u);
        wa = off + 12u;
    }
    DCSB_HD uint32_t pos() const { return (wa - 12u) * 8u + s - bias; }
    DCSB_HD uint32_t peek32() const
    {
#if DCSB_DEVICE_PASS
        return __funnelshift_l(w1, w0, s);
#else
        return s ? ((w0 << s) | (w1 >> (32 - s))) : w0;
#endif
    }
    // valid while s + bits <= 64: two short reads (<= 15 bits each) per refill
    DCSB_HD uint32_t peek_wide() const { return (uint32_t)((((((uint64_t)w0) << 32) | w1) << s) >> 32); }
    DCSB_HD void advance(uint32_t n) { s += n; }
    // select-style on purpose: a compare + predicated block costs a 13-cycle predicate latency on
    // the position chain; masks cost one ALU hop (the ring is always readable, so the load is
    // unconditional)
    DCSB_HD void refill()
    {
        const uint32_t r = s >> 5;                  // 0 or 1
        const uint32_t mask = 0u - r;
        const uint32_t nw1 = DcsbBits::be(nx), ld = ring_word(wa);
        s &= 31u;
        w0 ^= (w0 ^ w1) & mask;
        w1 ^= (w1 ^ nw1) & mask;
        nx ^= (nx ^ ld) & mask;
        wa += 4u * r;
    }
    DCSB_HD void skip(uint32_t n) { s += n; refill(); }      // n <= 32
    // the same with the position kept as t = 50 - s (the Huffman loop's form)
    DCSB_HD void refill_t(int &t)
    {
        const uint32_t mask = (uint32_t)((t - 19) >> 31);       // all ones when s >= 32
        const uint32_t nw1 = DcsbBits::be(nx), ld = ring_word(wa);
        t += (int)(32u & mask);
        w0 ^= (w0 ^ w1) & mask;
        w1 ^= (w1 ^ nw1) & mask;
        nx ^= (nx ^ ld) & mask;
        wa += 4u & mask;
    }
};

// Scan table in shared memory (DcsbTables::tx): the kernel hands over a 32-bit shared-window
// address that is a multiple of 16 KB (one codebook's table), so that a lookup address is
// table | byte offset -- one LOP3 -- and the load is one LDS.U16; the simulator passes a pointer.
#if DCSB_DEVICE_PASS
typedef uint32_t DcsbTxBase;
// v: the window shifted so that bits 1..13 are the next 13 stream bits
DCSB_HD uint32_t dcsb_tx_load(DcsbTxBase tb, uint32_t v)
{
    uint32_t a, r;
    asm("lop3.b32 %0, %1, 0x3FFE, %2, 0xEA;" : "=r"(a) : "r"(v), "r"(tb));      // (v & 0x3FFE) | tb
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r) : "r"(a));
    return r;
}
typedef uint32_t DcsbRingPtr;
#else
typedef const uint8_t *DcsbTxBase;
DCSB_HD uint32_t dcsb_tx_load(DcsbTxBase tb, uint32_t v)
{
    uint16_t r;
    memcpy(&r, tb + (v & 0x3FFEu), 2);
    return r;
}
typedef uint8_t *DcsbRingPtr;
#endif

// Band descriptor: what the band loop needs to know about a band of the current frame.
//   bit 31 Huffman band: bits 24..26 = codebook - 1, bits 0..17 = Rs = (16 * slots + 15) << 8 | 0xFF
//   bit 30 fixed-width band: bits 16..25 = slots * width (bits to skip)
//   0      empty band
// A descriptor depends on (stream type, half-density flag of the band, band, band type) only:
// dtab[((type1 * 2 + half) * 16 + band) * 16 + type], 4 KB, built once per CTA.
#define DCSB_DESC_HUFF  0x80000000u
#define DCSB_DESC_FIXED 0x40000000u
#define DCSB_DTAB_WORDS 1024
DCSB_HD uint32_t dcsb_band_desc94(const uint16_t *lut, int type1, int b, int nib, int count)
{
    int code = nib;
    if (type1) code = (int)(lut[DCSB_LUT_XLAT + (b < 3 ? 0 : (b < 6 ? 16 : 32)) + code] >> 8);      // :1926-1955
    if (code >= 1 && code <= 6) return DCSB_DESC_HUFF | ((uint32_t)(code - 1) << 24) | ((uint32_t)(count * 16 + 15) << 8) | 0xFFu;
    if (code > 6 && count) return DCSB_DESC_FIXED | ((uint32_t)(count * code) << 16);
    return 0;
}
DCSB_HD uint32_t dcsb_dtab_entry(const uint16_t *lut, int i)
{
    const int nib = i & 15, b = (i >> 4) & 15, half = (i >> 8) & 1, type1 = i >> 9;
    return dcsb_band_desc94(lut, type1, b, nib, dcsb_band_count94(b) >> half);
}
DCSB_HD int dcsb_nib32(uint32_t lo, uint32_t hi, int b) { return (int)(((b < 8 ? lo : hi) >> (4 * (b & 7))) & 15u); }
DCSB_HD int dcsb_ctz(uint32_t v)
{
#if DCSB_DEVICE_PASS
    return __ffs((int)v) - 1;
#else
    int n = 0;
    while (n < 32 && !(v & (1u << n))) ++n;
    return n;
#endif
}

// [f0, f1) = the frames this call walks (0, ~0u = the whole stream).  A call with f0 > 0 resumes
// from the end checkpoint the previous call left at frame f0 (status DCSB_SCAN_RUNNING); that is
// what lets dcsb_decode_streams cut a chunk into time slices whose PCM drains over PCIe while
// the later slices are still being scanned.
DCSB_HD void dcsb_scan94_stream(const uint8_t *slab, const DcsbStreamRec *streams, int si, const DcsbTables *tab,
                                const uint16_t *lut, DcsbTxBase tx, const uint32_t *dtab, DcsbRingPtr ring, uint32_t *desc,
                                const DcsbScanOut &out, uint32_t f0 = 0, uint32_t f1 = 0xFFFFFFFFu)
{
    if (f0 && out.status[si] != DCSB_SCAN_RUNNING) return;      // finished (or failed) in an earlier slice
    const DcsbStreamRec s = streams[si];
    const uint8_t *hdr = streams[si].hdr;
    const int type1 = hdr[0] >> 7;
    int nb = 0;
    while (nb < 16 && (hdr[nb] & 0x7F) != 0x7F) ++nb;
    // bands at half density (:1858-1862) select the other half of the descriptor table
    uint32_t halfmask = 0;
    for (int b = 0; b < nb; ++b) halfmask |= (uint32_t)((hdr[b] >> 6) & 1) << b;
    const uint32_t dsel = (uint32_t)type1 * 512u;
    const uint32_t dbytes = s.nbytes > 2u + s.hdr_len ? s.nbytes - 2u - s.hdr_len : 0u;   // (short streams have nframes == 0)
    const uint32_t nbits = dbytes * 8u;
    const uint64_t start = s.data_off + 2 + s.hdr_len;
    DcsbRingWin win;
    win.g = slab + (start & ~15ull);
    win.bias = (uint32_t)(start & 15) * 8u;
    win.ring = ring;
    win.fill = 0;
    // chunks worth reading: the stream, plus the bytes a frame that starts inside it may still
    // reach into the zero padding (the slab keeps >= 1 KB of slack behind the last stream)
    win.limit = s.nframes ? (uint32_t)(((start & 15) + dbytes + 64u + 15u) >> 4) : 0u;
    win.wa = 12u;
    uint32_t pos = 0;
    uint32_t bt_lo = 0, bt_hi = 0;                 // band types, 16 x 4 bits; InitStreamPlayback zeroes them (:1640)
    if (f0) {
        pos = out.bitpos[s.frame_base + f0];
        const uint2 b2 = out.bt[s.frame_base + f0];
        bt_lo = b2.x;
        bt_hi = b2.y;
        const uint32_t off = ((pos + win.bias) >> 5) * 4u;
        win.wa = off + 12u;
        win.fill = off >> 4;
    }
    win.topup();
    win.seek(pos);
    // plain reader on global memory for the rare long header codes
    DcsbBits rd;
    rd.w = reinterpret_cast<const uint32_t *>(slab + (start & ~3ull));
    rd.bias = (uint32_t)(start & 3) * 8;

#if defined(DCSB_SCAN_DEBUG) && DCSB_DEVICE_PASS
    const long long dbg_t0 = clock64();
    uint32_t dbg_steps = 0, dbg_hdr = 0, dbg_c[4] = { 0, 0, 0, 0 };
    long long dbg_t = dbg_t0;
#define DCSB_DBG(x) x
#define DCSB_DBG_LAP(k) { const long long t_ = clock64(); dbg_c[k] += (uint32_t)(t_ - dbg_t); dbg_t = t_; }
#else
#define DCSB_DBG(x)
#define DCSB_DBG_LAP(k)
#endif
    uint32_t queued = 0, qnext = DCSB_QITEM;       // output frames already handed to the decode kernel / next hand-over
    int status = s.nframes ? 0 : -1, stopband = 0xFF;    // -1 = DCSB_E_EMPTY (the host refines DCSB_E_SHORT)
    uint32_t nplay = s.nframes, f = f0;
    const uint32_t fe = f1 < s.nframes ? f1 : s.nframes;
    // band descriptors of the current band types: only a band whose type changes in a frame header
    // is looked up again, and the band loop walks the non-empty bands only
    uint32_t live = 0;
    for (int b = 0; b < nb; ++b) {
        const uint32_t d = dtab[dsel + ((halfmask >> b) & 1u) * 256u + (uint32_t)b * 16u + (uint32_t)dcsb_nib32(bt_lo, bt_hi, b)];
        desc[b] = d;
        live |= (d ? 1u : 0u) << b;
    }
    desc[16] = 0;
    for (; f < fe; ++f) {
        out.bitpos[s.frame_base + f] = pos;
        out.bt[s.frame_base + f] = make_uint2(bt_lo, bt_hi);
        DCSB_DBG_LAP(3)
        if (f != f0) win.topup();
        DCSB_DBG_LAP(0)
        // ---- frame header (:1780-1834): per iteration a run of 1-bit "unchanged" codes (count
        // leading ones) and the code behind it
        int rc = 0;
        for (int b = 0; b < nb;) {
            DCSB_DBG(++dbg_hdr;)
            const uint32_t v = win.peek32();
            const int ones = dcsb_clz(~v);
            const int left = nb - b;
            const int run = ones < left ? ones : left;          // <= 16
            b += run;
            if (b >= nb) { win.skip((uint32_t)run); break; }
            const uint32_t e = lut[DCSB_LUT_HDR94 + ((v << run) >> 24)];
            int delta;
            if (e == 0) {
                // codes longer than 8 bits: rare, matched bit-serially
                uint32_t q = win.pos() + (uint32_t)run;
                const int val = dcsb_long_code(rd, q, tab->long94, tab->n_long94);
                win.seek(q);
                if (val < 0) { rc = DCSB_WALK_BANDTYPE; break; }
                delta = val - 0x2E;
            } else {
                win.skip((uint32_t)run + (e >> 8));             // <= 24 bits
                delta = (int)(e & 0xFF) - 0x2E;
            }
            const uint32_t sh = (uint32_t)(b & 7) * 4u;
            const uint32_t w = b < 8 ? bt_lo : bt_hi;
            const int nbt = (int)((w >> sh) & 15u) + delta;
            if (nbt & ~15) { rc = DCSB_WALK_BANDTYPE; break; }
            const uint32_t nw = w + ((uint32_t)delta << sh);    // stays inside the nibble: 0 <= nbt <= 15
            if (b < 8) bt_lo = nw; else bt_hi = nw;
            const uint32_t d = dtab[dsel + ((halfmask >> b) & 1u) * 256u + (uint32_t)b * 16u + (uint32_t)nbt];
            desc[b] = d;
            live = (live & ~(1u << b)) | ((d ? 1u : 0u) << b);
            ++b;
        }
        // (a code that reaches into the bytes behind the stream is a truncation, whatever those bytes are)
        if (rc) { status = win.pos() > nbits ? -2 : rc; nplay = f; break; }
        DCSB_DBG_LAP(1)
        const uint32_t hpos = win.pos();
        out.hdrbits[s.frame_base + f] = (uint16_t)(hpos - pos);
        // ---- bands: lengths only
        int sb = 99;
        uint32_t m = live;
        int b = m ? dcsb_ctz(m) : 16;
        uint32_t d = desc[b];
        while (m) {
            m &= m - 1;
            const int bn = m ? dcsb_ctz(m) : 16;
            const uint32_t dn = desc[bn];                       // the next band's descriptor is on its way while this band is walked
            if (d & DCSB_DESC_HUFF) {
                // Huffman band (:2186-2225).  A table entry is {m8, m1}: m8 = as many whole codewords as
                // fit in the next 13 bits (at most 8 slots), m1 = the first codeword alone; each byte is
                // slots << 4 | bits.  Rs = (16 * slots left + 15) << 8 | 0xFF, so "the multi-symbol step
                // covers more slots than are left" is one compare on the raw entry, and the single
                // codeword is taken instead (a 'two zeros' codeword with one slot left leaves Rs < 0:
                // the reference's error case, :2213-2218).  t = 50 - s: the table index is the 64-bit
                // window shifted right by t, so the chain per step is shift, mask|base, load, compare,
                // select, subtract.
                const DcsbTxBase tb = tx + (((d >> 24) & 7u) << 14);
                int Rs = (int)(d & 0x3FFFFu);
                int t = 50 - (int)win.s;
                while (Rs > 0x0FFF) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        DCSB_DBG(++dbg_steps;)
                        const uint32_t v = (uint32_t)((((uint64_t)win.w0 << 32) | win.w1) >> t);
                        const uint32_t m16 = dcsb_tx_load(tb, v);
                        const uint32_t act = (uint32_t)((0x0FFF - Rs) >> 31);      // all ones while slots are left
                        const uint32_t b8 = (m16 >> 8) & act, b1 = m16 & 0xFFu & act;
                        const uint32_t mm = Rs >= (int)m16 ? b8 : b1;
                        t -= (int)(mm & 15u);
                        Rs -= (int)((mm & 0xF0u) << 8);
                    }
                    win.refill_t(t);
                }
                win.s = (uint32_t)(50 - t);
                if (Rs < 0 && sb > b) sb = b;        // 'two zeros' with one slot left (:2213-2218)
            } else {
                // fixed-width band (:2227-2234): count * width bits, closed form
                const uint32_t fbits = (d >> 16) & 0x3FFu;
                if (fbits <= 32) win.skip(fbits);
                else win.seek(win.pos() + fbits);
            }
            b = bn;
            d = dn;
        }
        DCSB_DBG_LAP(2)
        pos = win.pos();
        if (pos > nbits) { status = -2; nplay = f; break; }                       // DCSB_E_TRUNCATED
        if (sb != 99) { status = -5; nplay = f + 1; stopband = sb; ++f; break; }    // DCSB_E_STOPPED
        if ((f & 15) == 15) {
            // let the decode warps have the frames so far (entry f + 1 = this frame's own band types)
            out.bitpos[s.frame_base + f + 1] = pos;
            out.bt[s.frame_base + f + 1] = make_uint2(bt_lo, bt_hi);
            dcsb_publish(out.progress, si, f + 2);
        }
        if (f + 1 == qnext) {
            out.bitpos[s.frame_base + f + 1] = pos;
            out.bt[s.frame_base + f + 1] = make_uint2(bt_lo, bt_hi);
            dcsb_queue_push(out, si, queued, f + 1, false);
            queued = f + 1;
            qnext += DCSB_QITEM;
        }
    }
    // end checkpoint: band types after the last decodable frame (decode lanes read bt[f + 1]).
    // After a truncated / undecodable frame f the checkpoint of f itself already is the end.
    if ((status == 0 && s.nframes) || status == -5) {
        out.bitpos[s.frame_base + f] = pos;
        out.bt[s.frame_base + f] = make_uint2(bt_lo, bt_hi);
    }
#if DCSB_DEVICE_PASS
    asm volatile("cp.async.wait_group 0;" ::: "memory");    // nothing in flight when the ring is reused
#endif
    if (status == 0 && f < s.nframes) status = DCSB_SCAN_RUNNING;       // the next slice carries on from checkpoint f
#if defined(DCSB_SCAN_DEBUG) && DCSB_DEVICE_PASS
    if (out.dbg) {
        const long long dt = clock64() - dbg_t0;
        out.dbg[8 * si] = (uint32_t)dt; out.dbg[8 * si + 1] = (uint32_t)(dt >> 32);
        out.dbg[8 * si + 2] = dbg_steps; out.dbg[8 * si + 3] = dbg_hdr;
        for (int k = 0; k < 4; ++k) out.dbg[8 * si + 4 + k] = dbg_c[k];     // topup, header, huffman loops, rest
    }
#endif
    out.status[si] = status;
    out.nplay[si] = nplay;
    out.endbits[si] = status == -2 ? nbits : pos;      // a truncated stream occupies all of its bytes
    out.stopband[si] = (uint8_t)stopband;
    if (status == DCSB_SCAN_RUNNING) return;
    dcsb_publish(out.progress, si, DCSB_SCAN_DONE);
    dcsb_queue_push(out, si, queued, s.out_frames, true);
}
