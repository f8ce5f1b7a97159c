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

// Warp votes: on the device all 32 lanes of a warp walk one stream each in LOCK STEP (one
// instruction stream serves 32 streams; loops run until every lane is through); the simulator
// plays a warp of one lane, so a vote is the lane's own predicate.
#if DCSB_DEVICE_PASS
#define DCSB_ANY(p) (__any_sync(0xffffffffu, (p)) != 0)
#else
#define DCSB_ANY(p) (p)
#endif

// [f0, f1) = the frames this call walks (0, ~0u = the whole stream).  A call with f0 > 0 resumes
// from the end checkpoint the previous call left at frame f0 (status DCSB_SCAN_RUNNING); that is
// what lets dcsb_decode_streams cut a chunk into time slices whose PCM drains over PCIe while
// the later slices are still being scanned.
//
// LOCK STEP.  Called by all lanes of a warp together, lane = stream (si < 0: idle lane).  The
// lanes run ONE instruction stream: per frame a header loop (one run of "unchanged" codes plus
// one code per iteration) and a band loop whose iteration is one table step for whichever
// Huffman band the lane is in -- the switch to the next band is folded in by selects, a lane
// that has finished its frame executes no-ops until the last lane is through.  What is rare and
// long (a fixed-width band's closed-form skip needs the bit window re-seeked in the ring; header
// codes longer than 8 bits) sits behind warp-uniform branches on votes taken one iteration
// earlier, so the branch resolves at once and the lanes that do not need it pay nothing on their
// dependent chain.  The window refill is decided from the position BEFORE the current step
// (the 64-bit window has the room: s <= 44), which takes it off the chain position -> table
// entry -> next position.  Streams are ordered by cost (dcsb_scan_order), so a warp holds alike
// streams and waits little for its slowest lane.
DCSB_HD void dcsb_scan94_stream(const uint8_t *slab, const DcsbStreamRec *streams, int si, const DcsbTables *tab,
                                const uint16_t *lut, DcsbTxBase tx, const uint32_t *dtab, DcsbRingPtr ring, uint32_t *desc,
                                const DcsbScanOut &out, uint32_t f0 = 0, uint32_t f1 = 0xFFFFFFFFu)
{
    bool mine = si >= 0;
    if (mine && f0 && out.status[si] != DCSB_SCAN_RUNNING) mine = false;      // finished (or failed) in an earlier slice
    DcsbStreamRec s;
    if (mine) s = streams[si];
    else { s.nframes = 0; s.nbytes = 0; s.hdr_len = 16; s.data_off = 0; s.frame_base = 0; s.out_frames = 0; }
    const uint8_t *hdr = mine ? streams[si].hdr : streams[0].hdr;
    const int type1 = mine ? hdr[0] >> 7 : 0;
    int nb = 0;
    if (mine) while (nb < 16 && (hdr[nb] & 0x7F) != 0x7F) ++nb;
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
    if (mine && f0) {
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
    for (int b = nb; b <= 16; ++b) desc[b] = 0;
    bool run = mine && f < fe;                     // this lane still walks frames
    while (DCSB_ANY(run)) {
        if (run) {
            out.bitpos[s.frame_base + f] = pos;
            out.bt[s.frame_base + f] = make_uint2(bt_lo, bt_hi);
            if (f != f0) win.topup();
        }
        // ---- frame header (:1780-1834): per iteration a run of 1-bit "unchanged" codes (count
        // leading ones) and the code behind it
        int rc = 0;
        int hb = run ? 0 : nb;
        while (DCSB_ANY(hb < nb)) {
            const bool on = hb < nb;
            const uint32_t v = win.peek32();
            const int ones = dcsb_clz(~v);
            const int left = nb - hb;
            const int runl = ones < left ? ones : left;         // <= 16
            const int b = hb + runl;
            const bool code = on && b < nb;                     // a code follows the run
            uint32_t e = lut[DCSB_LUT_HDR94 + ((v << runl) >> 24)];
            int delta = (int)(e & 0xFF) - 0x2E;
            uint32_t adv = (uint32_t)runl + (code ? (e >> 8) : 0u);      // <= 24 bits
            if (DCSB_ANY(code && e == 0)) {
                // codes longer than 8 bits: rare, matched bit-serially
                if (code && e == 0) {
                    uint32_t q = win.pos() + (uint32_t)runl;
                    const int val = dcsb_long_code(rd, q, tab->long94, tab->n_long94);
                    win.seek(q);
                    adv = 0;
                    if (val < 0) { rc = DCSB_WALK_BANDTYPE; hb = nb; }
                    delta = val - 0x2E;
                }
            }
            if (on) win.skip(adv);
            if (code && !rc) {
                const uint32_t sh = (uint32_t)(b & 7) * 4u;
                const uint32_t w = b < 8 ? bt_lo : bt_hi;
                const int nbt = (int)((w >> sh) & 15u) + delta;
                if (nbt & ~15) { rc = DCSB_WALK_BANDTYPE; hb = nb; }
                else {
                    const uint32_t nw = w + ((uint32_t)delta << sh);    // stays inside the nibble: 0 <= nbt <= 15
                    if (b < 8) bt_lo = nw; else bt_hi = nw;
                    const uint32_t d = dtab[dsel + ((halfmask >> b) & 1u) * 256u + (uint32_t)b * 16u + (uint32_t)nbt];
                    desc[b] = d;
                    live = (live & ~(1u << b)) | ((d ? 1u : 0u) << b);
                    hb = b + 1;
                }
            } else if (on && !rc) hb = nb;                      // the run reached the last band
        }
        const uint32_t hpos = win.pos();
        // (a code that reaches into the bytes behind the stream is a truncation, whatever those bytes are)
        if (run && rc) { status = hpos > nbits ? -2 : rc; nplay = f; run = false; }
        if (run) out.hdrbits[s.frame_base + f] = (uint16_t)(hpos - pos);
        // ---- bands: lengths only.  One table step per iteration for whichever Huffman band the lane
        // is in (:2186-2225).  A table entry is {m8, m1}: m8 = as many whole codewords as fit in the
        // next 13 bits (at most 8 slots), m1 = the first codeword alone; each byte is slots << 4 | bits.
        // Rs = (16 * slots left + 15) << 8 | 0xFF, so "the multi-symbol step covers more slots than are
        // left" is one compare on the raw entry, and the single codeword is taken instead (a 'two
        // zeros' codeword with one slot left leaves Rs < 0: the reference's error case, :2213-2218).
        // t = 50 - s: the table index is the 64-bit window shifted right by t.
        int sb = 99;
        {
            // dn / bn: descriptor and index of the next band to take; m: the bands behind it.  The
            // descriptor behind dn is loaded every iteration (desc[16] = 0 ends the list), so that a
            // band switch is a handful of selects on values that are already there.
            uint32_t m = run ? live : 0u;
            int bn = m ? dcsb_ctz(m) : 16;
            m &= m - 1;
            uint32_t dn = desc[bn];
            int bcur = 16, Rs = 0;
            DcsbTxBase tb = tx;
            int t = 50 - (int)win.s;
            uint32_t fix = 0;                                   // bits of a fixed-width band waiting for the re-seek
            bool fin = dn == 0;                                 // nothing (left) to walk in this frame
            for (;;) {
                // votes on the state the previous iteration left, consumed at the END of this one: the
                // branches resolve long before they are reached (one idle iteration per frame and per
                // fixed-width band is the price)
                const bool any_fix = DCSB_ANY(fix != 0);
                const bool alive = DCSB_ANY(!fin);
                const int bnn = m ? dcsb_ctz(m) : 16;
                const uint32_t dnn = desc[bnn];
                const DcsbTxBase tbn = tx + (((dn >> 24) & 7u) << 14);
                const int Rsn = (int)dn < 0 ? (int)(dn & 0x3FFFFu) : 0;         // Huffman band: its slot budget
                const uint32_t fixn = (int)dn < 0 ? 0u : (dn >> 16) & 0x3FFu;   // fixed-width band (:2227-2234): count * width bits
                // -- one table step (a no-op once the band has no slots left)
                const uint32_t v = (uint32_t)((((uint64_t)win.w0 << 32) | win.w1) >> t);
                const uint32_t m16 = dcsb_tx_load(tb, v);
                const uint32_t rmask = (uint32_t)((t - 19) >> 31);              // all ones when s >= 32: drop a word
                {
                    const uint32_t nw1 = DcsbBits::be(win.nx), ld = win.ring_word(win.wa);
                    win.w0 ^= (win.w0 ^ win.w1) & rmask;
                    win.w1 ^= (win.w1 ^ nw1) & rmask;
                    win.nx ^= (win.nx ^ ld) & rmask;
                    win.wa += 4u & rmask;
                }
                const uint32_t act = (uint32_t)((0x0FFF - Rs) >> 31);           // all ones while slots are left
                const uint32_t b8 = (m16 >> 8) & act, b1 = m16 & 0xFFu & act;
                const uint32_t mm = Rs >= (int)m16 ? b8 : b1;
                t = t - (int)(mm & 15u) + (int)(32u & rmask);
                Rs -= (int)((mm & 0xF0u) << 8);
                sb = (Rs < 0 && sb > bcur) ? bcur : sb;                         // 'two zeros' with one slot left (:2213-2218)
                // -- band switch: take the next band once this one has no slots left
                const bool take = Rs <= 0x0FFF && fix == 0 && dn != 0;
                tb = take ? tbn : tb;
                Rs = take ? Rsn : Rs;
                fix = take ? fixn : fix;
                bcur = take ? bn : bcur;
                bn = take ? bnn : bn;
                dn = take ? dnn : dn;
                m = take ? (m & (m - 1)) : m;
                fin = Rs <= 0x0FFF && fix == 0 && dn == 0;
                // -- closed-form skip of a fixed-width band: re-seek the window in the ring
                if (any_fix) {
                    if (fix) {
                        win.s = (uint32_t)(50 - t);
                        win.seek(win.pos() + fix);
                        t = 50 - (int)win.s;
                        fix = 0;
                        fin = Rs <= 0x0FFF && dn == 0;
                    }
                }
                if (!alive) break;
            }
            win.s = (uint32_t)(50 - t);
            win.refill();
        }
        if (run) {
            pos = win.pos();
            if (pos > nbits) { status = -2; nplay = f; run = false; }                       // DCSB_E_TRUNCATED
            else if (sb != 99) { status = -5; nplay = f + 1; stopband = sb; ++f; run = false; }    // DCSB_E_STOPPED
            else {
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
                ++f;
                run = f < fe;
            }
        }
    }
#if DCSB_DEVICE_PASS
    asm volatile("cp.async.wait_group 0;" ::: "memory");    // nothing in flight when the ring is reused
#endif
    if (!mine) return;
    // end checkpoint: band types after the last decodable frame (decode lanes read bt[f + 1]).
    // After a truncated / undecodable frame f the checkpoint of f itself already is the end.
    if ((status == 0 && s.nframes) || status == -5) {
        out.bitpos[s.frame_base + f] = pos;
        out.bt[s.frame_base + f] = make_uint2(bt_lo, bt_hi);
    }
    if (status == 0 && f < s.nframes) status = DCSB_SCAN_RUNNING;       // the next slice carries on from checkpoint f
    out.status[si] = status;
    out.nplay[si] = nplay;
    out.endbits[si] = status == -2 ? nbits : pos;      // a truncated stream occupies all of its bytes
    out.stopband[si] = (uint8_t)stopband;
    if (status == DCSB_SCAN_RUNNING) return;
    dcsb_publish(out.progress, si, DCSB_SCAN_DONE);
    dcsb_queue_push(out, si, queued, s.out_frames, true);
}
