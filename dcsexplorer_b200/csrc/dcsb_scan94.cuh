// dcsb200 K1 body for the 1994+ frame layout: the frame-boundary scan, walked in LOCK STEP.
//
// A warp walks 32 streams, lane = stream (lengths only), and writes one checkpoint per frame plus
// the end checkpoint:
//   bitpos[f]  bit position of the frame start,   bt[f] band types carried INTO frame f,
//   hdrbits[f] length of the frame header (the decode lanes start at the first band and take
//              the frame's own band types from bt[f + 1]).
// It replaces the frame walk implicit in DecodeStream / GetStreamInfo (DCSDecoderNative.cpp:
// 1486-1589) and the header part of DecompressFrame (:1780-1834).
//
// Every stream is one dependent chain (position -> bits at that position -> next position) and
// the 32 lanes of a warp run ONE instruction stream, so what the scan costs is (iterations of
// the slowest lane) x (instructions per iteration): a warp alone on its scheduler issues one
// ALU instruction every other cycle.  The band loop's iteration is therefore kept to two
// dozen instructions:
//  * state of a lane in ONE register S = slots left << 12 | 0xF00 | t, t = 50 - bit offset in the
//    64-bit window.  A table entry is {y1 << 16 | y8}, y = slots << 12 | bits: y8 = as many whole
//    codewords as fit in the next 12 bits (at most 15 slots), y1 = the first codeword alone, taken
//    when y8 would cover more slots than the band has left (one compare S >= y8; the 0xF00 filler
//    makes the comparison independent of t).  The step is S = S - y + refill: position and slot
//    budget move with one add.  (A 'two zeros' codeword with one slot left leaves S < 0: the
//    reference's error case, :2213-2218);
//  * the window refill is decided from the position BEFORE the step (the 64-bit window has the
//    room: bit offset <= 43) and done with predicated moves, off the dependent chain;
//  * the bands of a lane are 16-byte entries {table, slots, mask, fixed bits} in a lane-private
//    array, pre-expanded when a frame header changes a band's type; the next entry is loaded an
//    iteration ahead, so the band switch is five predicated moves folded into the iteration;
//  * a lane without slots (frame finished, or waiting) looks up a zero entry (mask 0): a no-op;
//  * a fixed-width band's closed-form skip needs the window re-seeked in the ring: that block
//    is only compiled into the loop variant used for frames in which some lane has such a band;
//  * the loop's exit vote is taken on the state of the previous iteration, so the branch
//    resolves long before it is reached.
// The stream bytes are staged in a per-stream 1 KB shared-memory ring that cp.async fills a whole
// frame ahead.  Streams are ordered by cost (dcsb_scan_order), so a warp holds alike streams.
//
// Compiled by nvcc for sm_100a (the product) and by g++ for the CPU-side kernel simulator
// (tests/hostsim, test infrastructure only: a warp of ONE lane; cp.async becomes an immediate copy).
#pragma once
#include <string.h>
#include "dcsb_core.cuh"
#include "dcsb_fast94.cuh"

#define DCSB_RING_BYTES  1024u
#define DCSB_RING_CHUNKS (DCSB_RING_BYTES / 16u)
// the farthest a frame can reach past its first chunk: 16 header codes of up to 23 bits plus
// 255 samples of up to 15 bits (543 bytes), plus the window's words and chunk rounding
#define DCSB_RING_FRAME_CHUNKS 36u

// Shared-memory addresses: 32-bit shared-window addresses on the device (so that table | index
// is one LOP3 and loads are plain LDS), pointers in the simulator.
#if DCSB_DEVICE_PASS
typedef uint32_t DcsbSA;
DCSB_HD uint32_t dcsb_lds32(DcsbSA a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
DCSB_HD DcsbSA dcsb_sa_or(DcsbSA base, uint32_t off) { return base | off; }     // base is aligned beyond off's range
#else
typedef uintptr_t DcsbSA;
DCSB_HD uint32_t dcsb_lds32(DcsbSA a) { uint32_t v; memcpy(&v, reinterpret_cast<const void *>(a), 4); return v; }
DCSB_HD DcsbSA dcsb_sa_or(DcsbSA base, uint32_t off) { return base + off; }
#endif
typedef DcsbSA DcsbTxBase;
typedef DcsbSA DcsbRingPtr;
template <bool RING> struct DcsbWinT;
// (a & b) | c in one LOP3
DCSB_HD uint32_t dcsb_and_or(uint32_t a, uint32_t b, uint32_t c)
{
#if DCSB_DEVICE_PASS
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    return (a & b) | c;
#endif
}

// One band of a lane's current frame, as the band loop takes it.
struct DcsbBandEnt {
    DcsbSA tb;          // length table of the band's codebook (16 KB aligned), or the zero word
    uint32_t sinit;     // slots << 12 | 0xF00: the lane state the band starts with (the low byte, t, is kept)
    uint32_t amask;     // 0x3FFC (table index bits) or 0 (no table: the lookup reads the zero word)
    int32_t fix;        // > 0: fixed-width band, bits to skip; -1: end of the frame's list; else 0
};
#if DCSB_DEVICE_PASS
static_assert(sizeof(DcsbBandEnt) == 16, "band entries are loaded with one LDS.128");
#endif
#define DCSB_ENT_BYTES ((uint32_t)sizeof(DcsbBandEnt))

DCSB_HD DcsbBandEnt dcsb_ent_load(DcsbSA a)
{
    DcsbBandEnt e;
#if DCSB_DEVICE_PASS
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e.tb), "=r"(e.sinit), "=r"(e.amask), "=r"(e.fix) : "r"(a));
#else
    memcpy(&e, reinterpret_cast<const void *>(a), sizeof(e));
#endif
    return e;
}
DCSB_HD void dcsb_ent_store(DcsbSA a, const DcsbBandEnt &e)
{
#if DCSB_DEVICE_PASS
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(e.tb), "r"(e.sinit), "r"(e.amask), "r"(e.fix) : "memory");
#else
    memcpy(reinterpret_cast<void *>(a), &e, sizeof(e));
#endif
}

// Bit window over the stream: w0:w1 = 64 stream bits (big-endian order), s = bit offset of the next
// bit inside them, wa = byte offset (from chunk 0, cumulative) of the next word to take.
// RING = true: the words come from the lane's shared-memory ring (cp.async keeps it a frame ahead):
// the single-wave case, where the scan is as long as its slowest warp's dependent chain and a load's
// latency sits on it.  RING = false: the words are read from global memory through L1 (a lane reads
// its stream word by word, so 31 of 32 loads hit the line it has already fetched; the frame's next lines
// are prefetched at the frame start): no 32 KB of rings per warp, so many more warps fit an SM -- the
// multi-wave case, where what counts is how many warps an SM keeps in flight.
template <bool RING>
struct DcsbWinT {
    uint32_t w0, w1;
    uint32_t s;
    uint32_t wa;
    uint32_t bias;          // bit offset of the stream's first data bit from chunk 0
    uint32_t fill;          // chunks issued so far
    uint32_t limit;         // chunks that exist (stream bytes + slack)
    const uint8_t *g;       // global address of chunk 0 (16-byte aligned)
    DcsbRingPtr ring;       // this stream's ring (1 KB aligned in the shared window)

    DCSB_HD uint32_t ring_word(uint32_t off) const
    {
        if (RING) return dcsb_lds32(dcsb_sa_or(ring, off & (DCSB_RING_BYTES - 1)));
#if DCSB_DEVICE_PASS
        uint32_t v;
        asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(g + off));
        return v;
#else
        uint32_t v;
        memcpy(&v, g + off, 4);
        return v;
#endif
    }
    // issue the chunks up to DCSB_RING_CHUNKS - 1 ahead of the window, then make sure everything
    // the next frame can touch has landed.  Call at a frame start only (s < 32).
    DCSB_HD void topup()
    {
        if (!RING) {
#if DCSB_DEVICE_PASS
            // the two lines behind the one the window is in: what an average frame reaches
            asm volatile("prefetch.global.L1 [%0];" ::"l"(g + wa + 128u));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(g + wa + 256u));
#endif
            return;
        }
        const uint32_t cc = (wa - 8u) >> 4;                        // chunk holding w0
        uint32_t target = cc + DCSB_RING_CHUNKS - 1u;
        if (target > limit) target = limit;
        const bool behind = fill < cc + DCSB_RING_FRAME_CHUNKS && fill < limit;
        while (fill < target) {
#if DCSB_DEVICE_PASS
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(ring | ((fill * 16u) & (DCSB_RING_BYTES - 1))),
                         "l"(g + (size_t)fill * 16u) : "memory");
#else
            memcpy(reinterpret_cast<void *>(ring + ((fill * 16u) & (DCSB_RING_BYTES - 1))), g + (size_t)fill * 16u, 16);
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
        wa = off + 8u;
    }
    DCSB_HD uint32_t pos() const { return (wa - 8u) * 8u + s - bias; }
    DCSB_HD uint32_t peek32() const      // s < 32
    {
#if DCSB_DEVICE_PASS
        return __funnelshift_l(w1, w0, s);
#else
        return s ? ((w0 << s) | (w1 >> (32 - s))) : w0;
#endif
    }
    DCSB_HD void refill()                // s < 64
    {
        if (s >= 32u) {
            s -= 32u;
            w0 = w1;
            w1 = DcsbBits::be(ring_word(wa));
            wa += 4u;
        }
    }
    DCSB_HD void skip(uint32_t n) { s += n; refill(); }      // s < 32 on entry, n <= 32
};

// Band descriptor (compressed; 4 KB table per CTA): what a band of a frame is, by (stream type,
// half-density flag of the band, band, band type): dtab[((type1 * 2 + half) * 16 + band) * 16 + type]
//   bit 15 Huffman band: bits 10..12 = codebook - 1, bits 0..5 = slots
//   bit 14 fixed-width band: bits 0..9 = slots * width (bits to skip)
//   0      empty band
#define DCSB_DESC_HUFF  0x8000u
#define DCSB_DESC_FIXED 0x4000u
#define DCSB_DTAB_WORDS 1024
DCSB_HD uint32_t dcsb_band_desc94(const uint16_t *lut, int type1, int b, int nib, int count)
{
    int code = nib;
    if (type1) code = (int)(lut[DCSB_LUT_XLAT + (b < 3 ? 0 : (b < 6 ? 16 : 32)) + code] >> 8);      // :1926-1955
    if (code >= 1 && code <= 6 && count) return DCSB_DESC_HUFF | ((uint32_t)(code - 1) << 10) | (uint32_t)count;
    if (code > 6 && count) return DCSB_DESC_FIXED | (uint32_t)(count * code);
    return 0;
}
DCSB_HD uint32_t dcsb_dtab_entry(const uint16_t *lut, int i)
{
    const int nib = i & 15, b = (i >> 4) & 15, half = (i >> 8) & 1, type1 = i >> 9;
    return dcsb_band_desc94(lut, type1, b, nib, dcsb_band_count94(b) >> half);
}
// expand a descriptor into the entry the band loop takes
DCSB_HD DcsbBandEnt dcsb_band_entry(uint32_t d, DcsbTxBase tx, DcsbSA zero)
{
    DcsbBandEnt e;
    const bool huff = (d & DCSB_DESC_HUFF) != 0;
    e.tb = huff ? tx + (DcsbSA)(((d >> 10) & 7u) << 14) : zero;
    const bool fixed = (d & DCSB_DESC_FIXED) != 0;
    e.sinit = ((huff ? (d & 0x3Fu) : (fixed ? 1u : 0u)) << 12) | 0xF00u;            // a fixed-width band parks the lane (one slot, no table) until the re-seek
    e.amask = huff ? 0x3FFCu : 0u;
    e.fix = fixed ? (int32_t)(d & 0x3FFu) : 0;
    return e;
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

// Warp votes: the simulator plays a warp of one lane, so a vote is the lane's own predicate.
#if DCSB_DEVICE_PASS
#define DCSB_ANY(p) (__any_sync(0xffffffffu, (p)) != 0)
#define DCSB_REDUCE_OR(x) __reduce_or_sync(0xffffffffu, (x))
#else
#define DCSB_ANY(p) (p)
#define DCSB_REDUCE_OR(x) (x)
#endif

// an unconditional 16-bit table load (the compiler would otherwise predicate the second of two
// dependent-looking loads on the first one's result and serialise them)
DCSB_HD uint32_t dcsb_lut_load(const uint16_t *lut, uint32_t i)
{
#if DCSB_DEVICE_PASS
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(lut + i)));
    return v;
#else
    return lut[i];
#endif
}

// Small device/host helpers for the predicate-free inner loop
DCSB_HD uint32_t dcsb_lds16(DcsbSA a)
{
#if DCSB_DEVICE_PASS
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
#else
    uint16_t v;
    memcpy(&v, reinterpret_cast<const void *>(a), 2);
    return v;
#endif
}
DCSB_HD int dcsb_umin(int a, int b) { return (uint32_t)a < (uint32_t)b ? a : b; }
// m ? a : b for an all-ones / all-zeros mask m, one LOP3
DCSB_HD uint32_t dcsb_msel(uint32_t m, uint32_t a, uint32_t b)
{
#if DCSB_DEVICE_PASS
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xAC;" : "=r"(r) : "r"(m), "r"(b), "r"(a));       // F(m=0xF0, b=0xCC, a=0xAA) = (m & a) | (~m & b)
    return r;
#else
    return (m & a) | (~m & b);
#endif
}

// The band loop of one frame (:2186-2234, lengths only).  ents = the lane's band entries, the list
// ends with an entry whose fix is -1.  FIX: some lane of the warp has a fixed-width band in this frame.
// Returns the OR of the lane states the steps left (negative: some band overran its slot budget,
// :2213-2218 -- the caller then walks the frame again the slow way, whatever this loop made of it).
//
// One iteration, free of predicates on the dependent chain:
//   e8 / e1   the two halves of the table entry, loaded as two LDS.U16 (no unpacking)
//   S         = min_unsigned(S - e8, S - e1): the multi-codeword step when it fits the slot budget,
//             else the single codeword (S - e8 is negative = huge exactly when it does not fit)
//   dm        all ones when the band has no slots left: the switch to the next entry (reloaded
//             every iteration from ptr, which only then advances) is four mask selects
//   t         for the next lookup comes from S BEFORE the switch (the switch keeps the low byte)
// Four iterations per exit vote and fixed-band check (eight measured the same: fewer votes, more idle iterations at the frame end): a lane that is through parks on its end entry.
#define DCSB_SCAN_UNROLL 4
template <bool FIX, bool RING>
DCSB_HD int dcsb_scan94_bands(DcsbWinT<RING> &win, DcsbSA ents, DcsbSA ents_end, DcsbSA zero, bool run)
{
    DcsbSA ptr = run ? ents : ents_end;
    DcsbSA tb = zero;
    uint32_t amask = 0;
    int t = 50 - (int)win.s;
    int S = 0xF00 | t;                                      // no slots: the first iteration takes the first band
    int fix = 0, err = 0;
    (void)fix;
    uint32_t w0 = win.w0, w1 = win.w1, wa = win.wa;
    for (;;) {
        // (state the previous group left: resolves early.  A lane is through once it has taken its end entry)
        const bool alive = DCSB_ANY(FIX ? fix >= 0 : ptr <= ents_end);
#pragma unroll
        for (int u = 0; u < DCSB_SCAN_UNROLL; ++u) {
            // (the ring word a refill would take and the next band entry: addresses off the chain)
#if DCSB_DEVICE_PASS
            const uint32_t ld = RING ? dcsb_lds32(dcsb_and_or(wa, DCSB_RING_BYTES - 1, win.ring)) : win.ring_word(wa);
#else
            const uint32_t ld = win.ring_word(wa);
#endif
            const DcsbBandEnt en = dcsb_ent_load(ptr);
            // -- one table step (a no-op for a lane without table: the zero word)
#if DCSB_DEVICE_PASS
            const uint32_t v = (uint32_t)((((uint64_t)w0 << 32) | w1) >> t);
            const DcsbSA ea = dcsb_and_or(v, amask, tb);
#else
            const uint32_t v = (uint32_t)((((uint64_t)w0 << 32) | w1) >> (t & 63));     // (t < 64 unless the frame is damaged)
            const DcsbSA ea = tb + (v & amask);
#endif
            const int e8 = (int)dcsb_lds16(ea), e1 = (int)dcsb_lds16(ea + 2);
            // -- window refill, decided before the step: drop a word once the bit offset has reached 32
            const bool rf = t < 19;
            const int Sr = S + (rf ? 32 : 0);
            w0 = rf ? w1 : w0;
            w1 = rf ? DcsbBits::be(ld) : w1;
            wa = rf ? wa + 4u : wa;
            S = dcsb_umin(Sr - e8, Sr - e1);
            err |= S;
            t = S & 0xFF;
            // -- band switch: take the next band once this one has no slots left.  The state is REPLACED
            // (not topped up): whatever a damaged frame made of S, the lane has a sane slot budget again
            // and parks on the end entry at the latest.  (The end entry and a fixed-width band's entry
            // have a slot and no table: the lane parks on them.)
            const int dm = (S - 0x1000) >> 31;
            S = (int)dcsb_msel((uint32_t)dm, dcsb_and_or((uint32_t)S, 0xFFu, en.sinit), (uint32_t)S);
#if DCSB_DEVICE_PASS
            tb = dcsb_msel((uint32_t)dm, en.tb, tb);
#else
            tb = dm ? en.tb : tb;
#endif
            amask = dcsb_msel((uint32_t)dm, en.amask, amask);
            if (FIX) fix = (int)dcsb_msel((uint32_t)dm, (uint32_t)en.fix, (uint32_t)fix);
            ptr = ptr - (DcsbSA)dm * (DcsbSA)DCSB_ENT_BYTES;        // dm = -1: the next entry
        }
        // -- closed-form skip of a fixed-width band: re-seek the window in the ring, then take the next band
        if (FIX) {
            if (fix > 0) {
                const uint32_t a = (wa - 8u) * 8u + (uint32_t)(50 - t) + (uint32_t)fix;
                const uint32_t off = (a >> 5) * 4u;
                w0 = DcsbBits::be(win.ring_word(off));
                w1 = DcsbBits::be(win.ring_word(off + 4u));
                wa = off + 8u;
                t = 50 - (int)(a & 31u);
                S = 0xF00 | t;
                fix = 0;
            }
        }
        if (!alive) break;
    }
    win.w0 = w0;
    win.w1 = w1;
    win.wa = wa;
    win.s = (uint32_t)(50 - t);
    win.refill();
    return err;
}

// Rare path: a band of frame f overran its slot budget (a 'two zeros' codeword with one slot left:
// the decode kernel zeroes that band's contribution and the channel stops, :2213-2218).  Walks the
// frame's bands one codeword at a time on global memory, the way the reference does: the overrunning
// codeword is consumed, the band ends, the walk goes on.  pos = first band's bit position on entry,
// the frame's end on return; returns the first such band (99 if none).
DCSB_HD int dcsb_find_stopband94(const DcsbBits &rd, uint32_t &pos, const uint8_t *hdr, const uint16_t *lut, uint32_t bt_lo, uint32_t bt_hi)
{
    const int type1 = hdr[0] >> 7;
    int sb = 99;
    for (int b = 0; b < 16; ++b) {
        const int hb = hdr[b] & 0x7F;
        if (hb == 0x7F) break;
        int count = dcsb_band_count94(b);
        if (hb & 0x40) count >>= 1;
        int code = dcsb_nib32(bt_lo, bt_hi, b);
        if (type1) code = (int)(lut[DCSB_LUT_XLAT + (b < 3 ? 0 : (b < 6 ? 16 : 32)) + code] >> 8);
        if (code == 0) continue;
        if (code > 6) { pos += (uint32_t)(count * code); continue; }
        const uint16_t *cb = lut + DCSB_LUT_CB + dcsb_cb_ofs(code);
        const int mw = dcsb_cb_maxw(code);
        int rem = count;
        while (rem > 0) {
            const uint32_t e = cb[rd.peek(pos, mw)];
            pos += e >> 12;
            const int st = (e & 0x800u) ? 2 : 1;
            if (st > rem && sb > b) sb = b;
            rem -= st;
        }
    }
    return sb;
}

// One frame-header code: band b's type moves by delta (:1822-1834); the band's entry is expanded
// again.  next = the band the header walk carries on with (nb after an error: rc is set).
DCSB_HD void dcsb_hdr_apply94(int b, int delta, int nb, uint32_t dsel, uint32_t halfmask, const uint16_t *dtab, DcsbTxBase tx, DcsbSA zero,
                              DcsbSA ents, uint32_t &bt_lo, uint32_t &bt_hi, uint32_t &fixmask, int &rc, int &next)
{
    const uint32_t sh = (uint32_t)(b & 7) * 4u;
    const bool hi = b >= 8;
    const uint32_t w = hi ? bt_hi : bt_lo;
    const int nbt = (int)((w >> sh) & 15u) + delta;
    const bool bad = (nbt & ~15) != 0;
    const uint32_t nw = w + ((uint32_t)delta << sh);            // stays inside the nibble when 0 <= nbt <= 15
    bt_lo = (!bad && !hi) ? nw : bt_lo;
    bt_hi = (!bad && hi) ? nw : bt_hi;
    const uint32_t d = dtab[dsel + ((halfmask >> b) & 1u) * 256u + (uint32_t)b * 16u + (uint32_t)(nbt & 15)];
    if (!bad) dcsb_ent_store(ents + (DcsbSA)b * DCSB_ENT_BYTES, dcsb_band_entry(d, tx, zero));
    fixmask = bad ? fixmask : ((fixmask & ~(1u << b)) | (((d >> 14) & 1u) << b));
    rc = bad ? DCSB_WALK_BANDTYPE : rc;
    next = bad ? nb : b + 1;
}

// [f0, f1) = the frames this call walks (0, ~0u = the whole stream).  A call with f0 > 0 resumes
// from the end checkpoint the previous call left at frame f0 (status DCSB_SCAN_RUNNING); that is
// what lets dcsb_decode_streams cut a chunk into time slices whose PCM drains over PCIe while
// the later slices are still being scanned.
// Called by all lanes of a warp together, lane = stream (si < 0: idle lane).
// ents: the lane's 18 band entries (16 bands, the end entry, one more that is only loaded); zero: address of a zero word in shared memory.
template <bool RING>
DCSB_HD void dcsb_scan94_stream(const uint8_t *slab, const DcsbStreamRec *streams, int si, const DcsbTables *tab,
                                const uint16_t *lut, DcsbTxBase tx, const uint16_t *dtab, DcsbRingPtr ring, DcsbSA ents, DcsbSA zero,
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
    DcsbWinT<RING> win;
    win.g = slab + (start & ~15ull);
    win.bias = (uint32_t)(start & 15) * 8u;
    win.ring = ring;
    win.fill = 0;
    // chunks worth reading: the stream, plus the bytes a frame that starts inside it may still
    // reach into the zero padding (the slab keeps >= 1 KB of slack behind the last stream)
    win.limit = s.nframes ? (uint32_t)(((start & 15) + dbytes + 64u + 15u) >> 4) : 0u;
    win.wa = 8u;
    win.s = 0;
    uint32_t pos = 0;
    uint32_t bt_lo = 0, bt_hi = 0;                 // band types, 16 x 4 bits; InitStreamPlayback zeroes them (:1640)
    if (mine && f0) {
        pos = out.bitpos[s.frame_base + f0];
        const uint2 b2 = out.bt[s.frame_base + f0];
        bt_lo = b2.x;
        bt_hi = b2.y;
        const uint32_t off = ((pos + win.bias) >> 5) * 4u;
        win.wa = off + 8u;
        win.fill = off >> 4;
    }
    win.topup();
    win.seek(pos);
    // plain reader on global memory for the rare paths (long header codes, the band that overran)
    DcsbBits rd;
    rd.w = reinterpret_cast<const uint32_t *>(slab + (start & ~3ull));
    rd.bias = (uint32_t)(start & 3) * 8;

    uint32_t queued = 0, qnext = DCSB_QITEM;       // output frames already handed to the decode kernel / next hand-over
    int status = s.nframes ? 0 : -1, stopband = 0xFF;    // -1 = DCSB_E_EMPTY (the host refines DCSB_E_SHORT)
    uint32_t nplay = s.nframes, f = f0;
    const uint32_t fe = f1 < s.nframes ? f1 : s.nframes;
    // band entries of the current band types: only a band whose type changes in a frame header is
    // expanded again.  fixmask: the bands that are fixed-width right now.
    uint32_t fixmask = 0;
    for (int b = 0; b < nb; ++b) {
        const uint32_t d = dtab[dsel + ((halfmask >> b) & 1u) * 256u + (uint32_t)b * 16u + (uint32_t)dcsb_nib32(bt_lo, bt_hi, b)];
        dcsb_ent_store(ents + (DcsbSA)b * DCSB_ENT_BYTES, dcsb_band_entry(d, tx, zero));
        fixmask |= ((d >> 14) & 1u) << b;
    }
    const DcsbSA ents_end = ents + (DcsbSA)nb * DCSB_ENT_BYTES;
    {
        DcsbBandEnt term;
        term.tb = zero; term.sinit = 0x1F00u; term.amask = 0; term.fix = -1;       // one slot, no table: a lane parks here
        dcsb_ent_store(ents_end, term);
        dcsb_ent_store(ents_end + DCSB_ENT_BYTES, term);                            // (loaded behind the end entry, never taken)
    }
    bool run = mine && f < fe;                     // this lane still walks frames
    while (DCSB_ANY(run)) {
        if (run) {
            out.bitpos[s.frame_base + f] = pos;
            out.bt[s.frame_base + f] = make_uint2(bt_lo, bt_hi);
            if (f != f0) win.topup();
        }
        // ---- frame header (:1780-1834): per iteration a run of 1-bit "unchanged" codes (count
        // leading ones) and the code behind it.  Branch-free for the lanes (selects, predicated
        // stores); what is rare -- a code longer than 16 bits -- makes its lane wait one
        // iteration and is then served behind a warp-uniform branch.
        int rc = 0;
        int hb = run ? 0 : nb;
        int lng = 0;                                            // 1: a long code waits at band hb
        while (DCSB_ANY(hb < nb || lng)) {
            if (lng) {
                // codes longer than 16 bits: rare, matched bit-serially
                uint32_t q = win.pos();
                const int val = dcsb_long_code(rd, q, tab->long94, tab->n_long94);
                win.seek(q);
                lng = 0;
                if (val < 0) { rc = DCSB_WALK_BANDTYPE; hb = nb; }
                else dcsb_hdr_apply94(hb, val - 0x2E, nb, dsel, halfmask, dtab, tx, zero, ents, bt_lo, bt_hi, fixmask, rc, hb);
            }
            const uint32_t ld = win.ring_word(win.wa);
            const uint32_t v = win.peek32();
            const int ones = dcsb_clz(~v);
            const int left = nb - hb;
            const int runl = ones < left ? ones : left;         // <= 16; 0 for a lane that is through
            const int b = hb + runl;
            const bool code = b < nb;                           // a code follows the run
            // (codes of 9..16 bits all start with the same 8 bits: a second 8-bit LUT on the bits behind them)
            const uint32_t e1 = dcsb_lut_load(lut, DCSB_LUT_HDR94 + ((v << runl) >> 24));
            const uint32_t e2 = dcsb_lut_load(lut, DCSB_LUT_HDR94B + ((v << (runl + 8)) >> 24));     // (both loads in flight together)
            const uint32_t e = e1 ? e1 : e2;
            const bool islong = code && e == 0;
            const bool ok = code && e != 0;
            // consume the run and the code (a long code: the run only; its lane waits at band b)
            win.s += (uint32_t)runl + (ok ? (e >> 8) : 0u);      // <= 32 bits
            {
                const bool rf = win.s >= 32u;
                win.w0 = rf ? win.w1 : win.w0;
                win.w1 = rf ? DcsbBits::be(ld) : win.w1;
                win.wa = rf ? win.wa + 4u : win.wa;
                win.s &= 31u;
            }
            lng = islong ? 1 : 0;
            int nhb = nb;                                       // no code: the run reached the last band
            if (ok) dcsb_hdr_apply94(b, (int)(e & 0xFF) - 0x2E, nb, dsel, halfmask, dtab, tx, zero, ents, bt_lo, bt_hi, fixmask, rc, nhb);
            hb = islong ? b : nhb;
        }
        const uint32_t hpos = win.pos();
        // (a code that reaches into the bytes behind the stream is a truncation, whatever those bytes are)
        if (run && rc) { status = hpos > nbits ? -2 : rc; nplay = f; run = false; }
        if (run) out.hdrbits[s.frame_base + f] = (uint16_t)(hpos - pos);
        // ---- bands: lengths only
        int err;
        if (DCSB_ANY(run && fixmask != 0)) err = dcsb_scan94_bands<true, RING>(win, ents, ents_end, zero, run);
        else err = dcsb_scan94_bands<false, RING>(win, ents, ents_end, zero, run);
        if (run) {
            pos = win.pos();
            int sb = 99;
            if (err < 0) {                      // some band overran: the loop's position is not to be trusted
                pos = hpos;
                sb = dcsb_find_stopband94(rd, pos, hdr, lut, bt_lo, bt_hi);
                if (sb == 99 && pos <= nbits) win.seek(pos);
            }
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
    if (RING) asm volatile("cp.async.wait_group 0;" ::: "memory");    // nothing in flight when the ring is reused
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
