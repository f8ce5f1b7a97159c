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
//  * Huffman bands advance with a multi-symbol length table t8[codebook][next 13 bits] =
//    {bits consumed, output slots covered} that chains as many whole codewords as fit in the
//    peek (at most 8 slots).  When that step would cover more slots than the band has left, the
//    single-codeword table t1[codebook][next 9 bits] is used instead, so a step never overruns
//    the band and the table address depends on the bit position only, not on the slot count.
//    (A 'two zeros' codeword with one slot left leaves rem < 0: the reference's error case,
//    :2213-2218);
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
};

// Scan tables in shared memory (DcsbTables::t8 / t1): the kernel hands over 32-bit shared-window
// addresses so that a lookup is one add + LDS; the simulator passes plain pointers.
#if DCSB_DEVICE_PASS
typedef uint32_t DcsbSmemU8;
#define DCSB_SMEM_U8(ptr) ((uint32_t)__cvta_generic_to_shared(ptr))
DCSB_HD uint32_t dcsb_lds8(DcsbSmemU8 base, uint32_t idx)
{
    uint32_t v;
    // volatile: both table loads of a step are issued back to back instead of the second one
    // being sunk under the compare on the first
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(base + idx));
    return v;
}
typedef uint32_t DcsbRingPtr;
#else
typedef const uint8_t *DcsbSmemU8;
#define DCSB_SMEM_U8(ptr) (ptr)
DCSB_HD uint32_t dcsb_lds8(DcsbSmemU8 base, uint32_t idx) { return base[idx]; }
typedef uint8_t *DcsbRingPtr;
#endif

// [f0, f1) = the frames this call walks (0, ~0u = the whole stream).  A call with f0 > 0 resumes
// from the end checkpoint the previous call left at frame f0 (status DCSB_SCAN_RUNNING); that is
// what lets dcsb_decode_streams cut a chunk into time slices whose PCM drains over PCIe while
// the later slices are still being scanned.
DCSB_HD void dcsb_scan94_stream(const uint8_t *slab, const DcsbStreamRec *streams, int si, const DcsbTables *tab,
                                const uint16_t *lut, DcsbSmemU8 t8, DcsbSmemU8 t1, DcsbRingPtr ring, const DcsbScanOut &out,
                                uint32_t f0 = 0, uint32_t f1 = 0xFFFFFFFFu)
{
    if (f0 && out.status[si] != DCSB_SCAN_RUNNING) return;      // finished (or failed) in an earlier slice
    const DcsbStreamRec s = streams[si];
    const uint8_t *hdr = streams[si].hdr;
    const int type1 = hdr[0] >> 7;
    int nb = 0;
    while (nb < 16 && (hdr[nb] & 0x7F) != 0x7F) ++nb;
    // per-band slot count, 8 bits each (:1848-1862)
    uint64_t cnt_lo = 0, cnt_hi = 0;
    for (int b = 0; b < nb; ++b) {
        uint64_t c = (uint64_t)(dcsb_band_count94(b) >> ((hdr[b] >> 6) & 1));
        if (b < 8) cnt_lo |= c << (8 * b); else cnt_hi |= c << (8 * (b - 8));
    }
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
    uint64_t bt = 0;                               // InitStreamPlayback zeroes the band types (:1640)
    if (f0) {
        pos = out.bitpos[s.frame_base + f0];
        const uint2 b2 = out.bt[s.frame_base + f0];
        bt = ((uint64_t)b2.y << 32) | b2.x;
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
    for (; f < fe; ++f) {
        out.bitpos[s.frame_base + f] = pos;
        out.bt[s.frame_base + f] = make_uint2((uint32_t)bt, (uint32_t)(bt >> 32));
        DCSB_DBG_LAP(3)
        if (f != f0) win.topup();
        DCSB_DBG_LAP(0)
        // ---- frame header (:1780-1834)
        int rc = 0;
        for (int b = 0; b < nb;) {
            DCSB_DBG(++dbg_hdr;)
            const uint32_t v = win.peek32();
            const int ones = dcsb_clz(~v);
            const uint32_t e = lut[DCSB_LUT_HDR94 + (v >> 24)];
            if (ones == 0 && e == 0) {
                // codes longer than 8 bits: rare, matched bit-serially
                uint32_t q = win.pos();
                const int val = dcsb_long_code(rd, q, tab->long94, tab->n_long94);
                win.seek(q);
                const int nbt = dcsb_nib(bt, b) + val - 0x2E;
                if (val < 0 || nbt < 0 || nbt > 15) { rc = DCSB_WALK_BANDTYPE; break; }
                bt = (bt & ~(15ull << (4 * b))) | ((uint64_t)nbt << (4 * b));
                ++b;
                continue;
            }
            const int left = nb - b;
            const int run = ones < left ? ones : left;
            const bool unchanged = ones > 0;
            win.skip(unchanged ? (uint32_t)run : (e >> 8));
            const int nbt = dcsb_nib(bt, b) + (unchanged ? 0 : (int)(e & 0xFF) - 0x2E);
            if (nbt < 0 || nbt > 15) { rc = DCSB_WALK_BANDTYPE; break; }
            bt = (bt & ~(15ull << (4 * b))) | ((uint64_t)nbt << (4 * b));
            b += unchanged ? run : 1;
        }
        // (a code that reaches into the bytes behind the stream is a truncation, whatever those bytes are)
        if (rc) { status = win.pos() > nbits ? -2 : rc; nplay = f; break; }
        DCSB_DBG_LAP(1)
        const uint32_t hpos = win.pos();
        out.hdrbits[s.frame_base + f] = (uint16_t)(hpos - pos);
        // ---- bands: lengths only
        int sb = 99;
        for (int b = 0; b < nb; ++b) {
            int code = dcsb_nib(bt, b);
            const int count = (int)(((b < 8 ? cnt_lo >> (8 * b) : cnt_hi >> (8 * (b - 8)))) & 0xFF);
            if (type1) code = (int)(lut[DCSB_LUT_XLAT + (b < 3 ? 0 : (b < 6 ? 16 : 32)) + code] >> 8);
            if (code >= 1 && code <= 6) {
                DCSB_DBG_LAP(3)
                // Huffman band (:2186-2225).  R = 16 * (slots left) + 15, so that "the multi-symbol
                // step covers more slots than are left" is one compare on the raw table byte.
                const DcsbSmemU8 b8 = t8 + (uint32_t)(code - 1) * DCSB_T8_CB, b1 = t1 + (uint32_t)(code - 1) * DCSB_T1_CB;
                int R = count * 16 + 15;
                while (R > 15) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        DCSB_DBG(++dbg_steps;)
                        const uint32_t x = win.peek_wide();
                        const uint32_t m8 = dcsb_lds8(b8, x >> (32 - DCSB_T8_PEEK));
                        const uint32_t m1 = dcsb_lds8(b1, x >> (32 - DCSB_T1_PEEK));
                        const int active = (15 - R) >> 31;                  // all ones while slots are left
                        const int over = (R - (int)m8) >> 31;               // all ones: the multi-symbol step would overrun
                        const uint32_t m = (m8 ^ ((m8 ^ m1) & (uint32_t)over)) & (uint32_t)active;
                        win.advance(m & 15u);
                        R -= (int)(m & 0xF0u);
                    }
                    win.refill();
                }
                if (R < 0 && sb > b) sb = b;        // 'two zeros' with one slot left (:2213-2218)
            } else if (code > 6) {
                // fixed-width band (:2227-2234): count * code bits, closed form
                const uint32_t fbits = (uint32_t)(count * code);
                if (fbits <= 32) win.skip(fbits);
                else win.seek(win.pos() + fbits);
            }
        }
        pos = win.pos();
        if (pos > nbits) { status = -2; nplay = f; break; }                       // DCSB_E_TRUNCATED
        if (sb != 99) { status = -5; nplay = f + 1; stopband = sb; ++f; break; }    // DCSB_E_STOPPED
        if ((f & 15) == 15) {
            // let the decode warps have the frames so far (entry f + 1 = this frame's own band types)
            out.bitpos[s.frame_base + f + 1] = pos;
            out.bt[s.frame_base + f + 1] = make_uint2((uint32_t)bt, (uint32_t)(bt >> 32));
            dcsb_publish(out.progress, si, f + 2);
        }
        if (f + 1 == qnext) {
            out.bitpos[s.frame_base + f + 1] = pos;
            out.bt[s.frame_base + f + 1] = make_uint2((uint32_t)bt, (uint32_t)(bt >> 32));
            dcsb_queue_push(out, si, queued, f + 1, false);
            queued = f + 1;
            qnext += DCSB_QITEM;
        }
    }
    // end checkpoint: band types after the last decodable frame (decode lanes read bt[f + 1]).
    // After a truncated / undecodable frame f the checkpoint of f itself already is the end.
    if ((status == 0 && s.nframes) || status == -5) {
        out.bitpos[s.frame_base + f] = pos;
        out.bt[s.frame_base + f] = make_uint2((uint32_t)bt, (uint32_t)(bt >> 32));
    }
#if DCSB_DEVICE_PASS
    asm volatile("cp.async.wait_group 0;" ::: "memory");    // nothing in flight when the ring is reused
#endif
    if (status == 0 && f < s.nframes) status = DCSB_SCAN_RUNNING;       // the next slice carries on from checkpoint f
#if defined(DCSB_SCAN_DEBUG) && DCSB_DEVICE_PASS
    if (out.dbg) {
        const long long dt = clock64() - dbg_t0;
        out.dbg[4 * si] = (uint32_t)dt; out.dbg[4 * si + 1] = (uint32_t)(dt >> 32);
        out.dbg[4 * si + 2] = dbg_steps; out.dbg[4 * si + 3] = dbg_hdr;
        for (int k = 0; k < 4; ++k) out.dbg[4 * si + k] = dbg_c[k];     // topup, header, huffman loops, rest
        out.dbg[4 * si + 3] |= 0;
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
