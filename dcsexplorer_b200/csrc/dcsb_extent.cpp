// dcsb200 host side: the size of a stream whose caller does not know it (dcsb_stream_extent,
// include/dcsb200.h).  The reference's clients hand the decoder an unsized ROMPointer and let it
// read as far as the bits go (DCSEncoder.cpp:547-571); its GetStreamInfo finds a stream's size by
// walking every frame (DCSDecoderNative.cpp:1486-1537).  This walks the same frames with the
// frame-boundary scan's own body (dcsb_scan94.cuh / dcsb_core.cuh, written once for nvcc and for the
// host compiler), lengths only, reading the caller's memory only as far as the bits go (+ the bit
// window's look-ahead, DCSB_EXTENT_SLACK).  No PCM is produced here: decoding stays on the GPU.
#include <stdint.h>
#include <string.h>
#include <mutex>
#include <vector>
#include "dcsb_core.cuh"
#include "dcsb_fast94.cuh"
#include "dcsb_scan94.cuh"

extern "C" size_t dcsb_stream_extent(const uint8_t *data, int os_version)
{
    if (!data) return 0;
    const uint32_t nf = ((uint32_t)data[0] << 8) | data[1];
    if (nf == 0) return 0;
    static DcsbTables *tab = nullptr;
    static uint16_t dtab[DCSB_DTAB_WORDS];
    static std::once_flag once;
    std::call_once(once, [] {
        tab = new DcsbTables();
        dcsb_build_tables(tab);
        for (int i = 0; i < DCSB_DTAB_WORDS; ++i) dtab[i] = (uint16_t)dcsb_dtab_entry(tab->lut, i);
    });
    DcsbStreamRec r;
    memset(&r, 0, sizeof(r));
    r.nframes = nf;
    r.out_frames = nf;
    r.nbytes = 0x7FFFFFF0u;                     // not known: the walk is bounded by the frame count
    if (os_version == DCSB_OS94 || os_version == DCSB_OS95) { r.fmt = DCSB_FMT_94; r.hdr_len = 16; }
    else if (os_version == DCSB_OS93B) { r.fmt = DCSB_FMT_93; r.hdr_len = 16; }
    else if (os_version == DCSB_OS93A) {
        // OS93a: header byte 0 bit 7 selects the type-1 layout with its one-byte header (DCSDecoderNative.cpp:2850-2859)
        const bool t1 = (data[2] & 0x80) != 0;
        r.fmt = t1 ? DCSB_FMT_93A1 : DCSB_FMT_93;
        r.hdr_len = t1 ? 1 : 16;
    } else return 0;
    for (int k = 0; k < r.hdr_len; ++k) r.hdr[k] = data[2 + k];
    std::vector<uint32_t> bitpos(nf + 2);
    std::vector<uint2> bt(nf + 2);
    std::vector<uint16_t> hdrbits(nf + 2);
    int32_t status = 0;
    uint32_t nplay = 0, endbits = 0;
    uint8_t stopband = 0;
    DcsbScanOut so{ bitpos.data(), bt.data(), hdrbits.data(), &status, &nplay, &endbits, &stopband, nullptr, nullptr, nullptr, nullptr, nullptr };
    if (r.fmt == DCSB_FMT_94) {
        DcsbBandEnt ents[18];
        static const uint32_t zero_word[4] = { 0, 0, 0, 0 };
        dcsb_scan94_stream<false>(data, &r, 0, tab, tab->lut, (DcsbSA)tab->tx, dtab, (DcsbSA)0, (DcsbSA)ents, (DcsbSA)zero_word, so);
    } else dcsb_scan_stream(data, &r, 0, tab, tab->lut, so);
    if (status != 0 || nplay != nf) return 0;
    return 2u + r.hdr_len + (endbits + 7u) / 8u;
}
