// TEST INFRASTRUCTURE ONLY -- CPU-side simulator of the dcsb200 kernels.
//
// Compiles the SAME kernel bodies the GPU runs (dcsexplorer_b200/csrc/dcsb_core.cuh:
// dcsb_scan_stream = one K1 thread, dcsb_decode_tile<T93> = one K2 warp) as plain C++,
// with the warp played by a loop over lanes, so index/arith bugs are found by the CPU
// test-suite (pytest -m "not gpu") before any GPU time is spent.  It is NOT a product
// path: nothing in dcsexplorer_b200/ loads this library, and the product has no CPU
// fallback (dcsb_create fails without a CUDA device).
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../../dcsexplorer_b200/csrc/dcsb_core.cuh"
#include "../../dcsexplorer_b200/csrc/dcsb_fast94.cuh"
#include "../../dcsexplorer_b200/csrc/dcsb_scan94.cuh"

extern "C" int hostsim_decode_streams(const dcsb_stream_desc *descs, size_t n, int16_t *pcm_out,
                                      dcsb_result *results, uint32_t *bitpos_out, uint8_t *bt_out)
{
    DcsbPrepared p;
    int rc = dcsb_prepare(descs, n, &p, nullptr, 0);
    if (rc != DCSB_OK) return rc;
    std::vector<uint8_t> slab(p.slab_bytes + 64);
    dcsb_pack_slab(descs, n, &p, slab.data());
    static DcsbTables tab;
    dcsb_build_tables(&tab);

    std::vector<uint32_t> bitpos(p.total_checkpoints + 1), nplay(n + 1), endbits(n + 1);
    std::vector<uint2> bt(p.total_checkpoints + 1);
    std::vector<uint16_t> hdrbits(p.total_checkpoints + 1);
    std::vector<int32_t> status(n + 1);
    std::vector<uint8_t> stopband(n + 1);
    DcsbScanOut so{ bitpos.data(), bt.data(), hdrbits.data(), status.data(), nplay.data(), endbits.data(), stopband.data(), nullptr };
    std::vector<uint8_t> ring(DCSB_RING_BYTES, 0xA5);                // one K1 thread's shared-memory ring
    for (size_t i = 0; i < n; ++i)                                   // K1 grid
        if (p.recs[i].fmt == DCSB_FMT_94) dcsb_scan94_stream(slab.data(), p.recs.data(), (int)i, &tab, tab.lut, tab.t8, tab.t1, ring.data(), so);
        else dcsb_scan_stream(slab.data(), p.recs.data(), (int)i, &tab, tab.lut, so);

    std::vector<unsigned long long> csum(n + 1, 0);
    std::vector<uint32_t> rows(DcsbWarpSmem<true>::WORDS + DCSB_WARP94_WORDS);
    static DcsbTw94 tw;
    memcpy(tw.tw_c2, tab.tw_c2, sizeof(tw.tw_c2)); memcpy(tw.tw_s2, tab.tw_s2, sizeof(tw.tw_s2));
    memcpy(tw.pre_c0, tab.pre_c0, sizeof(tw.pre_c0)); memcpy(tw.pre_c1, tab.pre_c1, sizeof(tw.pre_c1));
    for (size_t t = 0; t < p.tiles.size(); ++t) {                    // K2 grid, one warp per tile
        if ((int)t < p.ntiles94)
            csum[p.tiles[t].stream] += dcsb_decode94_item(slab.data(), p.recs.data(), p.tiles[t], &tab, tab.lut, &tw,
                                                          p.recs[p.tiles[t].stream].hdr, so, pcm_out, rows.data());
        else
            csum[p.tiles[t].stream] += dcsb_decode_tile<true>(slab.data(), p.recs.data(), p.tiles[t], &tab, tab.lut, so, pcm_out, rows.data());
    }
    for (size_t i = 0; i < n; ++i) {
        if (results) {
            results[i].status = p.host_status[i] ? p.host_status[i] : status[i];
            results[i].frames = p.recs[i].out_frames;
            results[i].frames_decoded = nplay[i];
            results[i].stream_bytes = p.host_status[i] ? 0 : 2 + p.recs[i].hdr_len + (endbits[i] + 7) / 8;
            results[i].checksum = csum[i];
        }
    }
    uint64_t o = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t fb = p.recs[i].frame_base;
        for (uint32_t f = 0; f < p.recs[i].nframes; ++f, ++o) {
            if (bitpos_out) bitpos_out[o] = bitpos[fb + f];
            if (bt_out) {
                const uint64_t v = ((uint64_t)bt[fb + f].y << 32) | bt[fb + f].x;
                for (int k = 0; k < 16; ++k) bt_out[o * 16 + k] = (uint8_t)((v >> (4 * k)) & 15);
            }
        }
    }
    return DCSB_OK;
}
