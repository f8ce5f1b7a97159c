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

extern "C" int hostsim_decode_streams(const dcsb_stream_desc *descs, size_t n, int16_t *pcm_out,
                                      dcsb_result *results, uint32_t *bitpos_out, uint8_t *bt_out)
{
    DcsbPrepared p;
    int rc = dcsb_prepare(descs, n, &p);
    if (rc != DCSB_OK) return rc;
    std::vector<uint8_t> slab(p.slab_bytes + 64);
    dcsb_pack_slab(descs, n, &p, slab.data());
    static DcsbTables tab;
    dcsb_build_tables(&tab);

    std::vector<uint32_t> bitpos(p.total_frames_in + 1), nplay(n + 1), endbits(n + 1);
    std::vector<uint2> bt(p.total_frames_in + 1);
    std::vector<int32_t> status(n + 1);
    std::vector<uint8_t> stopband(n + 1);
    DcsbScanOut so{ bitpos.data(), bt.data(), status.data(), nplay.data(), endbits.data(), stopband.data() };
    for (size_t i = 0; i < n; ++i)                                   // K1 grid
        dcsb_scan_stream(slab.data(), p.recs.data(), (int)i, &tab, tab.lut, so);

    std::vector<unsigned long long> csum(n + 1, 0);
    std::vector<uint32_t> rows(DcsbWarpSmem<true>::WORDS);
    for (size_t t = 0; t < p.tiles.size(); ++t) {                    // K2 grid, one warp per tile
        if ((int)t < p.ntiles94)
            csum[p.tiles[t].stream] += dcsb_decode_tile<false>(slab.data(), p.recs.data(), p.tiles[t], &tab, tab.lut, so, pcm_out, rows.data());
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
    if (bitpos_out) memcpy(bitpos_out, bitpos.data(), p.total_frames_in * 4);
    if (bt_out)
        for (uint64_t f = 0; f < p.total_frames_in; ++f) {
            const uint64_t v = ((uint64_t)bt[f].y << 32) | bt[f].x;
            for (int k = 0; k < 16; ++k) bt_out[f * 16 + k] = (uint8_t)((v >> (4 * k)) & 15);
        }
    return DCSB_OK;
}
