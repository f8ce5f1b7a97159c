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
#include <algorithm>
#include "../../dcsexplorer_b200/csrc/dcsb_core.cuh"
#include "../../dcsexplorer_b200/csrc/dcsb_fast94.cuh"
#include "../../dcsexplorer_b200/csrc/dcsb_scan94.cuh"
#include "../../dcsexplorer_b200/csrc/dcsb_mix.cuh"

// slice_frames > 0: the time-sliced form dcsb_decode_streams uses for uniform chunks -- scan frames
// [k * slice, (k + 1) * slice) of every stream (resuming from the previous slice's end checkpoint),
// then decode the work items of that slice, slice after slice
static int sim_decode_streams(const dcsb_stream_desc *descs, size_t n, int16_t *pcm_out,
                              dcsb_result *results, uint32_t *bitpos_out, uint8_t *bt_out, uint32_t slice_frames)
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
    DcsbScanOut so{ bitpos.data(), bt.data(), hdrbits.data(), status.data(), nplay.data(), endbits.data(), stopband.data(), nullptr, nullptr, nullptr, nullptr, nullptr };
    std::vector<uint8_t> ring(DCSB_RING_BYTES, 0xA5);                // one K1 lane's shared-memory ring
    DcsbBandEnt ents[18];                                            // ... and its band entries
    static const uint32_t zero_word[4] = { 0, 0, 0, 0 };
    static uint16_t dtab[DCSB_DTAB_WORDS];                           // the CTA's descriptor table
    for (int i = 0; i < DCSB_DTAB_WORDS; ++i) dtab[i] = (uint16_t)dcsb_dtab_entry(tab.lut, i);
    std::vector<unsigned long long> csum(n + 1, 0);
    std::vector<uint32_t> rows(DcsbWarpSmem<true>::WORDS + DCSB_WARP94_WORDS);
    static DcsbTw94 tw;
    for (int i = 0; i < 64; ++i) dcsb_tw94_fill(&tw, &tab, i);
    uint32_t max_out = 0;
    for (size_t i = 0; i < n; ++i) max_out = std::max(max_out, p.recs[i].out_frames);
    for (uint32_t fa = 0; fa == 0 || fa < max_out; fa += slice_frames ? slice_frames : 0xFFFFFFFFu) {
        const uint32_t fb = slice_frames && fa + slice_frames < max_out ? fa + slice_frames : 0xFFFFFFFFu;
        for (size_t i = 0; i < n; ++i)                                   // K1 grid
            if (p.recs[i].fmt == DCSB_FMT_94) {
                // both staging variants of the scan (ring in shared memory / straight from global memory): alternate by stream
                if (i & 1) dcsb_scan94_stream<false>(slab.data(), p.recs.data(), (int)i, &tab, tab.lut, (DcsbSA)tab.tx, dtab, (DcsbSA)ring.data(), (DcsbSA)ents, (DcsbSA)zero_word, so, fa, fb);
                else dcsb_scan94_stream<true>(slab.data(), p.recs.data(), (int)i, &tab, tab.lut, (DcsbSA)tab.tx, dtab, (DcsbSA)ring.data(), (DcsbSA)ents, (DcsbSA)zero_word, so, fa, fb);
            }
            else dcsb_scan_stream(slab.data(), p.recs.data(), (int)i, &tab, tab.lut, so, fa, fb);
        std::vector<DcsbTile> t94, t93;
        if (slice_frames) dcsb_build_tiles(&p, fa, fb, &t94, &t93);
        else { t94.assign(p.tiles.begin(), p.tiles.begin() + p.ntiles94); t93.assign(p.tiles.begin() + p.ntiles94, p.tiles.end()); }
        for (const DcsbTile &tl : t94)                                   // K2 grid, one warp per item
            csum[tl.stream] += dcsb_decode94_item(slab.data(), p.recs.data(), tl, &tab, tab.lut, &tw, p.recs[tl.stream].hdr, so,
                                                  nplay[tl.stream], stopband[tl.stream], pcm_out, rows.data());
        for (const DcsbTile &tl : t93)
            csum[tl.stream] += dcsb_decode_tile<true>(slab.data(), p.recs.data(), tl, &tab, tab.lut, so, pcm_out, rows.data());
        if (fb == 0xFFFFFFFFu) break;
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

extern "C" int hostsim_decode_streams(const dcsb_stream_desc *descs, size_t n, int16_t *pcm_out,
                                      dcsb_result *results, uint32_t *bitpos_out, uint8_t *bt_out)
{
    return sim_decode_streams(descs, n, pcm_out, results, bitpos_out, bt_out, 0);
}
extern "C" int hostsim_decode_streams_sliced(const dcsb_stream_desc *descs, size_t n, int16_t *pcm_out,
                                             dcsb_result *results, uint32_t *bitpos_out, uint8_t *bt_out, uint32_t slice_frames)
{
    return sim_decode_streams(descs, n, pcm_out, results, bitpos_out, bt_out, slice_frames);
}

// ---- track playback: the host sequencer (product code, dcsb_rom.cpp) + K1 / K4 bodies on the CPU ----
extern "C" int hostsim_rom_render(const uint8_t *const *imgs, const size_t *sizes, const int *chipnos, int nchips,
                                  const dcsb_timeline *timelines, size_t n, int16_t *pcm_out, dcsb_timeline_result *results,
                                  dcsb_rom_info *info_out, uint8_t *host_bytes_out, size_t host_bytes_cap)
{
    dcsb_rom rom;
    for (int i = 0; i < nchips; ++i) rom.add(chipnos[i], imgs[i], sizes[i]);
    rom.check();
    if (info_out) {
        memset(info_out, 0, sizeof(*info_out));
        info_out->os_version = rom.os > 1 ? (uint16_t)rom.os : 0;
        info_out->hw_version = (uint8_t)rom.hw;
        info_out->n_channels = (uint8_t)rom.num_channels();
        info_out->n_tracks = rom.n_tracks;
        info_out->catalog_offset = rom.catalog_ofs;
        info_out->post_code = rom.post;
        info_out->version_number = rom.nominal_version;
    }
    if (rom.os <= 1) return DCSB_E_ARG;
    // stream table, as rom_prepare (dcsb_player.cu) builds it
    std::vector<uint32_t> addrs = rom.list_streams(true);
    {
        const std::vector<uint32_t> more = rom.list_streams(false);
        addrs.insert(addrs.end(), more.begin(), more.end());
        std::sort(addrs.begin(), addrs.end());
        addrs.erase(std::unique(addrs.begin(), addrs.end()), addrs.end());
    }
    for (int c = 0; c < 8; ++c) {
        rom.image_ofs[c] = (uint32_t)rom.image.size();
        if (!rom.chip[c].present) continue;
        rom.image.insert(rom.image.end(), rom.chip[c].bytes.begin(), rom.chip[c].bytes.begin() + rom.chip[c].size);
        rom.image.resize(rom.image.size() + 1024, 0);
    }
    if (rom.image.empty()) rom.image.resize(1024, 0);
    std::vector<dcsb_stream_desc> descs;
    for (uint32_t a : addrs) {
        const DcsbRomPtr p = rom.make_ptr(a);
        DcsbStreamFacts sf;
        sf.linear = a;
        sf.at = p;
        dcsb_stream_desc d;
        memset(&d, 0, sizeof(d));
        d.os_version = (uint16_t)rom.os;
        d.reserved = DCSB_STREAM_WRAP_EMPTY;
        if (rom.chip[p.chip].present && p.ofs < rom.chip[p.chip].size) {
            d.data = rom.image.data() + rom.image_ofs[p.chip] + p.ofs;
            d.nbytes = rom.chip[p.chip].size - p.ofs;
            sf.nframes = (uint16_t)rom.be(p, 2);
        } else { d.data = rom.image.data(); d.nbytes = 0; }
        rom.stream_by_addr[a & 0xFFFFFFu] = (uint32_t)rom.streams.size();
        rom.streams.push_back(sf);
        descs.push_back(d);
    }
    DcsbPrepared p;
    int rc = dcsb_prepare(descs.data(), descs.size(), &p, rom.image.data(), rom.image.size());
    if (rc != DCSB_OK) return rc;
    std::vector<uint8_t> slab(p.slab_bytes + 64, 0);
    memcpy(slab.data(), rom.image.data(), rom.image.size());
    static DcsbTables tab;
    dcsb_build_tables(&tab);
    const size_t ns = descs.size();
    std::vector<uint32_t> bitpos(p.total_checkpoints + 1), nplay(ns + 1), endbits(ns + 1);
    std::vector<uint2> bt(p.total_checkpoints + 1);
    std::vector<uint16_t> hdrbits(p.total_checkpoints + 1);
    std::vector<int32_t> status(ns + 1);
    std::vector<uint8_t> stopband(ns + 1);
    DcsbScanOut so{ bitpos.data(), bt.data(), hdrbits.data(), status.data(), nplay.data(), endbits.data(), stopband.data(), nullptr, nullptr, nullptr, nullptr, nullptr };
    std::vector<uint8_t> ring(DCSB_RING_BYTES, 0xA5);
    DcsbBandEnt ents[18];
    static const uint32_t zero_word[4] = { 0, 0, 0, 0 };
    static uint16_t dtab[DCSB_DTAB_WORDS];
    for (int i = 0; i < DCSB_DTAB_WORDS; ++i) dtab[i] = (uint16_t)dcsb_dtab_entry(tab.lut, i);
    for (size_t i = 0; i < ns; ++i) {
        if (p.recs[i].fmt == DCSB_FMT_94) dcsb_scan94_stream<true>(slab.data(), p.recs.data(), (int)i, &tab, tab.lut, (DcsbSA)tab.tx, dtab, (DcsbSA)ring.data(), (DcsbSA)ents, (DcsbSA)zero_word, so);
        else dcsb_scan_stream(slab.data(), p.recs.data(), (int)i, &tab, tab.lut, so);
        rom.streams[i].status = p.host_status[i] ? p.host_status[i] : status[i];
        rom.streams[i].nplay = p.host_status[i] ? 0 : nplay[i];
        if (p.host_status[i]) nplay[i] = 0;
    }
    rom.build_seq_streams();
    // schedules
    std::vector<DcsbSchedFrame> frames;
    std::vector<DcsbSchedEntry> entries;
    std::vector<uint32_t> first(n), count(n);
    size_t nhost = 0;
    for (size_t t = 0; t < n; ++t) {
        const dcsb_timeline &tl = timelines[t];
        DcsbSequencer seq(&rom);
        seq.soft_boot();
        seq.set_master_volume(tl.master_volume);
        first[t] = (uint32_t)frames.size();
        uint32_t w = 0;
        for (uint32_t f = 0; f < tl.n_frames; ++f) {
            while (w < tl.n_writes && tl.writes[w].frame <= f) seq.write_port(tl.writes[w++].byte);
            seq.frame(frames, entries);
        }
        count[t] = tl.n_frames;
        if (results) {
            results[t].status = seq.fatal ? DCSB_E_STOPPED : DCSB_OK;
            results[t].frames = tl.n_frames;
            results[t].n_host_bytes = (uint32_t)seq.host_bytes.size();
            results[t].checksum = 0;
        }
        for (uint8_t hb : seq.host_bytes) if (host_bytes_out && nhost < host_bytes_cap) host_bytes_out[nhost++] = hb;
    }
    // K4 grid
    const bool fam93 = rom.os == DCSB_OS93A || rom.os == DCSB_OS93B;
    DcsbMixSched sc{ frames.data(), entries.data() };
    static DcsbTw94 tw;
    for (int i = 0; i < 64; ++i) dcsb_tw94_fill(&tw, &tab, i);
    std::vector<uint32_t> rows(DcsbWarpSmem<true>::WORDS + DCSB_WARP94_WORDS);
    std::vector<uint8_t> hdrs(32 * 16);
    for (size_t t = 0; t < n; ++t) {
        const uint32_t item_len = fam93 ? DCSB_TILE_OUT : 63;          // (63: exercises the tile-to-tile tail carry)
        unsigned long long cs = 0;
        for (uint32_t f = 0; f < count[t]; f += item_len) {
            DcsbMixItem it{ first[t] + f, std::min<uint32_t>(item_len, count[t] - f), first[t], (uint32_t)t };
            if (fam93) cs += dcsb_mix93_tile(slab.data(), p.recs.data(), it, sc, &tab, tab.lut, so, pcm_out, rows.data());
            else cs += dcsb_mix94_item(slab.data(), p.recs.data(), it, sc, &tab, tab.lut, &tw, so, pcm_out, rows.data(), hdrs.data());
        }
        if (results) results[t].checksum = cs;
    }
    return DCSB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// The forward path on the CPU: the kernel bodies of dcsb_encode.cuh with loops playing the grids, in the order
// dcsb_encode_streams launches them (explicit stream types; the wildcard is host logic of the product's entry point).
#include "../../dcsexplorer_b200/csrc/dcsb_encode.cuh"
extern "C" int hostsim_encode_streams(const float *const *pcm, const uint64_t *n_samples, size_t n, const dcsb_encode_params *params,
                                      uint8_t *out, uint64_t out_capacity, uint64_t *out_offsets, float *frames_out)
{
    enc_build_tables(&g_enc_host);
    std::vector<EncStream> hs(n);
    std::vector<float> all;
    uint64_t total_frames = 0;
    for (size_t i = 0; i < n; ++i) {
        const dcsb_encode_params &p = params[i];
        const bool f93 = p.format_version == DCSB_OS93A || p.format_version == DCSB_OS93B;
        if (n_samples[i] == 0 || (p.stream_type != 0 && p.stream_type != 1)) return DCSB_E_ARG;
        EncStream &s = hs[i];
        memset(&s, 0, sizeof(s));
        s.pcm_off = all.size();
        s.n_samples = n_samples[i];
        s.frame0 = (uint32_t)total_frames;
        s.n_frames = (uint32_t)((n_samples[i] + 239) / 240);
        s.type = p.stream_type;
        s.subtype = p.stream_subtype;
        s.max_err2 = p.max_quantization_error * p.max_quantization_error;
        s.min_range = p.min_dynamic_range;
        s.fmt93 = f93 ? 1 : 0;
        all.insert(all.end(), pcm[i], pcm[i] + n_samples[i]);
        total_frames += s.n_frames;
    }
    const uint32_t nfr = (uint32_t)total_frames;
    std::vector<uint32_t> frame_stream(nfr), frame_bits(nfr, 0);
    for (size_t i = 0; i < n; ++i)
        for (uint32_t k = 0; k < hs[i].n_frames; ++k) frame_stream[hs[i].frame0 + k] = (uint32_t)i;
    std::vector<float> f((size_t)nfr * 256), power((size_t)nfr * 16), lo((size_t)nfr * 16), hi((size_t)nfr * 16), stats(n * 48);
    for (uint32_t t = 0; t < nfr; ++t) dcsb_enc_transform_body(t, all.data(), hs.data(), frame_stream.data(), nfr, f.data(), power.data(), lo.data(), hi.data());
    for (uint32_t t = 0; t < n * 16; ++t) dcsb_enc_stats_body(t, hs.data(), (int)n, power.data(), lo.data(), hi.data(), stats.data());
    if (frames_out) memcpy(frames_out, f.data(), f.size() * sizeof(float));
    for (size_t i = 0; i < n; ++i) enc_stream_header(&stats[i * 48], params[i], g_enc_host, &hs[i]);
    std::vector<uint8_t> best((size_t)nfr * 16 * ENC_NV * 2, 0), codes((size_t)nfr * 16, 0), padj((size_t)nfr * 4, 0);
    std::vector<uint16_t> dec((size_t)nfr * 16, 0);
    std::vector<uint64_t> frame_pos(nfr, 0), sbits(n, 0), word0(n + 1, 0);
    for (uint32_t t = 0; t < nfr * 16u; ++t) dcsb_enc_search_body(t, hs.data(), frame_stream.data(), nfr, f.data(), lo.data(), hi.data(), best.data());
    for (uint32_t t = 0; t < n; ++t) dcsb_enc_resolve_body(t, hs.data(), (int)n, best.data(), codes.data(), padj.data());
    for (uint32_t t = 0; t < nfr; ++t) dcsb_enc_emit_body<false>(t, hs.data(), frame_stream.data(), nfr, f.data(), codes.data(), padj.data(), frame_bits.data(), nullptr, nullptr, nullptr);
    for (uint32_t t = 0; t < nfr * 16u; ++t) dcsb_enc_search93_body(t, hs.data(), frame_stream.data(), nfr, f.data(), best.data());
    for (uint32_t t = 0; t < n; ++t) dcsb_enc_resolve93_body(t, hs.data(), (int)n, f.data(), best.data(), dec.data());
    for (uint32_t t = 0; t < nfr; ++t) dcsb_enc_frame93_body<false>(t, hs.data(), frame_stream.data(), nfr, f.data(), dec.data(), frame_bits.data(), nullptr, nullptr, nullptr);
    for (uint32_t t = 0; t < n; ++t) dcsb_enc_scan_body(t, hs.data(), (int)n, frame_bits.data(), frame_pos.data(), sbits.data());
    uint64_t need = 0;
    for (size_t i = 0; i < n; ++i) { word0[i + 1] = word0[i] + (sbits[i] + 31) / 32 + 1; need += 18 + (sbits[i] + 7) / 8; }
    if (need > out_capacity) return DCSB_E_NOMEM;
    std::vector<uint32_t> words(word0[n], 0);
    for (uint32_t t = 0; t < nfr; ++t) dcsb_enc_emit_body<true>(t, hs.data(), frame_stream.data(), nfr, f.data(), codes.data(), padj.data(), nullptr, frame_pos.data(), words.data(), word0.data());
    for (uint32_t t = 0; t < nfr; ++t) dcsb_enc_frame93_body<true>(t, hs.data(), frame_stream.data(), nfr, f.data(), dec.data(), nullptr, frame_pos.data(), words.data(), word0.data());
    uint64_t o = 0;
    for (size_t i = 0; i < n; ++i) {
        out_offsets[i] = o;
        out[o++] = (uint8_t)(hs[i].n_frames >> 8);
        out[o++] = (uint8_t)(hs[i].n_frames & 0xFF);
        memcpy(out + o, hs[i].hdr, 16);
        o += 16;
        const uint64_t nb = (sbits[i] + 7) / 8;
        memcpy(out + o, reinterpret_cast<const uint8_t *>(words.data() + word0[i]), nb);
        o += nb;
    }
    out_offsets[n] = o;
    return DCSB_OK;
}
