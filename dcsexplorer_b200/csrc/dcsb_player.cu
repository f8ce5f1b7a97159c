// dcsb200 C-ABI, part 2: ROM sets, players and timelines (include/dcsb200.h).  The host runs the
// reference's control plane (dcsb_rom.cpp); this file uploads the ROM's streams once, scans them
// on the GPU, turns sequencer output into mix work items and launches K4 (dcsb_mix.cuh).
#include <string.h>
#include <algorithm>
#include <chrono>
#include <functional>
#include <thread>
#include "dcsb_ctx.h"
#include "dcsb_rom.h"
#include "dcsb_mix.cuh"

// ======================================================================================
// ROM object
extern "C" int dcsb_rom_create(dcsb_rom **out)
{
    if (!out) return DCSB_E_ARG;
    *out = new dcsb_rom();
    return DCSB_OK;
}

static void rom_drop_batch(dcsb_rom *rom)
{
    if (rom->batch) { dcsb_batch_destroy(rom->batch); rom->batch = nullptr; }
    if (rom->d_seq_streams) { cudaFree(rom->d_seq_streams); rom->d_seq_streams = nullptr; }
    rom->streams.clear();
    rom->seq_streams.clear();
    rom->stream_by_addr.clear();
}

extern "C" void dcsb_rom_destroy(dcsb_rom *rom)
{
    if (!rom) return;
    rom_drop_batch(rom);
    delete rom;
}

extern "C" int dcsb_rom_add(dcsb_rom *rom, int chip_number, const uint8_t *image, size_t nbytes)
{
    if (!rom || !image || chip_number < 2 || chip_number > 9 || nbytes == 0 || (nbytes & (nbytes - 1))) return DCSB_E_ARG;
    rom_drop_batch(rom);
    rom->add(chip_number, image, nbytes);
    rom->hw = DCSB_HW_UNKNOWN;          // versions are re-detected by the next check
    return DCSB_OK;
}

extern "C" int dcsb_rom_load_zip(dcsb_rom *rom, const char *zip_path, const char *explicit_u2)
{
    if (!rom || !zip_path) return DCSB_ZIP_E_OPEN;
    rom_drop_batch(rom);
    const int rc = dcsb_rom_load_zip_impl(rom, zip_path, explicit_u2);
    rom->hw = DCSB_HW_UNKNOWN;
    return rc;
}

extern "C" int dcsb_rom_check(dcsb_rom *rom) { return rom ? rom->check() : 2; }
extern "C" const char *dcsb_rom_last_error(const dcsb_rom *rom) { return rom ? rom->err.c_str() : "no rom"; }

extern "C" size_t dcsb_rom_zip_files(const dcsb_rom *rom, dcsb_zip_file *files, size_t max)
{
    if (!rom) return 0;
    for (size_t i = 0; i < rom->zip_files.size() && i < max && files; ++i) {
        files[i].name = rom->zip_files[i].name.c_str();
        files[i].data = rom->zip_files[i].data.data();
        files[i].size = rom->zip_files[i].data.size();
        files[i].chip_number = i < rom->zip_chip.size() ? rom->zip_chip[i] : -1;
    }
    return rom->zip_files.size();
}

extern "C" int dcsb_rom_get_info(const dcsb_rom *rom, dcsb_rom_info *info)
{
    if (!rom || !info) return DCSB_E_ARG;
    memset(info, 0, sizeof(*info));
    info->os_version = rom->os > 1 ? (uint16_t)rom->os : 0;
    info->hw_version = (uint8_t)rom->hw;
    info->n_channels = (uint8_t)rom->num_channels();
    // GetVersionNumber (DCSDecoder.cpp:506-512)
    info->version_number = rom->nominal_version ? rom->nominal_version
                           : (rom->os == DCSB_OS93A || rom->os == DCSB_OS93B) ? 0x0100 : rom->os == DCSB_OS94 ? 0x0101 : 0;
    info->n_tracks = rom->n_tracks;
    info->catalog_offset = rom->catalog_ofs;
    info->post_code = rom->post;
    snprintf(info->signature, sizeof(info->signature), "%s", rom->signature.c_str());
    return DCSB_OK;
}

extern "C" size_t dcsb_rom_decompile_track(const dcsb_rom *rom, uint16_t track, dcsb_opcode *steps, size_t max)
{
    if (!rom) return 0;
    const std::vector<dcsb_opcode> v = rom->decompile_track(track);
    for (size_t i = 0; i < v.size() && i < max && steps; ++i) steps[i] = v[i];
    return v.size();
}

extern "C" int dcsb_rom_track_info(const dcsb_rom *rom, uint16_t track, dcsb_track_info *info)
{
    if (!rom || !info) return 0;
    return rom->track_info(track, info) ? 1 : 0;
}

extern "C" size_t dcsb_rom_list_streams(const dcsb_rom *rom, uint32_t *addresses, size_t max)
{
    if (!rom) return 0;
    const std::vector<uint32_t> l = rom->list_streams();
    for (size_t i = 0; i < l.size() && i < max && addresses; ++i) addresses[i] = l[i];
    return l.size();
}

extern "C" const uint8_t *dcsb_rom_pointer(const dcsb_rom *rom, uint32_t linear_address, uint32_t *bytes_left)
{
    if (bytes_left) *bytes_left = 0;
    if (!rom) return nullptr;
    const DcsbRomPtr p = rom->make_ptr(linear_address);
    const dcsb_rom::Chip &c = rom->chip[p.chip];
    if (!c.present || p.ofs >= c.size) return nullptr;
    if (bytes_left) *bytes_left = c.size - p.ofs;
    return c.bytes.data() + p.ofs;
}

// ======================================================================================
// stream table: every stream of the ROM resident in HBM + scanned once
static int rom_prepare(dcsb_ctx *ctx, dcsb_rom *rom)
{
    if (rom->batch && rom->batch->ctx == ctx) return DCSB_OK;
    rom_drop_batch(rom);
    if (rom->hw == DCSB_HW_UNKNOWN) rom->check();          // SoftBoot detects the versions if nobody has (DCSDecoder.cpp:1533-1534)
    if (rom->os <= 1) return fail(ctx, DCSB_E_ARG, "ROM set not recognised (no catalog / checksum match in U2)");
    std::vector<uint32_t> addrs = rom->list_streams(true);
    {   // (plus whatever only the reference's own listing finds: dcsb_player_load_audio_stream accepts those too)
        const std::vector<uint32_t> more = rom->list_streams(false);
        addrs.insert(addrs.end(), more.begin(), more.end());
        std::sort(addrs.begin(), addrs.end());
        addrs.erase(std::unique(addrs.begin(), addrs.end()), addrs.end());
    }
    // all chips back to back, 1 KB of zeros behind each
    rom->image.clear();
    for (int c = 0; c < 8; ++c) {
        rom->image_ofs[c] = (uint32_t)rom->image.size();
        if (!rom->chip[c].present) continue;
        rom->image.insert(rom->image.end(), rom->chip[c].bytes.begin(), rom->chip[c].bytes.begin() + rom->chip[c].size);
        rom->image.resize(rom->image.size() + 1024, 0);
    }
    if (rom->image.empty()) rom->image.resize(1024, 0);
    std::vector<dcsb_stream_desc> descs;
    for (uint32_t a : addrs) {
        const DcsbRomPtr p = rom->make_ptr(a);
        DcsbStreamFacts sf;
        sf.linear = a;
        sf.at = p;
        dcsb_stream_desc d;
        memset(&d, 0, sizeof(d));
        d.os_version = (uint16_t)rom->os;
        d.master_volume = 255;
        d.mixing_level = 0x64;
        d.reserved = DCSB_STREAM_WRAP_EMPTY;
        if (rom->chip[p.chip].present && p.ofs < rom->chip[p.chip].size) {
            d.data = rom->image.data() + rom->image_ofs[p.chip] + p.ofs;
            d.nbytes = rom->chip[p.chip].size - p.ofs;
            sf.nframes = (uint16_t)rom->be(p, 2);
        } else {
            d.data = rom->image.data();         // stream in a missing chip: rejected as too short
            d.nbytes = 0;
        }
        rom->stream_by_addr[a & 0xFFFFFFu] = (uint32_t)rom->streams.size();
        rom->streams.push_back(sf);
        descs.push_back(d);
    }
    int rc = dcsb_batch_create_impl(ctx, descs.data(), descs.size(), rom->image.data(), rom->image.size(), &rom->batch);
    if (rc != DCSB_OK) return rc;
    dcsb_batch *b = rom->batch;
    if (b->n) {
        CK(dcsb_launch_scan(b->d_slab, b->d_recs, b->d_order, (int)b->n, (int)b->n94, 0, ctx->d_tables, b->scan, nullptr), "scan kernel launch");
        CK(cudaDeviceSynchronize(), "ROM stream scan");
        std::vector<int32_t> st(b->n);
        std::vector<uint32_t> np(b->n), eb(b->n);
        CK(cudaMemcpy(st.data(), b->scan.status, b->n * 4, cudaMemcpyDeviceToHost), "D2H status");
        CK(cudaMemcpy(np.data(), b->scan.nplay, b->n * 4, cudaMemcpyDeviceToHost), "D2H nplay");
        CK(cudaMemcpy(eb.data(), b->scan.endbits, b->n * 4, cudaMemcpyDeviceToHost), "D2H endbits");
        for (size_t i = 0; i < b->n; ++i) {
            rom->streams[i].status = b->host_status[i] ? b->host_status[i] : st[i];
            rom->streams[i].nplay = b->host_status[i] ? 0 : np[i];
            rom->streams[i].nbytes = b->host_status[i] ? 0 : 2 + b->recs[i].hdr_len + (eb[i] + 7) / 8;
        }
    }
    rom->build_seq_streams();
    if (!rom->seq_streams.empty()) {
        CK(cudaMalloc(&rom->d_seq_streams, rom->seq_streams.size() * sizeof(DcsbSeqStream)), "cudaMalloc(stream table)");
        CK(cudaMemcpy(rom->d_seq_streams, rom->seq_streams.data(), rom->seq_streams.size() * sizeof(DcsbSeqStream), cudaMemcpyHostToDevice), "H2D stream table");
    }
    return DCSB_OK;
}

// ======================================================================================
// rendering a schedule
struct DcsbRenderBufs {
    DcsbBuf d_frames, d_entries, d_items, d_pcm, d_csum, h_pcm;
    void release()
    {
        for (DcsbBuf *b : { &d_frames, &d_entries, &d_items, &d_pcm, &d_csum }) b->release(false);
        h_pcm.release(true);
    }
};

// frames / entries: the whole schedule; tl_first[t], tl_frames[t]: each timeline's slice; skip[t]:
// leading frames that are only there to warm the overlap up (not rendered).  PCM lands in
// bufs.d_pcm at 240 * (global frame index) samples.
// work items: 1994 layout 31 + 32k frames per warp, 1993 layouts tiles of 31
static uint32_t mix_item_len(bool fam93, uint64_t total_frames)
{
    uint32_t item_len = DCSB_TILE_OUT;
    if (!fam93) {
        item_len = (uint32_t)std::min<uint64_t>(255, std::max<uint64_t>(31, total_frames / 8192));
        item_len = item_len < 63 ? 31 : 31 + ((item_len - 31) / 32) * 32;
    }
    return item_len;
}

static int render_schedule(dcsb_ctx *ctx, dcsb_rom *rom, DcsbRenderBufs &bufs, const std::vector<DcsbSchedFrame> &frames,
                           const std::vector<DcsbSchedEntry> &entries, const std::vector<uint32_t> &tl_first,
                           const std::vector<uint32_t> &tl_frames, const std::vector<uint32_t> &skip, cudaStream_t st)
{
    dcsb_batch *b = rom->batch;
    const bool fam93 = rom->os == DCSB_OS93A || rom->os == DCSB_OS93B;
    const size_t nt = tl_first.size();
    uint64_t total = 0;
    for (size_t t = 0; t < nt; ++t) total += tl_frames[t];
    const uint32_t item_len = mix_item_len(fam93, total);
    std::vector<DcsbMixItem> items;
    for (size_t t = 0; t < nt; ++t)
        for (uint32_t f = skip[t]; f < tl_frames[t]; f += item_len)
            items.push_back(DcsbMixItem{ tl_first[t] + f, std::min<uint32_t>(item_len, tl_frames[t] - f), tl_first[t], (uint32_t)t });
    if (items.empty()) return DCSB_OK;
#define ENS(buf, bytes, what) do { cudaError_t e_ = (buf).ensure((bytes), false); if (e_ != cudaSuccess) return fail(ctx, DCSB_E_NOMEM, what, e_); } while (0)
    ENS(bufs.d_frames, frames.size() * sizeof(DcsbSchedFrame), "cudaMalloc(schedule frames)");
    ENS(bufs.d_entries, std::max<size_t>(1, entries.size()) * sizeof(DcsbSchedEntry), "cudaMalloc(schedule entries)");
    ENS(bufs.d_items, items.size() * sizeof(DcsbMixItem), "cudaMalloc(mix items)");
    ENS(bufs.d_pcm, frames.size() * 480, "cudaMalloc(pcm)");
    ENS(bufs.d_csum, nt * 8, "cudaMalloc(checksums)");
#undef ENS
    CK(cudaMemcpyAsync(bufs.d_frames.p, frames.data(), frames.size() * sizeof(DcsbSchedFrame), cudaMemcpyHostToDevice, st), "H2D frames");
    if (!entries.empty())
        CK(cudaMemcpyAsync(bufs.d_entries.p, entries.data(), entries.size() * sizeof(DcsbSchedEntry), cudaMemcpyHostToDevice, st), "H2D entries");
    CK(cudaMemcpyAsync(bufs.d_items.p, items.data(), items.size() * sizeof(DcsbMixItem), cudaMemcpyHostToDevice, st), "H2D items");
    CK(cudaMemsetAsync(bufs.d_csum.p, 0, nt * 8, st), "memset checksums");
    CK(dcsb_launch_mix(fam93, b->d_slab, b->d_recs, bufs.d_items.p, (int)items.size(), bufs.d_frames.p, bufs.d_entries.p,
                       ctx->d_tables, b->scan, (int16_t *)bufs.d_pcm.p, (unsigned long long *)bufs.d_csum.p, st), "mix kernel launch");
    return DCSB_OK;
}

// ======================================================================================
// player
struct dcsb_player {
    dcsb_ctx *ctx;
    dcsb_rom *rom;
    DcsbSequencer seq;
    DcsbRenderBufs bufs;
    std::vector<DcsbSchedEntry> prev_entries;       // the last rendered frame: warms the overlap of the next chunk up
    DcsbSchedFrame prev_frame{ 0, 0, 8, 0, 0 };
    bool have_prev = false;
    size_t host_read = 0;
    // look-ahead (dcsb_player_set_lookahead): frames are rendered `lookahead` at a time and handed out from
    // `ahead`; `snap*` is the decoder state at the start of that block, so that an input arriving in the middle
    // takes effect at the frame it arrives at (the block's unconsumed frames are dropped: state restored, the
    // consumed frames replayed on the host sequencer -- about 200 ns each --, the rest rendered anew)
    uint32_t lookahead = 0;
    std::vector<int16_t> ahead;
    std::vector<uint8_t> ahead_playing;             // per pre-rendered frame: channels with a stream playing after it
    uint32_t ahead_n = 0, ahead_pos = 0;
    DcsbSequencer snap;
    std::vector<DcsbSchedEntry> snap_prev_entries;
    DcsbSchedFrame snap_prev_frame{ 0, 0, 8, 0, 0 };
    bool snap_have_prev = false;
    uint8_t snap_playing = 0;
    uint64_t served = 0;                            // frames handed out so far
    size_t host_visible = 0;                        // host bytes of frames handed out (and of inputs taken) so far
    dcsb_player(dcsb_ctx *c, dcsb_rom *r) : ctx(c), rom(r), seq(r), snap(r) {}
};

static uint8_t playing_mask(const DcsbSequencer &q)
{
    uint8_t m = 0;
    for (int ch = 0; ch < DCSB_MAX_CHANNELS; ++ch) m |= (uint8_t)((q.stream_playing(ch) ? 1 : 0) << ch);
    return m;
}

// bring the sequencer back to the frame the caller has consumed up to (drops what was rendered ahead)
static void player_sync(dcsb_player *p)
{
    if (p->ahead_pos < p->ahead_n) {
        p->seq = p->snap;
        p->prev_entries = p->snap_prev_entries;
        p->prev_frame = p->snap_prev_frame;
        p->have_prev = p->snap_have_prev;
        std::vector<DcsbSchedFrame> frames;
        std::vector<DcsbSchedEntry> entries;
        for (uint32_t f = 0; f < p->ahead_pos; ++f) p->seq.frame(frames, entries);
        if (!frames.empty()) {
            const DcsbSchedFrame last = frames.back();
            p->prev_entries.assign(entries.begin() + last.first_entry, entries.begin() + last.first_entry + last.n_entries);
            p->prev_frame = last;
            p->have_prev = true;
        }
    }
    p->ahead_n = p->ahead_pos = 0;
    p->host_visible = p->seq.host_bytes.size();     // the sequencer stands exactly where the caller is
}

extern "C" int dcsb_player_create(dcsb_ctx *ctx, dcsb_rom *rom, dcsb_player **out)
{
    if (!ctx || !rom || !out) return DCSB_E_ARG;
    *out = nullptr;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const int rc = rom_prepare(ctx, rom);
    if (rc != DCSB_OK) return rc;
    dcsb_player *p = new dcsb_player(ctx, rom);
    p->seq.soft_boot();
    *out = p;
    return DCSB_OK;
}

extern "C" void dcsb_player_destroy(dcsb_player *p)
{
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    p->bufs.release();
    delete p;
}

extern "C" int dcsb_player_set_lookahead(dcsb_player *p, uint32_t n_frames)
{
    if (!p || n_frames > 4096) return DCSB_E_ARG;
    player_sync(p);
    p->lookahead = n_frames;
    return DCSB_OK;
}

extern "C" void dcsb_player_set_master_volume(dcsb_player *p, int vol) { if (p) { player_sync(p); p->seq.set_master_volume(vol); p->host_visible = p->seq.host_bytes.size(); } }
extern "C" void dcsb_player_write_data_port(dcsb_player *p, uint8_t byte) { if (p) { player_sync(p); p->seq.write_port(byte); p->host_visible = p->seq.host_bytes.size(); } }
extern "C" void dcsb_player_add_track_command(dcsb_player *p, uint16_t track) { if (p) { player_sync(p); p->seq.add_track_command(track); } }
extern "C" void dcsb_player_clear_tracks(dcsb_player *p) { if (p) { player_sync(p); p->seq.clear_tracks(); } }
extern "C" int dcsb_player_is_stream_playing(const dcsb_player *p, int channel)
{
    if (!p || channel < 0 || channel >= DCSB_MAX_CHANNELS) return 0;
    if (p->ahead_pos < p->ahead_n)          // the sequencer has run ahead: answer for the frame the caller is at
        return ((p->ahead_pos ? p->ahead_playing[p->ahead_pos - 1] : p->snap_playing) >> channel) & 1;
    return p->seq.stream_playing(channel) ? 1 : 0;
}

extern "C" int dcsb_player_load_audio_stream(dcsb_player *p, int channel, uint32_t stream_address, int mixing_level)
{
    if (!p || channel < 0 || channel >= DCSB_MAX_CHANNELS) return DCSB_E_ARG;
    if (p->rom->stream_by_addr.find(stream_address & 0xFFFFFFu) == p->rom->stream_by_addr.end())
        return fail(p->ctx, DCSB_E_ARG, "dcsb_player_load_audio_stream: not a stream any track of this ROM plays");
    player_sync(p);
    p->seq.load_stream(channel, stream_address, mixing_level);
    return DCSB_OK;
}

extern "C" int dcsb_player_stream_info(const dcsb_player *p, uint32_t stream_address, dcsb_stream_info *info)
{
    if (!p || !info) return DCSB_E_ARG;
    memset(info, 0, sizeof(*info));
    auto it = p->rom->stream_by_addr.find(stream_address & 0xFFFFFFu);
    if (it == p->rom->stream_by_addr.end()) return DCSB_E_ARG;
    const DcsbStreamFacts &sf = p->rom->streams[it->second];
    const DcsbStreamRec &r = p->rom->batch->recs[it->second];
    info->n_frames = sf.nframes;
    info->n_bytes = (int32_t)sf.nbytes;
    info->status = sf.status;
    memcpy(info->header, r.hdr, r.hdr_len);
    info->stream_type = r.hdr[0] >> 7;
    if (r.fmt == DCSB_FMT_94) info->stream_subtype = ((r.hdr[1] & 0x80) >> 6) | ((r.hdr[1] & 0x80) >> 7);   // as the reference computes it (:1516)
    return DCSB_OK;
}

extern "C" size_t dcsb_player_host_bytes(dcsb_player *p, uint8_t *out, size_t max)
{
    if (!p) return 0;
    const size_t visible = p->lookahead ? p->host_visible : p->seq.host_bytes.size();
    size_t n = 0;
    while (p->host_read < visible && n < max) out[n++] = p->seq.host_bytes[p->host_read++];
    return n;
}

// n_frames main-loop passes: schedule on the host sequencer, one mix launch, PCM to host memory
static int player_render_block(dcsb_player *p, uint32_t n_frames, int16_t *pcm_out, uint8_t *playing)
{
    dcsb_ctx *ctx = p->ctx;
    std::vector<DcsbSchedFrame> frames;
    std::vector<DcsbSchedEntry> entries;
    const uint32_t lead = p->have_prev ? 1u : 0u;
    if (lead) {
        entries = p->prev_entries;
        frames.push_back(DcsbSchedFrame{ 0, (uint8_t)entries.size(), p->prev_frame.vs, p->prev_frame.flags, 0 });
    }
    for (uint32_t f = 0; f < n_frames; ++f) {
        p->seq.frame(frames, entries);
        if (playing) playing[f] = playing_mask(p->seq);
    }
    const DcsbSchedFrame last = frames.back();
    p->prev_entries.assign(entries.begin() + last.first_entry, entries.begin() + last.first_entry + last.n_entries);
    p->prev_frame = last;
    p->have_prev = true;
    std::vector<uint32_t> first{ 0 }, count{ n_frames + lead }, skip{ lead };
    int rc = render_schedule(ctx, p->rom, p->bufs, frames, entries, first, count, skip, nullptr);
    if (rc != DCSB_OK) return rc;
    CK(cudaMemcpy(pcm_out, (const int16_t *)p->bufs.d_pcm.p + (size_t)lead * 240, (size_t)n_frames * 480, cudaMemcpyDeviceToHost), "D2H pcm");
    return DCSB_OK;
}

extern "C" int dcsb_player_render(dcsb_player *p, uint32_t n_frames, int16_t *pcm_out)
{
    if (!p || (!pcm_out && n_frames)) return DCSB_E_ARG;
    dcsb_ctx *ctx = p->ctx;
    if (n_frames == 0) return DCSB_OK;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    if (p->lookahead == 0 || n_frames >= p->lookahead) {
        // nothing to gain from rendering ahead: exactly what is asked for, straight into the caller's buffer
        player_sync(p);
        const int rc = player_render_block(p, n_frames, pcm_out, nullptr);
        if (rc == DCSB_OK) { p->served += n_frames; p->host_visible = p->seq.host_bytes.size(); }
        return rc;
    }
    while (n_frames) {
        if (p->ahead_pos == p->ahead_n) {
            p->snap = p->seq;
            p->snap_prev_entries = p->prev_entries;
            p->snap_prev_frame = p->prev_frame;
            p->snap_have_prev = p->have_prev;
            p->snap_playing = playing_mask(p->seq);
            p->ahead.resize((size_t)p->lookahead * 240);
            p->ahead_playing.resize(p->lookahead);
            p->ahead_n = p->ahead_pos = 0;
            const int rc = player_render_block(p, p->lookahead, p->ahead.data(), p->ahead_playing.data());
            if (rc != DCSB_OK) return rc;
            p->ahead_n = p->lookahead;
        }
        const uint32_t k = std::min(n_frames, p->ahead_n - p->ahead_pos);
        memcpy(pcm_out, p->ahead.data() + (size_t)p->ahead_pos * 240, (size_t)k * 480);
        p->ahead_pos += k;
        p->served += k;
        pcm_out += (size_t)k * 240;
        n_frames -= k;
    }
    // host bytes of the frames handed out so far (a frame's bytes carry its number, DcsbSequencer::to_host)
    const uint32_t upto = p->snap.frame_no + p->ahead_pos;
    size_t v = p->host_visible;
    while (v < p->seq.host_bytes.size() && p->seq.host_byte_frames[v] < upto) ++v;
    p->host_visible = v;
    return DCSB_OK;
}

// ======================================================================================
// many timelines at once
// The call is a pipeline over chunks of timelines: host threads run the sequencers of chunk k + 1
// (the reference's MainLoop control plane, one instance per timeline) while the GPU renders chunk k
// and its PCM goes down the link.  Per chunk: schedules -> pinned staging (every thread copies its
// own timelines into place) -> one upload -> K4 mix kernel -> one PCM download (asynchronous into
// a page-locked pcm_out, else a blocking copy).  Device and staging buffers live in the context
// and are reused by the next call (no allocation in the steady state).
static bool is_pinned_host(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
#define DCSB_RT_MAX_CHUNKS 8
struct DcsbTimelineCache {
    cudaStream_t st = nullptr, copy = nullptr;
    DcsbBuf d_frames, d_pcm, d_csum, h_frames;
    DcsbBuf d_all_entries, d_all_items, d_tls, d_writes, d_seqout, h_meta;     // device-side sequencer path
    cudaEvent_t ev[DCSB_RT_MAX_CHUNKS] = { nullptr };
    DcsbBuf d_entries[DCSB_RT_MAX_CHUNKS], d_items[DCSB_RT_MAX_CHUNKS], h_entries[DCSB_RT_MAX_CHUNKS], h_items[DCSB_RT_MAX_CHUNKS];
};
static void timeline_cache_free(void *p)
{
    DcsbTimelineCache *c = static_cast<DcsbTimelineCache *>(p);
    if (!c) return;
    if (c->st) { cudaStreamSynchronize(c->st); cudaStreamDestroy(c->st); }
    if (c->copy) { cudaStreamSynchronize(c->copy); cudaStreamDestroy(c->copy); }
    for (cudaEvent_t &e : c->ev) if (e) cudaEventDestroy(e);
    for (DcsbBuf *b : { &c->d_frames, &c->d_pcm, &c->d_csum, &c->d_all_entries, &c->d_all_items, &c->d_tls, &c->d_writes, &c->d_seqout }) b->release(false);
    c->h_frames.release(true);
    c->h_meta.release(true);
    for (int k = 0; k < DCSB_RT_MAX_CHUNKS; ++k) {
        c->d_entries[k].release(false); c->d_items[k].release(false);
        c->h_entries[k].release(true); c->h_items[k].release(true);
    }
    delete c;
}

// The device-side route (the default): the track interpreter runs on the GPU, one thread per timeline
// (dcsb_seq_kernel), straight into the schedule arrays the mix kernel reads.  Host work: flatten the port writes
// (a few bytes per event), lay out the work items (known from the frame counts alone).  The mix kernel runs chunk
// by chunk on one stream while the PCM of the chunk before goes down the link on another.
static int render_timelines_device(dcsb_ctx *ctx, dcsb_rom *rom, DcsbTimelineCache &tc, const dcsb_timeline *timelines, size_t n,
                                   const std::vector<uint64_t> &first, bool fam93, bool packed, bool pinned,
                                   int16_t *pcm_out, const uint64_t *pcm_offsets, dcsb_timeline_result *results)
{
    dcsb_batch *b = rom->batch;
    const uint64_t total = first[n];
    const uint32_t item_len = mix_item_len(fam93, total);
    if (total * DCSB_MAX_CHANNELS >= 0xFFFFFFFFull) return fail(ctx, DCSB_E_ARG, "dcsb_render_timelines: more than 2^29 frames in one call");
    size_t nwrites = 0, nitems = 0;
    for (size_t t = 0; t < n; ++t) { nwrites += timelines[t].n_writes; nitems += (timelines[t].n_frames + item_len - 1) / item_len; }
#define ENS(buf, bytes, host, what) do { cudaError_t e_ = (buf).ensure((bytes), (host)); if (e_ != cudaSuccess) return fail(ctx, DCSB_E_NOMEM, what, e_); } while (0)
    const size_t b_tls = n * sizeof(DcsbSeqTimeline), b_wr = std::max<size_t>(1, nwrites) * sizeof(dcsb_port_write), b_it = std::max<size_t>(1, nitems) * sizeof(DcsbMixItem);
    ENS(tc.h_meta, b_tls + b_wr + b_it + 64, true, "cudaMallocHost(timeline descriptors)");
    ENS(tc.d_tls, b_tls, false, "cudaMalloc(timelines)");
    ENS(tc.d_writes, b_wr, false, "cudaMalloc(port writes)");
    ENS(tc.d_all_items, b_it, false, "cudaMalloc(mix items)");
    ENS(tc.d_frames, total * sizeof(DcsbSchedFrame), false, "cudaMalloc(schedule frames)");
    ENS(tc.d_all_entries, total * DCSB_MAX_CHANNELS * sizeof(DcsbSchedEntry), false, "cudaMalloc(schedule entries)");
    ENS(tc.d_seqout, n * 8, false, "cudaMalloc(sequencer results)");
    ENS(tc.d_pcm, total * 480, false, "cudaMalloc(pcm)");
    ENS(tc.d_csum, n * 8, false, "cudaMalloc(checksums)");
#undef ENS
    if (!tc.copy) CK(cudaStreamCreateWithFlags(&tc.copy, cudaStreamNonBlocking), "cudaStreamCreate");
    for (cudaEvent_t &ev : tc.ev) if (!ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate");
    uint8_t *hm = static_cast<uint8_t *>(tc.h_meta.p);
    DcsbSeqTimeline *htl = reinterpret_cast<DcsbSeqTimeline *>(hm);
    dcsb_port_write *hw = reinterpret_cast<dcsb_port_write *>(hm + b_tls);
    DcsbMixItem *hi = reinterpret_cast<DcsbMixItem *>(hm + b_tls + b_wr);
    // chunks of about equal frame counts (whole timelines); small calls are one chunk
    const size_t nchunks = (size_t)std::max<uint64_t>(1, std::min<uint64_t>(std::min<uint64_t>(DCSB_RT_MAX_CHUNKS, n), total / 65536));
    std::vector<size_t> cut(nchunks + 1, n), icut(nchunks + 1, 0);
    cut[0] = 0;
    for (size_t k = 1, t = 0; k < nchunks; ++k) {
        while (t < n && first[t] < total * k / nchunks) ++t;
        cut[k] = t;
    }
    size_t w = 0, it = 0, k = 0;
    for (size_t t = 0; t < n; ++t) {
        while (k + 1 <= nchunks && t == cut[k]) icut[k++] = it;
        const dcsb_timeline &tl = timelines[t];
        htl[t] = DcsbSeqTimeline{ (uint32_t)w, tl.n_writes, tl.n_frames, (uint32_t)first[t], tl.master_volume };
        for (uint32_t i = 0; i < tl.n_writes; ++i) hw[w++] = tl.writes[i];
        const uint32_t f0 = (uint32_t)first[t];
        for (uint32_t f = 0; f < tl.n_frames; f += item_len) hi[it++] = DcsbMixItem{ f0 + f, std::min<uint32_t>(item_len, tl.n_frames - f), f0, (uint32_t)t };
    }
    while (k <= nchunks) icut[k++] = it;
    CK(cudaMemsetAsync(tc.d_csum.p, 0, n * 8, tc.st), "memset checksums");
    CK(cudaMemcpyAsync(tc.d_tls.p, htl, b_tls, cudaMemcpyHostToDevice, tc.st), "H2D timelines");
    if (nwrites) CK(cudaMemcpyAsync(tc.d_writes.p, hw, nwrites * sizeof(dcsb_port_write), cudaMemcpyHostToDevice, tc.st), "H2D port writes");
    CK(cudaMemcpyAsync(tc.d_all_items.p, hi, nitems * sizeof(DcsbMixItem), cudaMemcpyHostToDevice, tc.st), "H2D mix items");
    DcsbRomView dv = rom->view();
    dv.image = b->d_slab;
    dv.streams = rom->d_seq_streams;
    CK(dcsb_launch_seq(&dv, static_cast<const DcsbSeqTimeline *>(tc.d_tls.p), (int)n, tc.d_writes.p, tc.d_frames.p, tc.d_all_entries.p,
                       static_cast<uint32_t *>(tc.d_seqout.p), tc.st), "sequencer kernel launch");
    cudaError_t e = cudaSuccess;
    for (size_t c = 0; c < nchunks && e == cudaSuccess; ++c) {
        const size_t t0 = cut[c], t1 = cut[c + 1];
        if (t0 == t1 || icut[c + 1] == icut[c]) continue;
        e = dcsb_launch_mix(fam93, b->d_slab, b->d_recs, static_cast<DcsbMixItem *>(tc.d_all_items.p) + icut[c], (int)(icut[c + 1] - icut[c]),
                            tc.d_frames.p, tc.d_all_entries.p, ctx->d_tables, b->scan, (int16_t *)tc.d_pcm.p, (unsigned long long *)tc.d_csum.p, tc.st);
        if (e == cudaSuccess && pinned) {
            const uint64_t cf0 = first[t0], cfn = first[t1] - first[t0];
            e = cudaEventRecord(tc.ev[c], tc.st);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(tc.copy, tc.ev[c], 0);
            if (e == cudaSuccess && cfn)
                e = cudaMemcpyAsync(pcm_out + cf0 * 240, (const int16_t *)tc.d_pcm.p + cf0 * 240, cfn * 480, cudaMemcpyDeviceToHost, tc.copy);
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(tc.st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(tc.copy);
    if (e == cudaSuccess && !pinned) {
        if (packed) e = cudaMemcpy(pcm_out, tc.d_pcm.p, total * 480, cudaMemcpyDeviceToHost);
        else
            for (size_t t = 0; t < n && e == cudaSuccess; ++t)
                if (timelines[t].n_frames)
                    e = cudaMemcpy(pcm_out + pcm_offsets[t], (const int16_t *)tc.d_pcm.p + first[t] * 240,
                                   (size_t)timelines[t].n_frames * 480, cudaMemcpyDeviceToHost);
    }
    if (e == cudaSuccess && results) {
        std::vector<unsigned long long> cs(n);
        std::vector<uint32_t> so(2 * n);
        e = cudaMemcpy(cs.data(), tc.d_csum.p, n * 8, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(so.data(), tc.d_seqout.p, n * 8, cudaMemcpyDeviceToHost);
        for (size_t t = 0; t < n; ++t) {
            results[t].status = so[2 * t] ? DCSB_E_STOPPED : DCSB_OK;
            results[t].frames = timelines[t].n_frames;
            results[t].checksum = cs[t];
            results[t].n_host_bytes = so[2 * t + 1];
            results[t].reserved = 0;
        }
    }
    if (e != cudaSuccess) return fail(ctx, DCSB_E_CUDA, "dcsb_render_timelines", e);
    return DCSB_OK;
}

extern "C" int dcsb_render_timelines(dcsb_ctx *ctx, dcsb_rom *rom, const dcsb_timeline *timelines, size_t n,
                                     int16_t *pcm_out, const uint64_t *pcm_offsets, dcsb_timeline_result *results)
{
    if (!ctx || !rom || (!timelines && n) || (!pcm_out && n)) return fail(ctx, DCSB_E_ARG, "dcsb_render_timelines: bad argument");
    if (n == 0) return DCSB_OK;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    int rc = rom_prepare(ctx, rom);
    if (rc != DCSB_OK) return rc;
    const auto t_begin = std::chrono::steady_clock::now();
    const bool trace = getenv("DCSB_TRACE") != nullptr;
    auto lap = [&](const char *what, int k) {      // DCSB_TRACE=1: host-side phases of the call
        if (trace)
            fprintf(stderr, "[dcsb trace] render_timelines chunk %d %-20s at %8.3f ms\n", k, what,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };
    if (!ctx->timeline_cache) { ctx->timeline_cache = new DcsbTimelineCache(); ctx->timeline_cache_free = timeline_cache_free; }
    DcsbTimelineCache &tc = *static_cast<DcsbTimelineCache *>(ctx->timeline_cache);
    if (!tc.st) CK(cudaStreamCreateWithFlags(&tc.st, cudaStreamNonBlocking), "cudaStreamCreate");
    const bool fam93 = rom->os == DCSB_OS93A || rom->os == DCSB_OS93B;
    dcsb_batch *b = rom->batch;

    // every timeline yields exactly n_frames schedule frames: the layout of frames and PCM is known up front
    std::vector<uint64_t> first(n + 1, 0);
    for (size_t t = 0; t < n; ++t) first[t + 1] = first[t] + timelines[t].n_frames;
    const uint64_t total = first[n];
    if (total == 0) return DCSB_OK;
    if (total >= 0xFFFFFFFFull) return fail(ctx, DCSB_E_ARG, "dcsb_render_timelines: more than 2^32 frames in one call");
    bool packed = true;
    for (size_t t = 0; t < n && pcm_offsets; ++t) if (pcm_offsets[t] != first[t] * 240) packed = false;
    const bool pinned = packed && is_pinned_host(pcm_out) && is_pinned_host(pcm_out + total * 240 - 1);
    // DCSB_SEQ_HOST=1: the track interpreters on host threads (the route of round 1; kept for comparison and tests)
    if (!getenv("DCSB_SEQ_HOST")) {
        const int drc = render_timelines_device(ctx, rom, tc, timelines, n, first, fam93, packed, pinned, pcm_out, pcm_offsets, results);
        lap("device route done", -1);
        return drc;
    }
    const uint32_t item_len = mix_item_len(fam93, total);
#define ENS(buf, bytes, host, what) do { cudaError_t e_ = (buf).ensure((bytes), (host)); if (e_ != cudaSuccess) return fail(ctx, DCSB_E_NOMEM, what, e_); } while (0)
    ENS(tc.d_frames, total * sizeof(DcsbSchedFrame), false, "cudaMalloc(schedule frames)");
    ENS(tc.h_frames, total * sizeof(DcsbSchedFrame), true, "cudaMallocHost(schedule frames)");
    ENS(tc.d_pcm, total * 480, false, "cudaMalloc(pcm)");
    ENS(tc.d_csum, n * 8, false, "cudaMalloc(checksums)");
    CK(cudaMemsetAsync(tc.d_csum.p, 0, n * 8, tc.st), "memset checksums");
    DcsbSchedFrame *hf = static_cast<DcsbSchedFrame *>(tc.h_frames.p);

    // chunks of about equal frame counts; small calls are one chunk
    const size_t nchunks = (size_t)std::max<uint64_t>(1, std::min<uint64_t>(std::min<uint64_t>(DCSB_RT_MAX_CHUNKS, n), total / 65536));
    std::vector<size_t> cut(nchunks + 1, n);
    cut[0] = 0;
    for (size_t k = 1, t = 0; k < nchunks; ++k) {
        while (t < n && first[t] < total * k / nchunks) ++t;
        cut[k] = t;
    }
    struct Part { std::vector<DcsbSchedFrame> frames; std::vector<DcsbSchedEntry> entries; bool fatal = false; uint32_t nhost = 0; };
    std::vector<Part> parts(n);
    std::vector<uint64_t> ebase(n + 1, 0);
    const unsigned nth_max = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    auto run_threads = [&](size_t t0, size_t t1, const std::function<void(size_t)> &fn) {
        const unsigned nth = (unsigned)std::min<size_t>(t1 - t0, nth_max);
        if (nth <= 1) { for (size_t t = t0; t < t1; ++t) fn(t); return; }
        std::vector<std::thread> th;
        for (unsigned j = 0; j < nth; ++j) th.emplace_back([&, j] { for (size_t t = t0 + j; t < t1; t += nth) fn(t); });
        for (auto &x : th) x.join();
    };
    cudaError_t e = cudaSuccess;
    for (size_t k = 0; k < nchunks && e == cudaSuccess; ++k) {
        const size_t t0 = cut[k], t1 = cut[k + 1];
        if (t0 == t1) continue;
        // host: one sequencer per timeline
        run_threads(t0, t1, [&](size_t t) {
            const dcsb_timeline &tl = timelines[t];
            Part &pt = parts[t];
            DcsbSequencer seq(rom);
            seq.soft_boot();
            seq.set_master_volume(tl.master_volume);
            uint32_t w = 0;
            pt.frames.reserve(tl.n_frames);
            pt.entries.reserve((size_t)tl.n_frames * 2);
            for (uint32_t f = 0; f < tl.n_frames; ++f) {
                while (w < tl.n_writes && tl.writes[w].frame <= f) seq.write_port(tl.writes[w++].byte);
                seq.frame(pt.frames, pt.entries);
            }
            pt.fatal = seq.fatal;
            pt.nhost = (uint32_t)seq.host_bytes.size();
        });
        lap("sequencers done", (int)k);
        // entries of the chunk are numbered from the chunk's own base; items from 0
        ebase[t0] = 0;
        size_t nitems = 0;
        std::vector<size_t> ibase(t1 - t0 + 1, 0);
        for (size_t t = t0; t < t1; ++t) {
            ebase[t + 1] = ebase[t] + parts[t].entries.size();
            ibase[t - t0 + 1] = ibase[t - t0] + (timelines[t].n_frames + item_len - 1) / item_len;
        }
        nitems = ibase[t1 - t0];
        const size_t nent = (size_t)ebase[t1];
        ENS(tc.h_entries[k], std::max<size_t>(1, nent) * sizeof(DcsbSchedEntry), true, "cudaMallocHost(schedule entries)");
        ENS(tc.d_entries[k], std::max<size_t>(1, nent) * sizeof(DcsbSchedEntry), false, "cudaMalloc(schedule entries)");
        ENS(tc.h_items[k], std::max<size_t>(1, nitems) * sizeof(DcsbMixItem), true, "cudaMallocHost(mix items)");
        ENS(tc.d_items[k], std::max<size_t>(1, nitems) * sizeof(DcsbMixItem), false, "cudaMalloc(mix items)");
        DcsbSchedEntry *he = static_cast<DcsbSchedEntry *>(tc.h_entries[k].p);
        DcsbMixItem *hi = static_cast<DcsbMixItem *>(tc.h_items[k].p);
        run_threads(t0, t1, [&](size_t t) {
            Part &pt = parts[t];
            DcsbSchedFrame *dst = hf + first[t];
            const uint32_t eb = (uint32_t)ebase[t];
            for (size_t f = 0; f < pt.frames.size(); ++f) { dst[f] = pt.frames[f]; dst[f].first_entry += eb; }
            if (!pt.entries.empty()) memcpy(he + ebase[t], pt.entries.data(), pt.entries.size() * sizeof(DcsbSchedEntry));
            DcsbMixItem *it = hi + ibase[t - t0];
            const uint32_t nf = timelines[t].n_frames, f0 = (uint32_t)first[t];
            for (uint32_t f = 0; f < nf; f += item_len) *it++ = DcsbMixItem{ f0 + f, std::min<uint32_t>(item_len, nf - f), f0, (uint32_t)t };
            std::vector<DcsbSchedFrame>().swap(pt.frames);
            std::vector<DcsbSchedEntry>().swap(pt.entries);
        });
        lap("staged", (int)k);
        const uint64_t cf0 = first[t0], cfn = first[t1] - first[t0];
        if (cfn == 0 || nitems == 0) continue;
        e = cudaMemcpyAsync(static_cast<DcsbSchedFrame *>(tc.d_frames.p) + cf0, hf + cf0, cfn * sizeof(DcsbSchedFrame), cudaMemcpyHostToDevice, tc.st);
        if (e == cudaSuccess && nent) e = cudaMemcpyAsync(tc.d_entries[k].p, he, nent * sizeof(DcsbSchedEntry), cudaMemcpyHostToDevice, tc.st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(tc.d_items[k].p, hi, nitems * sizeof(DcsbMixItem), cudaMemcpyHostToDevice, tc.st);
        if (e == cudaSuccess)
            e = dcsb_launch_mix(fam93, b->d_slab, b->d_recs, tc.d_items[k].p, (int)nitems, tc.d_frames.p, tc.d_entries[k].p, ctx->d_tables,
                                b->scan, (int16_t *)tc.d_pcm.p, (unsigned long long *)tc.d_csum.p, tc.st);
        if (e == cudaSuccess && pinned)
            e = cudaMemcpyAsync(pcm_out + cf0 * 240, (const int16_t *)tc.d_pcm.p + cf0 * 240, cfn * 480, cudaMemcpyDeviceToHost, tc.st);
        lap("queued", (int)k);
    }
#undef ENS
    if (e == cudaSuccess) e = cudaStreamSynchronize(tc.st);
    lap("GPU done", -1);
    if (e == cudaSuccess && !pinned) {
        if (packed) e = cudaMemcpy(pcm_out, tc.d_pcm.p, total * 480, cudaMemcpyDeviceToHost);
        else
            for (size_t t = 0; t < n && e == cudaSuccess; ++t)
                if (timelines[t].n_frames)
                    e = cudaMemcpy(pcm_out + pcm_offsets[t], (const int16_t *)tc.d_pcm.p + first[t] * 240,
                                   (size_t)timelines[t].n_frames * 480, cudaMemcpyDeviceToHost);
        lap("PCM on the host", -1);
    }
    if (e == cudaSuccess && results) {
        std::vector<unsigned long long> cs(n);
        e = cudaMemcpy(cs.data(), tc.d_csum.p, n * 8, cudaMemcpyDeviceToHost);
        for (size_t t = 0; t < n; ++t) {
            results[t].status = parts[t].fatal ? DCSB_E_STOPPED : DCSB_OK;
            results[t].frames = timelines[t].n_frames;
            results[t].checksum = cs[t];
            results[t].n_host_bytes = parts[t].nhost;
            results[t].reserved = 0;
        }
    }
    if (e != cudaSuccess) return fail(ctx, DCSB_E_CUDA, "dcsb_render_timelines", e);
    return DCSB_OK;
}
