// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// extern "C" wrapper around the UNMODIFIED reference classes (DCSDecoderNative,
// DCSEncoder), compiled by oracle/Makefile from the sources where they lie under
// /root/reference into oracle/_ref/libdcsref.so.  It exists so that tests/, smoke()
// and bench.py's cpu_baseline / --impl reference legs can (a) pin oracle/dcs_oracle.c
// against the real reference, (b) generate synthetic streams with the reference's own
// encoder, and (c) time the reference's CPU decoder.  Nothing under dcsexplorer_b200/
// links or loads this file.
//
// Protocol followed for "one stream -> PCM" (SURVEY.md section 3B):
//   DCSExplorer/DCSExplorer.cpp:1655-1721 and DCSEncoder/DCSEncoder.cpp:547-571:
//   fresh DCSDecoderNative -> InitStandalone(os) -> SoftBoot() -> SetMasterVolume(v)
//   -> LoadAudioStream(0, ROMPointer(0,data), level) -> GetNextSample() x frames*240.
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <vector>
#include <thread>
#include <chrono>
#include <string>
#include <list>
#include <memory>
#include "DCSDecoderNative.h"
#include "DCSEncoder.h"

namespace {

DCSDecoder::OSVersion ToOS(int v)
{
    switch (v) {
    case 0x9301: return DCSDecoder::OSVersion::OS93a;
    case 0x9302: return DCSDecoder::OSVersion::OS93b;
    case 0x9500: return DCSDecoder::OSVersion::OS95;
    default:     return DCSDecoder::OSVersion::OS94;
    }
}

// host that records DCS->host data port bytes
struct RecHost : DCSDecoder::Host {
    std::vector<uint8_t> bytes;
    void ReceiveDataPort(uint8_t d) override { bytes.push_back(d); }
    void ClearDataPort() override { }
    void BootTimerControl(bool) override { }
};

// subclass that exposes protected decoder state for frame-level probing
struct Probe : DCSDecoderNative {
    Probe(Host *h) : DCSDecoderNative(h) { }
    using DCSDecoderNative::channel;
    using DCSDecoderNative::decoderImpl;
    using DCSDecoderNative::frameBuffer;
    using DCSDecoderNative::overlapBuffer;
    using DCSDecoderNative::outputBuffer;
    using DCSDecoderNative::volumeMultiplier;
    void InitPlay(Channel &ch, const uint8_t *p) {
        InitChannelStream(ch, ROMPointer(0, p));
        InitStreamPlayback(ch);
    }
    uint32_t BitPos(Channel &ch) {
        auto &bp = ch.audioStream.playbackBitPtr;
        return static_cast<uint32_t>((bp.p.p - ch.audioStream.startPtr.p) * 8 - bp.nBits);
    }
    bool &StopFlag(int c) { return channel[c].stop; }
    uint16_t MixMult(int c) { return channel[c].mixingMultiplier; }
};

struct RomCtx {
    RecHost host;
    std::unique_ptr<Probe> dec;
    std::list<DCSDecoder::ZipFileData> zipData;
    std::vector<std::vector<uint8_t>> images;
};

} // namespace

extern "C" {

// Decode one stream with a fresh decoder; pulls n_frames*240 samples. Returns samples written.
int dcsref_decode_stream(const uint8_t *data, size_t nbytes, int os_version,
    int master_volume, int mixing_level, int n_frames, int16_t *pcm)
{
    std::vector<uint8_t> buf(nbytes + 64, 0);
    memcpy(buf.data(), data, nbytes);
    DCSDecoder::MinHost host;
    DCSDecoderNative dec(&host);
    dec.InitStandalone(ToOS(os_version));
    dec.SoftBoot();
    dec.SetMasterVolume(master_volume);
    dec.LoadAudioStream(0, DCSDecoder::ROMPointer(0, buf.data()), mixing_level);
    int n = n_frames * 240;
    for (int i = 0; i < n; ++i)
        pcm[i] = dec.GetNextSample();
    return n;
}

// GetStreamInfo (DCSDecoderNative.cpp:1486-1537)
int dcsref_stream_info(const uint8_t *data, size_t nbytes, int os_version,
    int *n_frames, int *n_bytes, int *type, int *subtype)
{
    std::vector<uint8_t> buf(nbytes + 64, 0);
    memcpy(buf.data(), data, nbytes);
    DCSDecoder::MinHost host;
    DCSDecoderNative dec(&host);
    dec.InitStandalone(ToOS(os_version));
    dec.SoftBoot();
    auto si = dec.GetStreamInfo(DCSDecoder::ROMPointer(0, buf.data()));
    *n_frames = si.nFrames; *n_bytes = si.nBytes; *type = si.formatType; *subtype = si.formatSubType;
    return 0;
}

// Walk the stream with DecompressFrame and report, for every frame f, the bit offset
// (relative to the first byte after the stream header) at which frame f starts
// [bitpos has n_frames+1 entries], the band-type buffer AFTER frame f [16 u16 each],
// the 256 (or 512) raw frequency bins the frame adds into a zeroed frame buffer with
// the given mixing multiplier, and whether the channel stop flag was raised.
int dcsref_probe_frames(const uint8_t *data, size_t nbytes, int os_version, uint16_t mix_mult,
    int max_frames, uint32_t *bitpos, uint16_t *bandtypes, uint16_t *bins, uint8_t *stopflags)
{
    std::vector<uint8_t> buf(nbytes + 64, 0);
    memcpy(buf.data(), data, nbytes);
    DCSDecoder::MinHost host;
    Probe dec(&host);
    dec.InitStandalone(ToOS(os_version));
    dec.SoftBoot();
    auto &ch = dec.channel[0];
    dec.InitPlay(ch, buf.data());
    ch.mixingMultiplier = mix_mult;
    int nf = ch.audioStream.numFrames;
    if (nf > max_frames) nf = max_frames;
    for (int f = 0; f < nf; ++f) {
        if (bitpos) bitpos[f] = dec.BitPos(ch);
        uint16_t fb[0x200];
        memset(fb, 0, sizeof(fb));
        ch.stop = false;
        dec.decoderImpl->DecompressFrame(ch, fb);
        if (bins) memcpy(bins + (size_t)f * 512, fb, sizeof(fb));
        if (bandtypes) memcpy(bandtypes + (size_t)f * 16, ch.audioStream.bandTypeBuf, 32);
        if (stopflags) stopflags[f] = ch.stop ? 1 : 0;
    }
    if (bitpos) bitpos[nf] = dec.BitPos(ch);
    return nf;
}

// Run only TransformFrame on caller-supplied bins/overlap (for per-stage parity tests).
// bins: 512 u16 in/out (frame buffer), overlap: 16 u16 in/out, pcm: 240 out.
int dcsref_transform(int os_version, int vol_shift, uint16_t *bins, uint16_t *overlap, int16_t *pcm)
{
    DCSDecoder::MinHost host;
    Probe dec(&host);
    dec.InitStandalone(ToOS(os_version));
    dec.SoftBoot();
    memcpy(dec.frameBuffer, bins, sizeof(dec.frameBuffer));
    memcpy(dec.overlapBuffer, overlap, sizeof(dec.overlapBuffer));
    dec.decoderImpl->TransformFrame(vol_shift);
    memcpy(bins, dec.frameBuffer, sizeof(dec.frameBuffer));
    memcpy(overlap, dec.overlapBuffer, sizeof(dec.overlapBuffer));
    memcpy(pcm, dec.outputBuffer, 480);
    return 0;
}

// Encode float PCM with the reference DCSEncoder (float WriteStream path,
// DCSEncoder.cpp:650).  Returns bytes written (0 on failure).
size_t dcsref_encode(const float *pcm, size_t n, int sample_rate, int format_version,
    int type, int subtype, int bit_rate, float power_cut, uint8_t *out, size_t cap, int *n_frames)
{
    DCSEncoder enc;
    enc.compressionParams.formatVersion = static_cast<uint16_t>(format_version);
    enc.compressionParams.streamFormatType = type;
    enc.compressionParams.streamFormatSubType = subtype;
    enc.compressionParams.targetBitRate = bit_rate;
    enc.compressionParams.powerBandCutoff = power_cut;
    std::string err;
    auto *s = enc.OpenStream(sample_rate, err);
    if (s == nullptr) return 0;
    enc.WriteStream(s, pcm, n);
    DCSEncoder::DCSAudio obj;
    if (!enc.CloseStream(s, obj, err)) return 0;
    if (obj.nBytes > cap) return 0;
    memcpy(out, obj.data.get(), obj.nBytes);
    if (n_frames) *n_frames = obj.nFrames;
    return obj.nBytes;
}

// The reference encoder WITHOUT its resampler: the samples are framed directly (16 samples of overlap, 240 new ones
// per frame, the last frame zero padded -- what WriteStream does with the resampler's output, DCSEncoder.cpp:693-703)
// and handed to the reference's own TransformFrame / CloseStream.  This is the oracle for the GPU encoder, which has no
// resampler either.  frames_out (may be NULL): the 256 floats the reference keeps per frame (Stream::Frame::f,
// DCSEncoder.h:322), n_frames_cap of them at most.  Returns the stream's byte count (0 on failure).
struct EncProbe : DCSEncoder {
    size_t encode_framed(const float *pcm, size_t n, uint8_t *out, size_t cap, int *n_frames, float *frames_out, size_t n_frames_cap)
    {
        std::string err;
        Stream *s = OpenStream(31250, err);
        if (s == nullptr) return 0;
        for (size_t i = 0; i < n; ++i) {
            s->inputBuf[s->nInputBuf++] = pcm[i];
            if (s->nInputBuf == 256) TransformFrame(s);
        }
        DCSAudio obj;
        if (!CloseStream(s, obj, err)) return 0;
        if (frames_out) {
            size_t k = 0;
            for (auto &f : s->frames) {
                if (k >= n_frames_cap) break;
                memcpy(frames_out + 256 * k, f.f, 256 * sizeof(float));
                ++k;
            }
        }
        if (n_frames) *n_frames = obj.nFrames;
        if (obj.nBytes > cap) return 0;
        memcpy(out, obj.data.get(), obj.nBytes);
        return obj.nBytes;
    }
};
size_t dcsref_encode_framed(const float *pcm, size_t n, int type, int subtype, int bit_rate, float power_cut,
    float max_quant_error, float min_dynamic_range, uint8_t *out, size_t cap, int *n_frames, float *frames_out, size_t n_frames_cap,
    int format_version)
{
    EncProbe enc;
    enc.compressionParams.formatVersion = static_cast<uint16_t>(format_version ? format_version : 0x9400);
    enc.compressionParams.streamFormatType = type;
    enc.compressionParams.streamFormatSubType = subtype;
    enc.compressionParams.targetBitRate = bit_rate;
    enc.compressionParams.powerBandCutoff = power_cut;
    if (max_quant_error >= 0) enc.compressionParams.maximumQuantizationError = max_quant_error;
    if (min_dynamic_range >= 0) enc.compressionParams.minimumDynamicRange = min_dynamic_range;
    return enc.encode_framed(pcm, n, out, cap, n_frames, frames_out, n_frames_cap);
}

// The reference's reader of raw ".dcs" stream files (DCSEncoder::IsDCSFile / EncodeDCSFile,
// DCSEncoder.cpp:358-571): is `path` a DCS file, which format version does its header name, and which
// stream bytes does the reference take from it when the target format is the file's own (the pass-through
// branch, :507-519).  Returns the stream's byte count (copied to out), -1 not a DCS file, -2 read error.
long dcsref_read_dcs_file(const char *path, int *format_version, uint8_t *out, size_t cap, int *n_frames)
{
    int fmt = 0;
    if (!DCSEncoder::IsDCSFile(path, &fmt)) return -1;
    if (format_version) *format_version = fmt;
    DCSEncoder enc;
    enc.compressionParams.formatVersion = static_cast<uint16_t>(fmt);
    DCSEncoder::DCSAudio obj;
    std::string err;
    if (!enc.EncodeDCSFile(path, obj, err)) return -2;
    if (obj.nBytes > cap) return -2;
    memcpy(out, obj.data.get(), obj.nBytes);
    if (n_frames) *n_frames = obj.nFrames;
    return (long)obj.nBytes;
}

// CPU baseline: decode a batch of streams with one DCSDecoderNative per thread,
// streams partitioned round-robin by index; returns wall seconds around the
// GetNextSample loops only (streams pre-loaded, padded copies made before timing).
// pcm_out may be null (samples are still pulled and folded into a checksum).
// per_stream_cs (may be null): for every stream the checksum dcsb_result::checksum is defined as
// (include/dcsb200.h): sum over samples i of (uint16)s[i] * (2 i + 1) mod 2^64 -- so that a caller can
// compare the reference's PCM with another decoder's per-stream checksums without keeping the PCM.
double dcsref_decode_batch_timed(const uint8_t *const *datas, const uint32_t *nbytes,
    const uint32_t *nframes_to_pull, size_t n, int os_version, int master_volume, int mixing_level,
    int n_threads, int16_t *const *pcm_out, uint64_t *checksum_out, uint64_t *per_stream_cs)
{
    std::vector<std::vector<uint8_t>> bufs(n);
    for (size_t i = 0; i < n; ++i) {
        bufs[i].assign(nbytes[i] + 64, 0);
        memcpy(bufs[i].data(), datas[i], nbytes[i]);
    }
    if (n_threads < 1) n_threads = 1;
    std::vector<uint64_t> sums(n_threads, 0);
    auto work = [&](int t) {
        uint64_t h = 0;
        for (size_t i = t; i < n; i += n_threads) {
            DCSDecoder::MinHost host;
            DCSDecoderNative dec(&host);
            dec.InitStandalone(ToOS(os_version));
            dec.SoftBoot();
            dec.SetMasterVolume(master_volume);
            dec.LoadAudioStream(0, DCSDecoder::ROMPointer(0, bufs[i].data()), mixing_level);
            size_t ns = (size_t)nframes_to_pull[i] * 240;
            int16_t *o = pcm_out ? pcm_out[i] : nullptr;
            uint64_t cs = 0;
            for (size_t k = 0; k < ns; ++k) {
                int16_t s = dec.GetNextSample();
                h = h * 1099511628211ULL + (uint16_t)s;
                cs += (uint64_t)(uint16_t)s * (2 * (uint64_t)k + 1);
                if (o) o[k] = s;
            }
            if (per_stream_cs) per_stream_cs[i] = cs;
        }
        sums[t] = h;
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    uint64_t h = 0;
    for (auto s : sums) h ^= s;
    if (checksum_out) *checksum_out = h;
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- ROM / track playback (config 4) -------------------------------------------------

void *dcsref_rom_open_zip(const char *zip_path, int master_volume, char *err, size_t errlen)
{
    auto *ctx = new RomCtx();
    ctx->dec.reset(new Probe(&ctx->host));
    std::string msg;
    auto st = ctx->dec->LoadROMFromZipFile(zip_path, ctx->zipData, nullptr, &msg);
    if (st != DCSDecoder::ZipLoadStatus::Success) {
        if (err && errlen) { strncpy(err, msg.c_str(), errlen - 1); err[errlen - 1] = 0; }
        delete ctx;
        return nullptr;
    }
    ctx->dec->SoftBoot();
    ctx->dec->SetMasterVolume(master_volume);
    return ctx;
}

// rom images given directly: chip numbers 2..9
void *dcsref_rom_open_images(const uint8_t *const *imgs, const size_t *sizes, const int *chipnos, int n, int master_volume)
{
    auto *ctx = new RomCtx();
    ctx->dec.reset(new Probe(&ctx->host));
    for (int i = 0; i < n; ++i) {
        ctx->images.emplace_back(imgs[i], imgs[i] + sizes[i]);
        ctx->images.back().resize(sizes[i] + 64, 0xFF);
    }
    int i = 0;
    for (auto &im : ctx->images) { ctx->dec->AddROM(chipnos[i], im.data(), sizes[i]); ++i; }
    ctx->dec->SoftBoot();
    ctx->dec->SetMasterVolume(master_volume);
    return ctx;
}

void dcsref_rom_close(void *h) { delete static_cast<RomCtx *>(h); }

int dcsref_rom_info(void *h, int *os_version, int *hw_version, int *n_tracks, int *n_channels, int *check)
{
    auto *c = static_cast<RomCtx *>(h);
    if (check) *check = c->dec->CheckROMs();
    DCSDecoder::HWVersion hw; DCSDecoder::OSVersion os;
    c->dec->GetVersionInfo(&hw, &os);
    if (os_version) *os_version = static_cast<int>(os);
    if (hw_version) *hw_version = static_cast<int>(hw);
    if (n_tracks) *n_tracks = c->dec->GetMaxTrackNumber();
    if (n_channels) *n_channels = c->dec->GetNumChannels();
    return 0;
}

void dcsref_rom_write_port(void *h, uint8_t b) { static_cast<RomCtx *>(h)->dec->WriteDataPort(b); }
void dcsref_rom_set_master_volume(void *h, int v) { static_cast<RomCtx *>(h)->dec->SetMasterVolume(v); }

// pull n_frames*240 samples
int dcsref_rom_render(void *h, int n_frames, int16_t *pcm)
{
    auto *c = static_cast<RomCtx *>(h);
    for (int i = 0; i < n_frames * 240; ++i) pcm[i] = c->dec->GetNextSample();
    return n_frames * 240;
}

// bytes the decoder sent back to the host since the last call
int dcsref_rom_host_bytes(void *h, uint8_t *out, int cap)
{
    auto *c = static_cast<RomCtx *>(h);
    int n = (int)c->host.bytes.size();
    if (n > cap) n = cap;
    memcpy(out, c->host.bytes.data(), n);
    c->host.bytes.clear();
    return n;
}

int dcsref_rom_list_streams(void *h, uint32_t *addrs, int cap)
{
    auto *c = static_cast<RomCtx *>(h);
    auto l = c->dec->ListStreams();
    int n = 0;
    for (auto a : l) { if (n < cap) addrs[n] = a; ++n; }
    return n;
}

// GetTrackInfo (DCSDecoder.h:416): out = {valid, address, channel, type, deferCode, time, looping}
int dcsref_rom_track_info(void *h, int track, uint32_t *out)
{
    auto *c = static_cast<RomCtx *>(h);
    DCSDecoder::TrackInfo ti;
    const bool ok = c->dec->GetTrackInfo(static_cast<uint16_t>(track), ti);
    out[0] = ok; out[1] = ti.address; out[2] = (uint32_t)ti.channel; out[3] = (uint32_t)ti.type;
    out[4] = ti.deferCode; out[5] = ti.time; out[6] = ti.looping;
    return ok;
}

// DecompileTrackProgram (DCSDecoder.h:481): per step 5 ints {offset, nestingLevel, loopParent, delayCount, opcode},
// nOperandBytes + 8 operand bytes, desc (64 chars), hexDesc (40 chars) -- the layout of dcsb_opcode
struct RefOpcode { int32_t offset, nesting, parent; uint16_t delay; uint8_t opcode, nops; uint8_t ops[8]; char desc[64]; char hex[40]; };
int dcsref_rom_decompile(void *h, int track, RefOpcode *out, int cap)
{
    auto *c = static_cast<RomCtx *>(h);
    auto v = c->dec->DecompileTrackProgram(static_cast<uint16_t>(track));
    int n = 0;
    for (auto &op : v) {
        if (n < cap) {
            RefOpcode &o = out[n];
            memset(&o, 0, sizeof(o));
            o.offset = op.offset; o.nesting = op.nestingLevel; o.parent = op.loopParent;
            o.delay = op.delayCount; o.opcode = op.opcode; o.nops = static_cast<uint8_t>(op.nOperandBytes);
            memcpy(o.ops, op.operandBytes, 8);
            snprintf(o.desc, sizeof(o.desc), "%s", op.desc.c_str());
            snprintf(o.hex, sizeof(o.hex), "%s", op.hexDesc.c_str());
        }
        ++n;
    }
    return n;
}

} // extern "C"
