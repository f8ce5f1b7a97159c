// DCSDecoderB200 -- the DCSDecoder C++ API surface (mjrgh/DCSExplorer, DCSDecoder/DCSDecoder.h
// and DCSDecoderNative.h) on top of libdcsb200.so, so that code written against the reference's
// decoder classes can switch to the B200 decode path: ROM load (AddROM / LoadROMFromZipFile /
// CheckROMs), stream and track lookup (GetTrackInfo, ListStreams, MakeROMPointer, GetStreamInfo),
// the data port (WriteDataPort, host callbacks) and PCM output (GetNextSample).
//
// Header-only C++17 over the extern "C" boundary of dcsb200.h; link with -ldcsb200.  Method
// names, argument meaning and error behaviour follow the reference (file:line in the comments);
// what differs is stated where it does:
//   * PCM is rendered on the GPU `chunkFrames` frames (7.68 ms each) at a time.  A byte written
//     to the data port takes effect at the next chunk boundary; with chunkFrames = 1 the timing
//     is the reference's (data port drained before every main-loop pass, DCSDecoder.cpp:1625).
//   * there is no ADSP-2105 boot code to run: HardBoot() / StartSelfTests() report the POST code
//     like the reference (0x79, status) and go straight to the decoder (its fast-boot mode,
//     DCSDecoder.cpp:1477-1516); no startup "bong" is synthesised.
//   * no CPU fallback: without a usable sm_100 device the object is in the error state
//     (IsOK() == false, GetErrorMessage() says why), as DCSDecoder reports its own fatal errors
//     (DCSDecoder.h:213-222).
#pragma once
#include <stdint.h>
#include <functional>
#include <list>
#include <string>
#include <vector>
#include "dcsb200.h"

class DCSDecoderB200
{
public:
    // DCSDecoder.h:127-201: the host interface, reduced to what a batch decoder can call
    struct Host {
        virtual ~Host() {}
        virtual void ReceiveDataPort(uint8_t) {}        // a byte from the sound board to the host
        virtual void BootTimerControl(bool) {}
    };
    enum class HWVersion { Unknown, Invalid, DCS93, DCS95 };                    // DCSDecoder.h:800-846
    enum class OSVersion { Unknown, Invalid, OS93a, OS93b, OS94, OS95 };
    enum class ZipLoadStatus { Success, OpenFileError, ExtractError, NoU2 };    // DCSDecoder.h:278-284
    struct TrackInfo {                                                          // DCSDecoder.h:384-414
        uint32_t address = 0;
        int channel = 0, type = 0;
        uint16_t deferCode = 0xFFFF;
        uint32_t time = 0;
        bool looping = false;
    };
    struct StreamInfo {                                                         // DCSDecoderNative.h:106-123
        int nFrames = 0, nBytes = 0, streamType = 0, streamSubType = 0;
        uint8_t header[16] = { 0 };
    };
    struct ROMPointer {                                                         // DCSDecoder.h:730-785
        int chipSelect = 0;
        const uint8_t *p = nullptr;
        uint32_t linearAddress = 0;
        size_t bytesAvailable = 0;      // standalone mode only: readable bytes at p (the reference's decoder reads as far as the bits go)
        ROMPointer() {}
        ROMPointer(int chipSelect, const uint8_t *p, size_t bytesAvailable = 0) : chipSelect(chipSelect), p(p), bytesAvailable(bytesAvailable) {}
        bool IsNull() const { return p == nullptr; }
        int NominalChipNumber() const { return chipSelect + 2; }
    };
    static const int MAX_CHANNELS = 8;
    static const int SAMPLE_RATE = 31250;                                       // DCSDecoder.h:123

    explicit DCSDecoderB200(Host *host = nullptr, int cudaDevice = 0, int chunkFrames = 16)
        : host(host), chunk(chunkFrames < 1 ? 1 : chunkFrames)
    {
        if (dcsb_create(cudaDevice, &ctx) != DCSB_OK) Fail("dcsb200: no usable sm_100 CUDA device (there is no CPU fallback)");
        else if (dcsb_rom_create(&rom) != DCSB_OK) Fail("dcsb200: out of memory");
    }
    ~DCSDecoderB200()
    {
        if (player) dcsb_player_destroy(player);
        if (rom) dcsb_rom_destroy(rom);
        if (ctx) dcsb_destroy(ctx);
    }
    DCSDecoderB200(const DCSDecoderB200 &) = delete;
    DCSDecoderB200 &operator=(const DCSDecoderB200 &) = delete;

    const char *Name() const { return "b200"; }                                 // DCSDecoder.h:210
    bool IsOK() const { return errorMessage.empty(); }                          // DCSDecoder.h:213-222
    bool IsRunning() const { return IsOK() && (player != nullptr || standalone); }
    std::string GetErrorMessage() const { return errorMessage; }

    // ---- ROM load (DCSDecoder.h:285-347).  Images are copied: the caller may free them.
    ZipLoadStatus LoadROMFromZipFile(const char *zipFileName, const char *explicitU2 = nullptr, std::string *errorDetails = nullptr)
    {
        DropPlayer();
        const int rc = rom ? dcsb_rom_load_zip(rom, zipFileName, explicitU2) : DCSB_ZIP_E_OPEN;
        if (rc != DCSB_ZIP_OK && errorDetails) *errorDetails = rom ? dcsb_rom_last_error(rom) : "no ROM object";
        return static_cast<ZipLoadStatus>(rc);
    }
    void AddROM(int n, const uint8_t *data, size_t size) { DropPlayer(); if (rom) dcsb_rom_add(rom, n, data, size); }
    uint8_t CheckROMs() { return rom ? static_cast<uint8_t>(dcsb_rom_check(rom)) : 2; }

    // ---- catalog / track / stream lookup
    uint16_t GetMaxTrackNumber() const { return static_cast<uint16_t>(Info().n_tracks - 1); }          // DCSDecoder.h:360
    bool GetTrackInfo(uint16_t trackNumber, TrackInfo &ti) const                                        // DCSDecoder.h:416
    {
        dcsb_track_info t;
        ti = TrackInfo();
        if (!rom || !dcsb_rom_track_info(rom, trackNumber, &t)) return false;
        ti.address = t.address; ti.channel = t.channel; ti.type = t.type;
        ti.deferCode = t.defer_code; ti.time = t.time; ti.looping = t.looping != 0;
        return true;
    }
    struct Opcode {                                                                                     // DCSDecoder.h:432-480
        int offset = 0, nestingLevel = 0, loopParent = -1;
        uint16_t delayCount = 0;
        uint8_t opcode = 0;
        int nOperandBytes = 0;
        uint8_t operandBytes[8] { 0, 0, 0, 0, 0, 0, 0, 0 };
        std::string desc, hexDesc;
    };
    std::vector<Opcode> DecompileTrackProgram(uint16_t trackNumber) const                               // DCSDecoder.h:481
    {
        std::vector<dcsb_opcode> raw(rom ? dcsb_rom_decompile_track(rom, trackNumber, nullptr, 0) : 0);
        if (!raw.empty()) dcsb_rom_decompile_track(rom, trackNumber, raw.data(), raw.size());
        std::vector<Opcode> v(raw.size());
        for (size_t i = 0; i < raw.size(); ++i) {
            v[i].offset = raw[i].offset; v[i].nestingLevel = raw[i].nesting_level; v[i].loopParent = raw[i].loop_parent;
            v[i].delayCount = raw[i].delay_count; v[i].opcode = raw[i].opcode; v[i].nOperandBytes = raw[i].n_operand_bytes;
            for (int k = 0; k < 8; ++k) v[i].operandBytes[k] = raw[i].operand_bytes[k];
            v[i].desc = raw[i].desc; v[i].hexDesc = raw[i].hex_desc;
        }
        return v;
    }
    // DCSDecoder.h:418-425 / DCSDecoder.cpp:1137-1228: the listing DCSExplorer prints for a track.  (The reference's
    // format string for a deferred track, "%Deferred", has a stray conversion; the intended text is produced here.)
    std::string ExplainTrackProgram(uint16_t trackNumber, const char *linePrefix) const
    {
        TrackInfo ti;
        if (!GetTrackInfo(trackNumber, ti)) return "[Invalid track]";
        char buf[256];
        if (ti.type == 2) { snprintf(buf, sizeof(buf), "%sDeferred ($%04x)", linePrefix, ti.deferCode); return buf; }
        if (ti.type == 3) { snprintf(buf, sizeof(buf), "%sDeferred Indirect ($%02x[$%02x])", linePrefix, ti.deferCode & 0xFF, ti.deferCode >> 8); return buf; }
        std::string program, loopIndent;
        for (auto &ele : DecompileTrackProgram(trackNumber)) {
            if (!program.empty()) program += "\n";
            std::string wait;
            if (ele.delayCount == 0xFFFFu) wait = "Wait(Forever) ";
            else if (ele.delayCount != 0) { snprintf(buf, sizeof(buf), "Wait(%u) ", ele.delayCount); wait = buf; }
            std::string comment = "// " + ele.hexDesc;
            if (ele.opcode == 0x0F) {
                if (ele.delayCount != 0 && !loopIndent.empty()) {
                    snprintf(buf, sizeof(buf), "%-60s    %s\n", (loopIndent + wait).c_str(), comment.c_str());
                    program += std::string(linePrefix) + buf;
                    wait.clear(); comment.clear();
                }
                if (!loopIndent.empty()) loopIndent = loopIndent.substr(2);
                else comment += " Unmatched loop end opcode (0x0F)";
            }
            snprintf(buf, sizeof(buf), "%-60s    %s", (loopIndent + wait + ele.desc).c_str(), comment.c_str());
            program += std::string(linePrefix) + buf;
            if (ele.opcode == 0x0E) loopIndent += "  ";
        }
        return program;
    }
    std::list<uint32_t> ListStreams() const                                                             // DCSDecoder.h:486
    {
        std::vector<uint32_t> v(rom ? dcsb_rom_list_streams(rom, nullptr, 0) : 0);
        if (!v.empty()) dcsb_rom_list_streams(rom, v.data(), v.size());
        return std::list<uint32_t>(v.begin(), v.end());
    }
    ROMPointer MakeROMPointer(uint32_t linearAddress) const                                             // DCSDecoder.h:792
    {
        ROMPointer rp;
        rp.linearAddress = linearAddress;
        rp.chipSelect = static_cast<int>((linearAddress >> (Info().hw_version == 3 ? 21 : 20)) & 7);
        rp.p = rom ? dcsb_rom_pointer(rom, linearAddress, nullptr) : nullptr;
        return rp;
    }
    std::string GetSignature() const { return Info().signature; }
    int GetNumChannels() const { return Info().n_channels; }                                            // DCSDecoder.h:1110
    int GetVersionNumber() const { return Info().version_number; }                                      // DCSDecoder.h:1096
    HWVersion GetHWVersion() const { return HWFrom(Info().hw_version); }
    OSVersion GetOSVersion() const { return OSFrom(Info().os_version); }
    std::string GetVersionInfo(HWVersion *hw = nullptr, OSVersion *os = nullptr) const                  // DCSDecoder.h:1081
    {
        const dcsb_rom_info i = Info();
        if (hw) *hw = HWFrom(i.hw_version);
        if (os) *os = OSFrom(i.os_version);
        const char *h = i.hw_version == 2 ? "DCS audio board" : i.hw_version == 3 ? "DCS-95 A/V board"
                      : i.hw_version == 1 ? "Hardware type not detected" : "Unknown hardware type";
        char s[64] = "Unknown";
        switch (i.os_version) {
        case DCSB_OS93A: snprintf(s, sizeof(s), "Software 1.0a (1993)"); break;
        case DCSB_OS93B: snprintf(s, sizeof(s), "Software 1.0b (1993)"); break;
        case DCSB_OS94: snprintf(s, sizeof(s), "Software 1.01 (1993)"); break;
        case DCSB_OS95:
            if (i.version_number) snprintf(s, sizeof(s), "Software %d.%02d (%s)", i.version_number >> 8, i.version_number & 0xFF,
                                           i.version_number == 0x0103 ? "1995" : (i.version_number == 0x0104 || i.version_number == 0x0105) ? "1997" : "1995+");
            else snprintf(s, sizeof(s), "Software 1.02 (1995)");
            break;
        default: if (i.hw_version == 1) snprintf(s, sizeof(s), "Not detected"); break;
        }
        return std::string(h) + ", " + s;
    }

    // ---- standalone mode (DCSDecoderNative.h:19-34): no ROMs, streams played from the caller's memory.  It implements the
    // protocol the reference's own clients use it for (DCSEncoder.cpp:547-571, DCSExplorer.cpp:1655-1721): InitStandalone(os),
    // SoftBoot(), SetMasterVolume(v), LoadAudioStream(0, ptr, level), then GetNextSample() for as long as wanted -- one stream,
    // channel 0, decoded as by a freshly booted decoder (dcsb_decode_streams' contract); after the stream and the overlap
    // tail of its last frame the output is silence, as the reference's is.
    void InitStandalone(OSVersion os)
    {
        DropPlayer();
        standalone = true;
        standaloneOS = os == OSVersion::OS93a ? DCSB_OS93A : os == OSVersion::OS93b ? DCSB_OS93B : os == OSVersion::OS95 ? DCSB_OS95 : DCSB_OS94;
        standaloneVolume = defaultVolume;
        buf.clear();
        bufPos = 0;
    }
    void LoadAudioStream(int channel, const uint8_t *data, size_t nbytes, int mixingLevel)
    {
        if (!standalone || !IsOK()) return;
        buf.clear();
        bufPos = 0;
        standaloneFrames = 0;
        if (channel != 0) { Fail("dcsb200: standalone mode plays one stream, on channel 0"); return; }
        dcsb_stream_desc d = dcsb_stream_desc();
        d.data = data;
        d.nbytes = static_cast<uint32_t>(nbytes);
        d.os_version = static_cast<uint16_t>(standaloneOS);
        d.master_volume = static_cast<uint8_t>(standaloneVolume < 0 ? 0 : standaloneVolume > 255 ? 255 : standaloneVolume);
        d.mixing_level = static_cast<uint8_t>(mixingLevel);
        d.tail_frames = 1;                                  // the frame that carries the last frame's overlap tail
        const uint32_t nf = nbytes >= 2 ? ((static_cast<uint32_t>(data[0]) << 8) | data[1]) : 0;
        buf.assign(static_cast<size_t>(nf + 1) * 240, 0);
        dcsb_result r;
        const uint64_t off = 0;
        if (dcsb_decode_streams(ctx, &d, 1, buf.data(), &off, &r) != DCSB_OK) { Fail(std::string("dcsb200: ") + dcsb_last_error(ctx)); buf.clear(); return; }
        standaloneFrames = nf;
    }

    // ---- boot (DCSDecoder.h:577-594)
    void SetDefaultVolume(int vol) { defaultVolume = vol; }
    void SoftBoot()
    {
        if (!IsOK()) return;
        if (standalone) { standaloneVolume = defaultVolume; buf.clear(); bufPos = 0; standaloneFrames = 0; return; }
        DropPlayer();
        if (Info().hw_version == 0) CheckROMs();
        if (dcsb_player_create(ctx, rom, &player) != DCSB_OK) { Fail(std::string("dcsb200: ") + dcsb_last_error(ctx)); return; }
        dcsb_player_set_master_volume(player, defaultVolume);
        buf.clear();
        bufPos = 0;
    }
    void HardBoot() { StartSelfTests(); }
    void StartSelfTests()
    {
        const uint8_t post = CheckROMs();
        if (host) { host->ReceiveDataPort(0x79); host->ReceiveDataPort(post); }
        SoftBoot();
    }

    // ---- run time
    void SetFastBootMode(bool) {}                                                                       // DCSDecoder.h:541 (always fast here)
    void SetMasterVolume(int vol)                                                                       // DCSDecoder.h:546
    {
        if (standalone) standaloneVolume = vol;
        else if (player) dcsb_player_set_master_volume(player, vol);
    }
    void WriteDataPort(uint8_t b) { if (player) dcsb_player_write_data_port(player, b); }               // DCSDecoder.h:663
    int16_t GetNextSample()                                                                             // DCSDecoder.h:565
    {
        if (standalone) return bufPos < buf.size() ? buf[bufPos++] : 0;
        if (!player) return 0;
        if (bufPos >= buf.size()) {
            buf.resize(static_cast<size_t>(chunk) * 240);
            if (dcsb_player_render(player, static_cast<uint32_t>(chunk), buf.data()) != DCSB_OK) {
                Fail(std::string("dcsb200: ") + dcsb_last_error(ctx));
                DropPlayer();
                return 0;
            }
            bufPos = 0;
            if (host) {
                uint8_t hb[256];
                for (size_t n; (n = dcsb_player_host_bytes(player, hb, sizeof(hb))) != 0;)
                    for (size_t i = 0; i < n; ++i) host->ReceiveDataPort(hb[i]);
            }
        }
        return buf[bufPos++];
    }
    // whole frames at once (no per-sample call overhead): n_frames * 240 samples
    bool GetFrames(uint32_t n_frames, int16_t *pcm)
    {
        for (uint32_t i = 0; i < n_frames * 240u; ++i) pcm[i] = GetNextSample();
        return IsOK();
    }

    // ---- DCSDecoderNative extras (DCSDecoderNative.h:34-129)
    void LoadAudioStream(int channel, const ROMPointer &streamPtr, int mixingLevel)
    {
        if (standalone) {
            if (streamPtr.p && streamPtr.bytesAvailable) LoadAudioStream(channel, streamPtr.p, streamPtr.bytesAvailable, mixingLevel);
            else Fail("dcsb200: a standalone stream needs its size: ROMPointer(0, data, nbytes)");
        } else if (player) dcsb_player_load_audio_stream(player, channel, streamPtr.linearAddress, mixingLevel);
    }
    bool IsStreamPlaying(int channel) const
    {
        if (standalone) return channel == 0 && bufPos < static_cast<size_t>(standaloneFrames) * 240;
        return player && dcsb_player_is_stream_playing(player, channel) != 0;
    }
    StreamInfo GetStreamInfo(const ROMPointer &streamPtr) const
    {
        StreamInfo si;
        dcsb_stream_info i;
        if (standalone) {
            // the reference finds a stream's size by walking all of its frames (DCSDecoderNative.cpp:1486-1537); so does this
            if (!streamPtr.p || streamPtr.bytesAvailable < 3 || !IsOK()) return si;
            dcsb_stream_desc d = dcsb_stream_desc();
            d.data = streamPtr.p;
            d.nbytes = static_cast<uint32_t>(streamPtr.bytesAvailable);
            d.os_version = static_cast<uint16_t>(standaloneOS);
            d.master_volume = 255;
            d.mixing_level = 0x64;
            const uint32_t nf = (static_cast<uint32_t>(streamPtr.p[0]) << 8) | streamPtr.p[1];
            std::vector<int16_t> scratch(static_cast<size_t>(nf ? nf : 1) * 240);
            dcsb_result r;
            const uint64_t off = 0;
            if (dcsb_decode_streams(ctx, &d, 1, scratch.data(), &off, &r) != DCSB_OK || r.stream_bytes == 0) return si;
            si.nFrames = static_cast<int>(nf);
            si.nBytes = static_cast<int>(r.stream_bytes);
            for (size_t k = 0; k < 16 && 2 + k < streamPtr.bytesAvailable; ++k) si.header[k] = streamPtr.p[2 + k];
            si.streamType = si.header[0] >> 7;
            if (standaloneOS == DCSB_OS94 || standaloneOS == DCSB_OS95) si.streamSubType = ((si.header[1] & 0x80) >> 6) | ((si.header[1] & 0x80) >> 7);
            return si;
        }
        if (player && dcsb_player_stream_info(player, streamPtr.linearAddress, &i) == DCSB_OK) {
            si.nFrames = i.n_frames; si.nBytes = i.n_bytes; si.streamType = i.stream_type; si.streamSubType = i.stream_subtype;
            for (int k = 0; k < 16; ++k) si.header[k] = i.header[k];
        }
        return si;
    }
    void ClearTracks() { if (player) dcsb_player_clear_tracks(player); }
    void AddTrackCommand(uint16_t trackNum) { if (player) dcsb_player_add_track_command(player, trackNum); }

private:
    Host *host;
    int chunk;
    int defaultVolume = 0x67;                           // DCSDecoder.h:1146
    bool standalone = false;
    int standaloneOS = DCSB_OS94, standaloneVolume = 0x67;
    uint32_t standaloneFrames = 0;
    dcsb_ctx *ctx = nullptr;
    dcsb_rom *rom = nullptr;
    dcsb_player *player = nullptr;
    std::vector<int16_t> buf;
    size_t bufPos = 0;
    std::string errorMessage;

    void Fail(const std::string &m) { if (errorMessage.empty()) errorMessage = m; }
    void DropPlayer() { if (player) { dcsb_player_destroy(player); player = nullptr; } }
    dcsb_rom_info Info() const
    {
        dcsb_rom_info i;
        if (!rom || dcsb_rom_get_info(rom, &i) != DCSB_OK) { i = dcsb_rom_info(); }
        return i;
    }
    static HWVersion HWFrom(int v) { return v == 2 ? HWVersion::DCS93 : v == 3 ? HWVersion::DCS95 : v == 1 ? HWVersion::Invalid : HWVersion::Unknown; }
    static OSVersion OSFrom(int v)
    {
        return v == DCSB_OS93A ? OSVersion::OS93a : v == DCSB_OS93B ? OSVersion::OS93b : v == DCSB_OS94 ? OSVersion::OS94
             : v == DCSB_OS95 ? OSVersion::OS95 : OSVersion::Unknown;
    }
};
