// DCSDecoderB200 -- the DCSDecoder C++ API surface (mjrgh/DCSExplorer, DCSDecoder/DCSDecoder.h
// and DCSDecoderNative.h) on top of libdcsb200.so, so that code written against the reference's
// decoder classes can switch to the B200 decode path: ROM load (AddROM / LoadROMFromZipFile /
// CheckROMs), stream and track lookup (GetTrackInfo, ListStreams, MakeROMPointer, GetStreamInfo),
// the data port (WriteDataPort, host callbacks) and PCM output (GetNextSample).
//
// Header-only C++17 over the extern "C" boundary of dcsb200.h; link with -ldcsb200.  Method
// names, argument meaning and error behaviour follow the reference (file:line in the comments);
// what differs is stated where it does:
//   * PCM is rendered on the GPU `chunkFrames` frames (7.68 ms each) ahead (dcsb_player_set_lookahead).
//     The timing of inputs is the reference's all the same (data port drained before every main-loop
//     pass, DCSDecoder.cpp:1625): a byte written to the data port, a volume change or a track command
//     takes effect at the frame it arrives at -- what was rendered ahead of it is dropped and rendered anew.
//   * there is no ADSP-2105 boot code to run, but the boot sequence a client sees is the
//     reference's (DCSDecoder.cpp:1233-1246, :1477-1516, :1579-1619, :1697-1728): HardBoot() gives
//     250 ms of silence during which a data-port byte soft-boots at once, then the POST code
//     (0x79, status) goes to the host and the startup "bong" plays once per status count (a decaying
//     195 Hz square wave, synthesised on the host as the reference does) unless SetFastBootMode(true).
//   * no CPU fallback: without a usable sm_100 device the object is in the error state
//     (IsOK() == false, GetErrorMessage() says why), as DCSDecoder reports its own fatal errors
//     (DCSDecoder.h:213-222).
#pragma once
#include <stdint.h>
#include <string.h>
#include <functional>
#include <list>
#include <memory>
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
    struct ZipFileData {                                                        // DCSDecoder.h:225-235
        ZipFileData(const char *filename, size_t uncompressedSize)
            : filename(filename), data(new uint8_t[uncompressedSize ? uncompressedSize : 1]), dataSize(uncompressedSize) {}
        std::string filename;
        int chipNum = -1;
        std::unique_ptr<uint8_t[]> data;
        size_t dataSize;
    };
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

    // ---- ROM load (DCSDecoder.h:285-347).  The reference's signature: the zip's files are handed back in
    // zipFileData (one entry per file, chipNum = 2..9 for the ones taken as ROM images, else -1).  The reference
    // keeps pointers into that list; this decoder copies the images, so the caller MAY free the list (keeping it,
    // as the reference requires, is harmless).
    ZipLoadStatus LoadROMFromZipFile(const char *zipFileName, std::list<ZipFileData> &zipFileData,
                                     const char *explicitU2 = nullptr, std::string *errorDetails = nullptr)
    {
        DropPlayer();
        const int rc = rom ? dcsb_rom_load_zip(rom, zipFileName, explicitU2) : DCSB_ZIP_E_OPEN;
        if (rc != DCSB_ZIP_OK && errorDetails) *errorDetails = rom ? dcsb_rom_last_error(rom) : "no ROM object";
        if (rom) {
            std::vector<dcsb_zip_file> files(dcsb_rom_zip_files(rom, nullptr, 0));
            if (!files.empty()) dcsb_rom_zip_files(rom, files.data(), files.size());
            for (const dcsb_zip_file &f : files) {
                zipFileData.emplace_back(f.name, f.size);
                zipFileData.back().chipNum = f.chip_number;
                if (f.size) memcpy(zipFileData.back().data.get(), f.data, f.size);
            }
        }
        return static_cast<ZipLoadStatus>(rc);
    }
    // convenience form for callers that do not want the file list
    ZipLoadStatus LoadROMFromZipFile(const char *zipFileName, const char *explicitU2 = nullptr, std::string *errorDetails = nullptr)
    {
        std::list<ZipFileData> files;
        return LoadROMFromZipFile(zipFileName, files, explicitU2, errorDetails);
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
        if (host) host->BootTimerControl(false);
        state = State::Running;
        if (standalone) { standaloneVolume = defaultVolume; buf.clear(); bufPos = 0; standaloneFrames = 0; return; }
        DropPlayer();
        if (Info().hw_version == 0) CheckROMs();
        if (dcsb_player_create(ctx, rom, &player) != DCSB_OK) { Fail(std::string("dcsb200: ") + dcsb_last_error(ctx)); return; }
        dcsb_player_set_lookahead(player, chunk > 1 ? static_cast<uint32_t>(chunk) : 0u);
        dcsb_player_set_master_volume(player, defaultVolume);
        buf.clear();
        bufPos = 0;
    }
    void HardBoot()                                                                                     // DCSDecoder.cpp:1233-1246
    {
        DropPlayer();
        state = State::HardBoot;
        modeSampleCounter = 0;
        if (host) host->BootTimerControl(true);
    }
    void StartSelfTests()                                                                               // DCSDecoder.cpp:1477-1516
    {
        if (host) host->BootTimerControl(false);
        if (state != State::HardBoot) return;
        const uint8_t post = CheckROMs();
        if (host) { host->ReceiveDataPort(0x79); host->ReceiveDataPort(post); }
        if (fastBootMode) { SoftBoot(); return; }
        BongStart();
        state = State::Bong;
        modeSampleCounter = 0;
        bongCount = post;
    }

    // ---- run time
    void SetFastBootMode(bool fast) { fastBootMode = fast; }                                            // DCSDecoder.h:541
    void SetMasterVolume(int vol)                                                                       // DCSDecoder.h:546
    {
        if (standalone) standaloneVolume = vol;
        else if (player) dcsb_player_set_master_volume(player, vol);
    }
    void WriteDataPort(uint8_t b)                                                                       // DCSDecoder.h:663, DCSDecoder.cpp:1542-1558
    {
        if (state == State::HardBoot) { SoftBoot(); return; }       // a byte during the 250 ms boot wait cancels the self test; it is not queued
        if (player) dcsb_player_write_data_port(player, b);
    }
    int16_t GetNextSample()                                                                             // DCSDecoder.h:565, DCSDecoder.cpp:1579-1690
    {
        if (state == State::HardBoot) {                             // 250 ms = 7812 samples of silence, then the self tests
            if (++modeSampleCounter >= 7812) StartSelfTests();
            return 0;
        }
        if (state == State::Bong) {                                 // 750 ms per bong, one bong per count of the POST code
            if (++modeSampleCounter >= 23437) {
                if (--bongCount <= 0) SoftBoot();
                else { BongStart(); modeSampleCounter = 0; }
            }
            return BongSample();
        }
        if (standalone) return bufPos < buf.size() ? buf[bufPos++] : 0;
        if (!player) return 0;
        if (bufPos >= buf.size()) {
            // one frame at a time from the player, which renders `chunk` frames ahead on the GPU
            buf.resize(240);
            if (dcsb_player_render(player, 1u, buf.data()) != DCSB_OK) {
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
            // ROMPointer(0, data) without a size is what the reference's clients pass (DCSEncoder.cpp:553): the stream's
            // extent is then found the way GetStreamInfo finds it, by walking its frames (host side, lengths only; needs
            // DCSB_EXTENT_SLACK readable bytes behind the stream where the reference needs one)
            size_t nbytes = streamPtr.bytesAvailable;
            if (streamPtr.p && !nbytes) nbytes = dcsb_stream_extent(streamPtr.p, standaloneOS);
            if (streamPtr.p && nbytes) LoadAudioStream(channel, streamPtr.p, nbytes, mixingLevel);
            else Fail("dcsb200: not a decodable stream");
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
            if (!streamPtr.p || !IsOK()) return si;
            const size_t avail = streamPtr.bytesAvailable ? streamPtr.bytesAvailable : dcsb_stream_extent(streamPtr.p, standaloneOS);
            if (avail < 3) return si;
            dcsb_stream_desc d = dcsb_stream_desc();
            d.data = streamPtr.p;
            d.nbytes = static_cast<uint32_t>(avail);
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
            for (size_t k = 0; k < 16 && 2 + k < avail; ++k) si.header[k] = streamPtr.p[2 + k];
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
    enum class State { HardBoot, Bong, Running };                               // DCSDecoder.h (decoder state machine)
    State state = State::Running;
    int modeSampleCounter = 0, bongCount = 0;
    bool fastBootMode = false;                                                  // DCSDecoder.h:1290
    // startup bong (DCSDecoder.cpp:1697-1728): a square wave of about 195 Hz under an exponential decay envelope
    int bongEnvelopeSamples = 0, bongSignSamples = 0, bongSign = -1;
    uint16_t bongLevel = 0x0FFF;
    void BongStart() { bongEnvelopeSamples = 0; bongSignSamples = 0; bongLevel = 0x0FFF; }
    int16_t BongSample()
    {
        if (bongEnvelopeSamples++ >= 31) {                                      // about every millisecond: level *= 0.996 (1.15 fixed point)
            bongLevel = static_cast<uint16_t>(((static_cast<uint32_t>(bongLevel) * 0x7f80u) << 1) >> 16);
            bongEnvelopeSamples = 0;
        }
        if (bongSignSamples++ >= 80) { bongSign = -bongSign; bongSignSamples = 0; }
        return static_cast<int16_t>(bongSign * static_cast<int16_t>(bongLevel));
    }
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
