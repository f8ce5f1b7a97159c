// dcsb200 host control plane for ROM sets: the ROM model (chips, catalog, track index, version
// detection), track / stream lookup, and the sequencer that runs the track byte-code programs,
// the command queue, the mixer (levels, fades) and the per-frame gain staging -- everything
// DCSDecoderNative::MainLoop does EXCEPT touching audio bits.  Its product is a per-frame mix
// schedule {channel -> (stream, frame, effective multiplier)} + volume shift that the GPU
// renders (dcsb_mix.cuh).  No CUDA here: linked into libdcsb200.so and, for the CPU-side kernel
// simulator, into tests/hostsim.
//
// Reference behaviour followed (restated, not copied): DCSDecoder.cpp:26-76 (AddROM,
// MakeROMPointer), :207-234 (FindCatalog), :236-504 (CheckROMs), :622-651 (GetNumChannels),
// :671-884 (GetTrackInfo), :1248-1293 (ListStreams), :1734-1908 (SearchForOpcodes);
// DCSDecoderZipLoader.cpp:61-203; DCSDecoderNative.cpp:89-306 (MainLoop), :826-1228 (tracks),
// :1241-1371 (loops, mixing level ops), :1387-1463 (stream load), :1546-1589 (DecodeStream),
// :3042-3135 (UpdateMixingLevels), :3250-3282 (SetMasterVolume), :3297-3437 (IRQ2Handler).
#pragma once
#include <stdint.h>
#include <deque>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/dcsb200.h"

#include "dcsb_seq.cuh"      // DcsbRomPtr, DcsbRomView, the sequencer core (shared with the device build)

// what the sequencer needs to know about a stream without decoding it: its length and where
// (if anywhere) the decoder's error path stops the channel.  Filled from the GPU scan.
struct DcsbStreamFacts {
    uint32_t linear = 0;        // 24-bit ROM address
    DcsbRomPtr at;              // chip / offset of the U16 frame count
    uint16_t nframes = 0;       // frame count from the stream preamble
    int32_t status = 0;         // scan status (DCSB_OK, DCSB_E_STOPPED, DCSB_E_TRUNCATED, DCSB_E_BANDTYPE)
    uint32_t nplay = 0;         // frames that decode (the last one partially when status == DCSB_E_STOPPED)
    uint32_t nbytes = 0;        // stream size: count + header + frame bits, rounded up to a byte
};

struct DcsbZipEntry { std::string name; std::vector<uint8_t> data; };
struct dcsb_rom {
    std::vector<DcsbZipEntry> zip_files;    // the files of the zip last loaded (dcsb_rom_zip_files)
    std::vector<int> zip_chip;              // ... and the chip each one was taken for (-1: none)
    struct Chip {
        std::vector<uint8_t> bytes;     // image + 64 bytes of 0xFF slack
        uint32_t size = 0, mask = 0;
        bool present = false;
    } chip[8];
    uint32_t catalog_ofs = 0, track_index = 0, indirect_index = 0;     // offsets inside U2
    uint16_t n_tracks = 0;
    int hw = DCSB_HW_UNKNOWN;
    int os = 0;                         // 0 unknown, 1 invalid, else DCSB_OS93A / OS93B / OS94 / OS95
    uint16_t nominal_version = 0;
    bool totan = false;
    int post = 0;                       // last CheckROMs result (1 = all good, 2..9 = failing chip)
    std::string signature, err;
    // stream table (dcsb_rom_prepare): every stream a track program can start
    std::vector<DcsbStreamFacts> streams;
    std::unordered_map<uint32_t, uint32_t> stream_by_addr;
    struct dcsb_batch *batch = nullptr; // the streams resident in HBM (owned; dcsb_api.cu)
    std::vector<uint8_t> image;         // all chips back to back (what the device slab mirrors)
    uint32_t image_ofs[8] = { 0 };
    std::vector<DcsbSeqStream> seq_streams;     // `streams` as the sequencer core looks them up: sorted by address
    DcsbRomView view() const;           // the ROM set as the sequencer core reads it (host pointers)
    void build_seq_streams();           // seq_streams from streams (call when the scan results are in)
    // the same on the device (dcsb_player.cu; owned: freed with the batch)
    DcsbSeqStream *d_seq_streams = nullptr;

    void add(int n, const uint8_t *data, size_t size);
    int check();
    DcsbRomPtr make_ptr(uint32_t linear) const;
    uint8_t u8(const DcsbRomPtr &p, uint32_t d = 0) const;
    uint32_t be(const DcsbRomPtr &p, int nbytes, uint32_t d = 0) const;
    uint32_t u2_be(uint32_t ofs, int nbytes) const;
    int num_channels() const;
    bool track_info(uint16_t track, dcsb_track_info *ti) const;
    std::vector<dcsb_opcode> decompile_track(uint16_t track) const;
    std::vector<uint32_t> list_streams(bool as_executed = false) const;
    int search_opcodes(const char *pattern, uint32_t from, uint32_t nbytes, std::unordered_map<char, uint32_t> *vars) const;
    int opcode_operand_bytes(int opcode) const;     // as GetTrackInfo / the decompiler count them
};

// The decoder instance the player drives: the sequencer core (dcsb_seq.cuh) plus the host-side conveniences --
// host bytes kept with the frame they belong to, copyable as a whole (the render-ahead player snapshots it).
struct DcsbSequencer {
    explicit DcsbSequencer(const dcsb_rom *rom);
    void soft_boot();                               // Initialize(): channel defaults, default volume
    void set_master_volume(int vol);
    void write_port(uint8_t b);                     // WriteDataPort + IRQ2Handler (taken before the next frame)
    void add_track_command(uint16_t track) { dcsb_seq_q_push(st, track); }
    void load_stream(int ch, uint32_t linear, int level);   // LoadAudioStream(ch, ptr, level)
    void clear_tracks();
    bool stream_playing(int ch) const { return ch >= 0 && ch < DCSB_MAX_CHANNELS && st.chan[ch].st.active; }
    // run one main-loop pass; appends the frame to `frames` / `entries`.  Returns false once the
    // decoder is in its fatal-error state (the frame is then silent).
    bool frame(std::vector<DcsbSchedFrame> &frames, std::vector<DcsbSchedEntry> &entries);
    std::vector<uint8_t> host_bytes;                // bytes sent back to the host (ReceiveDataPort)
    std::vector<uint32_t> host_byte_frames;         // frame number of each
    bool fatal = false;
    uint32_t frame_no = 0;

private:
    const dcsb_rom *rom;
    DcsbSeqState st;
    void drain_host_bytes();
};

// zip container (stored / deflate entries) -> named files; false + err on failure
bool dcsb_unzip(const char *path, std::vector<DcsbZipEntry> &out, std::string &err);
int dcsb_rom_load_zip_impl(dcsb_rom *rom, const char *path, const char *explicit_u2);
