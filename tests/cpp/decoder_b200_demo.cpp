// TEST: drives the DCSDecoder-compatible C++ front (include/DCSDecoderB200.h) the way a client
// of the reference drives DCSDecoderNative (DCSExplorer.cpp:1308-1341): load a ROM zip, soft-boot,
// write data-port bytes at given frames, pull samples one at a time, dump raw PCM.
//   decoder_b200_demo <rom.zip> <timeline.txt> <n_frames> <master_volume> <chunk_frames> <out.pcm>
// timeline.txt: lines "<frame> <byte>" sorted by frame.  Exit code 3 = no usable GPU.
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include "../../include/DCSDecoderB200.h"

struct RecHost : DCSDecoderB200::Host {
    std::vector<uint8_t> bytes;
    void ReceiveDataPort(uint8_t b) override { bytes.push_back(b); }
};

// standalone mode, the reference's stream-extraction protocol (DCSExplorer.cpp:1655-1721):
//   decoder_b200_demo --standalone <os: 9301|9302|9400|9500> <stream.bin> <master_volume> <mixing_level> <n_frames> <out.pcm>
static int standalone_main(int argc, char **argv)
{
    if (argc < 8) return 2;
    DCSDecoderB200 dec(nullptr, 0, 1);
    if (!dec.IsOK()) { fprintf(stderr, "%s\n", dec.GetErrorMessage().c_str()); return 3; }
    std::vector<uint8_t> data;
    if (FILE *f = fopen(argv[3], "rb")) { int c; while ((c = fgetc(f)) != EOF) data.push_back((uint8_t)c); fclose(f); }
    const unsigned os = (unsigned)strtoul(argv[2], nullptr, 16);
    dec.InitStandalone(os == 0x9301 ? DCSDecoderB200::OSVersion::OS93a : os == 0x9302 ? DCSDecoderB200::OSVersion::OS93b
                     : os == 0x9500 ? DCSDecoderB200::OSVersion::OS95 : DCSDecoderB200::OSVersion::OS94);
    dec.SoftBoot();
    dec.SetMasterVolume(atoi(argv[4]));
    // "unsized" as a ninth argument: the pointer the reference's own clients pass (DCSEncoder.cpp:553), no size
    const bool unsized = argc > 8 && std::string(argv[8]) == "unsized";
    data.resize(data.size() + DCSB_EXTENT_SLACK, 0);
    const DCSDecoderB200::ROMPointer rp = unsized ? DCSDecoderB200::ROMPointer(0, data.data())
                                                  : DCSDecoderB200::ROMPointer(0, data.data(), data.size() - DCSB_EXTENT_SLACK);
    dec.LoadAudioStream(0, rp, atoi(argv[5]));
    if (!dec.IsOK()) { fprintf(stderr, "%s\n", dec.GetErrorMessage().c_str()); return 6; }
    const unsigned nframes = (unsigned)atoi(argv[6]);
    std::vector<int16_t> pcm((size_t)nframes * 240);
    unsigned playing = 0;
    for (size_t i = 0; i < pcm.size(); ++i) { playing += dec.IsStreamPlaying(0) ? 1 : 0; pcm[i] = dec.GetNextSample(); }
    FILE *o = fopen(argv[7], "wb");
    fwrite(pcm.data(), 2, pcm.size(), o);
    fclose(o);
    auto si = dec.GetStreamInfo(rp);
    printf("standalone: %d frames, %d bytes, type %d.%d, playing for %u samples\n", si.nFrames, si.nBytes, si.streamType, si.streamSubType, playing);
    return 0;
}

// the boot sequence (DCSDecoder.cpp:1233-1246, :1477-1516, :1579-1619):
//   decoder_b200_demo --hardboot <rom.zip> <fast: 0|1> <port write at sample, -1 = none> <n_samples> <out.pcm>
static int hardboot_main(int argc, char **argv)
{
    if (argc < 7) return 2;
    RecHost host;
    DCSDecoderB200 dec(&host, 0, 1);
    if (!dec.IsOK()) { fprintf(stderr, "%s\n", dec.GetErrorMessage().c_str()); return 3; }
    std::list<DCSDecoderB200::ZipFileData> files;
    std::string err;
    if (dec.LoadROMFromZipFile(argv[2], files, nullptr, &err) != DCSDecoderB200::ZipLoadStatus::Success) { fprintf(stderr, "%s\n", err.c_str()); return 4; }
    dec.SetFastBootMode(atoi(argv[3]) != 0);
    dec.HardBoot();
    const long at = atol(argv[4]), n = atol(argv[5]);
    std::vector<int16_t> pcm((size_t)n);
    for (long i = 0; i < n; ++i) {
        if (i == at) dec.WriteDataPort(0x55);
        pcm[(size_t)i] = dec.GetNextSample();
    }
    FILE *o = fopen(argv[6], "wb");
    fwrite(pcm.data(), 2, pcm.size(), o);
    fclose(o);
    printf("zip files:");
    for (auto &f : files) printf(" %s=%d(%zu)", f.filename.c_str(), f.chipNum, f.dataSize);
    printf(" | host bytes");
    for (uint8_t b : host.bytes) printf(" %02x", b);
    printf(" | running %d\n", (int)dec.IsRunning());
    return 0;
}

int main(int argc, char **argv)
{
    if (argc > 1 && std::string(argv[1]) == "--standalone") return standalone_main(argc, argv);
    if (argc > 1 && std::string(argv[1]) == "--hardboot") return hardboot_main(argc, argv);
    if (argc < 7) { fprintf(stderr, "usage: %s rom.zip timeline.txt n_frames volume chunk out.pcm\n", argv[0]); return 2; }
    RecHost host;
    DCSDecoderB200 dec(&host, 0, atoi(argv[5]));
    if (!dec.IsOK()) { fprintf(stderr, "%s\n", dec.GetErrorMessage().c_str()); return 3; }
    std::string err;
    if (dec.LoadROMFromZipFile(argv[1], nullptr, &err) != DCSDecoderB200::ZipLoadStatus::Success) { fprintf(stderr, "%s\n", err.c_str()); return 4; }
    if (dec.CheckROMs() != 1) { fprintf(stderr, "ROM check failed\n"); return 5; }
    dec.SoftBoot();
    if (!dec.IsOK()) { fprintf(stderr, "%s\n", dec.GetErrorMessage().c_str()); return 6; }
    dec.SetMasterVolume(atoi(argv[4]));
    std::vector<std::pair<unsigned, unsigned>> writes;
    if (FILE *f = fopen(argv[2], "r")) {
        unsigned fr, b;
        while (fscanf(f, "%u %u", &fr, &b) == 2) writes.emplace_back(fr, b);
        fclose(f);
    }
    const unsigned nframes = (unsigned)atoi(argv[3]);
    std::vector<int16_t> pcm((size_t)nframes * 240);
    size_t w = 0;
    for (unsigned fr = 0; fr < nframes; ++fr) {
        while (w < writes.size() && writes[w].first <= fr) dec.WriteDataPort((uint8_t)writes[w++].second);
        for (int i = 0; i < 240; ++i) pcm[(size_t)fr * 240 + i] = dec.GetNextSample();
    }
    FILE *o = fopen(argv[6], "wb");
    fwrite(pcm.data(), 2, pcm.size(), o);
    fclose(o);
    DCSDecoderB200::HWVersion hw; DCSDecoderB200::OSVersion os;
    printf("%s | tracks 0..%u | %d channels | %zu streams | host bytes", dec.GetVersionInfo(&hw, &os).c_str(), dec.GetMaxTrackNumber(),
           dec.GetNumChannels(), dec.ListStreams().size());
    for (uint8_t b : host.bytes) printf(" %02x", b);
    printf("\n");
    DCSDecoderB200::TrackInfo ti;
    if (dec.GetTrackInfo(0, ti)) {
        auto rp = dec.MakeROMPointer(ti.address);
        printf("track 0: type %d channel %d time %u looping %d first byte %02x\n", ti.type, ti.channel, ti.time, (int)ti.looping, rp.p ? rp.p[0] : 0);
    }
    printf("%s\n", dec.ExplainTrackProgram(0, "  ").c_str());
    printf("track 0: %zu steps\n", dec.DecompileTrackProgram(0).size());
    for (uint32_t a : dec.ListStreams()) {
        auto si = dec.GetStreamInfo(dec.MakeROMPointer(a));
        printf("stream $%06x: %d frames, %d bytes, type %d.%d\n", a, si.nFrames, si.nBytes, si.streamType, si.streamSubType);
        break;
    }
    return 0;
}
