// TEST: a client of the REFERENCE's decoder interface (the reference's own DCSDecoder.o, compiled
// unmodified into oracle/_ref) that picks its decoder implementation by registry name the way
// DCSExplorer does (DCSExplorer.cpp:459-537), loads ROM images with AddROM, plays a data-port
// timeline and dumps the PCM that GetNextSample() returns.  Run with "native" it is the
// reference; run with "b200" it is dcsexplorer_b200/plugin/DCSDecoderB200Plugin.cpp behind the
// same base class -- the test compares the two outputs bit for bit.
//   ref_plugin_host --list
//   ref_plugin_host <decoder> <timeline.txt> <n_frames> <master_volume> <out.pcm> <chip>=<image file> ...
// Exit code 3 = the decoder did not initialise (b200 without a GPU: no CPU fallback).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <memory>
#include <vector>
#include "DCSDecoder.h"

struct RecHost : DCSDecoder::Host {
    std::vector<uint8_t> bytes;
    void ReceiveDataPort(uint8_t d) override { bytes.push_back(d); }
    void ClearDataPort() override {}
    void BootTimerControl(bool) override {}
};

static std::vector<uint8_t> slurp(const char *path)
{
    std::vector<uint8_t> v;
    if (FILE *f = fopen(path, "rb")) {
        fseek(f, 0, SEEK_END);
        v.resize((size_t)ftell(f));
        fseek(f, 0, SEEK_SET);
        if (fread(v.data(), 1, v.size(), f) != v.size()) v.clear();
        fclose(f);
    }
    return v;
}

int main(int argc, char **argv)
{
    if (argc >= 2 && !strcmp(argv[1], "--list")) {
        for (auto &r : DCSDecoder::GetRegistrationMap()) printf("%s\t%s\n", r.second.name, r.second.desc);
        return 0;
    }
    if (argc < 7) { fprintf(stderr, "usage: %s decoder timeline.txt n_frames volume out.pcm chip=file...\n", argv[0]); return 2; }
    RecHost host;
    auto &map = DCSDecoder::GetRegistrationMap();
    auto it = map.find(argv[1]);
    if (it == map.end()) { fprintf(stderr, "no decoder named '%s' is registered\n", argv[1]); return 2; }
    std::unique_ptr<DCSDecoder> dec(it->second.factory(&host));
    std::vector<std::vector<uint8_t>> images;
    for (int i = 6; i < argc; ++i) {
        const char *eq = strchr(argv[i], '=');
        if (!eq) continue;
        images.push_back(slurp(eq + 1));
        if (images.back().empty()) { fprintf(stderr, "cannot read %s\n", eq + 1); return 4; }
        dec->AddROM(atoi(argv[i]), images.back().data(), images.back().size());
    }
    if (dec->CheckROMs() != 1) { fprintf(stderr, "ROM check failed\n"); return 5; }
    dec->SoftBoot();
    if (!dec->IsOK()) { fprintf(stderr, "%s: %s\n", dec->Name(), dec->GetErrorMessage().c_str()); return 3; }
    dec->SetMasterVolume(atoi(argv[4]));
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
        while (w < writes.size() && writes[w].first <= fr) dec->WriteDataPort((uint8_t)writes[w++].second);
        for (int i = 0; i < 240; ++i) pcm[(size_t)fr * 240 + i] = dec->GetNextSample();
    }
    if (!dec->IsOK()) { fprintf(stderr, "%s: %s\n", dec->Name(), dec->GetErrorMessage().c_str()); return 6; }
    FILE *o = fopen(argv[5], "wb");
    if (!o) return 7;
    fwrite(pcm.data(), 2, pcm.size(), o);
    fclose(o);
    printf("%s | %s | tracks 0..%u | %d channels | %zu streams | host bytes", dec->Name(), dec->GetVersionInfo().c_str(),
           dec->GetMaxTrackNumber(), dec->GetNumChannels(), dec->ListStreams().size());
    for (uint8_t b : host.bytes) printf(" %02x", b);
    printf("\n");
    return 0;
}
