// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Command-line driver around the UNMODIFIED reference ROM compiler (DCSEncoder/DCSCompiler.cpp,
// DCSTokenizer.cpp, DCSEncoder.cpp), built by oracle/Makefile into oracle/_ref/ref_dcscompile.
// It does what the reference's own DCSEncoder/Main.cpp does with its three positional inputs
// (Main.cpp:154-250): LoadPrototypeROM(proto zip) -> ParseScript(script) -> GenerateROM(out zip),
// and stands in for the two pieces of the reference's build that cannot be compiled here:
//   * DCSEncoder::EncodeFile (DCSEncodeFile.cpp) needs libnyquist for MP3/Ogg/FLAC; this one
//     takes raw .dcs streams (IsDCSFile -> EncodeDCSFile, as the reference does, :50-51) and
//     16-bit PCM WAV files, fed to the encoder as floats through OpenStream / WriteStream(float) /
//     CloseStream exactly like the rest of that function (:72-103);
//   * OSInit (OSSpecificWin32.cpp) is Windows console set-up: empty here.
// It is used by tests/golden/make_compiled_rom_golden.py to build BASELINE config 4's "ROM image
// built by DCSCompiler"; the resulting ROM images are committed as a fixture.
//   ref_dcscompile <prototype.zip> <script> <out.zip> [rom-size-bytes]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <memory>
#include <string>
#include <vector>
#include <list>
#include "DCSEncoder.h"
#include "DCSCompiler.h"
#include "DCSTokenizer.h"

void OSInit() {}

static bool ReadWav16(const char *filename, std::vector<float> &mono, int &rate, std::string &err)
{
    FILE *f = fopen(filename, "rb");
    if (!f) { err = std::string("cannot open ") + filename; return false; }
    std::vector<uint8_t> b;
    fseek(f, 0, SEEK_END);
    b.resize((size_t)ftell(f));
    fseek(f, 0, SEEK_SET);
    const bool ok = fread(b.data(), 1, b.size(), f) == b.size();
    fclose(f);
    if (!ok || b.size() < 44 || memcmp(&b[0], "RIFF", 4) || memcmp(&b[8], "WAVE", 4)) { err = "not a WAV file"; return false; }
    int channels = 0, bits = 0;
    rate = 0;
    for (size_t p = 12; p + 8 <= b.size();) {
        const uint32_t len = b[p + 4] | (b[p + 5] << 8) | (b[p + 6] << 16) | ((uint32_t)b[p + 7] << 24);
        if (!memcmp(&b[p], "fmt ", 4) && len >= 16) {
            channels = b[p + 10] | (b[p + 11] << 8);
            rate = b[p + 12] | (b[p + 13] << 8) | (b[p + 14] << 16) | (b[p + 15] << 24);
            bits = b[p + 22] | (b[p + 23] << 8);
        } else if (!memcmp(&b[p], "data", 4)) {
            if (bits != 16 || (channels != 1 && channels != 2)) { err = "only 16-bit mono/stereo PCM WAV"; return false; }
            const size_t n = std::min<size_t>(len, b.size() - p - 8) / 2;
            for (size_t i = 0; i + channels <= n; i += channels) {
                float s = 0;
                for (int c = 0; c < channels; ++c) s += (int16_t)(b[p + 8 + 2 * (i + c)] | (b[p + 9 + 2 * (i + c)] << 8)) / 32768.0f;
                mono.push_back(s / channels);
            }
            return true;
        }
        p += 8 + len + (len & 1);
    }
    err = "no data chunk";
    return false;
}

bool DCSEncoder::EncodeFile(const char *filename, DCSAudio &dcsObj, std::string &errorMessage, OpenStreamStatus *statusPtr)
{
    auto Status = [&](OpenStreamStatus s) { if (statusPtr) *statusPtr = s; return s == OpenStreamStatus::OK; };
    if (IsDCSFile(filename))
        return EncodeDCSFile(filename, dcsObj, errorMessage);
    std::vector<float> mono;
    int rate = 0;
    if (!ReadWav16(filename, mono, rate, errorMessage)) return Status(OpenStreamStatus::Error);
    std::unique_ptr<Stream> stream(OpenStream(rate, errorMessage));
    if (stream == nullptr) return Status(OpenStreamStatus::Error);
    for (size_t p = 0; p < mono.size(); p += 256)
        WriteStream(stream.get(), mono.data() + p, std::min<size_t>(256, mono.size() - p));
    if (!CloseStream(stream.get(), dcsObj, errorMessage)) return Status(OpenStreamStatus::Error);
    return Status(OpenStreamStatus::OK);
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s prototype.zip script out.zip [rom-size]\n", argv[0]); return 1; }
    DCSCompiler compiler;
    std::string err;
    if (!compiler.LoadPrototypeROM(argv[1], false, err)) { fprintf(stderr, "prototype ROM: %s\n", err.c_str()); return 2; }
    struct Logger : DCSTokenizer::ErrorLogger {
        void Status(const char *, bool) override {}
    } logger;
    compiler.ParseScript(argv[2], logger);
    if (logger.errors != 0 || logger.fatal != 0) { fprintf(stderr, "script: %d errors\n", logger.errors + logger.fatal); return 3; }
    std::list<DCSCompiler::ROMDesc> roms;
    const uint32_t romSize = argc > 4 ? (uint32_t)strtoul(argv[4], nullptr, 0) : 1024u * 1024u;
    if (!compiler.GenerateROM(argv[3], romSize, "snd_", err, &roms)) { fprintf(stderr, "GenerateROM: %s\n", err.c_str()); return 4; }
    for (auto &r : roms) printf("U%d %u bytes, %u free, %s\n", r.chipNum, (unsigned)r.size, (unsigned)r.bytesFree, r.filename.c_str());
    return 0;
}
