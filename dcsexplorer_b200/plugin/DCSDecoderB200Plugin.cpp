// DCSDecoderB200Plugin -- the reference-side binding: a DCSDecoder subclass, registered in the
// reference's decoder registry as "b200", that renders on the GPU through libdcsb200.so.
//
// This is the one file a maintainer of mjrgh/DCSExplorer adds to the DCSDecoder project to get
// `--decoder=b200` (DCSExplorer/DCSExplorer.cpp:459-537 picks implementations by registry name).
// It is compiled against the reference's OWN DCSDecoder.h (include path given at build time;
// the header is not part of this repo) and linked with the reference's DCSDecoder.o and
// -ldcsb200.  It implements exactly the subclass contract of the abstract class:
//     Name()            DCSDecoder.h:210
//     SetMasterVolume() DCSDecoder.h:546
//     Initialize()      DCSDecoder.h:1137   called by SoftBoot (DCSDecoder.cpp:1518-1542)
//     IRQ2Handler()     DCSDecoder.h:1140   one data-port byte (DCSDecoder.cpp:1625-1626)
//     MainLoop()        DCSDecoder.h:1143   one 240-sample frame into the autobuffer
// and self-registers like DCSDecoderNative does (DCSDecoderNative.cpp:18-19, DCSDecoder.h:
// 1115-1128).  Everything else the clients call (AddROM, LoadROMFromZipFile, CheckROMs,
// GetTrackInfo, ListStreams, WriteDataPort, GetNextSample, HardBoot, ...) stays the base
// class's own code, so a client cannot tell the two implementations apart except by Name().
//
// There is no CPU fallback: without a usable sm_100 device Initialize() fails, the base class
// enters State::InitializationError (IsOK() == false) and GetErrorMessage() says why.
#include <stdint.h>
#include <string>
#include "DCSDecoder.h"             // the reference's header
#include "DCSDecoderB200.h"         // include/DCSDecoderB200.h of this repo

class DCSDecoderB200Plugin : public DCSDecoder
{
public:
    explicit DCSDecoderB200Plugin(Host *host, int cudaDevice = 0)
        : DCSDecoder(host), impl(&fwd, cudaDevice, /*chunkFrames: rendered ahead, inputs still land on their frame*/ 16)
    {
        fwd.host = host;
        for (auto &s : frame) s = 0;
    }
    const char *Name() const override { return "b200"; }
    void SetMasterVolume(int vol) override { impl.SetMasterVolume(vol); }

protected:
    bool Initialize() override
    {
        if (!impl.IsOK()) { errorMessage = impl.GetErrorMessage(); return false; }
        // hand over the images the base class holds (AddROM / LoadROMFromZipFile stored them in
        // ROM[], DCSDecoder.h:669-724); CheckROMs has filled the empty slots with dummies
        for (int i = 0; i < 8; ++i)
            if (ROM[i].data != nullptr && !ROM[i].isDummy) impl.AddROM(i + 2, ROM[i].data, ROM[i].size);
        if (impl.CheckROMs() != 1) { errorMessage = "dcsb200: the ROM set failed the power-on checks"; return false; }
        impl.SetDefaultVolume(defaultVolume);           // what the decoder applies after a soft reset (DCSDecoder.h:1146)
        impl.SoftBoot();
        if (!impl.IsOK()) { errorMessage = impl.GetErrorMessage(); return false; }
        // GetNextSample drains length/2 = 240 samples per MainLoop pass (DCSDecoder.cpp:1629-1677),
        // the same autobuffer shape DCSDecoderNative sets up (DCSDecoderNative.cpp:3203)
        autobuffer.Set(frame, 0x1E0, 1);
        return true;
    }
    void IRQ2Handler() override { impl.WriteDataPort(ReadDataPort()); }
    void MainLoop() override
    {
        if (!impl.GetFrames(1, reinterpret_cast<int16_t *>(frame))) {
            // a CUDA failure mid-stream: silence from here on, reported like a decoder fatal error
            for (auto &s : frame) s = 0;
            errorMessage = impl.GetErrorMessage();
            state = State::DecoderFatalError;
        }
    }

private:
    struct Fwd : DCSDecoderB200::Host {
        DCSDecoder::Host *host = nullptr;
        void ReceiveDataPort(uint8_t b) override { if (host) host->ReceiveDataPort(b); }
        void BootTimerControl(bool set) override { if (host) host->BootTimerControl(set); }
    } fwd;
    DCSDecoderB200 impl;
    uint16_t frame[0x1E0];
};

static DCSDecoder::Registration registrationB200("b200", "B200 batch decoder (CUDA sm_100a, libdcsb200)",
    [](DCSDecoder::Host *host) -> DCSDecoder * { return new DCSDecoderB200Plugin(host); });
