// DEV TOOL (not product, not test): statistics of the 1994+ frame walk on a corpus, used to
// choose the step scheme of the lock-step scan.  Reads blob.bin / offs.bin (a corpus dumped
// from corpus_cache/*.npz), walks every frame codeword by codeword, and reports the number of
// scan iterations per frame under several step schemes, alone and in lock-step groups.
//   g++ -O2 -std=c++17 -I dcsexplorer_b200/csrc tools/scan_stats.cpp dcsexplorer_b200/csrc/dcsb_host.cpp -o /tmp/st/scan_stats -lpthread
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "dcsb_internal.h"

static DcsbTables T;
static int cb_maxw(int k) { return k <= 2 ? k + 1 : (k == 3 ? 5 : k + 3); }
static int cb_ofs(int k) { return k == 1 ? 0 : k == 2 ? 4 : k == 3 ? 12 : k == 4 ? 44 : k == 5 ? 172 : 428; }
static int band_count(int b) { return b == 0 ? 7 : (b == 1 ? 8 : (b == 15 ? 32 : 16)); }

struct Bits {
    const uint8_t *p;
    size_t n;
    uint32_t peek(uint64_t pos, int k) const
    {
        uint64_t v = 0;
        size_t by = pos >> 3;
        for (int i = 0; i < 8; ++i) v = (v << 8) | (by + i < n ? p[by + i] : 0);
        v <<= (pos & 7);
        return (uint32_t)(v >> (64 - k));
    }
};

// one codeword of codebook k at pos: returns len, sets slots
static int cw(const Bits &b, uint64_t pos, int k, int &slots)
{
    const uint16_t e = T.lut[DCSB_LUT_CB + cb_ofs(k) + b.peek(pos, cb_maxw(k))];
    slots = (e & 0x800) ? 2 : 1;
    return e >> 12;
}

// scheme: greedy chain of whole codewords inside a peek of P bits, at most `cap` slots (first always taken)
static void chain(const Bits &b, uint64_t pos, int k, int P, int cap, int &bits, int &slots)
{
    bits = 0; slots = 0;
    for (;;) {
        int s;
        const int l = cw(b, pos + bits, k, s);
        if (l == 0) break;
        if (bits + l > P) break;
        if (slots && slots + s > cap) break;
        bits += l; slots += s;
        if (slots >= cap) break;
    }
    if (!slots) { int s; bits = cw(b, pos, k, s); slots = s; }
}

enum { NSCH = 10 };
static double cbsteps[7], cbbits[7], cbbands[7];
static const char *SCHN[NSCH] = { "A m8/m1 P13", "B m8/m4/m2/m1 P12", "C exact P13 cap8", "D exact P12 cap8", "E exact P15 cap8", "F m8/m4/m2/m1 P13", "G 13/12/12/9", "H 14/12/12/9", "I 14/13/13/9", "J P12 cap15 y8/y1" };

// steps to walk a Huffman band of `count` slots at pos under scheme sc; returns steps, sets end pos
static int band_steps(const Bits &b, uint64_t &pos, int k, int count, int sc)
{
    int rem = count, steps = 0;
    while (rem > 0) {
        int bits, slots;
        switch (sc) {
        case 0:
            chain(b, pos, k, 13, 8, bits, slots);
            if (slots > rem) chain(b, pos, k, 9, 1, bits, slots);
            break;
        case 1: case 5: {
            const int P = sc == 1 ? 12 : 13;
            const int cap = rem >= 8 ? 8 : rem >= 4 ? 4 : rem >= 2 ? 2 : 1;
            chain(b, pos, k, P, cap, bits, slots);
            break;
        }
        case 9:
            chain(b, pos, k, 12, 15, bits, slots);
            if (slots > rem) chain(b, pos, k, 9, 1, bits, slots);
            break;
        case 2: chain(b, pos, k, 13, rem < 8 ? rem : 8, bits, slots); break;
        case 3: chain(b, pos, k, 12, rem < 8 ? rem : 8, bits, slots); break;
        case 4: chain(b, pos, k, 15, rem < 8 ? rem : 8, bits, slots); break;
        case 6: case 7: case 8: {
            const int Pa = sc == 6 ? 13 : 14, Pb = sc == 8 ? 13 : 12;
            if (rem >= 8) chain(b, pos, k, Pa, 8, bits, slots);
            else if (rem >= 4) chain(b, pos, k, Pb, 4, bits, slots);
            else if (rem >= 2) chain(b, pos, k, Pb, 2, bits, slots);
            else chain(b, pos, k, 9, 1, bits, slots);
            break;
        }
        }
        pos += bits;
        rem -= slots;       // a two-zeros codeword with one slot left leaves rem < 0 (error case)
        ++steps;
    }
    return steps;
}

struct FrameStat { uint16_t hdr_iters, hdr_codes, bands, fixed; uint16_t steps[NSCH]; uint16_t bits; uint8_t bsteps[16]; };

int main(int argc, char **argv)
{
    const char *dir = argc > 1 ? argv[1] : "/tmp/st";
    char path[512];
    snprintf(path, sizeof path, "%s/offs.bin", dir);
    FILE *f = fopen(path, "rb");
    std::vector<int64_t> offs(1 << 20);
    offs.resize(fread(offs.data(), 8, offs.size(), f));
    fclose(f);
    snprintf(path, sizeof path, "%s/blob.bin", dir);
    f = fopen(path, "rb");
    std::vector<uint8_t> blob(offs.back() + 64);
    if (fread(blob.data(), 1, offs.back(), f) != (size_t)offs.back()) return 1;
    fclose(f);
    dcsb_build_tables(&T);
    const int ns = (int)offs.size() - 1;
    std::vector<std::vector<FrameStat>> st(ns);
    std::vector<double> bpf(ns);
    for (int si = 0; si < ns; ++si) {
        const uint8_t *d = blob.data() + offs[si];
        const size_t nbytes = offs[si + 1] - offs[si];
        const int nframes = (d[0] << 8) | d[1];
        const uint8_t *hdr = d + 2;
        const int type1 = hdr[0] >> 7;
        int nb = 0;
        while (nb < 16 && (hdr[nb] & 0x7F) != 0x7F) ++nb;
        Bits b{ d + 18, nbytes - 18 };
        uint64_t pos = 0;
        int bt[16] = { 0 };
        st[si].resize(nframes);
        for (int fr = 0; fr < nframes; ++fr) {
            FrameStat &fs = st[si][fr];
            memset(&fs, 0, sizeof fs);
            const uint64_t p0 = pos;
            // header: per iteration a run of 1-bits + the code behind it
            for (int bi = 0; bi < nb;) {
                ++fs.hdr_iters;
                int run = 0;
                while (bi < nb && b.peek(pos, 1)) { ++pos; ++bi; ++run; }
                if (bi >= nb) break;
                const uint32_t e = T.lut[DCSB_LUT_HDR94 + b.peek(pos, 8)];
                int delta;
                if (e == 0) {
                    int hit = -1;
                    for (int i = 0; i < T.n_long94 && hit < 0; ++i)
                        if (b.peek(pos, T.long94[i].len) == T.long94[i].code) hit = i;
                    if (hit < 0) { fprintf(stderr, "bad hdr code\n"); return 2; }
                    pos += T.long94[hit].len; delta = T.long94[hit].val - 0x2E;
                } else { pos += e >> 8; delta = (int)(e & 0xFF) - 0x2E; }
                bt[bi] += delta;
                ++fs.hdr_codes;
                ++bi;
            }
            for (int bi = 0; bi < nb; ++bi) {
                int count = band_count(bi);
                if (hdr[bi] & 0x40) count >>= 1;
                int code = bt[bi];
                if (type1) code = T.lut[DCSB_LUT_XLAT + (bi < 3 ? 0 : (bi < 6 ? 16 : 32)) + code] >> 8;
                if (code == 0 || count == 0) continue;
                ++fs.bands;
                if (code > 6) { ++fs.fixed; pos += (uint64_t)count * code; continue; }
                uint64_t pe = 0;
                for (int sc = 0; sc < NSCH; ++sc) {
                    uint64_t q = pos;
                    const int n_ = band_steps(b, q, code, count, sc);
                    fs.steps[sc] += n_;
                    if (sc == 0) { cbsteps[code] += n_; cbbits[code] += q - pos; cbbands[code] += 1; }
                    if (sc == 9) fs.bsteps[bi] = (uint8_t)n_;
                    if (sc == 0) pe = q; else if (q != pe) { fprintf(stderr, "scheme %d disagrees\n", sc); return 3; }
                }
                pos = pe;
            }
            fs.bits = (uint16_t)(pos - p0);
        }
        bpf[si] = nframes ? (double)pos / nframes : 0;
        if (((pos + 7) >> 3) + 18 > nbytes + 1) fprintf(stderr, "stream %d overruns\n", si);
    }
    // ---- per-stream averages
    double tb = 0, tf = 0, th = 0, thc = 0, tbands = 0, tfixed = 0, tsteps[NSCH] = { 0 };
    for (int si = 0; si < ns; ++si)
        for (auto &fs : st[si]) {
            tf += 1; tb += fs.bits; th += fs.hdr_iters; thc += fs.hdr_codes; tbands += fs.bands; tfixed += fs.fixed;
            for (int sc = 0; sc < NSCH; ++sc) tsteps[sc] += fs.steps[sc];
        }
    printf("streams %d frames %.0f bits/frame %.1f hdr iters %.2f hdr codes %.2f bands %.2f (fixed %.2f)\n", ns, tf, tb / tf, th / tf, thc / tf,
           tbands / tf, tfixed / tf);
    for (int sc = 0; sc < NSCH; ++sc) printf("  scheme %-22s steps/frame %.1f\n", SCHN[sc], tsteps[sc] / tf);
    for (int k = 1; k <= 6; ++k) printf("  codebook %d: bands/frame %.2f steps/frame %.1f bits/frame %.1f bits/step %.1f\n", k, cbbands[k] / tf, cbsteps[k] / tf, cbbits[k] / tf, cbbits[k] / cbsteps[k]);
    // ---- lock-step groups: streams sorted by bits per frame, L lanes per warp; iterations per frame =
    //      max over lanes (merged: header iterations + band switches folded + steps)
    std::vector<int> ord(ns);
    for (int i = 0; i < ns; ++i) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](int a, int b2) { return bpf[a] < bpf[b2]; });
    for (int L : { 32, 16, 8, 4 })
        for (int sc = 0; sc < NSCH; ++sc) {
            double worst = 0, sum = 0;
            int nw = 0;
            for (int w0 = 0; w0 + L <= ns; w0 += L, ++nw) {
                size_t nf = st[ord[w0]].size();
                for (int l = 1; l < L; ++l) nf = std::max(nf, st[ord[w0 + l]].size());
                double it = 0;
                for (size_t fr = 0; fr < nf; ++fr) {
                    int mh = 0, ms = 0;
                    for (int l = 0; l < L; ++l) {
                        const auto &v = st[ord[w0 + l]];
                        if (fr >= v.size()) continue;
                        mh = std::max<int>(mh, v[fr].hdr_iters);
                        ms = std::max<int>(ms, v[fr].steps[sc] + v[fr].fixed);
                    }
                    it += mh + ms;
                }
                worst = std::max(worst, it);
                sum += it;
            }
            if (sc == 9) {
                double worstb = 0, sumb = 0; int nwb = 0;
                for (int w0 = 0; w0 + L <= ns; w0 += L, ++nwb) {
                    size_t nf = 0;
                    for (int l = 0; l < L; ++l) nf = std::max(nf, st[ord[w0 + l]].size());
                    double it = 0;
                    for (size_t fr = 0; fr < nf; ++fr) {
                        int mh = 0;
                        for (int bb = 0; bb < 16; ++bb) {
                            int mb = 0;
                            for (int l = 0; l < L; ++l) { const auto &v = st[ord[w0 + l]]; if (fr < v.size()) mb = std::max<int>(mb, v[fr].bsteps[bb]); }
                            it += mb;
                        }
                        for (int l = 0; l < L; ++l) { const auto &v = st[ord[w0 + l]]; if (fr < v.size()) mh = std::max<int>(mh, v[fr].hdr_iters); }
                        it += mh;
                    }
                    worstb = std::max(worstb, it); sumb += it;
                }
                printf("L=%2d band-synchronous (scheme J): mean %.0f worst %.0f (per frame %.1f / %.1f)\n", L, sumb / nwb, worstb, sumb / nwb / 1303, worstb / 1303);
            }
            printf("L=%2d %-22s iterations per warp: mean %.0f worst %.0f  (per frame of 1303: %.1f / %.1f)\n", L, SCHN[sc], sum / nw, worst,
                   sum / nw / 1303, worst / 1303);
        }
    return 0;
}
