// dcsb200 host control plane for ROM sets -- see dcsb_rom.h for the map to the reference.
#include <ctype.h>
#include <stdio.h>
#include <string.h>
#include <strings.h>
#include <algorithm>
#include <set>
#include <zlib.h>
#include "dcsb_rom.h"

// ======================================================================================
// ROM model
static const uint32_t kCatalogOffsets[3] = { 0x3000, 0x4000, 0x6000 };

static bool is_jump(const uint8_t *p) { return (p[0] & 0xFC) == 0x18 && (p[2] & 0x0F) == 0x0F; }   // ADSP-2105 JUMP

static std::string read_signature(const uint8_t *u2, size_t size)
{
    // JUMP at offset 0, printable text from offset 4 up to a NUL (DCSDecoder.cpp:100-127)
    if (size < 8 || !is_jump(u2)) return "";
    size_t len = 0;
    while (len < 120 && 4 + len < size && u2[4 + len] >= 32 && u2[4 + len] < 127) ++len;
    if (4 + len >= size || u2[4 + len] != 0) return "";
    return std::string(reinterpret_cast<const char *>(u2 + 4), len);
}

static bool contains_nocase(const std::string &hay, const char *needle)
{
    const size_t n = strlen(needle);
    for (size_t i = 0; i + n <= hay.size(); ++i)
        if (strncasecmp(hay.c_str() + i, needle, n) == 0) return true;
    return false;
}

void dcsb_rom::add(int n, const uint8_t *data, size_t size)
{
    if (n < 2 || n > 9 || size == 0 || !data) return;
    Chip &c = chip[n - 2];
    c.bytes.assign(data, data + size);
    c.bytes.resize(size + 64, 0xFF);
    c.size = (uint32_t)size;
    c.mask = (uint32_t)(size - 1);
    c.present = true;
    batch = nullptr;        // (an already prepared stream table belongs to the old images; dcsb_api.cu frees it)
    if (n != 2) return;
    // the catalog: first entry describes U2 itself -- its size in 4 KB units, chip select 0,
    // checksum 0 (DCSDecoder.cpp:207-234)
    catalog_ofs = 0;
    for (uint32_t ofs : kCatalogOffsets) {
        if (ofs + 0x48 > size) continue;
        const uint32_t sz = u2_be(ofs, 2) * 4096u, sel = u2_be(ofs + 2, 2) >> 8, ck = u2_be(ofs + 4, 2);
        if (sel == 0 && ck == 0 && sz == size) { catalog_ofs = ofs; break; }
    }
    track_index = indirect_index = 0;
    n_tracks = 0;
    if (catalog_ofs) {
        track_index = u2_be(catalog_ofs + 0x40, 3);
        indirect_index = u2_be(catalog_ofs + 0x43, 3);
        n_tracks = (uint16_t)u2_be(catalog_ofs + 0x46, 2);
    }
    signature = read_signature(data, size);
    totan = contains_nocase(signature, "Arabian Nights");       // game-specific data port quirk (DCSDecoderNative.cpp:3345)
}

uint32_t dcsb_rom::u2_be(uint32_t ofs, int nbytes) const
{
    uint32_t v = 0;
    for (int i = 0; i < nbytes; ++i) v = (v << 8) | (ofs + i < chip[0].size ? chip[0].bytes[ofs + i] : 0xFFu);
    return v;
}

DcsbRomPtr dcsb_rom::make_ptr(uint32_t linear) const
{
    // chip select in bits 21-23 on the DCS-95 board, 20-22 on the original one (DCSDecoder.cpp:67-76)
    DcsbRomPtr p;
    p.chip = (int)((linear >> (hw == DCSB_HW_DCS95 ? 21 : 20)) & 7);
    p.ofs = linear & (chip[p.chip].present ? chip[p.chip].mask : 0x1FFFu);     // absent chips read as 8 KB of $FF
    return p;
}

uint8_t dcsb_rom::u8(const DcsbRomPtr &p, uint32_t d) const
{
    if (p.chip < 0) return 0xFF;
    const Chip &c = chip[p.chip];
    const uint64_t o = (uint64_t)p.ofs + d;
    return (c.present && o < c.size) ? c.bytes[o] : 0xFF;
}

uint32_t dcsb_rom::be(const DcsbRomPtr &p, int nbytes, uint32_t d) const
{
    uint32_t v = 0;
    for (int i = 0; i < nbytes; ++i) v = (v << 8) | u8(p, d + i);
    return v;
}

// Opcode pattern search over U2: 24-bit big-endian opcodes in 4-byte slots; pattern digits are
// literal hex, '*' is a wildcard nibble, any other letter names a variable (DCSDecoder.cpp:1734-1908)
int dcsb_rom::search_opcodes(const char *pattern, uint32_t from, uint32_t nbytes, std::unordered_map<char, uint32_t> *vars) const
{
    struct Op { uint32_t value, mask; };
    struct Var { char name; int op; int shift; uint32_t mask; };
    std::vector<Op> ops;
    std::vector<Var> vlist;
    for (const char *p = pattern; *p;) {
        while (*p == ' ') ++p;
        if (!*p) break;
        Op op{ 0, 0 };
        Var cur{ 0, 0, 0, 0 };
        for (int i = 0; i < 6 && *p && *p != ' '; ++i, ++p) {
            const char c = *p;
            const bool hex = isdigit((unsigned char)c) || (c >= 'a' && c <= 'f') || (c >= 'A' && c <= 'F');
            if (hex || c == '*') {
                const uint32_t d = !hex ? 0 : (uint32_t)(isdigit((unsigned char)c) ? c - '0' : (tolower(c) - 'a' + 10));
                op.value = (op.value << 4) | d;
                op.mask = (op.mask << 4) | (hex ? 0xFu : 0u);
                if (cur.name) { vlist.push_back(cur); cur = Var{ 0, 0, 0, 0 }; }
            } else {
                if (cur.name && cur.name != c) { vlist.push_back(cur); cur = Var{ 0, 0, 0, 0 }; }
                cur.name = c;
                cur.op = (int)ops.size();
                cur.shift = 20 - 4 * i;
                cur.mask = (cur.mask << 4) | 0xFu;
                op.value <<= 4;
                op.mask <<= 4;
            }
        }
        if (cur.name) vlist.push_back(cur);
        ops.push_back(op);
    }
    const uint32_t nops = nbytes / 4;
    auto fetch = [&](uint32_t i) { return u2_be(from + 4 * i, 3); };
    for (uint32_t a = 0; a + ops.size() < nops; ++a) {
        bool ok = true;
        for (size_t k = 0; k < ops.size() && ok; ++k) ok = (fetch(a + (uint32_t)k) & ops[k].mask) == ops[k].value;
        if (!ok) continue;
        if (vars)
            for (const Var &v : vlist) (*vars)[v.name] = (fetch(a + (uint32_t)v.op) >> v.shift) & v.mask;
        return (int)(a * 4);
    }
    return -1;
}

static uint16_t chip_checksum(const uint8_t *p, size_t n)
{
    // sum of even-offset bytes in the high byte, of odd-offset bytes in the low byte (DCSDecoder.cpp:653-669)
    unsigned even = 0, odd = 0;
    for (size_t i = 0; i + 1 < n + 1 && i < n; i += 2) { even += p[i]; if (i + 1 < n) odd += p[i + 1]; }
    return (uint16_t)(((even << 8) & 0xFF00u) | (odd & 0xFFu));
}

int dcsb_rom::check()
{
    hw = DCSB_HW_INVALID;
    os = 1;
    nominal_version = 0;
    if (!chip[0].present) return post = 2;
    uint16_t sums[8] = { 0 };
    int populated = 0;
    for (int i = 0; i < 8; ++i)
        if (chip[i].present) { sums[i] = chip_checksum(chip[i].bytes.data(), chip[i].size); ++populated; }
    for (uint32_t ofs : kCatalogOffsets) {
        if (ofs + 9 * 6 > chip[0].size) continue;
        int in_table = 0, validated = 0, first_failed = -1;
        for (int e = 0; e < 9; ++e) {
            const uint32_t size = u2_be(ofs + 6 * e, 2) * 4096u;
            uint32_t sel = u2_be(ofs + 6 * e + 2, 2) >> 8;
            const uint32_t ck = u2_be(ofs + 6 * e + 4, 2);
            if (size == 0) break;
            ++in_table;
            if (ofs == 0x6000) sel >>= 1;                       // DCS-95 bank numbers carry one more bit
            if (sel < 8 && chip[sel].present && chip[sel].size == size && sums[sel] == ck) ++validated;
            else { first_failed = e; break; }
        }
        if (validated == 0) continue;
        if (ofs == 0x6000) {
            hw = DCSB_HW_DCS95;
            os = DCSB_OS95;
            std::unordered_map<char, uint32_t> v;
            if (search_opcodes("4vvvvE 0F16F8 93300E 18***F 4wwwwE 0F1608 0F16F8 93300E 18***F", 0x2000 + 0x300 * 4, 0x180 * 4, &v) >= 0)
                nominal_version = (uint16_t)v['v'];
        } else {
            hw = DCSB_HW_DCS93;
            os = DCSB_OS94;
            if (search_opcodes("380026 3C1005 0C00C0", 0x1000 + 0x100 * 4, 0x180 * 4, nullptr) >= 0) {
                os = DCSB_OS93B;
                if (search_opcodes("47FFF2 47C946", 0x2000 + 0x200 * 4, 0x100 * 4, nullptr) >= 0) os = DCSB_OS93A;
            }
        }
        if (validated == populated && populated == in_table) return post = 1;
        return post = (uint8_t)(first_failed + 2);
    }
    return post = 2;
}

int dcsb_rom::num_channels() const
{
    std::unordered_map<char, uint32_t> v;
    if (chip[0].present &&
        search_opcodes("22200F 4000n4 26E20F 221800 9****A 8****A 400mm4 26E20F 18***1", 0, 0x6000, &v) >= 0) {
        const int n = (int)v['n'];
        if (v['m'] == (uint32_t)((1 << n) - 1)) return n;
    }
    return 0;
}

// operand bytes per opcode as the reference's track scanners count them (DCSDecoder.cpp:836-866)
int dcsb_rom::opcode_operand_bytes(int opcode) const
{
    switch (opcode) {
    case 0x01: return 5;
    case 0x02: case 0x05: case 0x0E: return 1;
    case 0x03: case 0x06: case 0x07: case 0x08: case 0x09: case 0x11: case 0x12: return 2;
    case 0x0A: case 0x0B: case 0x0C: return 4;
    case 0x04: return os == DCSB_OS93A ? 3 : 1;
    default: return 0;
    }
}

bool dcsb_rom::track_info(uint16_t track, dcsb_track_info *ti) const
{
    memset(ti, 0, sizeof(*ti));
    ti->defer_code = 0xFFFF;
    if (track >= n_tracks) return false;
    const uint32_t addr = u2_be(track_index + 3u * track, 3);
    if ((addr & 0xFF0000u) == 0xFF0000u) return false;
    DcsbRomPtr p = make_ptr(addr);
    const int type = u8(p), ch = u8(p, 1);
    p.ofs += 2;
    if (ch > 7) return false;
    bool done = false;
    uint16_t defer = 0xFFFF;
    if (type == 2 || type == 3) { defer = (uint16_t)be(p, 2); done = true; }
    else if (type != 1) return false;
    // playing time: wait prefixes summed over nested loops (DCSDecoder.cpp:727-884)
    struct Level { uint32_t time = 0, loop_stream = 0; uint8_t n = 1; bool forever = false; };
    std::vector<Level> st(1);
    for (int guard = 0; !done && guard < 100000; ++guard) {
        const uint32_t wait = be(p, 2);
        const int op = u8(p, 2);
        p.ofs += 3;
        if (wait == 0xFFFF) {                       // waits forever: what stays audible is the looping stream
            st.back().forever = true;
            st.back().time += st.back().loop_stream;
            break;
        }
        st.back().time += wait;
        if (op == 0x00) break;
        if (op == 0x01) {
            const DcsbRomPtr sp = make_ptr(be(p, 3, 1));
            const int repeat = u8(p, 4);
            st.back().loop_stream = repeat == 0 ? be(sp, 2) : 0;
        } else if (op == 0x0E) {
            Level l;
            l.n = u8(p);
            l.forever = l.n == 0;
            st.push_back(l);
        } else if (op == 0x0F && st.size() > 1) {
            const Level l = st.back();
            st.pop_back();
            st.back().time += (l.forever ? 1u : l.n) * l.time;
            if (l.forever) { st.back().forever = true; break; }
        }
        p.ofs += (uint32_t)opcode_operand_bytes(op);
    }
    while (st.size() > 1) {
        const Level l = st.back();
        st.pop_back();
        st.back().time += (l.n == 0 ? 1u : l.n) * l.time;
        if (l.forever) st.back().forever = true;
    }
    ti->address = addr;
    ti->channel = ch;
    ti->type = type;
    ti->defer_code = defer;
    ti->time = st.back().time;
    ti->looping = st.back().forever;
    return true;
}

// The steps of a type-1 track program, described the way the reference's decompiler describes them
// (DCSDecoder.cpp:885-1135): same step boundaries (its operand sizes, incl. the 3-byte opcode 4 of
// OS93a), same loop bookkeeping (loop_parent = the reference's "parentOffset": the number of steps
// up to and including the loop's own), same mnemonic and hex wording.  The walk ends behind End,
// an invalid opcode, or a step that waits forever.
std::vector<dcsb_opcode> dcsb_rom::decompile_track(uint16_t track) const
{
    std::vector<dcsb_opcode> v;
    dcsb_track_info ti;
    if (!track_info(track, &ti) || ti.type != 1) return v;
    DcsbRomPtr p = make_ptr(ti.address);
    const uint32_t start = p.ofs;
    p.ofs += 2;
    std::vector<int> loops;
    auto chtag = [&](int ch, const char *sep) { char b[24] = ""; if (ch != ti.channel) snprintf(b, sizeof(b), "channel %d,%s", ch, sep); return std::string(b); };
    for (bool done = false; !done && v.size() < 65536;) {
        dcsb_opcode e;
        memset(&e, 0, sizeof(e));
        e.nesting_level = (int32_t)loops.size();
        e.loop_parent = loops.empty() ? -1 : loops.back();
        e.offset = (int32_t)(p.ofs - start);
        e.delay_count = (uint16_t)be(p, 2);
        if (e.delay_count == 0xFFFF) done = true;
        e.opcode = u8(p, 2);
        p.ofs += 3;
        // (the decompiler's own operand sizes: unlike GetTrackInfo's walk it takes 2 bytes for $10 and 4 for $11 / $12)
        const int nops = e.opcode == 0x10 ? 2 : (e.opcode == 0x11 || e.opcode == 0x12) ? 4 : e.opcode < 0x10 ? opcode_operand_bytes(e.opcode) : 0;
        uint8_t o[8] = { 0 };
        for (int i = 0; i < nops && i < 8; ++i) o[i] = u8(p, (uint32_t)i);
        const unsigned w01 = (o[0] << 8) | o[1], w23 = (o[2] << 8) | o[3];
        char d[96] = "", h[64] = "";
        switch (e.opcode) {
        case 0x00: snprintf(d, sizeof(d), "End;"); done = true; break;
        case 0x01: {
            const unsigned sp = (o[1] << 16) | (o[2] << 8) | o[3];
            snprintf(h, sizeof(h), " %02X %06X %02X", o[0], sp, o[4]);
            const std::string tag = chtag(o[0], "");
            if (o[4] == 0) snprintf(d, sizeof(d), "Play(%sstream $%06X, repeat forever);", tag.c_str(), sp);
            else if (o[4] == 1) snprintf(d, sizeof(d), "Play(%sstream $%06X);", tag.c_str(), sp);
            else snprintf(d, sizeof(d), "Play(%sstream $%06X, repeat %d);", tag.c_str(), sp, o[4]);
            break;
        }
        case 0x02: snprintf(h, sizeof(h), " %02X", o[0]); snprintf(d, sizeof(d), "Stop(channel %d);", o[0]); break;
        case 0x03: snprintf(h, sizeof(h), " %04X", w01); snprintf(d, sizeof(d), "Queue(track $%0X);", w01); break;
        case 0x04:
            if (os == DCSB_OS93A) {
                const unsigned cnt = (o[1] << 8) | o[2];
                snprintf(h, sizeof(h), " %02X %04X", o[0], cnt);
                snprintf(d, sizeof(d), "SetChannelTimer(byte $%02X, counter $%04X);", o[0], cnt);
            } else {
                snprintf(h, sizeof(h), " %02X", o[0]);
                snprintf(d, sizeof(d), "WriteDataPort(byte $%02X);", o[0]);
            }
            break;
        case 0x05: snprintf(h, sizeof(h), " %02X", o[0]); snprintf(d, sizeof(d), "StartDeferred(channel %d);", o[0]); break;
        case 0x06: snprintf(h, sizeof(h), " %02X %02X", o[0], o[1]); snprintf(d, sizeof(d), "SetVariable(var $%02X, value $%02X);", o[0], o[1]); break;
        case 0x07: case 0x08: case 0x09:
            snprintf(h, sizeof(h), " %02X %02X", o[0], o[1]);
            snprintf(d, sizeof(d), "SetMixingLevel(%s%s %d);", chtag(o[0], " ").c_str(),
                     e.opcode == 7 ? "level" : e.opcode == 8 ? "increase" : "decrease", o[1]);
            break;
        case 0x0A: case 0x0B: case 0x0C:
            snprintf(h, sizeof(h), " %02X %02X %04X", o[0], o[1], w23);
            snprintf(d, sizeof(d), "SetMixingLevel(%s%s %u, steps %u);", chtag(o[0], " ").c_str(),
                     e.opcode == 0x0A ? "level" : e.opcode == 0x0B ? "increase" : "decrease", o[1], w23);
            break;
        case 0x0D: snprintf(d, sizeof(d), "NOP;"); break;
        case 0x0E:
            snprintf(h, sizeof(h), " %02X", o[0]);
            if (o[0]) snprintf(d, sizeof(d), "Loop (%d) {", o[0]); else snprintf(d, sizeof(d), "Loop {");
            loops.push_back((int)v.size() + 1);
            break;
        case 0x0F:
            if (!loops.empty()) { loops.pop_back(); snprintf(d, sizeof(d), "}"); } else snprintf(d, sizeof(d), "LoopEnd");
            break;
        case 0x10: snprintf(h, sizeof(h), " %02X %02X", o[0], o[1]); snprintf(d, sizeof(d), "Opcode$10($%02X,$%02X);", o[0], o[1]); break;
        case 0x11: case 0x12:
            snprintf(h, sizeof(h), " %02X %02X %04X", o[0], o[1], w23);
            snprintf(d, sizeof(d), "Opcode$%02x($%02X,$%02X,$%04X);", e.opcode, o[0], o[1], w23);
            break;
        default: snprintf(d, sizeof(d), "InvalidOpcode$%02X;", e.opcode); done = true; break;
        }
        p.ofs += (uint32_t)nops;
        e.n_operand_bytes = (uint8_t)nops;
        memcpy(e.operand_bytes, o, 8);
        snprintf(e.desc, sizeof(e.desc), "%s", d);
        snprintf(e.hex_desc, sizeof(e.hex_desc), "%04X %02X%s", e.delay_count, e.opcode, h);
        v.push_back(e);
    }
    return v;
}

// every stream a Play opcode (0x01) of a type-1 track refers to, ascending.
// as_executed = false: the reference's ListStreams (DCSDecoder.cpp:1248-1293), which walks the
// programs with DecompileTrackProgram's operand sizes (:886-1126) -- on 1993 software that
// miscounts opcode 6 (no operands there) and can miss streams behind it.
// as_executed = true: operand sizes as ExecTrack consumes them (DCSDecoderNative.cpp:848-1228);
// this is what the player's stream table is built from.
std::vector<uint32_t> dcsb_rom::list_streams(bool as_executed) const
{
    std::set<uint32_t> found;
    const bool os93 = os == DCSB_OS93A || os == DCSB_OS93B;
    for (uint32_t t = 0; t < n_tracks; ++t) {
        dcsb_track_info ti;
        if (!track_info((uint16_t)t, &ti) || ti.type != 1) continue;
        DcsbRomPtr p = make_ptr(ti.address);
        p.ofs += 2;
        for (int guard = 0; guard < 100000; ++guard) {
            const uint32_t wait = be(p, 2);
            const int op = u8(p, 2);
            p.ofs += 3;
            if (op == 0x01) found.insert(be(p, 3, 1));
            if (op == 0x00 || op > 0x12 || wait == 0xFFFF) break;
            static const uint8_t operands[0x13] = { 0, 5, 1, 2, 1, 1, 2, 2, 2, 2, 4, 4, 4, 0, 1, 0, 2, 4, 4 };
            uint32_t n = operands[op];
            if (op == 0x04 && os == DCSB_OS93A) n = 3;
            if (op == 0x06 && as_executed && os93) n = 0;
            p.ofs += n;
        }
    }
    return std::vector<uint32_t>(found.begin(), found.end());
}

// ======================================================================================
// zip container
static uint32_t le32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint32_t le16(const uint8_t *p) { return p[0] | (p[1] << 8); }

bool dcsb_unzip(const char *path, std::vector<DcsbZipEntry> &out, std::string &err)
{
    FILE *f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    std::vector<uint8_t> z;
    uint8_t buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) z.insert(z.end(), buf, buf + n);
    fclose(f);
    // end-of-central-directory record: last occurrence of PK\5\6
    if (z.size() < 22) { err = "not a zip file"; return false; }
    size_t eocd = std::string::npos;
    for (size_t i = z.size() - 22; i + 1 > 0 && z.size() - i < 66000; --i) {
        if (le32(&z[i]) == 0x06054b50u) { eocd = i; break; }
        if (i == 0) break;
    }
    if (eocd == std::string::npos) { err = "zip directory not found"; return false; }
    const uint32_t count = le16(&z[eocd + 10]);
    size_t p = le32(&z[eocd + 16]);
    // nothing in the directory is trusted: every length is checked against the file before it is used, and an
    // entry may not claim more than DCSB_ZIP_MAX_ENTRY bytes (the largest DCS sound ROM chip is 1 MB)
    const size_t DCSB_ZIP_MAX_ENTRY = 16u << 20;
    if (p > eocd || (size_t)count * 46 > eocd - p) { err = "corrupt zip directory"; return false; }
    for (uint32_t e = 0; e < count; ++e) {
        if (p + 46 > eocd || le32(&z[p]) != 0x02014b50u) { err = "corrupt zip directory"; return false; }
        const uint32_t method = le16(&z[p + 10]), csize = le32(&z[p + 20]), usize = le32(&z[p + 24]);
        const uint32_t nlen = le16(&z[p + 28]), xlen = le16(&z[p + 30]), clen = le16(&z[p + 32]), lho = le32(&z[p + 42]);
        if (p + 46 + (size_t)nlen + xlen + clen > eocd) { err = "corrupt zip directory"; return false; }
        std::string name(reinterpret_cast<const char *>(&z[p + 46]), nlen);
        p += 46 + nlen + xlen + clen;
        if (!name.empty() && name.back() == '/') continue;      // directory
        if ((size_t)lho + 30 > z.size() || le32(&z[lho]) != 0x04034b50u) { err = "corrupt zip entry " + name; return false; }
        const size_t data = (size_t)lho + 30 + le16(&z[lho + 26]) + le16(&z[lho + 28]);
        if (data > z.size() || (size_t)csize > z.size() - data) { err = "truncated zip entry " + name; return false; }
        if (usize > DCSB_ZIP_MAX_ENTRY) { err = "zip entry too large: " + name; return false; }
        DcsbZipEntry ent;
        ent.name = name;
        ent.data.resize(usize);
        if (method == 0) {
            if (csize != usize) { err = "bad stored entry " + name; return false; }
            memcpy(ent.data.data(), &z[data], usize);
        } else if (method == 8) {
            z_stream zs;
            memset(&zs, 0, sizeof(zs));
            if (inflateInit2(&zs, -15) != Z_OK) { err = "inflateInit2 failed"; return false; }
            zs.next_in = &z[data];
            zs.avail_in = csize;
            zs.next_out = ent.data.data();
            zs.avail_out = usize;
            const int rc = inflate(&zs, Z_FINISH);
            inflateEnd(&zs);
            if (rc != Z_STREAM_END || zs.total_out != usize) { err = "error uncompressing " + name; return false; }
        } else { err = "unsupported compression method in " + name; return false; }
        out.push_back(std::move(ent));
    }
    return true;
}

// "[SU]<non-digits><digit> ... <ws>dd/dd/dd" at the start of a sound ROM image names its chip
// (DCSDecoderZipLoader.cpp:168-186)
static int image_chip_digit(const std::vector<uint8_t> &d)
{
    size_t n = 0;
    while (n < d.size() && n < 256 && d[n]) ++n;
    if (n == d.size() || n == 256) return 0;
    const char *s = reinterpret_cast<const char *>(d.data());
    if (s[0] != 'S' && s[0] != 'U') return 0;
    size_t i = 1;
    while (i < n && !isdigit((unsigned char)s[i])) ++i;
    if (i >= n) return 0;
    const int digit = s[i];
    // the text must end with whitespace + a dd/dd/dd date
    if (n < i + 1 + 9) return 0;
    const char *t = s + n - 8;
    const bool date = isdigit((unsigned char)t[0]) && isdigit((unsigned char)t[1]) && t[2] == '/' && isdigit((unsigned char)t[3]) &&
                      isdigit((unsigned char)t[4]) && t[5] == '/' && isdigit((unsigned char)t[6]) && isdigit((unsigned char)t[7]);
    if (!date || !isspace((unsigned char)t[-1])) return 0;
    return digit;
}

int dcsb_rom_load_zip_impl(dcsb_rom *rom, const char *path, const char *explicit_u2)
{
    std::vector<DcsbZipEntry> files;
    FILE *probe = fopen(path, "rb");
    if (!probe) { rom->err = std::string("Error opening ROM Zip file \"") + path + "\""; return DCSB_ZIP_E_OPEN; }
    fclose(probe);
    if (!dcsb_unzip(path, files, rom->err)) return DCSB_ZIP_E_EXTRACT;
    // U2: the image that starts with a JUMP and has a '2' in its name, or the one named explicitly
    int u2 = -1;
    for (size_t i = 0; i < files.size() && u2 < 0; ++i) {
        const DcsbZipEntry &f = files[i];
        if ((f.data.size() >= 4 && is_jump(f.data.data()) && f.name.find('2') != std::string::npos) ||
            (explicit_u2 && strcasecmp(f.name.c_str(), explicit_u2) == 0))
            u2 = (int)i;
    }
    if (u2 < 0) {
        rom->err = std::string("No file in ") + path + " could be identified as ROM U2";
        rom->zip_chip.assign(files.size(), -1);
        rom->zip_files = std::move(files);
        return DCSB_ZIP_E_NOU2;
    }
    std::vector<int> used(files.size(), 0);
    used[u2] = 2;
    rom->add(2, files[u2].data.data(), files[u2].data.size());
    std::string base = path;
    const size_t slash = base.find_last_of("/\\");
    if (slash != std::string::npos) base = base.substr(slash + 1);
    const bool cactus = base.size() >= 4 && strncasecmp(base.c_str(), "cc_", 3) == 0 && isdigit((unsigned char)base[3]);
    for (int n = 3; n <= 9; ++n)
        for (size_t i = 0; i < files.size(); ++i) {
            if (used[i] || files[i].name.find((char)('0' + n)) == std::string::npos) continue;
            const int digit = image_chip_digit(files[i].data);
            bool load = digit == '0' + n;
            if (cactus && digit && n == 7 && digit == '6') load = true;     // mislabelled chip in the cc_ sets
            if (load) { rom->add(n, files[i].data.data(), files[i].data.size()); used[i] = n; break; }
        }
    rom->zip_chip.resize(files.size());
    for (size_t i = 0; i < files.size(); ++i) rom->zip_chip[i] = used[i] ? used[i] : -1;
    rom->zip_files = std::move(files);
    return DCSB_ZIP_OK;
}

// ======================================================================================
// sequencer: the core lives in dcsb_seq.cuh (shared with the device build); this is the player's instance of it
DcsbRomView dcsb_rom::view() const
{
    DcsbRomView v;
    memset(&v, 0, sizeof(v));
    v.image = image.data();
    for (int i = 0; i < 8; ++i) {
        v.chip_ofs[i] = image_ofs[i];
        v.chip_size[i] = chip[i].size;
        v.chip_mask[i] = chip[i].mask;
        v.present[i] = chip[i].present ? 1 : 0;
    }
    v.hw = hw;
    v.os = os;
    v.nominal_version = nominal_version;
    v.n_tracks = n_tracks;
    v.totan = totan ? 1 : 0;
    v.track_index = track_index;
    v.indirect_index = indirect_index;
    v.streams = seq_streams.data();
    v.n_streams = (uint32_t)seq_streams.size();
    return v;
}

void dcsb_rom::build_seq_streams()
{
    seq_streams.resize(streams.size());
    for (size_t i = 0; i < streams.size(); ++i)
        seq_streams[i] = DcsbSeqStream{ streams[i].linear & 0xFFFFFFu, streams[i].nplay, streams[i].status, (uint32_t)i };
    std::sort(seq_streams.begin(), seq_streams.end(), [](const DcsbSeqStream &a, const DcsbSeqStream &b) { return a.linear < b.linear; });
}

DcsbSequencer::DcsbSequencer(const dcsb_rom *r) : rom(r) { dcsb_seq_init(st); }

void DcsbSequencer::drain_host_bytes()
{
    // (a frame hands back at most DCSB_SEQ_FRAME_HOST_BYTES bytes with their values; a well-formed program sends a few)
    const uint32_t n = st.host_n < DCSB_SEQ_FRAME_HOST_BYTES ? st.host_n : DCSB_SEQ_FRAME_HOST_BYTES;
    for (uint32_t i = 0; i < n; ++i) { host_bytes.push_back(st.host_buf[i]); host_byte_frames.push_back(st.frame_no); }
    st.host_n = 0;
}

void DcsbSequencer::soft_boot()
{
    for (int i = 0; i < DCSB_MAX_CHANNELS; ++i) { st.chan[i].stop = false; st.chan[i].volume = 0xFF; }
    dcsb_seq_set_master_volume(st, 0x67);       // DCSDecoder's default volume until the host says otherwise
    st.port_bytes = 0;
    st.quiet = 0;
}

void DcsbSequencer::set_master_volume(int vol) { dcsb_seq_set_master_volume(st, vol); }
void DcsbSequencer::clear_tracks() { dcsb_seq_clear_tracks(st); }
void DcsbSequencer::write_port(uint8_t data) { dcsb_seq_write_port(st, rom->view(), data); drain_host_bytes(); }
void DcsbSequencer::load_stream(int ch, uint32_t linear, int level) { dcsb_seq_load_stream(st, rom->view(), ch, linear, level); }

bool DcsbSequencer::frame(std::vector<DcsbSchedFrame> &frames, std::vector<DcsbSchedEntry> &entries)
{
    DcsbSchedFrame fr{ (uint32_t)entries.size(), 0, 8, 0, 0 };
    DcsbSchedEntry e[DCSB_MAX_CHANNELS];
    // (host bytes of this frame carry its number: the core counts the frame at the end of the pass)
    const DcsbRomView v = rom->view();
    const uint32_t this_frame = st.frame_no;
    const bool ok = dcsb_seq_frame(st, v, &fr, e);
    for (int i = 0; i < fr.n_entries; ++i) entries.push_back(e[i]);
    frames.push_back(fr);
    const uint32_t n = st.host_n < DCSB_SEQ_FRAME_HOST_BYTES ? st.host_n : DCSB_SEQ_FRAME_HOST_BYTES;
    for (uint32_t i = 0; i < n; ++i) { host_bytes.push_back(st.host_buf[i]); host_byte_frames.push_back(this_frame); }
    st.host_n = 0;
    fatal = st.fatal;
    frame_no = st.frame_no;
    return ok;
}
