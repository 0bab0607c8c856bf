// JPEG decoder for RtTexture (the reference's scenes ship JPEG textures).
//
// The reference decodes textures with stb_image v2.16 (`stbi_load(path, .., 3)` with flip-on-load,
// reflectcuts/realtimetechniques/rtcommon.h:32,144).  Texel bytes feed pow(byte/255, gamma) and then every
// shaded pixel, so the decoder has to give the SAME bytes as stb, not merely "a valid decode": JPEG leaves the
// IDCT, the chroma upsampling filter and the YCbCr->RGB rounding to the implementation.  This file is an
// independent decoder (own bit reader, canonical Huffman tables, coefficient planes for baseline and
// progressive alike) whose three arithmetic stages restate stb's:
//   * IDCT: 12-bit fixed-point LL&M ("islow") butterflies, column pass keeps 2 extra bits (>>10), row pass
//     rounds with +65536+(128<<17) and >>17, clamp to 0..255                    (stb_image.h:2113-2212)
//   * upsampling: h2v1 / h1v2 triangle filters (3*near+far+2)>>2, h2v2 (3*t0+t1+8)>>4 with the row
//     pairing near/far alternating per output row, nearest for other ratios   (stb_image.h:3124-3213, 3549-3590)
//   * colour: y<<20 + (1<<19) + c*fixed(k)<<8, the Cb term of green masked to its high 16 bits, >>20, clamp
//                                                                               (stb_image.h:3317-3343)
// tests/test_cpu.py compares it bit-exactly with stb itself (oracle/_ref/libstb_ref.so, built from the
// reference tree's own stb_image.h) on the committed fixtures in tests/golden/jpeg/.
//
// Supported: SOF0/SOF1/SOF2 (baseline, extended sequential, progressive), 8-bit, Huffman, 1/3/4 components,
// any sampling factors 1..4, restart intervals, JFIF / Adobe colour transform flags.  Not supported (stb does
// not either): arithmetic coding, 12-bit, lossless, hierarchical.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace evplp_host {
namespace jpeg {

struct Image {
    int width = 0, height = 0;
    int fileChannels = 0;       // 1 (grey) or 3 (anything with >= 3 components), as stbi_load reports it
    std::vector<uint8_t> rgb;   // width*height*3, row 0 = top
};

namespace detail {

static const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// a run that overshoots the block (corrupt stream) lands on the last coefficient instead of outside the block
inline int zz(int k) { return kZigzag[k < 64 ? k : 63]; }

[[noreturn]] inline void fail(const char* what) { throw std::runtime_error(std::string("jpeg: ") + what); }

struct ByteSource {
    const uint8_t* p;
    const uint8_t* end;
    bool eof() const { return p >= end; }
    int u8() { return p < end ? *p++ : 0; }
    int u16() { int a = u8(); return (a << 8) | u8(); }
    void skip(int n) { p = (n > end - p) ? end : p + n; }
};

// Entropy-coded segment reader: MSB-first, 0xFF00 unstuffing; when a marker is met it is remembered and the
// stream continues as zero bits (a truncated scan then decodes to something rather than failing).
struct BitReader {
    ByteSource* src = nullptr;
    uint32_t acc = 0;   // left-aligned
    int count = 0;
    int marker = -1;

    void reset() { acc = 0; count = 0; marker = -1; }
    void fill() {
        while (count <= 24) {
            int b = 0;
            if (marker < 0) {
                b = src->u8();
                if (b == 0xFF) {
                    int c = src->u8();
                    while (c == 0xFF) c = src->u8();
                    if (c != 0) { marker = c; b = 0; }
                }
            }
            acc |= (uint32_t)b << (24 - count);
            count += 8;
        }
    }
    uint32_t peek16() { if (count < 16) fill(); return acc >> 16; }
    void drop(int n) { acc <<= n; count -= n; }
    int bits(int n) {
        if (n == 0) return 0;
        if (count < n) fill();
        uint32_t v = acc >> (32 - n);
        drop(n);
        return (int)v;
    }
    int bit() { return bits(1); }
    // JPEG "receive + extend" (ITU T.81 F.2.2.1)
    int receiveExtend(int n) {
        if (n == 0) return 0;
        int v = bits(n);
        return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
    }
};

struct HuffTable {
    bool defined = false;
    uint8_t values[256];
    uint32_t limit[18];   // limit[len] = (first code of the next length) << (16 - len); sentinel at 17
    int offset[17];       // symbol index = (code >> (16-len)) + offset[len]
    uint16_t lut[512];    // 9-bit prefix -> (len << 8) | value, 0 = longer code

    void build(const int counts[16], const uint8_t* syms, int n) {
        memcpy(values, syms, (size_t)n);
        int code = 0, k = 0;
        memset(lut, 0, sizeof lut);
        for (int len = 1; len <= 16; len++) {
            offset[len] = k - code;
            for (int i = 0; i < counts[len - 1]; i++, k++, code++) {
                if (code >= (1 << len)) fail("bad code lengths");
                if (len <= 9) {
                    int first = code << (9 - len), span = 1 << (9 - len);
                    for (int j = 0; j < span; j++) lut[first + j] = (uint16_t)((len << 8) | syms[k]);
                }
            }
            limit[len] = (uint32_t)code << (16 - len);
            code <<= 1;
        }
        limit[17] = 0xFFFFFFFFu;
        defined = true;
    }
    int decode(BitReader& br) const {
        if (!defined) fail("scan uses an undefined huffman table");
        uint32_t top = br.peek16();
        uint16_t e = lut[top >> 7];
        if (e) {
            int len = e >> 8;
            if (len > br.count) fail("bad huffman code");
            br.drop(len);
            return e & 255;
        }
        int len = 10;
        while (top >= limit[len]) len++;
        if (len > 16 || len > br.count) fail("bad huffman code");
        int idx = (int)(top >> (16 - len)) + offset[len];
        br.drop(len);
        return values[idx & 255];
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0;
    int td = 0, ta = 0;      // Huffman table selectors of the current scan
    int x = 0, y = 0;        // size in samples
    int blocksW = 0, blocksH = 0;   // allocated blocks (whole MCUs)
    int dcPred = 0;
    std::vector<int16_t> coeff;     // blocksW*blocksH*64, natural order, NOT dequantised
    std::vector<uint8_t> plane;     // (blocksW*8) x (blocksH*8)
    int stride() const { return blocksW * 8; }
};

// stb_image.h:2113: constants are round(x * 4096) computed as (int)(x*4096 + 0.5) on the float literal
constexpr int fix12(float x) { return (int)(x * 4096 + 0.5); }

// One 8-point pass of the LL&M inverse DCT in 12-bit fixed point.  even[k] +/- odd[k] are the outputs
// k and 7-k (before the pass's own rounding shift).  Arithmetic is modulo 2^32 (unsigned), which equals the
// reference's int arithmetic on every decodable stream and stays defined on corrupt ones.
typedef uint32_t u32;
constexpr u32 K(float x) { return (u32)fix12(x); }
inline void idct8(const u32 s[8], u32 even[4], u32 odd[4]) {
    // even part: s0, s4 through a plain butterfly; s2, s6 through the sqrt(2)cos(6pi/16) rotation
    const u32 z = (s[2] + s[6]) * K(0.5411961f);
    const u32 e2 = z + s[6] * K(-1.847759065f);
    const u32 e3 = z + s[2] * K(0.765366865f);
    const u32 e0 = (s[0] + s[4]) << 12;
    const u32 e1 = (s[0] - s[4]) << 12;
    even[0] = e0 + e3; even[3] = e0 - e3;
    even[1] = e1 + e2; even[2] = e1 - e2;
    // odd part
    const u32 a = s[7], b = s[5], c = s[3], d = s[1];
    const u32 z5 = (a + c + b + d) * K(1.175875602f);
    const u32 z1 = z5 + (a + d) * K(-0.899976223f);
    const u32 z2 = z5 + (b + c) * K(-2.562915447f);
    const u32 z3 = (a + c) * K(-1.961570560f);
    const u32 z4 = (b + d) * K(-0.390180644f);
    odd[3] = a * K(0.298631336f) + z1 + z3;   // pairs with even[3]  (outputs 3, 4)
    odd[2] = b * K(2.053119869f) + z2 + z4;   //            even[2]  (outputs 2, 5)
    odd[1] = c * K(3.072711026f) + z2 + z3;   //            even[1]  (outputs 1, 6)
    odd[0] = d * K(1.501321110f) + z1 + z4;   //            even[0]  (outputs 0, 7)
}

inline uint8_t clamp255(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

// Dequantised block (natural order) -> 8x8 samples.  stb_image.h:2156-2212.
inline void idctBlock(const int16_t in[64], uint8_t* out, int stride) {
    u32 mid[64], s[8], ev[4], od[4];
    for (int c = 0; c < 8; c++) {
        for (int k = 0; k < 8; k++) s[k] = (u32)(int)in[k * 8 + c];
        idct8(s, ev, od);
        for (int k = 0; k < 4; k++) {
            mid[k * 8 + c] = (u32)((int32_t)(ev[k] + 512 + od[k]) >> 10);
            mid[(7 - k) * 8 + c] = (u32)((int32_t)(ev[k] + 512 - od[k]) >> 10);
        }
    }
    const u32 bias = 65536 + (128 << 17);
    for (int r = 0; r < 8; r++) {
        idct8(mid + r * 8, ev, od);
        uint8_t* o = out + (size_t)r * stride;
        for (int k = 0; k < 4; k++) {
            o[k] = clamp255((int32_t)(ev[k] + bias + od[k]) >> 17);
            o[7 - k] = clamp255((int32_t)(ev[k] + bias - od[k]) >> 17);
        }
    }
}

struct Decoder {
    ByteSource src;
    BitReader br;
    HuffTable dcTab[4], acTab[4];
    uint16_t quant[4][64];   // natural order
    Component comp[4];
    int ncomp = 0, width = 0, height = 0;
    int hMax = 1, vMax = 1, mcuX = 0, mcuY = 0;
    bool progressive = false, haveFrame = false;
    bool jfif = false;
    int adobeTransform = -1;
    int rgbIds = 0;
    int restartInterval = 0;
    // scan state
    int scanN = 0, order[4];
    int ss = 0, se = 63, ah = 0, al = 0;
    int eobRun = 0;

    int nextMarker() {
        // markers are 0xFF + non-zero, possibly preceded by fill 0xFF bytes
        if (br.marker >= 0) { int m = br.marker; br.marker = -1; return m; }
        int x = src.u8();
        if (x != 0xFF) return -1;
        while (x == 0xFF) x = src.u8();
        return x;
    }

    void readDQT() {
        int L = src.u16() - 2;
        while (L > 0) {
            int q = src.u8();
            int prec = q >> 4, t = q & 15;
            if (prec > 1) fail("bad DQT type");
            if (t > 3) fail("bad DQT table");
            for (int i = 0; i < 64; i++) quant[t][kZigzag[i]] = (uint16_t)(prec ? src.u16() : src.u8());
            L -= prec ? 129 : 65;
        }
        if (L != 0) fail("bad DQT len");
    }
    void readDHT() {
        int L = src.u16() - 2;
        while (L > 0) {
            int q = src.u8();
            int tc = q >> 4, th = q & 15;
            if (tc > 1 || th > 3) fail("bad DHT header");
            int counts[16], n = 0;
            for (int i = 0; i < 16; i++) { counts[i] = src.u8(); n += counts[i]; }
            if (n > 256) fail("bad DHT counts");
            uint8_t syms[256];
            for (int i = 0; i < n; i++) syms[i] = (uint8_t)src.u8();
            (tc ? acTab[th] : dcTab[th]).build(counts, syms, n);
            L -= 17 + n;
        }
        if (L != 0) fail("bad DHT len");
    }
    void readAPPorCOM(int m) {
        int L = src.u16();
        if (L < 2) fail("bad APP/COM len");
        L -= 2;
        if (m == 0xE0 && L >= 5) {
            static const char tag[5] = {'J', 'F', 'I', 'F', 0};
            bool ok = true;
            for (int i = 0; i < 5; i++) ok &= src.u8() == (uint8_t)tag[i];
            L -= 5;
            if (ok) jfif = true;
        } else if (m == 0xEE && L >= 12) {
            static const char tag[6] = {'A', 'd', 'o', 'b', 'e', 0};
            bool ok = true;
            for (int i = 0; i < 6; i++) ok &= src.u8() == (uint8_t)tag[i];
            L -= 6;
            if (ok) {
                src.skip(5);   // version, flags0, flags1
                adobeTransform = src.u8();
                L -= 6;
            }
        }
        src.skip(L);
    }
    void readSOF(int m) {
        progressive = (m == 0xC2);
        int Lf = src.u16();
        if (Lf < 11) fail("bad SOF len");
        if (src.u8() != 8) fail("only 8-bit samples are supported");
        height = src.u16();
        width = src.u16();
        if (height == 0 || width == 0) fail("zero image size");
        ncomp = src.u8();
        if (ncomp != 1 && ncomp != 3 && ncomp != 4) fail("bad component count");
        if (Lf != 8 + 3 * ncomp) fail("bad SOF len");
        rgbIds = 0;
        for (int i = 0; i < ncomp; i++) {
            Component& c = comp[i];
            c.id = src.u8();
            if (ncomp == 3 && c.id == "RGB"[i]) rgbIds++;
            int q = src.u8();
            c.h = q >> 4; c.v = q & 15;
            if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4) fail("bad sampling factor");
            c.tq = src.u8();
            if (c.tq > 3) fail("bad TQ");
            if (c.h > hMax) hMax = c.h;
            if (c.v > vMax) vMax = c.v;
        }
        if ((uint64_t)width * height > (1ull << 28)) fail("image too large");
        mcuX = (width + hMax * 8 - 1) / (hMax * 8);
        mcuY = (height + vMax * 8 - 1) / (vMax * 8);
        for (int i = 0; i < ncomp; i++) {
            Component& c = comp[i];
            c.x = (width * c.h + hMax - 1) / hMax;
            c.y = (height * c.v + vMax - 1) / vMax;
            c.blocksW = mcuX * c.h;
            c.blocksH = mcuY * c.v;
            c.coeff.assign((size_t)c.blocksW * c.blocksH * 64, 0);
        }
        haveFrame = true;
    }
    void readSOS() {
        if (!haveFrame) fail("SOS before SOF");
        int Ls = src.u16();
        scanN = src.u8();
        if (scanN < 1 || scanN > 4 || scanN > ncomp) fail("bad SOS component count");
        if (Ls != 6 + 2 * scanN) fail("bad SOS len");
        for (int i = 0; i < scanN; i++) {
            int id = src.u8(), q = src.u8(), which = 0;
            while (which < ncomp && comp[which].id != id) which++;
            if (which == ncomp) fail("SOS names an unknown component");
            comp[which].td = q >> 4;
            comp[which].ta = q & 15;
            if (comp[which].td > 3 || comp[which].ta > 3) fail("bad huffman selector");
            order[i] = which;
        }
        ss = src.u8(); se = src.u8();
        int a = src.u8();
        ah = a >> 4; al = a & 15;
        if (progressive) {
            if (ss > 63 || se > 63 || ss > se || ah > 13 || al > 13) fail("bad SOS");
            if (ss == 0 && se != 0) fail("progressive scan mixes DC and AC");
            if (ss != 0 && scanN != 1) fail("interleaved AC scan");
        } else {
            if (ss != 0 || ah != 0 || al != 0) fail("bad SOS");
            se = 63;
        }
    }

    void resetEntropy() {
        br.reset();
        for (int i = 0; i < 4; i++) comp[i].dcPred = 0;
        eobRun = 0;
    }

    // ---- per-block entropy decoding (coefficients stay un-dequantised int16, natural order)
    void blockSequential(Component& c, int16_t* blk) {
        const HuffTable& hd = dcTab[c.td];
        const HuffTable& ha = acTab[c.ta];
        int t = hd.decode(br);
        if (t > 15) fail("bad DC category");
        c.dcPred += br.receiveExtend(t);
        blk[0] = (int16_t)c.dcPred;
        for (int k = 1; k < 64;) {
            int rs = ha.decode(br);
            int r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (rs != 0xF0) break;
                k += 16;
            } else {
                k += r;
                blk[zz(k++)] = (int16_t)br.receiveExtend(s);
            }
        }
    }
    void blockDCProgressive(Component& c, int16_t* blk) {
        if (ah == 0) {
            int t = dcTab[c.td].decode(br);
            if (t > 15) fail("bad DC category");
            c.dcPred += br.receiveExtend(t);
            blk[0] = (int16_t)((uint32_t)c.dcPred << al);
        } else if (br.bit()) {
            blk[0] = (int16_t)(blk[0] + (1 << al));
        }
    }
    // correction bit for an already non-zero coefficient (ITU T.81 G.1.2.3)
    void refine(int16_t& v, int16_t bit) {
        if (br.bit() && (v & bit) == 0) v = (int16_t)(v > 0 ? v + bit : v - bit);
    }
    void blockACProgressive(Component& c, int16_t* blk) {
        const HuffTable& ha = acTab[c.ta];
        if (ah == 0) {
            if (eobRun) { eobRun--; return; }
            for (int k = ss; k <= se;) {
                int rs = ha.decode(br);
                int r = rs >> 4, s = rs & 15;
                if (s == 0) {
                    if (r < 15) {
                        eobRun = (1 << r) - 1;
                        if (r) eobRun += br.bits(r);
                        break;
                    }
                    k += 16;
                } else {
                    k += r;
                    blk[zz(k++)] = (int16_t)((uint32_t)br.receiveExtend(s) << al);
                }
            }
            return;
        }
        const int16_t bit = (int16_t)(1 << al);
        if (eobRun) {
            eobRun--;
            for (int k = ss; k <= se; k++) {
                int16_t& v = blk[kZigzag[k]];
                if (v != 0) refine(v, bit);
            }
            return;
        }
        for (int k = ss; k <= se;) {
            int rs = ha.decode(br);
            int r = rs >> 4, s = rs & 15;
            int newValue = 0;
            if (s == 0) {
                if (r < 15) {
                    eobRun = (1 << r) - 1;
                    if (r) eobRun += br.bits(r);
                    r = 64;   // run to the end of the band, refining on the way
                }
            } else {
                if (s != 1) fail("bad refinement code");
                newValue = br.bit() ? bit : -bit;
            }
            while (k <= se) {
                int16_t& v = blk[kZigzag[k++]];
                if (v != 0) {
                    refine(v, bit);
                } else {
                    if (r == 0) { v = (int16_t)newValue; break; }
                    r--;
                }
            }
        }
    }
    void decodeBlock(Component& c, int bx, int by) {
        int16_t* blk = &c.coeff[((size_t)by * c.blocksW + bx) * 64];
        if (!progressive) blockSequential(c, blk);
        else if (ss == 0) blockDCProgressive(c, blk);
        else blockACProgressive(c, blk);
    }

    // returns false when the scan has to stop (restart expected but something else found)
    bool afterMCU(int& todo) {
        if (--todo > 0) return true;
        if (br.count < 24) br.fill();
        if (br.marker < 0xD0 || br.marker > 0xD7) return false;
        resetEntropy();
        todo = restartInterval ? restartInterval : 0x7FFFFFFF;
        return true;
    }
    void readScanData() {
        br.src = &src;
        resetEntropy();
        int todo = restartInterval ? restartInterval : 0x7FFFFFFF;
        if (scanN == 1) {
            Component& c = comp[order[0]];
            const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
            for (int by = 0; by < bh; by++)
                for (int bx = 0; bx < bw; bx++) {
                    decodeBlock(c, bx, by);
                    if (!afterMCU(todo)) return;
                }
        } else {
            for (int my = 0; my < mcuY; my++)
                for (int mx = 0; mx < mcuX; mx++) {
                    for (int k = 0; k < scanN; k++) {
                        Component& c = comp[order[k]];
                        for (int v = 0; v < c.v; v++)
                            for (int h = 0; h < c.h; h++) decodeBlock(c, mx * c.h + h, my * c.v + v);
                    }
                    if (!afterMCU(todo)) return;
                }
        }
    }

    void parse() {
        if (nextMarker() != 0xD8) fail("no SOI");
        for (;;) {
            int m = nextMarker();
            if (m < 0) {
                if (src.eof()) {
                    if (haveFrame) return;   // missing EOI: keep what was decoded
                    fail("no SOF");
                }
                continue;
            }
            if (m == 0xD9) return;
            switch (m) {
            case 0xC0: case 0xC1: case 0xC2:
                if (haveFrame) fail("second SOF");
                readSOF(m);
                break;
            case 0xC4: readDHT(); break;
            case 0xDB: readDQT(); break;
            case 0xDD:
                if (src.u16() != 4) fail("bad DRI len");
                restartInterval = src.u16();
                break;
            case 0xDA:
                readSOS();
                readScanData();
                if (br.marker < 0) {
                    // skip whatever is left of the segment up to the next marker
                    while (!src.eof()) {
                        if (src.u8() == 0xFF) {
                            int c = src.u8();
                            while (c == 0xFF) c = src.u8();
                            if (c != 0) { br.marker = c; break; }
                        }
                    }
                } else if (br.marker >= 0xD0 && br.marker <= 0xD7) {
                    br.marker = -1;
                }
                break;
            case 0xDC: src.skip(src.u16() - 2); break;   // DNL
            default:
                if ((m >= 0xE0 && m <= 0xEF) || m == 0xFE) readAPPorCOM(m);
                else if (m >= 0xC3 && m <= 0xCF) fail("unsupported JPEG process (arithmetic / lossless / hierarchical)");
                else fail("unknown marker");
            }
        }
    }

    void reconstructPlanes() {
        int16_t deq[64];
        for (int i = 0; i < ncomp; i++) {
            Component& c = comp[i];
            c.plane.assign((size_t)c.stride() * c.blocksH * 8, 0);
            const uint16_t* q = quant[c.tq];
            for (int by = 0; by < c.blocksH; by++)
                for (int bx = 0; bx < c.blocksW; bx++) {
                    const int16_t* blk = &c.coeff[((size_t)by * c.blocksW + bx) * 64];
                    for (int k = 0; k < 64; k++) deq[k] = (int16_t)(blk[k] * q[k]);   // wraps to 16 bits like the reference
                    idctBlock(deq, &c.plane[(size_t)by * 8 * c.stride() + bx * 8], c.stride());
                }
        }
    }

    // One full-resolution row of component c for output row j.  stb_image.h:3549-3590 pairs the rows through a
    // small state machine (ystep/ypos/line0/line1); in closed form: near = j / vs, and for vs == 2 the far row
    // is the previous source row on even j and the next one (clamped to the component height) on odd j.
    const uint8_t* upsampleRow(const Component& c, int j, uint8_t* tmp) const {
        const int hs = hMax / c.h, vs = vMax / c.v;
        const int w = (width + hs - 1) / hs;
        int nearRow = j / vs, farRow = nearRow;
        if (vs == 2) farRow = (j & 1) ? nearRow + 1 : nearRow - 1;
        if (farRow < 0) farRow = 0;
        if (nearRow > c.y - 1) nearRow = c.y - 1;
        if (farRow > c.y - 1) farRow = c.y - 1;
        const uint8_t* n = &c.plane[(size_t)nearRow * c.stride()];
        const uint8_t* f = &c.plane[(size_t)farRow * c.stride()];
        if (hs == 1 && vs == 1) return n;
        if (hs == 1 && vs == 2) {
            for (int i = 0; i < w; i++) tmp[i] = (uint8_t)((3 * n[i] + f[i] + 2) >> 2);
        } else if (hs == 2 && vs == 1) {
            if (w == 1) { tmp[0] = tmp[1] = n[0]; return tmp; }
            tmp[0] = n[0];
            tmp[1] = (uint8_t)((3 * n[0] + n[1] + 2) >> 2);
            for (int i = 1; i + 1 < w; i++) {
                tmp[2 * i] = (uint8_t)((3 * n[i] + n[i - 1] + 2) >> 2);
                tmp[2 * i + 1] = (uint8_t)((3 * n[i] + n[i + 1] + 2) >> 2);
            }
            // the reference weights the second-to-last output towards sample w-2, not w-1 (stb_image.h:3165)
            tmp[2 * w - 2] = (uint8_t)((3 * n[w - 2] + n[w - 1] + 2) >> 2);
            tmp[2 * w - 1] = n[w - 1];
        } else if (hs == 2 && vs == 2) {
            if (w == 1) { tmp[0] = tmp[1] = (uint8_t)((3 * n[0] + f[0] + 2) >> 2); return tmp; }
            int prev = 3 * n[0] + f[0];
            tmp[0] = (uint8_t)((prev + 2) >> 2);
            for (int i = 1; i < w; i++) {
                const int cur = 3 * n[i] + f[i];
                tmp[2 * i - 1] = (uint8_t)((3 * prev + cur + 8) >> 4);
                tmp[2 * i] = (uint8_t)((3 * cur + prev + 8) >> 4);
                prev = cur;
            }
            tmp[2 * w - 1] = (uint8_t)((prev + 2) >> 2);
        } else {
            for (int i = 0; i < w; i++)
                for (int k = 0; k < hs; k++) tmp[i * hs + k] = n[i];
        }
        return tmp;
    }

    // stb_image.h:3317: ((int)(x * 4096.0f + 0.5f)) << 8
    static constexpr int fix20(float x) { return ((int)(x * 4096.0f + 0.5f)) << 8; }
    static void ycc(uint8_t* out, int y, int cb, int cr) {
        const int yf = (y << 20) + (1 << 19);
        cb -= 128; cr -= 128;
        int r = yf + cr * fix20(1.40200f);
        int g = yf + cr * -fix20(0.71414f) + (int)((uint32_t)(cb * -fix20(0.34414f)) & 0xFFFF0000u);
        int b = yf + cb * fix20(1.77200f);
        out[0] = clamp255(r >> 20);
        out[1] = clamp255(g >> 20);
        out[2] = clamp255(b >> 20);
    }
    // (x*y)/255 rounded, stb_image.h:3519
    static uint8_t mul255(int x, int y) { unsigned t = (unsigned)(x * y + 128); return (uint8_t)((t + (t >> 8)) >> 8); }

    void toRGB(Image* img) {
        img->width = width; img->height = height;
        img->fileChannels = ncomp >= 3 ? 3 : 1;
        img->rgb.resize((size_t)width * height * 3);
        const bool isRGB = ncomp == 3 && (rgbIds == 3 || (adobeTransform == 0 && !jfif));
        std::vector<uint8_t> tmp[4];
        for (int i = 0; i < ncomp; i++) tmp[i].resize((size_t)width + 8 + 4 * 8);
        const uint8_t* row[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int j = 0; j < height; j++) {
            for (int i = 0; i < ncomp; i++) row[i] = upsampleRow(comp[i], j, tmp[i].data());
            uint8_t* out = &img->rgb[(size_t)j * width * 3];
            for (int x = 0; x < width; x++, out += 3) {
                if (ncomp == 1) {
                    out[0] = out[1] = out[2] = row[0][x];
                } else if (ncomp == 3) {
                    if (isRGB) { out[0] = row[0][x]; out[1] = row[1][x]; out[2] = row[2][x]; }
                    else ycc(out, row[0][x], row[1][x], row[2][x]);
                } else if (adobeTransform == 0) {          // CMYK
                    const int k = row[3][x];
                    out[0] = mul255(row[0][x], k); out[1] = mul255(row[1][x], k); out[2] = mul255(row[2][x], k);
                } else if (adobeTransform == 2) {          // YCCK
                    const int k = row[3][x];
                    ycc(out, row[0][x], row[1][x], row[2][x]);
                    out[0] = mul255(255 - out[0], k); out[1] = mul255(255 - out[1], k); out[2] = mul255(255 - out[2], k);
                } else {                                   // YCbCr + a fourth channel that is ignored
                    ycc(out, row[0][x], row[1][x], row[2][x]);
                }
            }
        }
    }
};

}  // namespace detail

inline bool IsJpeg(const uint8_t* data, size_t n) { return n >= 3 && data[0] == 0xFF && data[1] == 0xD8 && data[2] == 0xFF; }

// Throws std::runtime_error on a stream it cannot decode.
inline void Decode(const uint8_t* data, size_t n, Image* out) {
    std::unique_ptr<detail::Decoder> d(new detail::Decoder());
    d->src = detail::ByteSource{data, data + n};
    memset(d->quant, 0, sizeof d->quant);
    d->parse();
    if (!d->haveFrame) detail::fail("no SOF");
    d->reconstructPlanes();
    d->toRGB(out);
}

}  // namespace jpeg
}  // namespace evplp_host
