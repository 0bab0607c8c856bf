// pngdecode.h -- PNG decoder for the host loader (the conference error metric needs
// scene/conference/conference_mask.png, scene/conference/README.md:1-2; the reference reads images through stb_image,
// rtcommon.h:139-194).  Own code: zlib container + inflate (stored / fixed / dynamic Huffman blocks, RFC 1950 / 1951), the five
// PNG row filters, colour types 0 / 2 / 3 / 4 / 6 at 8 or 16 bits (1 / 2 / 4 bits for grey and palette), plain or Adam7-interlaced.
// Output: top-down 8-bit samples with the file's channel count expanded to `wantChannels` like stbi_load(path, .., want)
// does (grey -> rgb replication, alpha dropped or set to 255; 16-bit samples keep their high byte).
#pragma once
#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace evplp_host {
namespace png {

struct Image {
    int width = 0, height = 0, fileChannels = 0, channels = 0;
    std::vector<uint8_t> pixels;   // top-down, `channels` bytes per pixel
};

namespace detail {

struct BitReader {
    const uint8_t* p; size_t n, pos = 0; uint32_t buf = 0; int cnt = 0;
    BitReader(const uint8_t* data, size_t size) : p(data), n(size) {}
    uint32_t bits(int k) {
        while (cnt < k) {
            if (pos >= n) throw std::runtime_error("png: truncated zlib stream");
            buf |= (uint32_t)p[pos++] << cnt; cnt += 8;
        }
        const uint32_t v = buf & ((k == 32) ? 0xffffffffu : ((1u << k) - 1u));
        buf >>= k; cnt -= k;
        return v;
    }
    void alignByte() { buf = 0; cnt = 0; }
};

// canonical Huffman code over `n` symbols with the given lengths (RFC 1951 3.2.2), decoded bit by bit
struct Huffman {
    uint16_t count[16] = {0}, symbol[320];
    void build(const uint8_t* lengths, int n) {
        memset(count, 0, sizeof(count));
        for (int i = 0; i < n; i++) count[lengths[i]]++;
        count[0] = 0;
        uint16_t offs[16]; offs[1] = 0;
        for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
        for (int i = 0; i < n; i++) if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
    }
    int decode(BitReader& br) const {
        int code = 0, first = 0, index = 0;
        for (int l = 1; l <= 15; l++) {
            code |= (int)br.bits(1);
            const int c = count[l];
            if (code - c < first) return symbol[index + (code - first)];
            index += c; first += c; first <<= 1; code <<= 1;
        }
        throw std::runtime_error("png: bad Huffman code");
    }
};

inline std::vector<uint8_t> inflate(const uint8_t* data, size_t size, size_t expected) {
    if (size < 6) throw std::runtime_error("png: zlib stream too short");
    if ((data[0] & 0x0f) != 8 || ((data[0] << 8 | data[1]) % 31) != 0 || (data[1] & 0x20)) throw std::runtime_error("png: bad zlib header");
    BitReader br(data + 2, size - 2);
    std::vector<uint8_t> out;
    out.reserve(expected);
    static const uint16_t lenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t lenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t distBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t distExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    bool last = false;
    while (!last) {
        last = br.bits(1) != 0;
        const uint32_t type = br.bits(2);
        if (type == 0) {
            br.alignByte();
            if (br.pos + 4 > br.n) throw std::runtime_error("png: truncated stored block");
            const uint32_t len = br.p[br.pos] | (br.p[br.pos + 1] << 8), nlen = br.p[br.pos + 2] | (br.p[br.pos + 3] << 8);
            br.pos += 4;
            if ((len ^ 0xffffu) != nlen || br.pos + len > br.n) throw std::runtime_error("png: bad stored block");
            out.insert(out.end(), br.p + br.pos, br.p + br.pos + len);
            br.pos += len;
            continue;
        }
        if (type == 3) throw std::runtime_error("png: bad block type");
        Huffman lit, dist;
        if (type == 1) {
            uint8_t l[288];
            for (int i = 0; i < 144; i++) l[i] = 8;
            for (int i = 144; i < 256; i++) l[i] = 9;
            for (int i = 256; i < 280; i++) l[i] = 7;
            for (int i = 280; i < 288; i++) l[i] = 8;
            lit.build(l, 288);
            uint8_t d[30];
            for (int i = 0; i < 30; i++) d[i] = 5;
            dist.build(d, 30);
        } else {
            const int hlit = (int)br.bits(5) + 257, hdist = (int)br.bits(5) + 1, hclen = (int)br.bits(4) + 4;
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t cl[19] = {0};
            for (int i = 0; i < hclen; i++) cl[order[i]] = (uint8_t)br.bits(3);
            Huffman clh; clh.build(cl, 19);
            uint8_t lengths[320] = {0};
            int i = 0;
            while (i < hlit + hdist) {
                const int sym = clh.decode(br);
                if (sym < 16) { lengths[i++] = (uint8_t)sym; continue; }
                int rep; uint8_t val = 0;
                if (sym == 16) { if (i == 0) throw std::runtime_error("png: bad length repeat"); val = lengths[i - 1]; rep = 3 + (int)br.bits(2); }
                else if (sym == 17) rep = 3 + (int)br.bits(3);
                else rep = 11 + (int)br.bits(7);
                if (i + rep > hlit + hdist) throw std::runtime_error("png: length repeat overruns");
                while (rep--) lengths[i++] = val;
            }
            lit.build(lengths, hlit);
            dist.build(lengths + hlit, hdist);
        }
        for (;;) {
            const int sym = lit.decode(br);
            if (sym < 256) { out.push_back((uint8_t)sym); continue; }
            if (sym == 256) break;
            if (sym > 285) throw std::runtime_error("png: bad length symbol");
            const int len = lenBase[sym - 257] + (int)br.bits(lenExtra[sym - 257]);
            const int ds = dist.decode(br);
            if (ds > 29) throw std::runtime_error("png: bad distance symbol");
            const size_t d = distBase[ds] + br.bits(distExtra[ds]);
            if (d > out.size()) throw std::runtime_error("png: distance beyond the window");
            const size_t start = out.size() - d;
            for (int k = 0; k < len; k++) out.push_back(out[start + k]);
        }
    }
    return out;
}

inline uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

}  // namespace detail

inline bool IsPng(const uint8_t* data, size_t size) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    return size >= 8 && memcmp(data, sig, 8) == 0;
}

inline Image Decode(const uint8_t* data, size_t size, int wantChannels = 0) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (size < 8 || memcmp(data, sig, 8) != 0) throw std::runtime_error("png: not a PNG file");
    size_t p = 8;
    int w = 0, h = 0, depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, palette, trns;
    bool end = false;
    while (!end) {
        if (p + 12 > size) throw std::runtime_error("png: truncated chunk");
        const uint32_t len = detail::be32(data + p);
        const char* type = (const char*)data + p + 4;
        if (p + 12 + (size_t)len > size) throw std::runtime_error("png: chunk overruns the file");
        const uint8_t* body = data + p + 8;
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) throw std::runtime_error("png: bad IHDR");
            w = (int)detail::be32(body); h = (int)detail::be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
            if (w <= 0 || h <= 0 || (int64_t)w * h > (1 << 28) || body[10] != 0 || body[11] != 0) throw std::runtime_error("png: unsupported header");
        } else if (!memcmp(type, "PLTE", 4)) palette.assign(body, body + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(body, body + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(type, "IEND", 4)) end = true;
        p += 12 + (size_t)len;
    }
    if (ctype < 0 || idat.empty()) throw std::runtime_error("png: missing IHDR / IDAT");
    if (interlace > 1) throw std::runtime_error("png: unknown interlace method");
    int samples;
    switch (ctype) {
        case 0: samples = 1; break; case 2: samples = 3; break; case 3: samples = 1; break; case 4: samples = 2; break; case 6: samples = 4; break;
        default: throw std::runtime_error("png: bad colour type");
    }
    if (!(depth == 8 || depth == 16 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4))) || (ctype == 3 && depth == 16))
        throw std::runtime_error("png: unsupported bit depth");
    const size_t bpp = (size_t)(samples * depth + 7) / 8;                 // filter unit in bytes
    // the reduced images the stream holds: the whole image, or the seven Adam7 passes (x0, y0, dx, dy)
    static const int adam7[7][4] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    static const int whole[1][4] = {{0, 0, 1, 1}};
    const int (*passes)[4] = interlace ? adam7 : whole;
    const int numPasses = interlace ? 7 : 1;
    size_t expected = 0;
    for (int p = 0; p < numPasses; p++) {
        const size_t pw = (size_t)(w - passes[p][0] + passes[p][2] - 1) / passes[p][2], ph = (size_t)(h - passes[p][1] + passes[p][3] - 1) / passes[p][3];
        if (w > passes[p][0] && h > passes[p][1]) expected += (((pw * samples * depth + 7) / 8) + 1) * ph;
    }
    std::vector<uint8_t> raw = detail::inflate(idat.data(), idat.size(), expected);
    if (raw.size() < expected) throw std::runtime_error("png: image data too short");
    // un-filter pass by pass and scatter the samples (8 bits, or 16 for 16-bit files; sub-byte samples unscaled) to their pixels
    std::vector<uint16_t> samp((size_t)w * h * samples);
    size_t off = 0;
    for (int p = 0; p < numPasses; p++) {
        if (!(w > passes[p][0] && h > passes[p][1])) continue;
        const int pw = (w - passes[p][0] + passes[p][2] - 1) / passes[p][2], ph = (h - passes[p][1] + passes[p][3] - 1) / passes[p][3];
        const size_t rowBytes = ((size_t)pw * samples * depth + 7) / 8;
        std::vector<uint8_t> prev(rowBytes, 0);
        for (int r = 0; r < ph; r++) {
            uint8_t* row = raw.data() + off;
            off += rowBytes + 1;
            const int f = row[0];
            uint8_t* x = row + 1;
            for (size_t i = 0; i < rowBytes; i++) {
                const int a = i >= bpp ? x[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
                int pred;
                switch (f) {
                    case 0: pred = 0; break;
                    case 1: pred = a; break;
                    case 2: pred = b; break;
                    case 3: pred = (a + b) >> 1; break;
                    case 4: {
                        const int pp = a + b - c, pa = abs(pp - a), pb = abs(pp - b), pc = abs(pp - c);
                        pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                        break;
                    }
                    default: throw std::runtime_error("png: bad filter type");
                }
                x[i] = (uint8_t)(x[i] + pred);
            }
            memcpy(prev.data(), x, rowBytes);
            const int y = passes[p][1] + r * passes[p][3];
            for (int cx = 0; cx < pw; cx++) {
                uint16_t* o = samp.data() + ((size_t)y * w + (passes[p][0] + cx * passes[p][2])) * samples;
                for (int k = 0; k < samples; k++) {
                    if (depth == 8) o[k] = x[(size_t)cx * samples + k];
                    else if (depth == 16) o[k] = (uint16_t)(x[((size_t)cx * samples + k) * 2] << 8 | x[((size_t)cx * samples + k) * 2 + 1]);
                    else {
                        const size_t bit = (size_t)cx * depth;
                        o[k] = (uint16_t)((x[bit >> 3] >> (8 - depth - (int)(bit & 7))) & ((1 << depth) - 1));
                    }
                }
            }
        }
    }
    Image img;
    img.width = w; img.height = h;
    img.fileChannels = ctype == 3 ? (trns.empty() ? 3 : 4) : samples;
    img.channels = wantChannels ? wantChannels : img.fileChannels;
    if (img.channels < 1 || img.channels > 4) throw std::runtime_error("png: bad channel request");
    img.pixels.resize((size_t)w * h * img.channels);
    const bool greyFile = ctype == 0 || ctype == 4;
    const int sh = depth == 16 ? 8 : 0;
    for (size_t pix = 0; pix < (size_t)w * h; pix++) {
        const uint16_t* q = samp.data() + pix * samples;
        int px[4] = {0, 0, 0, depth == 16 ? 65535 : 255};   // r g b a of this pixel (16-bit files keep 16 bits until the end, like stb)
        auto sample = [&](int k) -> int {                    // grey samples below 8 bits are scaled to 0..255, palette indices are not
            return (depth < 8 && ctype == 0) ? q[k] * 255 / ((1 << depth) - 1) : q[k];
        };
        if (ctype == 0) { px[0] = px[1] = px[2] = sample(0); }
        else if (ctype == 4) { px[0] = px[1] = px[2] = sample(0); px[3] = sample(1); }
        else if (ctype == 2) { px[0] = sample(0); px[1] = sample(1); px[2] = sample(2); }
        else if (ctype == 6) { px[0] = sample(0); px[1] = sample(1); px[2] = sample(2); px[3] = sample(3); }
        else {
            const size_t idx = (size_t)sample(0);
            if (idx * 3 + 2 >= palette.size()) throw std::runtime_error("png: palette index out of range");
            px[0] = palette[idx * 3]; px[1] = palette[idx * 3 + 1]; px[2] = palette[idx * 3 + 2];
            px[3] = idx < trns.size() ? trns[idx] : 255;
        }
        uint8_t* o = img.pixels.data() + pix * img.channels;
        // stb's channel conversion (luma for rgb -> grey, replication for grey -> rgb) runs on the file's sample width;
        // 16-bit results are then cut to their high byte
        const int luma = greyFile ? px[0] : (((px[0] * 77 + px[1] * 150 + px[2] * 29) >> 8) & (depth == 16 ? 0xffff : 0xff));
        switch (img.channels) {
            case 1: o[0] = (uint8_t)(luma >> sh); break;
            case 2: o[0] = (uint8_t)(luma >> sh); o[1] = (uint8_t)(px[3] >> sh); break;
            case 3: o[0] = (uint8_t)(px[0] >> sh); o[1] = (uint8_t)(px[1] >> sh); o[2] = (uint8_t)(px[2] >> sh); break;
            default: o[0] = (uint8_t)(px[0] >> sh); o[1] = (uint8_t)(px[1] >> sh); o[2] = (uint8_t)(px[2] >> sh); o[3] = (uint8_t)(px[3] >> sh); break;
        }
    }
    return img;
}

inline Image DecodeFile(const std::string& path, int wantChannels = 0) {
    std::ifstream is(path, std::ios::binary);
    if (!is.is_open()) throw std::runtime_error("png: cannot open " + path);
    std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());
    return Decode(bytes.data(), bytes.size(), wantChannels);
}

}  // namespace png
}  // namespace evplp_host
