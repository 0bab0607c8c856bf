"""Generates tests/golden/png/*.png (synthetic images, written by the small encoder below so that every colour type, bit depth,
row filter and deflate block type appears) and expected.npz: the samples that the reference's own decoder -- stb_image.h v2.16
from /root/reference, built by `make -C oracle ref` into oracle/_ref/libstb_ref.so -- returns for
stbi_load_from_memory(.., want) with want = 0 (the file's channels) and want = 3 (what RtTexture asks for, rtcommon.h:144).

    make -C oracle ref && python tests/golden/make_png_golden.py
"""
import ctypes as C
import os
import struct
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "png")
ROOT = os.path.dirname(os.path.dirname(HERE))


def chunk(tag, body):
    return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xffffffff)


def paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def encode_png(rows, w, h, depth, ctype, filters, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, palette=None, trns=None, idat_split=0):
    """rows: h byte strings of packed samples; filters: one filter type per row (cycled)."""
    samples = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    bpp = max(1, samples * depth // 8)
    raw = bytearray()
    prev = bytes(len(rows[0]))
    for r in range(h):
        f = filters[r % len(filters)]
        x = rows[r]
        out = bytearray(len(x))
        for i in range(len(x)):
            a = x[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            pred = (0, a, b, (a + b) >> 1, paeth(a, b, c))[f]
            out[i] = (x[i] - pred) & 255
        raw.append(f)
        raw += out
        prev = x
    co = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
    z = co.compress(bytes(raw)) + co.flush()
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0))
    if palette is not None:
        png += chunk(b"PLTE", bytes(palette))
    if trns is not None:
        png += chunk(b"tRNS", bytes(trns))
    png += chunk(b"tEXt", b"Comment\x00synthetic test image")
    if idat_split:
        for i in range(0, len(z), idat_split):
            png += chunk(b"IDAT", z[i:i + idat_split])
    else:
        png += chunk(b"IDAT", z)
    return png + chunk(b"IEND", b"")


ADAM7 = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]


def encode_png_interlaced(img, depth, ctype, filters, level=6, palette=None, trns=None):
    """img: [h, w, samples] integer samples -> Adam7-interlaced PNG (the seven reduced images, each filtered on its own)."""
    h, w, s = img.shape
    bpp = max(1, s * depth // 8)
    raw = bytearray()
    n = 0
    for (x0, y0, dx, dy) in ADAM7:
        sub = img[y0::dy, x0::dx]
        if sub.shape[0] == 0 or sub.shape[1] == 0:
            continue
        rows = pack_rows(sub, depth)
        prev = bytes(len(rows[0]))
        for x in rows:
            f = filters[n % len(filters)]
            n += 1
            out = bytearray(len(x))
            for i in range(len(x)):
                a = x[i - bpp] if i >= bpp else 0
                b = prev[i]
                c = prev[i - bpp] if i >= bpp else 0
                out[i] = (x[i] - (0, a, b, (a + b) >> 1, paeth(a, b, c))[f]) & 255
            raw.append(f)
            raw += out
            prev = x
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 1))
    if palette is not None:
        png += chunk(b"PLTE", bytes(palette))
    if trns is not None:
        png += chunk(b"tRNS", bytes(trns))
    return png + chunk(b"IDAT", zlib.compress(bytes(raw), level)) + chunk(b"IEND", b"")


def pack_rows(img, depth):
    """img: [h, w, samples] integer samples of `depth` bits -> list of packed row byte strings"""
    h, w, s = img.shape
    rows = []
    for r in range(h):
        flat = img[r].reshape(-1)
        if depth == 8:
            rows.append(bytes(flat.astype(np.uint8)))
        elif depth == 16:
            rows.append(flat.astype(">u2").tobytes())
        else:
            bits = 0
            nb = 0
            out = bytearray()
            for v in flat:
                bits = (bits << depth) | int(v)
                nb += depth
                if nb == 8:
                    out.append(bits); bits = 0; nb = 0
            if nb:
                out.append(bits << (8 - nb))
            rows.append(bytes(out))
    return rows


def fixtures():
    rng = np.random.default_rng(20260101)
    out = {}

    def synth(w, h, s, depth):
        y, x = np.mgrid[0:h, 0:w]
        base = np.stack([(x * 7 + y * 3), (x * y + 5), (x * x + 11 * y), (255 - 9 * x - y)], -1)[:, :, :s]
        noise = rng.integers(0, 4, size=(h, w, s))
        img = (base * (1 << max(depth - 8, 0)) // (1 if depth >= 8 else (1 << (8 - depth))) + noise) % (1 << depth)
        return img.astype(np.int64)

    allf = [0, 1, 2, 3, 4]
    out["grey8_allfilters"] = encode_png(pack_rows(synth(37, 23, 1, 8), 8), 37, 23, 8, 0, allf)
    out["greyalpha8_paeth"] = encode_png(pack_rows(synth(19, 11, 2, 8), 8), 19, 11, 8, 4, [4])
    out["rgb8_allfilters"] = encode_png(pack_rows(synth(41, 29, 3, 8), 8), 41, 29, 8, 2, allf)
    out["rgba8_avg_sub"] = encode_png(pack_rows(synth(33, 17, 4, 8), 8), 33, 17, 8, 6, [3, 1])
    out["rgb8_stored_blocks"] = encode_png(pack_rows(synth(24, 24, 3, 8), 8), 24, 24, 8, 2, [0, 2], level=0)
    out["rgb8_fixed_huffman"] = encode_png(pack_rows(synth(24, 24, 3, 8), 8), 24, 24, 8, 2, allf, strategy=zlib.Z_FIXED)
    out["rgba8_split_idat"] = encode_png(pack_rows(synth(64, 48, 4, 8), 8), 64, 48, 8, 6, allf, idat_split=97)
    out["rgb16_up"] = encode_png(pack_rows(synth(21, 13, 3, 16), 16), 21, 13, 16, 2, [2, 4])
    out["grey16_sub"] = encode_png(pack_rows(synth(17, 9, 1, 16), 16), 17, 9, 16, 0, [1])
    out["rgba16_allfilters"] = encode_png(pack_rows(synth(15, 15, 4, 16), 16), 15, 15, 16, 6, allf)
    for d in (1, 2, 4):
        out[f"grey{d}_ragged"] = encode_png(pack_rows(synth(13, 7, 1, d), d), 13, 7, d, 0, [0, 2])
    pal = rng.integers(0, 256, size=16 * 3).tolist()
    out["palette4"] = encode_png(pack_rows(synth(29, 15, 1, 4), 4), 29, 15, 4, 3, [0], palette=pal)
    pal256 = rng.integers(0, 256, size=256 * 3).tolist()
    out["palette8_trns"] = encode_png(pack_rows(synth(31, 19, 1, 8), 8), 31, 19, 8, 3, allf, palette=pal256,
                                      trns=rng.integers(0, 256, size=100).tolist())
    out["palette1"] = encode_png(pack_rows(synth(9, 5, 1, 1), 1), 9, 5, 1, 3, [0], palette=[10, 20, 30, 200, 210, 220])
    out["one_pixel_rgba"] = encode_png([bytes([1, 2, 3, 4])], 1, 1, 8, 6, [4])
    # Adam7-interlaced files (seven reduced images; narrow ones leave some passes empty)
    out["interlaced_rgb8"] = encode_png_interlaced(synth(37, 23, 3, 8), 8, 2, allf)
    out["interlaced_grey8_narrow"] = encode_png_interlaced(synth(3, 19, 1, 8), 8, 0, allf)
    out["interlaced_rgba16"] = encode_png_interlaced(synth(18, 9, 4, 16), 16, 6, [4, 1])
    out["interlaced_grey2"] = encode_png_interlaced(synth(21, 13, 1, 2), 2, 0, [0, 2])
    out["interlaced_palette4_trns"] = encode_png_interlaced(synth(26, 11, 1, 4), 4, 3, [0, 1], palette=pal, trns=[0, 128, 255])
    out["interlaced_one_pixel"] = encode_png_interlaced(synth(1, 1, 3, 8), 8, 2, [0])
    big = synth(128, 96, 3, 8)
    big[:, :, 0] = (big[:, :, 0] // 16) * 16   # long matches: exercises length / distance codes beyond the short ones
    out["rgb8_large_dynamic"] = encode_png(pack_rows(big, 8), 128, 96, 8, 2, allf, level=9)
    return out


def load_stb():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libstb_ref.so"))
    lib.stbi_load_from_memory.restype = C.POINTER(C.c_ubyte)
    lib.stbi_failure_reason.restype = C.c_char_p
    return lib


def stb_decode(lib, data, want):
    w, h, ch = C.c_int(), C.c_int(), C.c_int()
    buf = (C.c_ubyte * len(data)).from_buffer_copy(data)
    lib.stbi_set_flip_vertically_on_load(0)
    p = lib.stbi_load_from_memory(buf, len(data), C.byref(w), C.byref(h), C.byref(ch), want)
    if not p:
        raise RuntimeError(lib.stbi_failure_reason())
    n = want or ch.value
    out = np.ctypeslib.as_array(p, (h.value, w.value, n)).copy()
    lib.stbi_image_free(p)
    return out, ch.value


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    lib = load_stb()
    exp = {}
    for name, data in fixtures().items():
        with open(os.path.join(OUT, name + ".png"), "wb") as f:
            f.write(data)
        for want in (0, 1, 3, 4):
            px, ch = stb_decode(lib, data, want)
            exp[f"{name}__want{want}"] = px
        exp[name + "__channels"] = np.int32(ch)
        print(name, len(data), "bytes", exp[f"{name}__want0"].shape)
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **exp)
