"""Generates tests/golden/jpeg/*.jpg (Pillow-encoded synthetic images) and expected.npz: the pixels that the
reference's own decoder -- stb_image.h v2.16 from /root/reference, built by `make -C oracle ref` into
oracle/_ref/libstb_ref.so -- returns for stbi_load(.., 3).  Run from the repo root in the build container:

    make -C oracle ref && python tests/golden/make_jpeg_golden.py
"""
import ctypes as C
import io
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "jpeg")
ROOT = os.path.dirname(os.path.dirname(HERE))


def stb_decode(lib, data):
    w, h, ch = C.c_int(), C.c_int(), C.c_int()
    buf = (C.c_ubyte * len(data)).from_buffer_copy(data)
    lib.stbi_set_flip_vertically_on_load(0)
    p = lib.stbi_load_from_memory(buf, len(data), C.byref(w), C.byref(h), C.byref(ch), 3)
    if not p:
        raise RuntimeError(lib.stbi_failure_reason())
    out = np.ctypeslib.as_array(p, (h.value, w.value, 3)).copy()
    lib.stbi_image_free(p)
    return out, ch.value


def load_stb():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libstb_ref.so"))
    lib.stbi_load_from_memory.restype = C.POINTER(C.c_ubyte)
    lib.stbi_failure_reason.restype = C.c_char_p
    return lib


def synth(rng, w, h, mode):
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([(x * 7 + y * 3) % 256, (x * y) % 256, (255 * np.sin(x / 5.0) * np.cos(y / 3.0)).astype(int) % 256], -1)
    img = (img + rng.integers(-40, 40, img.shape)).clip(0, 255).astype(np.uint8)
    im = Image.fromarray(img)
    return im if mode == "RGB" else im.convert(mode)


# name: (w, h, mode, save kwargs)
CASES = {
    "base_444_q90": (37, 29, "RGB", dict(quality=90, subsampling=0)),
    "base_422_q75": (37, 29, "RGB", dict(quality=75, subsampling=1)),
    "base_420_q75": (37, 29, "RGB", dict(quality=75, subsampling=2)),
    "base_420_q100_opt": (64, 48, "RGB", dict(quality=100, subsampling=2, optimize=True)),
    "base_420_q20": (129, 65, "RGB", dict(quality=20, subsampling=2)),
    "base_420_restart": (70, 50, "RGB", dict(quality=80, subsampling=2, restart_marker_blocks=3)),
    "base_444_restart_rows": (40, 40, "RGB", dict(quality=80, subsampling=0, restart_marker_rows=1)),
    "prog_444_q90": (37, 29, "RGB", dict(quality=90, subsampling=0, progressive=True)),
    "prog_422_q60": (37, 29, "RGB", dict(quality=60, subsampling=1, progressive=True)),
    "prog_420_q95": (129, 65, "RGB", dict(quality=95, subsampling=2, progressive=True)),
    "prog_420_restart": (70, 50, "RGB", dict(quality=80, subsampling=2, progressive=True, restart_marker_blocks=2)),
    "grey_base": (33, 17, "L", dict(quality=80)),
    "grey_prog": (33, 17, "L", dict(quality=80, progressive=True)),
    "cmyk_base": (20, 12, "CMYK", dict(quality=85)),
    "tiny_1x1": (1, 1, "RGB", dict(quality=90, subsampling=2)),
    "tiny_1x1_prog": (1, 1, "RGB", dict(quality=90, subsampling=2, progressive=True)),
    "thin_3x40_420": (3, 40, "RGB", dict(quality=70, subsampling=2)),
    "wide_40x3_422": (40, 3, "RGB", dict(quality=70, subsampling=1)),
    "exact_16x16_420": (16, 16, "RGB", dict(quality=50, subsampling=2)),
    "odd_17x33_420_prog": (17, 33, "RGB", dict(quality=88, subsampling=2, progressive=True)),
}


# Sampling layouts Pillow cannot write: re-label the luma sampling factors of a 32x32 file whose MCU count stays
# the same, so the entropy stream still decodes block for block but lands in a different geometry.  This is what
# reaches the nearest-neighbour and h1v2 upsamplers.
PATCHED = {
    "patched_440_h1v2": ("RGB", dict(quality=85, subsampling=1), 0x21, 0x12),
    "patched_411_h4v1": ("RGB", dict(quality=85, subsampling=2), 0x22, 0x41),
    "patched_h1v4": ("RGB", dict(quality=85, subsampling=2), 0x22, 0x14),
    "patched_440_prog": ("RGB", dict(quality=85, subsampling=1, progressive=True), 0x21, 0x12),
}


def patch_luma_sampling(data, old, new):
    data = bytearray(data)
    for marker in (b"\xff\xc0", b"\xff\xc2"):
        i = data.find(marker)
        if i >= 0:
            assert data[i + 11] == old, hex(data[i + 11])
            data[i + 11] = new
            return bytes(data)
    raise AssertionError("no SOF")


def main():
    os.makedirs(OUT, exist_ok=True)
    lib = load_stb()
    rng = np.random.default_rng(20260101)
    expected = {}
    for name, (w, h, mode, kw) in CASES.items():
        bio = io.BytesIO()
        synth(rng, w, h, mode).save(bio, "JPEG", **kw)
        data = bio.getvalue()
        with open(os.path.join(OUT, name + ".jpg"), "wb") as f:
            f.write(data)
        px, ch = stb_decode(lib, data)
        expected[name] = px
        expected[name + "__channels"] = np.int32(ch)
        print(f"{name:28s} {w}x{h} {len(data)} bytes, file channels {ch}")
    for name, (mode, kw, old, new) in PATCHED.items():
        bio = io.BytesIO()
        synth(rng, 32, 32, mode).save(bio, "JPEG", **kw)
        data = patch_luma_sampling(bio.getvalue(), old, new)
        with open(os.path.join(OUT, name + ".jpg"), "wb") as f:
            f.write(data)
        px, ch = stb_decode(lib, data)
        expected[name] = px
        expected[name + "__channels"] = np.int32(ch)
        print(f"{name:28s} 32x32 {len(data)} bytes, file channels {ch}")
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **expected)


if __name__ == "__main__":
    main()
