"""K10 (upg_png_encode, csrc/upg_png.cu): PNG encoding of the chunk textures on the device -- the third item of SURVEY.md 8f n2
(Image::exportFile in the generator's chunk tasks, sources/generator.cpp:135-139).  Parity bar: the file is a valid PNG and ANY decoder
returns the pixels bit for bit -- checked with PIL (libpng + zlib) and, independently, by walking the chunks by hand (CRC-32 of every
chunk, zlib.decompress of the concatenated IDAT data, Adler-32, un-filtering in numpy)."""
import io
import struct
import zlib

import numpy as np
import pytest

from unnatural_planets_b200 import Planet, png_encode, synth

pytestmark = pytest.mark.gpu


def decode_by_hand(data):
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    at, idat, ihdr, kinds = 8, [], None, []
    while at < len(data):
        (n,) = struct.unpack(">I", data[at:at + 4])
        kind, body = data[at + 4:at + 8], data[at + 8:at + 8 + n]
        (crc,) = struct.unpack(">I", data[at + 8 + n:at + 12 + n])
        assert zlib.crc32(kind + body) == crc, (kind, at)
        kinds.append(kind)
        if kind == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", body)
        elif kind == b"IDAT":
            idat.append(body)
        at += 12 + n
    assert at == len(data) and kinds[0] == b"IHDR" and kinds[-1] == b"IEND" and set(kinds[1:-1]) == {b"IDAT"}
    w, h, depth, ctype, comp, flt, lace = ihdr
    assert depth == 8 and comp == 0 and flt == 0 and lace == 0
    ch = {0: 1, 4: 2, 2: 3, 6: 4}[ctype]
    raw = zlib.decompress(b"".join(idat))  # checks the Adler-32 of the stream too
    rows = np.frombuffer(raw, np.uint8).reshape(h, w * ch + 1)
    out = np.zeros((h, w * ch), np.int32)
    types = rows[:, 0]
    for r in range(h):
        f = rows[r, 1:].astype(np.int32)
        up = out[r - 1] if r else np.zeros(w * ch, np.int32)
        t = types[r]
        if t == 0:
            cur = f
        elif t == 2:
            cur = (f + up) & 255
        else:
            cur = np.zeros(w * ch, np.int32)
            for i in range(w * ch):
                a = cur[i - ch] if i >= ch else 0
                b = up[i]
                c = up[i - ch] if i >= ch else 0
                if t == 1:
                    p = a
                elif t == 3:
                    p = (a + b) >> 1
                else:
                    pp = a + b - c
                    pa, pb, pc = abs(pp - a), abs(pp - b), abs(pp - c)
                    p = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[i] = (f[i] + p) & 255
        out[r] = cur
    return out.astype(np.uint8).reshape(h, w, ch), types


def pil_decode(data):
    from PIL import Image

    return np.asarray(Image.open(io.BytesIO(data)))


@pytest.mark.parametrize("shape", [(1, 1, 1), (3, 5, 3), (17, 33, 4), (64, 64, 2), (40, 2731, 3), (301, 1, 1)])
def test_small_images_decode_by_hand_and_with_pil(shape):
    rng = np.random.default_rng(sum(shape))
    h, w, ch = shape
    # smooth content (filters and matches matter) with noise on top (literals) and a constant band (long matches)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([(x * (k + 1) + y * 3 + rng.integers(0, 4, (h, w))) & 255 for k in range(ch)], axis=2).astype(np.uint8)
    img[h // 3: h // 2] = 77
    data = png_encode(img)
    got, types = decode_by_hand(data)
    assert np.array_equal(got, img)
    pil = pil_decode(data)
    assert np.array_equal(pil.reshape(h, w, ch), img)


@pytest.mark.parametrize("ch", [1, 3, 4])
def test_random_noise_and_constant_images(ch):
    rng = np.random.default_rng(ch)
    noise = rng.integers(0, 256, (777, 1000, ch), dtype=np.uint8)  # incompressible: literals only, 9-bit codes for half of them
    data = png_encode(noise)
    assert np.array_equal(pil_decode(data).reshape(noise.shape), noise)
    assert len(data) < noise.size * 1.15
    flat = np.full((512, 2048, ch), 200, np.uint8)  # maximal matches
    data = png_encode(flat)
    assert np.array_equal(pil_decode(data).reshape(flat.shape), flat)
    assert len(data) < flat.size // 20


def test_baked_textures_round_trip_and_compression():
    """the real payloads: albedo / pbr / normal map of a 2048^2 land bake, from host arrays and straight from device memory"""
    import torch

    p = Planet(shape="sphere", elevation="lakes", grid=200)
    m = p.field_meshes_clean(("land",))["land"]
    uv = synth.octahedral_uv(m["positions"])
    idx = synth.drop_folded_triangles(uv, m["indices"])
    T = 2048
    images = p.bake_export(0, m["positions"], m["normals"], uv, idx, T, T)
    for name, img in zip(("albedo", "pbr", "normal"), images):
        data = png_encode(img)
        assert np.array_equal(pil_decode(data), img), name
        dev = torch.from_numpy(img).cuda()
        assert png_encode(dev) == data, name  # device input: same file
        z1 = len(zlib.compress(img.tobytes(), 1))
        # a fixed-code 8 KiB-window deflate of filtered rows: never absurdly larger than zlib level 1 over the raw rows
        assert len(data) < 1.6 * z1 + 4096, (name, len(data), z1)


def test_large_image_and_errors():
    from unnatural_planets_b200 import UpgError, capi
    import ctypes as C

    rng = np.random.default_rng(9)
    y, x = np.mgrid[0:4096, 0:4096]
    img = np.stack([(x // 3 + y // 5) & 255, (x ^ y) & 255, rng.integers(0, 8, (4096, 4096))], axis=2).astype(np.uint8)
    data = png_encode(img)
    assert np.array_equal(pil_decode(data), img)
    L = capi.load_library()
    size = C.c_uint64(0)
    out = np.zeros(64, np.uint8)
    small = np.zeros((8, 8, 3), np.uint8)
    rc = L.upg_png_encode(0, small.ctypes.data_as(C.c_void_p), 8, 8, 3, out.ctypes.data_as(C.c_void_p), 64, C.byref(size))
    assert rc == 1 and b"too small" in L.upg_last_error()
    assert L.upg_png_encode(0, small.ctypes.data_as(C.c_void_p), 8, 8, 5, out.ctypes.data_as(C.c_void_p), 64, C.byref(size)) == 1
    assert L.upg_png_encode(0, small.ctypes.data_as(C.c_void_p), 0, 8, 3, out.ctypes.data_as(C.c_void_p), 64, C.byref(size)) == 1
