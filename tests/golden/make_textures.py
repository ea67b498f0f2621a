"""Texture-decoder fixtures: writes small image files in every PNG / TGA / JPEG variant the decoders claim
(tests/golden/textures/*), an OBJ + MTL whose materials name them, and tests/golden/textures.npz = the texels the
REFERENCE's loader (unmodified PT_sv5_/Model.cpp + its vendored stb_image, compiled into oracle/_ref) makes of them.

Run in the build container (needs /root/reference for oracle/_ref, PIL for the JPEG encoder):
    python tests/golden/make_textures.py
The files and the .npz are committed; the GPU box never regenerates them."""
import os
import struct
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "textures")


# ---------------------------------------------------------------- PNG writer (every colour type / depth / filter / Adam7)
def _png_rows(a, depth, rng):
    """a: (h, w, chan) integer samples -> filtered scanlines (bytes), filter type chosen per row at random"""
    h, w, chan = a.shape
    if depth == 16:
        raw = a.astype(">u2").tobytes()
        rb = w * chan * 2
    elif depth == 8:
        raw = a.astype(np.uint8).tobytes()
        rb = w * chan
    else:
        per = 8 // depth
        pad = (-w) % per
        b = np.pad(a[..., 0], ((0, 0), (0, pad))).reshape(h, -1, per).astype(np.uint32)
        sh = np.arange(per - 1, -1, -1) * depth
        raw = (b << sh).sum(-1).astype(np.uint8).tobytes()
        rb = (w + pad) // per
    bpp = max(1, chan * depth // 8)
    out = bytearray()
    prev = bytearray(rb)
    for y in range(h):
        cur = bytearray(raw[y * rb:(y + 1) * rb])
        ft = int(rng.integers(0, 5))
        line = bytearray(rb)
        for i in range(rb):
            A = cur[i - bpp] if i >= bpp else 0
            B = prev[i]
            Cc = prev[i - bpp] if i >= bpp else 0
            if ft == 0: p = 0
            elif ft == 1: p = A
            elif ft == 2: p = B
            elif ft == 3: p = (A + B) >> 1
            else:
                pp = A + B - Cc
                pa, pb, pc = abs(pp - A), abs(pp - B), abs(pp - Cc)
                p = A if (pa <= pb and pa <= pc) else (B if pb <= pc else Cc)
            line[i] = (cur[i] - p) & 255
        out.append(ft); out += line
        prev = cur
    return bytes(out)


def write_png(path, a, ctype, depth, interlace=False, palette=None, trns=None, level=6, seed=0, idat_split=0):
    rng = np.random.default_rng(seed)
    h, w = a.shape[:2]
    if a.ndim == 2: a = a[..., None]
    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    body = b""
    if interlace:
        x0, y0, dx, dy = [0, 4, 0, 2, 0, 1, 0], [0, 0, 4, 0, 2, 0, 1], [8, 8, 4, 4, 2, 2, 1], [8, 8, 8, 4, 4, 2, 2]
        for p in range(7):
            sub = a[y0[p]::dy[p], x0[p]::dx[p]]
            if sub.shape[0] and sub.shape[1]: body += _png_rows(sub, depth, rng)
    else:
        body = _png_rows(a, depth, rng)
    z = zlib.compress(body, level)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 1 if interlace else 0))
    out += chunk(b"gAMA", struct.pack(">I", 45455))                  # ancillary chunk: must be ignored
    if palette is not None: out += chunk(b"PLTE", np.asarray(palette, np.uint8).tobytes())
    if trns is not None: out += chunk(b"tRNS", trns)
    if idat_split:
        for i in range(0, len(z), idat_split): out += chunk(b"IDAT", z[i:i + idat_split])
    else:
        out += chunk(b"IDAT", z)
    out += chunk(b"IEND", b"")
    open(path, "wb").write(out)


# ---------------------------------------------------------------- TGA writer
def write_tga(path, px, kind, bits, rle=False, top=False, palette=None, pal_bits=24, idlen=0):
    """px: (h, w, k) uint8 channel values already in FILE byte order per pixel (or (h, w) indices / packed 16-bit)"""
    h, w = px.shape[:2]
    typ = {"rgb": 2, "grey": 3, "pal": 1}[kind] + (8 if rle else 0)
    hdr = struct.pack("<BBBHHBHHHHBB", idlen, 1 if kind == "pal" else 0, typ, 0, len(palette) if palette is not None else 0,
                      pal_bits if kind == "pal" else 0, 0, 0, w, h, bits, (0x20 if top else 0) | (8 if bits == 32 else 0))
    out = bytearray(hdr) + bytes(range(idlen))
    if palette is not None: out += np.asarray(palette, np.uint8).tobytes()
    pix = px.reshape(h * w, -1)
    if not rle:
        out += pix.tobytes()
    else:
        i, n = 0, len(pix)
        while i < n:
            run = 1
            while i + run < n and run < 128 and (pix[i + run] == pix[i]).all(): run += 1
            if run > 1:
                out.append(0x80 | (run - 1)); out += pix[i].tobytes(); i += run
            else:
                lit = 1
                while i + lit < n and lit < 128 and not (pix[i + lit] == pix[i + lit - 1]).all(): lit += 1
                out.append(lit - 1); out += pix[i:i + lit].tobytes(); i += lit
    open(path, "wb").write(bytes(out))


def picture(w, h, seed, chan=3, smooth=True):
    """smooth colour ramps with a few hard edges and some noise: exercises DCT, filters and run lengths alike"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    a = np.zeros((h, w, chan))
    for c in range(chan):
        f = rng.uniform(0.02, 0.3, 4)
        a[..., c] = 127 + 70 * np.sin(f[0] * x + f[1] * y + c) + 40 * np.cos(f[2] * x - f[3] * y)
    if smooth:
        a[h // 3:h // 2, w // 4:w // 2] = rng.uniform(0, 255, chan)
        a += rng.normal(0, 6, a.shape)
    else:
        a = rng.uniform(0, 255, a.shape)
    a[: max(1, h // 6), : max(1, w // 3)] = a[0, 0]                  # a flat region: runs for the RLE / LZ paths
    return np.clip(a, 0, 255).astype(np.uint8)


def main():
    from PIL import Image
    os.makedirs(OUT, exist_ok=True)
    for f in os.listdir(OUT): os.remove(os.path.join(OUT, f))
    names = []
    def P(name): names.append(name); return os.path.join(OUT, name)

    # PNG
    rgb = picture(37, 29, 1); rgba = picture(37, 29, 2, 4); grey = picture(37, 29, 3, 1)
    write_png(P("rgb8.png"), rgb, 2, 8, seed=1)
    write_png(P("rgba8.png"), rgba, 6, 8, seed=2)
    write_png(P("rgba8_adam7.png"), rgba, 6, 8, interlace=True, seed=3)
    write_png(P("grey8.png"), grey, 0, 8, seed=4)
    write_png(P("grey8_key.png"), grey // 32 * 32, 0, 8, trns=struct.pack(">H", 96), seed=5)
    write_png(P("greya8.png"), rgba[..., :2], 4, 8, seed=6)
    write_png(P("rgb8_key.png"), rgb // 64 * 64, 2, 8, trns=struct.pack(">HHH", *[int(v) for v in (rgb // 64 * 64)[5, 7]]), seed=7)
    r16 = (picture(23, 17, 8).astype(np.uint32) * 257 + np.random.default_rng(8).integers(0, 200, (17, 23, 3))) & 0xffff
    write_png(P("rgb16.png"), r16, 2, 16, seed=8)
    write_png(P("rgb16_key.png"), r16 // 4096 * 4096 + 5, 2, 16, trns=struct.pack(">HHH", *[int(v) for v in (r16 // 4096 * 4096 + 5)[3, 3]]), seed=9)
    write_png(P("rgba16_adam7.png"), np.concatenate([r16, r16[..., :1] ^ 0x5aa5], -1), 6, 16, interlace=True, seed=10)
    write_png(P("grey16.png"), r16[..., :1], 0, 16, seed=11)
    write_png(P("greya16.png"), r16[..., :2], 4, 16, seed=12)
    for d in (1, 2, 4):
        g = (picture(21, 11, 20 + d, 1)[..., 0] >> (8 - d)).astype(np.uint8)
        write_png(P(f"grey{d}.png"), g, 0, d, seed=20 + d)
        write_png(P(f"grey{d}_adam7_key.png"), g, 0, d, interlace=True, trns=struct.pack(">H", 1), seed=30 + d)
        pal = np.random.default_rng(40 + d).integers(0, 256, (1 << d, 3))
        write_png(P(f"pal{d}.png"), g, 3, d, palette=pal, seed=40 + d)
    pal = np.random.default_rng(50).integers(0, 256, (200, 3))
    idx = (picture(37, 29, 51, 1)[..., 0].astype(np.uint32) * 200 // 256).astype(np.uint8)
    write_png(P("pal8_trns.png"), idx, 3, 8, palette=pal, trns=bytes(np.random.default_rng(52).integers(0, 256, 120).astype(np.uint8)), seed=51)
    write_png(P("noise_stored.png"), picture(40, 30, 60, 3, smooth=False), 2, 8, level=0, seed=60)
    write_png(P("big_rgb8.png"), picture(151, 103, 61), 2, 8, level=9, seed=61, idat_split=777)      # several dynamic blocks, split IDAT
    write_png(P("one_px.png"), picture(1, 1, 62, 4), 6, 8, seed=62)
    Image.fromarray(rgba, "RGBA").save(P("pil_rgba.png"), optimize=True)

    # TGA (file byte order: B, G, R [, A])
    bgr = rgb[..., ::-1].copy(); bgra = np.concatenate([rgba[..., 2::-1], rgba[..., 3:]], -1).copy()
    write_tga(P("rgb24.tga"), bgr, "rgb", 24)
    write_tga(P("rgb24_top_rle.tga"), bgr // 32 * 32, "rgb", 24, rle=True, top=True)
    write_tga(P("rgba32.tga"), bgra, "rgb", 32, idlen=9)
    write_tga(P("rgba32_rle.tga"), bgra // 64 * 64, "rgb", 32, rle=True)
    p16 = np.random.default_rng(70).integers(0, 65536, (13, 19)).astype("<u2")
    write_tga(P("rgb16.tga"), p16.view(np.uint8).reshape(13, 19, 2), "rgb", 16)
    write_tga(P("rgb15_rle.tga"), (p16 // 4096 * 4096).view(np.uint8).reshape(13, 19, 2), "rgb", 15, rle=True)
    write_tga(P("grey8.tga"), grey, "grey", 8)
    write_tga(P("grey8_rle_top.tga"), grey // 64 * 64, "grey", 8, rle=True, top=True)
    write_tga(P("greya16.tga"), rgba[..., :2].copy(), "grey", 16)
    pal24 = np.random.default_rng(71).integers(0, 256, (64, 3)).astype(np.uint8)
    write_tga(P("pal8_24.tga"), (grey // 4).astype(np.uint8), "pal", 8, palette=pal24, pal_bits=24)
    write_tga(P("pal8_32_rle.tga"), (grey // 16).astype(np.uint8), "pal", 8, rle=True, palette=np.random.default_rng(72).integers(0, 256, (16, 4)).astype(np.uint8), pal_bits=32)
    write_tga(P("pal8_16.tga"), (grey // 8).astype(np.uint8), "pal", 8, palette=np.random.default_rng(73).integers(0, 256, (32, 2)).astype(np.uint8), pal_bits=16)

    # JPEG (PIL / libjpeg encoder)
    big = picture(83, 61, 80); tiny = picture(5, 3, 81)
    Image.fromarray(big).save(P("q90_444.jpg"), quality=90, subsampling=0)
    Image.fromarray(big).save(P("q85_422.jpg"), quality=85, subsampling=1)
    Image.fromarray(big).save(P("q75_420.jpg"), quality=75, subsampling=2)
    Image.fromarray(big).save(P("q60_411.jpg"), quality=60, subsampling="4:1:1")
    Image.fromarray(big).save(P("q80_420_opt.jpg"), quality=80, subsampling=2, optimize=True)
    Image.fromarray(big).save(P("q80_420_prog.jpg"), quality=80, subsampling=2, progressive=True)
    Image.fromarray(big).save(P("q95_444_prog.jpg"), quality=95, subsampling=0, progressive=True)
    Image.fromarray(big).save(P("q80_420_rst.jpg"), quality=80, subsampling=2, restart_marker_blocks=2)
    Image.fromarray(big).save(P("q70_422_prog_rst.jpg"), quality=70, subsampling=1, progressive=True, restart_marker_rows=1)
    Image.fromarray(big[..., 0]).save(P("grey_q85.jpg"), quality=85)
    Image.fromarray(big[..., 0]).save(P("grey_prog.jpg"), quality=85, progressive=True)
    Image.fromarray(tiny).save(P("tiny_420.jpg"), quality=90, subsampling=2)
    Image.fromarray(picture(17, 16, 82)).save(P("w17_420.jpg"), quality=100, subsampling=2)
    Image.fromarray(picture(16, 9, 83)).save(P("h9_422.jpg"), quality=30, subsampling=1)
    Image.fromarray(picture(151, 103, 84, smooth=False)).save(P("noise_q100_444.jpg"), quality=100, subsampling=0)

    # BMP / PPM keep working
    Image.fromarray(rgb).save(P("rgb24.bmp"))
    open(P("rgb.ppm"), "wb").write(b"P6\n# comment\n37 29\n255\n" + rgb.tobytes())
    open(P("broken.png"), "wb").write(open(os.path.join(OUT, "rgb8.png"), "rb").read()[:120])       # truncated: must be refused
    open(P("not_an_image.tga"), "wb").write(b"hello, this is text and not a picture at all")

    # OBJ + MTL naming every file: one shape per texture (the reference keeps its texture map per shape)
    with open(os.path.join(OUT, "textures.mtl"), "w") as m, open(os.path.join(OUT, "textures.obj"), "w") as o:
        o.write("mtllib textures.mtl\n")
        for i, n in enumerate(names):
            m.write(f"newmtl m{i}\nKd 0.5 0.5 0.5\nmap_Kd {n}\n\n")
            o.write(f"o s{i}\nv {i} 0 0\nv {i + 1} 0 0\nv {i + 1} 1 0\nv {i} 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nusemtl m{i}\nf -4/-4 -3/-3 -2/-2 -1/-1\n")

    from oracle import orc
    R = orc.Oracle("reference"); R.load_obj(os.path.join(OUT, "textures.obj"))
    ms = R.meshes(); tx = R.textures()
    ids = [m["texture_id"] for m in ms]
    assert len(ms) == len(names)
    arrays = {"names": np.array(names), "texture_ids": np.array(ids)}
    for n, t in zip(names, ids):
        if t >= 0: arrays["px_" + n] = tx[t]
        print(f"{n:28s} ->", "refused" if t < 0 else tx[t].shape)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "textures.npz"), **arrays)
    R.close()
    print("total bytes of image fixtures:", sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)))


if __name__ == "__main__":
    main()
