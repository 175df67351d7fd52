"""The PNG / TIFF readers on files written by an INDEPENDENT encoder (Pillow: libpng,
libtiff): compressed and international text chunks, the five scanline filters as libpng
picks them, deflate-compressed and multi-strip TIFFs. Skipped where Pillow is absent.

ref: src/turtle/io/png16.c:195-377, geotiff16.c:166-260 (what the reference reads through
libpng / libtiff; here tb_io.cpp parses the containers itself)."""
import os

import numpy as np
import pytest

import turtle_b200 as tb

PIL = pytest.importorskip("PIL")
from PIL import Image, PngImagePlugin, TiffImagePlugin  # noqa: E402

HEADER = ('{"topography" : {"x0" : %s, "y0" : %s, "z0" : %s, "x1" : %s, "y1" : %s, '
          '"z1" : %s, "projection" : "%s"}}')


def terrain(nx, ny, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:ny, 0:nx]
    v = 900. + 700. * np.sin(xx * 0.21) * np.cos(yy * 0.13) + 40. * rng.standard_normal((ny, nx))
    return np.clip(np.rint(v), 0, 65535).astype(np.uint16)


def nodes(m):
    info, tag = m.meta()
    return info, tag, np.array([[m.node(ix, iy)[2] for ix in range(info.nx)]
                                for iy in range(info.ny)])


@pytest.mark.parametrize("kind", ["tEXt", "zTXt", "iTXt"])
def test_png_written_by_libpng(tmp_path, kind):
    nx, ny = 53, 37
    raw = terrain(nx, ny, 3)  # raw[0] is the NORTH row in the file
    x0, x1, y0, y1, z0, z1 = 650000.25, 650520.25, 6860000.5, 6860360.5, -12.5, 3264.25
    text = HEADER % tuple([float(v).hex() for v in (x0, y0, z0, x1, y1, z1)] + ["Lambert 93"])
    meta = PngImagePlugin.PngInfo()
    if kind == "tEXt":
        meta.add_text("Comment", text)
    elif kind == "zTXt":
        meta.add_text("Comment", text, zip=True)
    else:
        meta.add_itxt("Comment", text, zip=True)
    path = str(tmp_path / ("libpng_%s.png" % kind))
    Image.fromarray(raw).save(path, pnginfo=meta)  # mode I;16, libpng chooses the filters
    info, tag, z = nodes(tb.Map(path=path))
    assert (info.nx, info.ny, tag) == (nx, ny, "Lambert 93")
    assert (info.x[0], info.x[1], info.y[0], info.y[1]) == (x0, x1, y0, y1)
    assert info.z[0] == z0 and abs(info.z[1] - z1) < 1e-9
    dz = (z1 - z0) / 65535
    want = z0 + raw[::-1].astype(np.float64) * dz  # node row 0 is the SOUTH row
    assert z.tobytes() == want.tobytes()


def test_png_without_topography_header(tmp_path):
    """No JSON header: the map loads with a null scale, as in the reference
    (png16.c:219-222: meta data initialised to zero, only nx / ny from the IHDR)."""
    path = str(tmp_path / "plain.png")
    Image.fromarray(terrain(9, 7, 1)).save(path)
    info, tag, z = nodes(tb.Map(path=path))
    assert (info.nx, info.ny, tag) == (9, 7, None) and (z == 0.).all()
    # 8-bit and RGB images are refused like the reference does
    Image.fromarray(np.zeros((4, 4), np.uint8)).save(str(tmp_path / "g8.png"))
    with pytest.raises(tb.TurtleError, match="invalid bit depth"):
        tb.Map(path=str(tmp_path / "g8.png"))
    Image.fromarray(np.zeros((4, 4, 3), np.uint8)).save(str(tmp_path / "rgb.png"))
    with pytest.raises(tb.TurtleError, match="invalid color scheme"):
        tb.Map(path=str(tmp_path / "rgb.png"))


@pytest.mark.parametrize("compression", [None, "tiff_adobe_deflate", "tiff_lzw", "tiff_lzw+predictor"])
def test_tiff_written_by_libtiff(tmp_path, compression):
    nx, ny = 47, 29
    raw = (terrain(nx, ny, 5).astype(np.int32) - 500).astype(np.int16)
    raw[3, 4] = -417  # a negative elevation survives as int16
    ifd = TiffImagePlugin.ImageFileDirectory_v2()
    ifd[33550] = (1. / 1200, 1. / 1200, 0.)          # ModelPixelScale
    ifd.tagtype[33550] = 12
    ifd[33922] = (0., 0., 0., 3.25, 44.75, 0.)       # ModelTiepoint: north-west corner
    ifd.tagtype[33922] = 12
    path = str(tmp_path / "libtiff.tif")
    kw = dict(tiffinfo=ifd)
    if compression and compression.endswith("+predictor"):
        ifd[317] = 2  # horizontal differencing
        ifd.tagtype[317] = 3
        compression = compression.split("+")[0]
    if compression:
        kw["compression"] = compression
    Image.fromarray(raw.view(np.uint16)).save(path, **kw)
    info, tag, z = nodes(tb.Map(path=path))
    assert (info.nx, info.ny, tag) == (nx, ny, None)
    assert info.x[0] == 3.25 and info.y[0] == 44.75 + (1 - ny) * (1. / 1200)
    assert (info.z[0], info.z[1]) == (-32767., 32768.)
    assert z.tobytes() == raw[::-1].astype(np.float64).tobytes()
    assert z[ny - 1 - 3, 4] == -417.


ADAM7 = ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2),
         (0, 1, 1, 2))


def _png_filter(kind, cur, up):
    """One scanline through PNG filter `kind` (2 bytes per pixel), as an encoder does."""
    cur, up = cur.astype(np.int32), up.astype(np.int32)
    a = np.concatenate([np.zeros(2, np.int32), cur[:-2]])
    c = np.concatenate([np.zeros(2, np.int32), up[:-2]])
    if kind == 0:
        pred = 0
    elif kind == 1:
        pred = a
    elif kind == 2:
        pred = up
    elif kind == 3:
        pred = (a + up) >> 1
    else:
        p = a + up - c
        pa, pb, pc = abs(p - a), abs(p - up), abs(p - c)
        pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, up, c))
    return ((cur - pred) & 0xff).astype(np.uint8)


def write_interlaced_png(path, raw, text):
    """16-bit greyscale, Adam7, every filter type in turn (Pillow cannot write interlaced
    files; it reads them, which is how this encoder is checked below)."""
    import struct
    import zlib

    def chunk(kind, data):
        return (struct.pack(">I", len(data)) + kind + data +
                struct.pack(">I", zlib.crc32(kind + data) & 0xffffffff))
    ny, nx = raw.shape
    be = raw.astype(">u2")
    stream, k = bytearray(), 0
    for x0, y0, dx, dy in ADAM7:
        sub = be[y0::dy, x0::dx]
        if sub.size == 0:
            continue
        up = np.zeros(2 * sub.shape[1], np.uint8)
        for row in sub:
            cur = np.frombuffer(row.tobytes(), np.uint8)
            stream.append(k % 5)
            stream += _png_filter(k % 5, cur, up).tobytes()
            up, k = cur, k + 1
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", nx, ny, 16, 0, 0, 0, 1))
    if text is not None:
        data += chunk(b"tEXt", b"Comment\x00" + text.encode())
    half = len(stream) // 2  # two IDAT chunks, split anywhere in the deflate stream
    z = zlib.compress(bytes(stream), 6)
    data += chunk(b"IDAT", z[:len(z) // 3]) + chunk(b"IDAT", z[len(z) // 3:]) + chunk(b"IEND", b"")
    assert half >= 0
    with open(path, "wb") as f:
        f.write(data)


@pytest.mark.parametrize("nx,ny", [(53, 37), (8, 8), (1, 1), (3, 2), (2, 9), (5, 1), (1, 6),
                                   (16, 17)])
def test_interlaced_png(tmp_path, nx, ny):
    """Adam7 files load like progressive ones: the reference reads them through
    png_read_image, which de-interlaces (png16.c:440). Narrow images have empty passes."""
    raw = terrain(nx, ny, nx * 100 + ny)
    x0, x1, y0, y1, z0, z1 = 0., float(max(nx - 1, 1)), 0., float(max(ny - 1, 1)), -100., 5000.
    text = HEADER % tuple([float(v).hex() for v in (x0, y0, z0, x1, y1, z1)] + ["UTM 31N"])
    path = str(tmp_path / "adam7.png")
    write_interlaced_png(path, raw, text)
    seen = np.array(Image.open(path))  # libpng agrees that this is the image
    assert seen.shape == raw.shape and (seen == raw).all()
    info, tag, z = nodes(tb.Map(path=path))
    assert (info.nx, info.ny, tag) == (nx, ny, "UTM 31N")
    dz = (z1 - z0) / 65535
    want = z0 + raw[::-1].astype(np.float64) * dz
    assert z.tobytes() == want.tobytes()
    # the same nodes as the progressive file of the same image
    flat = str(tmp_path / "flat.png")
    meta = PngImagePlugin.PngInfo()
    meta.add_text("Comment", text)
    Image.fromarray(raw).save(flat, pnginfo=meta)
    assert nodes(tb.Map(path=flat))[2].tobytes() == z.tobytes()


def test_interlaced_png_truncated(tmp_path):
    raw = terrain(20, 20, 5)
    path = str(tmp_path / "adam7.png")
    write_interlaced_png(path, raw, None)
    data = open(path, "rb").read()
    # a stream that holds fewer scanlines than the seven passes need
    import struct
    import zlib
    short = zlib.compress(b"\x00" * 300)
    i = data.index(b"IDAT") - 4
    body = data[:i] + struct.pack(">I", len(short)) + b"IDAT" + short + \
        struct.pack(">I", zlib.crc32(b"IDAT" + short) & 0xffffffff) + \
        struct.pack(">I", 0) + b"IEND" + struct.pack(">I", zlib.crc32(b"IEND") & 0xffffffff)
    (tmp_path / "short.png").write_bytes(body)
    with pytest.raises(tb.TurtleError, match="libpng error"):
        tb.Map(path=str(tmp_path / "short.png"))
