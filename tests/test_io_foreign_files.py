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
