"""File formats: the product's readers of .png / .tif / .grd /
.asc against what the UNMODIFIED reference makes of the same files -- fixtures written
and described by the reference in tests/golden/make_io_golden.py (meta data and every
node value, bit for bit) -- and the writers against the product's own readers. No GPU.

ref: src/turtle/io/png16.c, geotiff16.c, grd.c, asc.c, src/turtle/io.c, map.c:116-176."""
import os
import shutil

import numpy as np
import pytest

import turtle_b200 as tb

HERE = os.path.dirname(os.path.abspath(__file__))
IO = os.path.join(HERE, "golden", "io")
GOLD = np.load(os.path.join(HERE, "golden", "io_vectors.npz"))
FILES = [str(f) for f in GOLD["files"]]


def describe(m):
    info, tag = m.meta()
    z = np.array([[m.node(ix, iy)[2] for ix in range(info.nx)] for iy in range(info.ny)])
    x0, y0, _ = m.node(0, 0)
    x1, y1, _ = m.node(info.nx - 1, info.ny - 1)
    meta = np.array([info.nx, info.ny, info.x[0], info.x[1], info.y[0], info.y[1], info.z[0],
                     info.z[1], x0, y0, x1, y1])
    return meta, tag or "", info.encoding.decode() if info.encoding else "", z


@pytest.mark.parametrize("name", FILES)
def test_reader_matches_reference(name):
    key = name.replace(".", "_")
    meta, tag, enc, z = describe(tb.Map(path=os.path.join(IO, name)))
    assert meta.tobytes() == GOLD[key + "_meta"].tobytes(), (meta, GOLD[key + "_meta"])
    assert tag == str(GOLD[key + "_projection"])
    assert enc == str(GOLD[key + "_encoding"])
    assert z.tobytes() == GOLD[key + "_z"].tobytes()


@pytest.mark.parametrize("name,ext", [("utm31n.png", "png"), ("lambert93.png", "png"),
                                      ("n44e003.tif", "tif"), ("n44e003.tif", "png"),
                                      ("n44e003.png", "tif"), ("geoid.grd", "png"),
                                      ("esri.asc", "png")])
def test_dump_then_load_round_trip(tmp_path, name, ext):
    """turtle_map_dump writes what turtle_map_load reads back unchanged (png16.c:456-546,
    geotiff16.c:262-330), across formats too."""
    a = tb.Map(path=os.path.join(IO, name))
    out = str(tmp_path / ("copy." + ext))
    a.dump(out)
    ma, ta, _, za = describe(a)
    mb, tbg, eb, zb = describe(tb.Map(path=out))
    assert eb == ext and ta == tbg
    assert za.tobytes() == zb.tobytes()
    np.testing.assert_allclose(mb, ma, rtol=1e-15, atol=0)
    if name.endswith(ext):  # same format: the header survives bit for bit
        assert ma.tobytes() == mb.tobytes()


def test_created_map_dump_matches_reference_file(tmp_path):
    """A map created and filled through the C ABI, dumped to PNG, decodes to the same
    nodes and meta data as the reference's dump of the same map (fixture utm31n.png)."""
    ref = tb.Map(path=os.path.join(IO, "utm31n.png"))
    info, tag = ref.meta()
    z = np.array([[ref.node(ix, iy)[2] for ix in range(info.nx)] for iy in range(info.ny)])
    m = tb.Map(info.nx, info.ny, (info.x[0], info.x[1]), (info.y[0], info.y[1]),
               (info.z[0], info.z[1]), tag, z)
    out = str(tmp_path / "mine.png")
    m.dump(out)
    got = describe(tb.Map(path=out))
    want = describe(ref)
    assert got[0].tobytes() == want[0].tobytes() and got[1] == want[1]
    assert got[3].tobytes() == want[3].tobytes()


def test_errors(tmp_path):
    with pytest.raises(tb.TurtleError, match="no valid format") as e:
        tb.Map(path=str(tmp_path / "nothing.xyz"))
    assert e.value.code == 2  # TURTLE_RETURN_BAD_EXTENSION (turtle.h:39)
    with pytest.raises(tb.TurtleError):
        tb.Map(path=str(tmp_path / "missing.png"))
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"not a png at all")
    with pytest.raises(tb.TurtleError, match="invalid header for png"):
        tb.Map(path=str(bad))
    # a corrupted chunk is caught by its CRC
    data = bytearray(open(os.path.join(IO, "utm31n.png"), "rb").read())
    data[-30] ^= 0x55
    (tmp_path / "crc.png").write_bytes(bytes(data))
    with pytest.raises(tb.TurtleError):
        tb.Map(path=str(tmp_path / "crc.png"))
    # a header that asks for 2^31 x 2^31 nodes is refused, not a crash
    import struct
    tif = bytearray(open(os.path.join(IO, "n44e003.tif"), "rb").read())
    assert tif[:2] == b"MM"  # the reference writes big-endian TIFFs (libtiff mode "wb+")
    ifd = struct.unpack(">I", bytes(tif[4:8]))[0]
    for e in range(struct.unpack(">H", bytes(tif[ifd:ifd + 2]))[0]):
        at = ifd + 2 + 12 * e
        if struct.unpack(">H", bytes(tif[at:at + 2]))[0] in (256, 257):
            tif[at + 2:at + 4] = struct.pack(">H", 4)
            tif[at + 8:at + 12] = struct.pack(">I", 0x7fffffff)
    (tmp_path / "huge.tif").write_bytes(bytes(tif))
    with pytest.raises(tb.TurtleError, match="libtiff error|could not allocate memory"):
        tb.Map(path=str(tmp_path / "huge.tif"))
    # writers: GeoTIFF wants the integer metre scale and no projection; hgt/grd/asc have none
    m = tb.Map(path=os.path.join(IO, "utm31n.png"))
    with pytest.raises(tb.TurtleError, match="unsupported z scale"):
        m.dump(str(tmp_path / "x.tif"))
    with pytest.raises(tb.TurtleError, match="invalid write format"):
        m.dump(str(tmp_path / "x.grd"))
    with pytest.raises(tb.TurtleError, match="inconsistent data"):
        short = tmp_path / "short.asc"
        short.write_text("ncols 3\nnrows 2\nxllcorner 0\nyllcorner 0\ncellsize 1\nNODATA_value -1\n1 2 3\n")
        tb.Map(path=str(short))


def test_stack_of_png_and_tif_tiles(tmp_path):
    """turtle_stack_create takes tiles of any format (stack.c:73-91): two adjacent 1 x 1
    degree tiles, one GeoTIFF and one PNG, answer like the maps they hold."""
    d = tmp_path / "tiles"
    d.mkdir()
    shutil.copy(os.path.join(IO, "n44e003.tif"), str(d / "n44e003.tif"))
    src = tb.Map(path=os.path.join(IO, "n44e003.tif"))
    info, _ = src.meta()
    z = np.array([[src.node(ix, iy)[2] for ix in range(info.nx)] for iy in range(info.ny)])
    east = tb.Map(info.nx, info.ny, (4., 5.), (44., 45.), (-32767., 32768.), None, z[:, ::-1])
    east.dump(str(d / "n44e004.png"))
    stack = tb.Stack(str(d))
    for lat, lon in ((44.25, 3.5), (44.75, 3.01), (44.5, 4.5), (44.99, 4.98)):
        got, inside = stack.elevation(lat, lon)
        want, _ = (src if lon < 4. else east).elevation(lon, lat)
        assert inside == 1 and got == want
    assert stack.elevation(45.5, 3.5)[1] == 0
