"""The reference's own format tests, restated against this library:
tests/test-turtle.c:985-1046 (test_io_grd), :1049-1088 (test_io_hgt), :1094-1142
(test_io_tiff), :1148-1213 (test_io_asc) -- same synthetic files, same assertions and
tolerances, through the same calls (turtle_map_load / _node / _elevation / _fill / _dump)."""
import math
import struct

import numpy as np

import turtle_b200 as tb
from turtle_b200._lib import lib

DEG = math.pi / 180.


def fill(m, ix, iy, z):
    tb.api._check(lib.turtle_map_fill(m.handle, ix, iy, z))


def test_io_grd(tmp_path):
    path = tmp_path / "geoid.grd"
    with open(path, "w") as f:
        f.write("   -90.000000   90.000000     .000000  360.000000   15.000000   30.000000\n\n")
        k = 0
        for i in range(13):
            c = math.cos((i * 15 - 90) * DEG)
            for j in range(13):
                undulation = 100 * c * math.cos(j * 30 * DEG)
                if k % 8 == 0:
                    f.write(" ")
                f.write(" %8.3f" % undulation)
                if k % 8 == 7:
                    f.write("\n")
                if k % 170 == 169:
                    f.write("\n")
                k += 1
    geoid = tb.Map(path=str(path))
    for i in range(13):
        latitude = i * 15 - 90
        c = math.cos(latitude * DEG)
        for j in range(13):
            longitude = j * 30
            z, _ = geoid.elevation(longitude, latitude)
            assert abs(z - 100 * c * math.cos(longitude * DEG)) <= 1e-2
    fill(geoid, 0, 0, 1)  # writing to a GRD map
    assert abs(geoid.elevation(0, -90)[0] - 1) <= 1e-2


def test_io_hgt(tmp_path):
    path = tmp_path / "N45E003.hgt"
    k = np.arange(3601 * 3601, dtype=np.int64)
    z = np.where(k % 2 == 0, -1, 1).astype(">i2")
    path.write_bytes(z.tobytes())
    m = tb.Map(path=str(path))
    for kk in list(range(0, 3601 * 3601, 100 * 997)) + list(range(0, 3601 * 3601, 101 * 991)):
        i, j = divmod(kk, 3601)
        assert abs(m.node(j, i)[2] - (-1 if kk % 2 == 0 else 1)) <= 1e-2
    fill(m, 0, 0, 10)  # writing to an HGT map
    assert abs(m.elevation(3, 45)[0] - 10) <= 1e-2


def test_io_tiff(tmp_path):
    path = str(tmp_path / "map.tif")
    nx = ny = 101
    m = tb.Map(nx, ny, (3., 4.), (45., 46.), (-32767., 32768.), None)
    k = 0
    for i in range(ny):
        for j in range(nx):
            fill(m, j, i, -1. if k % 2 == 0 else 1.)
            k += 1
    m.dump(path)
    del m
    m = tb.Map(path=path)
    k = 0
    for i in range(ny):
        for j in range(nx):
            if k % 10 == 0 or k % 11 == 0:
                assert m.node(j, i)[2] == (-1. if k % 2 == 0 else 1.)
            k += 1
    fill(m, 0, 0, 10)  # writing to a GEOTIFF map
    assert abs(m.elevation(3, 45)[0] - 10.) <= 1e-2
    # what this writer emits is a baseline little-endian TIFF any reader takes
    head = open(path, "rb").read(8)
    assert head[:4] == b"II*\x00" and struct.unpack("<I", head[4:])[0] > 8


def test_io_asc(tmp_path):
    path = tmp_path / "bathymetry.asc"
    with open(path, "w") as f:
        f.write("ncols        10\nnrows        10\nxllcorner    142.000000000000\n"
                "yllcorner    35.000000000000\ncellsize     0.1\n"
                "NODATA_value  9.9692099683868690468e+36\n")
        k = 0
        for i in range(10):
            c = math.cos((35.05 + (9 - i) * 0.1) * DEG)
            for j in range(10):
                depth = -100 * abs(c * math.cos((142.05 + j * 0.1) * DEG))
                if k % 8 == 0:
                    f.write(" ")
                f.write(" %8.3f" % depth)
                if k % 8 == 7:
                    f.write("\n")
                k += 1
    bathymetry = tb.Map(path=str(path))
    checked = 0
    for i in range(10):
        latitude = 35.05 + i * 0.1
        c = math.cos(latitude * DEG)
        for j in range(10):
            longitude = 142.05 + j * 0.1
            z, inside = bathymetry.elevation(longitude, latitude)
            if inside:
                assert abs(z + 100 * abs(c * math.cos(longitude * DEG))) <= 1e-2
                checked += 1
    assert checked >= 80
    fill(bathymetry, 0, 0, -64)  # writing to an ASC map
    assert abs(bathymetry.elevation(142.05, 35.05)[0] + 64) <= 1e-2
