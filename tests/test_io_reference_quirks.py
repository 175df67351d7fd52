"""Where the product's WRITERS deliberately differ from the reference's.

turtle_map_dump(".tif"): the reference writes node row 0 -- the SOUTH row -- as scanline 0
(geotiff16.c:313-322) under a tiepoint that places scanline 0 at the north edge
(geotiff16.c:297-299), and its reader takes scanline 0 for the NORTH row (geotiff16.c:247-
254): a map dumped by the reference comes back upside down, in the reference and in any
GeoTIFF reader. (Its own round-trip test, tests/test-turtle.c:1094-1140, uses a checkerboard
that is symmetric under that flip.) The product writes the north row first, so that dump ->
load is the identity and the file agrees with its own tiepoint. The READERS agree on every
file (tests/test_io_formats.py); this test pins the reference's behaviour so that the claim
stays reproducible."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

import turtle_b200 as tb
from oracle import harness as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.exists(H.REF), reason="oracle/_ref not built")

WRITER = r"""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import make_io_golden as G
lib = G.reference()
vals = np.load(sys.argv[2])
ny, nx = vals.shape
m = G.make_map(lib, nx, ny, (3., 4.), (44., 45.), (-32767., 32768.), None, vals)
assert lib.turtle_map_dump(m, sys.argv[3].encode()) == 0
lib.turtle_map_destroy(C.byref(m))
back = G.describe(lib, sys.argv[3])[-1]
np.save(sys.argv[4], back)
"""


def test_reference_tiff_dump_comes_back_upside_down(tmp_path):
    PIL = pytest.importorskip("PIL")
    libs = os.path.join(os.path.dirname(os.path.dirname(PIL.__file__)), "pillow.libs")
    found = glob.glob(os.path.join(libs, "libtiff-*.so*"))
    if not found:
        pytest.skip("no libtiff for the reference")
    shim = tmp_path / "shim"
    shim.mkdir()
    os.symlink(found[0], str(shim / "libtiff.so"))
    env = dict(os.environ, TURTLE_IO_SHIM="1",
               LD_LIBRARY_PATH="%s:%s:%s" % (shim, libs, os.environ.get("LD_LIBRARY_PATH", "")))
    rng = np.random.default_rng(8)
    vals = np.rint(rng.uniform(-400., 8000., (31, 41)))
    np.save(str(tmp_path / "vals.npy"), vals)
    ref_file, back = str(tmp_path / "reference.tif"), str(tmp_path / "back.npy")
    subprocess.run([sys.executable, "-c", WRITER, os.path.join(ROOT, "tests", "golden"),
                    str(tmp_path / "vals.npy"), ref_file, back], env=env, check=True)
    seen_by_reference = np.load(back)
    assert np.array_equal(seen_by_reference, vals[::-1])  # the reference's own round trip
    assert not np.array_equal(seen_by_reference, vals)

    def nodes(path):
        m = tb.Map(path=path)
        info, _ = m.meta()
        return np.array([[m.node(ix, iy)[2] for ix in range(info.nx)] for iy in range(info.ny)])
    # the product reads the reference's file as the reference does ...
    assert np.array_equal(nodes(ref_file), seen_by_reference)
    # ... and writes a file that comes back as it went in
    ours = str(tmp_path / "product.tif")
    tb.Map(41, 31, (3., 4.), (44., 45.), (-32767., 32768.), None, vals).dump(ours)
    assert np.array_equal(nodes(ours), vals)
