"""GPU: batched frame transforms and elevation queries through the C ABI, against the
oracle on the same inputs. + - * / sqrt are IEEE on both sides (no FMA contraction), so
everything that does not go through a transcendental must be BIT-EXACT; the rest is
within a few ulp of glibc (tolerances written below)."""
import os

import numpy as np
import pytest

import turtle_b200 as tb
from oracle import harness as H
from tests.common import geoid_map, ulp_distance, utm_map

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


@pytest.fixture(scope="module")
def ora():
    return H.Driver(H.best_oracle())


def test_device_present():
    assert tb.device_count() >= 1
    assert tb.dfma_peak(1) > 1000.  # G FP64 FMA / s: a B200 does ~17000


def test_exact_division():
    """The kernels divide through a shared reciprocal (tb::divide): it must return the
    bits of the IEEE division for every operand pair (2^27 random pairs, 3 seeds)."""
    from turtle_b200._lib import lib
    for seed in (1, 0xC0FFEE, 2 ** 40 + 7):
        assert lib.turtle_b200_selftest_division(1 << 26, seed) == 0


def test_to_geodetic_vs_oracle(ora):
    rng = np.random.default_rng(11)
    n = 1 << 20
    lat, lon = rng.uniform(-90, 90, n), rng.uniform(-180, 180, n)
    alt = rng.uniform(-1000, 20000, n)
    ecef = ora.ecef_from_geodetic(lat, lon, alt)
    want = ora.ecef_to_geodetic(ecef)
    got = tb.ecef_to_geodetic_batch(ecef)
    assert np.array_equal(want[2], got[2])            # altitude: no transcendental, bit-exact
    assert ulp_distance(want[0], got[0]).max() <= 4   # latitude: asin / acos, <= 4 ulp
    assert ulp_distance(want[1], got[1]).max() <= 4   # longitude: atan2, <= 4 ulp
    assert np.abs(got[0] - lat).max() < 1e-8 and np.abs(got[2] - alt).max() < 1e-8


def test_to_geodetic_golden_and_poles():
    la, lo, al = tb.ecef_to_geodetic_batch(GOLD["geo_ecef_all"])
    assert np.array_equal(al, GOLD["geo_back_alt"])
    assert ulp_distance(la, GOLD["geo_back_lat"]).max() <= 4
    # longitude +-180 at the antimeridian may flip sign: compare modulo 360
    dlon = np.abs(((lo - GOLD["geo_back_lon"]) + 180.) % 360. - 180.)
    assert dlon.max() < 1e-13
    assert list(la[-5:-3]) == [90., -90.] and list(lo[-5:-2]) == [0., 0., 0.]  # ecef.c:77-84


def test_from_geodetic_and_horizontal_vs_oracle(ora):
    rng = np.random.default_rng(12)
    n = 1 << 18
    lat, lon = rng.uniform(-90, 90, n), rng.uniform(-180, 180, n)
    alt = rng.uniform(-1000, 20000, n)
    want, got = ora.ecef_from_geodetic(lat, lon, alt), tb.ecef_from_geodetic_batch(lat, lon, alt)
    # sin / cos differ by <= 2 ulp; products of three factors: 1e-15 relative of the radius
    assert np.abs(want - got).max() <= 1e-15 * 6.4e6 * 4
    az, el = rng.uniform(0, 360, n), rng.uniform(-90, 90, n)
    want = ora.ecef_from_horizontal(lat, lon, az, el)
    got = tb.ecef_from_horizontal_batch(lat, lon, az, el)
    assert np.abs(want - got).max() <= 4e-16 * 4


def test_map_elevation_bit_exact(ora):
    mp = utm_map(n=301)
    m = ora.map_create(mp["nx"], mp["ny"], mp["x"], mp["y"], mp["z"], mp["projection"], mp["values"])
    gm = tb.Map(mp["nx"], mp["ny"], mp["x"], mp["y"], mp["z"], mp["projection"], mp["values"])
    rng = np.random.default_rng(13)
    n = 1 << 19
    x = rng.uniform(mp["x"][0] - 100, mp["x"][1] + 100, n)
    y = rng.uniform(mp["y"][0] - 100, mp["y"][1] + 100, n)
    # closed domain: the exact corners and edges, NaN, just outside
    x[:6] = [mp["x"][0], mp["x"][1], mp["x"][1], np.nan, mp["x"][1] + 1e-7, mp["x"][0]]
    y[:6] = [mp["y"][0], mp["y"][1], mp["y"][0], mp["y"][0], mp["y"][0], np.nan]
    want_z, want_in = ora.map_elevation(m, x, y)
    got_z, got_in = gm.elevation_batch(x, y)
    assert np.array_equal(want_in, got_in) and list(got_in[:6]) == [1, 1, 1, 0, 0, 0]
    assert np.array_equal(want_z[want_in == 1], got_z[got_in == 1])  # bit-exact
    assert (got_z[got_in == 0] == 0.).all()  # untouched where outside
    # a re-fill must reach the device mirror
    v2 = np.asarray(mp["values"]) * 0.5
    gm.fill(v2)
    m2 = ora.map_create(mp["nx"], mp["ny"], mp["x"], mp["y"], mp["z"], mp["projection"], v2)
    w2, _ = ora.map_elevation(m2, x[6:1000], y[6:1000])
    g2, _ = gm.elevation_batch(x[6:1000], y[6:1000])
    assert np.array_equal(w2, g2)
    # the cell-packed copy of the grid (one 8-byte load per query): same bits, follows a
    # re-fill, and the gradient kernel keeps reading the row-major mirror
    from turtle_b200._lib import lib
    lib.turtle_map_gather_set(gm.handle, 1)
    w3, win3 = ora.map_elevation(m2, x, y)
    g3, gin3 = gm.elevation_batch(x, y)
    assert np.array_equal(win3, gin3) and np.array_equal(w3[win3 == 1], g3[gin3 == 1])
    gm.fill(np.asarray(mp["values"]))
    g4, gin4 = gm.elevation_batch(x, y)
    assert np.array_equal(want_in, gin4) and np.array_equal(want_z[want_in == 1], g4[gin4 == 1])
    wx, wy, wgin = ora.map_gradient(m, x[6:5000], y[6:5000])
    gx, gy, ggin = gm.gradient_batch(x[6:5000], y[6:5000])
    assert np.array_equal(wx, gx) and np.array_equal(wy, gy) and np.array_equal(wgin, ggin)


def test_fused_ecef_elevation(ora):
    """ECEF -> geodetic -> UTM -> bilinear in one kernel vs the same chain on the oracle."""
    mp = utm_map(n=301)
    m = ora.map_create(mp["nx"], mp["ny"], mp["x"], mp["y"], mp["z"], mp["projection"], mp["values"])
    gm = tb.Map(mp["nx"], mp["ny"], mp["x"], mp["y"], mp["z"], mp["projection"], mp["values"])
    rng = np.random.default_rng(14)
    n = 1 << 18
    x = rng.uniform(mp["x"][0] - 200, mp["x"][1] + 200, n)
    y = rng.uniform(mp["y"][0] - 200, mp["y"][1] + 200, n)
    la, lo = ora.project("UTM 31N", x, y, inverse=True)
    ecef = ora.ecef_from_geodetic(la, lo, rng.uniform(0, 5000, n))
    wla, wlo, wal = ora.ecef_to_geodetic(ecef)
    wx, wy = ora.project("UTM 31N", wla, wlo)
    wz, win = ora.map_elevation(m, wx, wy)
    gla, glo, gal, gz, gin = gm.elevation_ecef_batch(ecef)
    assert np.array_equal(gal, wal)
    assert ulp_distance(gla, wla).max() <= 4 and ulp_distance(glo, wlo).max() <= 4
    flips = int((gin != win).sum())  # points within 1e-8 m of the map border may flip
    assert flips <= 2
    both = (gin == 1) & (win == 1)
    assert np.abs(gz[both] - wz[both]).max() < 1e-6  # x, y move by ~1e-9 m; slope < 100


def test_map_gradient_bit_exact(ora):
    """turtle_map_gradient_batch vs the oracle: + - * / only, so bit-exact; gy is left
    untouched in the first row of cells as in the reference (map.c:353)."""
    mp = utm_map(n=301)
    m = ora.map_create(mp["nx"], mp["ny"], mp["x"], mp["y"], mp["z"], mp["projection"], mp["values"])
    gm = tb.Map(mp["nx"], mp["ny"], mp["x"], mp["y"], mp["z"], mp["projection"], mp["values"])
    rng = np.random.default_rng(16)
    n = 1 << 18
    x = rng.uniform(mp["x"][0] - 50, mp["x"][1] + 50, n)
    y = rng.uniform(mp["y"][0] - 50, mp["y"][1] + 50, n)
    y[:4000] = rng.uniform(mp["y"][0], mp["y"][0] + 6., 4000)
    wx, wy, win = ora.map_gradient(m, x, y)
    gx, gy, gin = gm.gradient_batch(x, y)
    assert np.array_equal(win, gin)
    assert np.array_equal(wx, gx) and np.array_equal(wy, gy)
    assert (np.abs(gx[gin == 1]) > 0).mean() > 0.9


def test_stack_elevation_and_gradient_bit_exact(small_stack):
    """turtle_stack_elevation_batch / turtle_stack_gradient_batch on the resident tiles of a
    plan vs the reference's scalar calls (stack.c:338-388): + - * / only, bit-exact, over
    tile interiors, shared tile edges, the missing tile and the outside of the stack."""
    ref = H.Driver(H.best_oracle())
    st = ref.stack_create(small_stack)
    stepper = tb.Stepper(range=0.)
    stepper.add_stack(tb.Stack(small_stack), 0.)
    plan = stepper.freeze(0)
    rng = np.random.default_rng(23)
    n = 1 << 17
    la = rng.uniform(44.9, 47.1, n)
    lo = rng.uniform(1.9, 4.1, n)
    k = 6000  # on and next to the tile edges (integer degrees) and the first rows of cells
    la[:k] = np.round(la[:k]) + rng.choice([0., 1e-13, -1e-13, 1. / 1200 * 0.3], k)
    lo[k:2 * k] = np.round(lo[k:2 * k]) + rng.choice([0., 1e-13, -1e-13, 1. / 1200 * 0.3], k)
    wz, win = ref.stack_elevation(st, la, lo)
    gz, gin = plan.stack_elevation(0, la, lo)
    assert np.array_equal(win, gin) and np.array_equal(wz, gz)
    wla, wlo, win = ref.stack_gradient(st, la, lo)
    gla, glo, gin = plan.stack_gradient(0, la, lo)
    assert np.array_equal(win, gin)
    assert np.array_equal(wla, gla) and np.array_equal(wlo, glo)
    assert 0.5 < gin.mean() < 0.9  # the missing tile and the margin are outside
    with pytest.raises(tb.TurtleError):
        plan.stack_gradient(1, la[:4], lo[:4])


def test_geoid_style_geodetic_map(ora):
    g = geoid_map()
    m = ora.map_create(g["nx"], g["ny"], g["x"], g["y"], g["z"], None, g["values"])
    gm = tb.Map(g["nx"], g["ny"], g["x"], g["y"], g["z"], None, g["values"])
    rng = np.random.default_rng(15)
    x, y = rng.uniform(-1, 361, 100000), rng.uniform(-91, 91, 100000)
    wz, win = ora.map_elevation(m, x, y)
    gz, gin = gm.elevation_batch(x, y)
    assert np.array_equal(win, gin) and np.array_equal(wz[win == 1], gz[gin == 1])


@pytest.mark.parametrize("tag", ["UTM 31N", "UTM 3.0S", "Lambert 93", "Lambert IIe"])
def test_projection_batches(ora, tag):
    """batched project / unproject vs the oracle. Everything here goes
    through transcendentals (CUDA libm vs glibc): x, y within 2e-8 m (a few ulp of 1e7),
    latitude / longitude within 1e-12 deg plus, for Lambert, the FLT_EPSILON stopping rule
    of the reference's inverse iteration (projection.c:265)."""
    rng = np.random.default_rng(17)
    n = 1 << 17
    la, lo = rng.uniform(41., 51., n), rng.uniform(-5., 9., n)
    p = tb.Projection(tag)
    wx, wy = ora.project(tag, la, lo)
    gx, gy = p.project_batch(la, lo)
    assert np.abs(wx - gx).max() < 2e-8 and np.abs(wy - gy).max() < 2e-8
    wla, wlo = ora.project(tag, wx, wy, inverse=True)
    gla, glo = p.unproject_batch(wx, wy)
    tol = 1e-12 if tag.startswith("UTM") else 2e-8
    assert np.abs(wla - gla).max() < tol and np.abs(wlo - glo).max() < 1e-12
    assert np.abs(gla - la).max() < 5e-8 and np.abs(glo - lo).max() < 1e-8  # round trip


def test_empty_batches():
    la, lo, al = tb.ecef_to_geodetic_batch(np.zeros((0, 3)))
    assert len(la) == 0
    assert tb.ecef_from_geodetic_batch([], [], []).shape == (0, 3)
