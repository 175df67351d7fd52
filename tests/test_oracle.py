"""Pin the oracle: the C restatement (oracle/turtle_oracle.c) against

  1. the golden vectors generated from the unmodified reference
     (tests/golden/reference_vectors.npz, made by tests/golden/make_golden.py);
  2. the compiled reference itself when oracle/_ref exists (this container);
  3. the closed-form assertions of the reference's own test-suite
     (tests/test-turtle.c, listed in SURVEY.md section 8c).

The product's scalar (host) calls are held to the same three checks: they are the
host instantiation of the expressions the CUDA kernels run.
"""
import os

import numpy as np
import pytest

from oracle import harness as H
from tests.common import Scene, geoid_map, lambert_map, utm_map
from tests.golden.make_golden import TAGS
from turtle_b200 import synth

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))
LIBS = [("port", H.PORT), ("product", H.PRODUCT)]
ids = [n for n, _ in LIBS]


@pytest.fixture(scope="module", params=LIBS, ids=ids)
def drv(request):
    return H.Driver(request.param[1])


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


# ---- 1. golden vectors ------------------------------------------------------------

def test_golden_geodesy(drv):
    assert same(drv.ecef_from_geodetic(GOLD["geo_lat"], GOLD["geo_lon"], GOLD["geo_alt"]),
                GOLD["geo_ecef"])
    la, lo, al = drv.ecef_to_geodetic(GOLD["geo_ecef_all"])
    assert same(la, GOLD["geo_back_lat"]) and same(lo, GOLD["geo_back_lon"])
    assert same(al, GOLD["geo_back_alt"])
    d = drv.ecef_from_horizontal(GOLD["geo_lat"], GOLD["geo_lon"], GOLD["hor_az"], GOLD["hor_el"])
    assert same(d, GOLD["hor_dir"])
    az, el = drv.ecef_to_horizontal(GOLD["geo_lat"], GOLD["geo_lon"], GOLD["hor_dir"])
    assert same(az, GOLD["hor_back_az"]) and same(el, GOLD["hor_back_el"])


def test_survey_appendix_d_vectors(drv):
    """Absolute known answers printed from the reference by the survey."""
    e = drv.ecef_from_geodetic([45.5], [3.5], [1000.])[0]
    assert [float.hex(v) for v in e] == ["0x1.10db28b573548p+22", "0x1.0b047a7eda615p+18",
                                         "0x1.145139d3398a4p+22"]
    la, lo, al = drv.ecef_to_geodetic(e[None])
    assert (float.hex(la[0]), float.hex(lo[0]), float.hex(al[0])) == (
        "0x1.6c00000000001p+5", "0x1.c000000000000p+1", "0x1.f400000000a3ep+9")
    x, y = drv.project("UTM 31N", [45.76415653], [2.95536402])
    assert (float.hex(x[0]), float.hex(y[0])) == ("0x1.e4e444fa02c40p+18", "0x1.355114b52a684p+22")
    x, y = drv.project("Lambert 93", [45.76415653], [2.95536402])
    assert (float.hex(x[0]), float.hex(y[0])) == ("0x1.541a56683803cp+19", "0x1.8dd831fef6c18p+22")


@pytest.mark.parametrize("tag", TAGS)
def test_golden_projections(drv, tag):
    key = tag.replace(" ", "_").replace(".", "p")
    x, y = drv.project(tag, GOLD["proj_lat"], GOLD["proj_lon"])
    assert same(x, GOLD["proj_%s_x" % key]) and same(y, GOLD["proj_%s_y" % key])
    la, lo = drv.project(tag, x, y, inverse=True)
    assert same(la, GOLD["proj_%s_lat" % key]) and same(lo, GOLD["proj_%s_lon" % key])
    # reference test_projection (tests/test-turtle.c:538-556): round trip to 1e-8 deg at
    # its test point; the Lambert inverse stops at FLT_EPSILON rad (projection.c:265),
    # which leaves up to ~2e-8 deg over this wider latitude range.
    assert np.abs(la - GOLD["proj_lat"]).max() < 5e-8
    assert np.abs(lo - GOLD["proj_lon"]).max() < 1e-8


def test_golden_map(drv):
    m = drv.map_create(5, 4, (10., 14.), (20., 23.), (0., 1000.), None, GOLD["map_vals"])
    z, inside = drv.map_elevation(m, GOLD["map_qx"], GOLD["map_qy"])
    assert same(inside, GOLD["map_inside"])
    assert same(z[inside == 1], GOLD["map_z"][GOLD["map_inside"] == 1])
    ix, iy = np.meshgrid(np.arange(5), np.arange(4))
    x, y, zz = drv.map_node(m, ix.ravel(), iy.ravel())
    assert same(x, GOLD["map_node_x"]) and same(y, GOLD["map_node_y"])
    assert same(zz, GOLD["map_node_z"])
    # digitisation: (zmax - zmin) / 65535 (ref: include/turtle.h:437-439)
    assert np.abs(zz - GOLD["map_vals"].ravel()).max() <= 0.5 * 1000. / 65535 + 1e-12


def _c1_scene(rg):
    return Scene(maps=[utm_map(n=201)], ops=[(H.ADD_FLAT, 0, -100.), (H.ADD_LAYER, 0, 0.),
                                             (H.ADD_MAP, 0, 0.)], range=rg)


def _c3_scene(library, stack_dir, rg, geoid):
    lm = lambert_map(H.Driver(library), n=201)
    return Scene(maps=[lm, geoid_map()], stacks=[stack_dir],
                 ops=[(H.ADD_FLAT, 0, 0.), (H.ADD_STACK, 0, 0.), (H.ADD_MAP, 0, 0.),
                      (H.ADD_LAYER, 0, 0.), (H.ADD_STACK, 0, 500.), (H.ADD_MAP, 0, 600.)],
                 geoid=geoid, range=rg)


@pytest.mark.parametrize("rg", [0., 10.])
def test_golden_trace_utm_map(drv, rg):
    d = _c1_scene(rg).oracle(drv.library)
    key = "c1_r%d" % int(rg)
    res, steps, _ = d.trace(GOLD[key + "_pos"], GOLD[key + "_dir"], H.rule(3100.))
    assert res.tobytes() == GOLD[key + "_res"].tobytes()
    assert steps == GOLD[key + "_res"]["n_steps"].sum()


@pytest.mark.parametrize("rg,geoid", [(0., -1), (10., 1)])
def test_golden_trace_layered(drv, small_stack, rg, geoid):
    d = _c3_scene(drv.library, small_stack, rg, geoid).oracle(drv.library)
    key = "c3_r%d" % int(rg)
    res, _, _ = d.trace(GOLD[key + "_pos"], GOLD[key + "_dir"], H.rule(9000., length_max=5e4))
    assert res.tobytes() == GOLD[key + "_res"].tobytes()
    one = d.step(GOLD[key + "_pos"], GOLD[key + "_dir"])
    for f in ("position", "latitude", "longitude", "altitude", "elevation", "step", "index"):
        assert same(one[f], GOLD[key + "_step_" + f]), f
    # a good mix of outcomes is covered
    st = GOLD[key + "_res"]["status"]
    assert (st == 0).any() and (st == 2).any() and GOLD[key + "_res"]["n_changes"].sum() > 100


def test_golden_flat_steps(drv):
    d = Scene(ops=[(H.ADD_FLAT, 0, 0.)], range=0.).oracle(drv.library)
    pos, idx = d.position([45.], [3.], [100.], 0)
    assert same(pos, GOLD["flat_pos"])
    w = d.walk(pos, np.repeat(GOLD["flat_dir"][None], 3, 0))
    assert same(w["step"], GOLD["flat_step"]) and same(w["altitude"], GOLD["flat_alt"])
    assert same(w["index"], GOLD["flat_index"]) and same(w["position"], GOLD["flat_final"])
    assert float.hex(float(w["step"][0, 0])) == "0x1.3ffffffff1c5cp+5"


# ---- 2. against the compiled reference ------------------------------------------------

needs_ref = pytest.mark.skipif(not os.path.exists(H.REF), reason="oracle/_ref not built")


@needs_ref
def test_reference_random_geodesy(drv):
    ref = H.Driver(H.REF)
    rng = np.random.default_rng(99)
    n = 100000
    lat, lon = rng.uniform(-90, 90, n), rng.uniform(-180, 180, n)
    alt = rng.uniform(-1000, 20000, n)
    e = ref.ecef_from_geodetic(lat, lon, alt)
    assert same(e, drv.ecef_from_geodetic(lat, lon, alt))
    for a, b in zip(ref.ecef_to_geodetic(e), drv.ecef_to_geodetic(e)):
        assert same(a, b)
    la, lo = rng.uniform(41, 51, n), rng.uniform(-5, 9, n)
    for tag in TAGS:
        a, b = ref.project(tag, la, lo), drv.project(tag, la, lo)
        assert same(a[0], b[0]) and same(a[1], b[1])


@needs_ref
@pytest.mark.parametrize("rg,geoid", [(0., -1), (1., 1), (100., 1)])
def test_reference_random_traces(drv, small_stack, rg, geoid):
    rng = np.random.default_rng(5)
    sc = _c3_scene(H.REF, small_stack, rg, geoid)
    ref, d = sc.oracle(H.REF, locked=True), sc.oracle(drv.library, locked=True)
    n = 1500
    pos = ref.ecef_from_geodetic(rng.uniform(44.9, 47.1, n), rng.uniform(1.9, 4.1, n),
                                 rng.uniform(-500, 5000, n))
    dirs = synth.random_unit(n, 3)
    r0, s0, _ = ref.trace(pos, dirs, H.rule(9000., length_max=1e5), threads=4)
    r1, s1, _ = d.trace(pos, dirs, H.rule(9000., length_max=1e5), threads=3)
    assert s0 == s1 and r0.tobytes() == r1.tobytes()
    la, lo = rng.uniform(44.9, 47.1, 20000), rng.uniform(1.9, 4.1, 20000)
    la[:6], lo[:6] = [45., 46., 47., 45.5, 46., 46.], [2., 3., 4., 3., 3., 2.5]  # tile edges
    z0, i0 = ref.stack_elevation(0, la, lo)
    z1, i1 = d.stack_elevation(0, la, lo)
    assert same(i0, i1) and same(z0, z1)
    p0, k0 = ref.position(la[:3000], lo[:3000], np.full(3000, 1.5), 1)
    p1, k1 = d.position(la[:3000], lo[:3000], np.full(3000, 1.5), 1)
    assert same(k0, k1) and same(p0, p1) and (k0 == -1).any()


@needs_ref
def test_reference_gradients(drv, small_stack):
    """turtle_map_gradient / turtle_stack_gradient, bit-exact with the
    reference, including the first-row behaviour of map.c:353 (y slope stored in *gx,
    *gy left untouched)."""
    ref = H.Driver(H.REF)
    mp = utm_map(n=101)
    rng = np.random.default_rng(8)
    n = 50000
    x = rng.uniform(mp["x"][0] - 30, mp["x"][1] + 30, n)
    y = rng.uniform(mp["y"][0] - 30, mp["y"][1] + 30, n)
    y[:2000] = rng.uniform(mp["y"][0], mp["y"][0] + 6., 2000)   # first row of cells
    x[2000:2010] = [mp["x"][0], mp["x"][1]] * 5                    # closed borders
    out = []
    for d in (ref, drv):
        m = d.map_create(mp["nx"], mp["ny"], mp["x"], mp["y"], mp["z"], mp["projection"],
                         mp["values"])
        out.append(d.map_gradient(m, x, y, fill=-7.5))
    for a, b in zip(*out):
        assert same(a, b)
    gx, gy, inside = out[0]
    first_row = (inside == 1) & (y < mp["y"][0] + 5.)
    assert first_row.sum() > 500 and (gy[first_row] == -7.5).all()  # untouched, map.c:353
    sr, sd = ref.stack_create(small_stack), drv.stack_create(small_stack)
    la, lo = rng.uniform(44.9, 47.1, 20000), rng.uniform(1.9, 4.1, 20000)
    la[:4], lo[:4] = [45., 46., 45.5, 46.], [2., 3., 3., 2.5]
    for a, b in zip(ref.stack_gradient(sr, la, lo), drv.stack_gradient(sd, la, lo)):
        assert same(a, b)


# ---- 3. the reference's own closed-form assertions (tests/test-turtle.c) ------------------

FLT_EPSILON = 1.1920929e-07
DBL_MAX = 1.7976931348623157e308


def test_ecef_round_trips_and_poles(drv):
    """ref: test_ecef, tests/test-turtle.c:582-623."""
    lat, lon, alt = np.array([45.5, -60., 10.]), np.array([3.5, 120., -70.]), np.array([1000., 0., 50.])
    la, lo, al = drv.ecef_to_geodetic(drv.ecef_from_geodetic(lat, lon, alt))
    assert np.abs(la - lat).max() < 1e-8 and np.abs(lo - lon).max() < 1e-8
    assert np.abs(al - alt).max() < 1e-8
    la, lo, al = drv.ecef_to_geodetic(np.array([[0., 0., 6356752.3142 + 10.],
                                                [0., 0., -6356752.3142 - 10.]]))
    assert list(la) == [90., -90.] and list(lo) == [0., 0.]
    assert np.allclose(al, 10., atol=1e-8)
    az, el = drv.ecef_to_horizontal([45.], [3.], drv.ecef_from_horizontal([45.], [3.], [60.], [30.]))
    assert abs(az[0] - 60.) < 1e-8 and abs(el[0] - 30.) < 1e-8


def _layered(drv, small_stack):
    """The geometry of test_stepper_layer (tests/test-turtle.c:255-330): two layers of
    flat / stack / map with offsets -0.5 and 0."""
    mp = utm_map(n=201, x0=486000., y0=5057000.)
    ops = []
    for off in (-0.5, 0.):
        ops += [(H.ADD_LAYER, 0, 0.), (H.ADD_FLAT, 0, off), (H.ADD_STACK, 0, off),
                (H.ADD_MAP, 0, off)]
    return Scene(maps=[mp], stacks=[small_stack], ops=ops, range=1.).oracle(drv.library), mp


def test_layer_index_semantics_and_step_rule(drv, small_stack):
    d, mp = _layered(drv, small_stack)
    cx, cy = 0.5 * (mp["x"][0] + mp["x"][1]), 0.5 * (mp["y"][0] + mp["y"][1])
    lat0, lon0 = [v[0] for v in d.project("UTM 31N", [cx], [cy], inverse=True)]
    values = np.zeros((3, 2))
    sites = [(lat0, lon0, 0), (45.5, 2.5, 1), (40., 10., 2)]  # on the map / stack only / flat only
    for i in (0, 1):
        for j, (la, lo, want) in enumerate(sites):
            pos, idx = d.position([la], [lo], [-0.25], i)
            assert idx[0] == want  # last added data has priority, tests/test-turtle.c:286-324
            q = d.step(pos)
            if i or j:
                assert tuple(q["index"][0]) == (i, want)
            else:
                assert q["index"][0][0] == 0
            values[j, i] = q["altitude"][0]
    for j in range(3):  # equal offsets => equal altitudes, :326-330
        assert abs((values[j, 0] + 0.5) - values[j, 1]) < FLT_EPSILON
    slope = 0.4
    # top medium: step = slope * distance to the layer below; upper bound is the sentinel
    pos, _ = d.position([lat0], [lon0], [0.5], 1)
    q = d.step(pos)
    assert q["index"][0][0] == 2 and q["elevation"][0][1] == DBL_MAX
    assert abs(q["elevation"][0][0] - (values[0, 1] + 0.25)) < FLT_EPSILON
    assert abs(q["step"][0] - 0.5 * slope) < FLT_EPSILON
    # middle medium: both bounds, the closest one rules, :356-367
    pos, _ = d.position([lat0], [lon0], [-0.1], 1)
    q = d.step(pos)
    assert tuple(q["index"][0]) == (1, 0)
    assert abs(q["elevation"][0][0] - (values[0, 0] + 0.25)) < FLT_EPSILON
    assert abs(q["elevation"][0][1] - (values[0, 1] + 0.25)) < FLT_EPSILON
    assert abs(q["step"][0] - 0.1 * slope) < FLT_EPSILON
    # bottom medium: lower bound is the sentinel, :369-378
    pos, _ = d.position([lat0], [lon0], [-0.5], 0)
    q = d.step(pos)
    assert tuple(q["index"][0]) == (0, 0) and q["elevation"][0][0] == -DBL_MAX
    assert abs(q["step"][0] - 0.5 * slope) < FLT_EPSILON


def test_boundary_bisection_and_resolution_step(small_stack):
    """ref: tests/test-turtle.c:381-401, on both scalar implementations."""
    for _, lib in LIBS:
        mp = utm_map(n=201)
        ops = []
        for off in (-0.5, 0.):
            ops += [(H.ADD_LAYER, 0, 0.), (H.ADD_FLAT, 0, off), (H.ADD_STACK, 0, off),
                    (H.ADD_MAP, 0, off)]
        d = Scene(maps=[mp], stacks=[small_stack], ops=ops, range=1., slope=2.).oracle(lib)
        cx, cy = 0.5 * (mp["x"][0] + mp["x"][1]), 0.5 * (mp["y"][0] + mp["y"][1])
        lat0, lon0 = [v[0] for v in d.project("UTM 31N", [cx], [cy], inverse=True)]
        top, _ = d.position([lat0], [lon0], [0.], 1)
        surface = d.step(top)["altitude"][0]
        pos, _ = d.position([lat0], [lon0], [-0.1], 1)
        up = d.ecef_from_horizontal([lat0], [lon0], [0.], [90.])
        w = d.walk(pos, np.repeat(up[None], 2, 0))
        # 1st step: slope 2 overshoots the surface, bisection lands on it (1e-5 m)
        assert w["index"][0, 0, 0] == 2 and abs(w["altitude"][0, 0] - surface) < 1e-5
        assert abs(w["step"][0, 0] - 0.1) < 1e-5
        # 2nd step, from the boundary: exactly one resolution long
        assert w["index"][1, 0, 0] == 2 and abs(w["step"][1, 0] - 1e-2) < 1e-5
        assert abs(w["altitude"][1, 0] - 1e-2 - surface) < 1e-5


def test_outside_domain_and_idempotence(drv, small_stack):
    """ref: tests/test-turtle.c:823-858, 923-931."""
    d = Scene(stacks=[small_stack], ops=[(H.ADD_STACK, 0, 0.)], range=1.).oracle(drv.library)
    pos0 = np.array([[1., 2., 3.]])
    pos, idx = d.position([10.], [10.], [1.], 0)  # no tile there
    assert idx[0] == -1 and same(pos, np.zeros((1, 3)))
    far = d.ecef_from_geodetic([10.], [10.], [100.])
    q = d.step(far, synth.random_unit(1, 1))
    assert tuple(q["index"][0]) == (-1, -1) and q["step"][0] == 0.
    assert same(q["elevation"], np.zeros((1, 2))) and same(q["position"], far)
    inside = d.ecef_from_geodetic([45.5], [2.5], [3000.])
    a, b = d.step(inside), d.step(inside)  # re-query: bit-for-bit the same
    for f in a:
        assert same(a[f], b[f])
    del pos0


def test_geoid_consistency(drv, small_stack):
    """ref: tests/test-turtle.c:906-921: elevation[1] + height == altitude with a geoid."""
    sc = Scene(maps=[geoid_map()], stacks=[small_stack], ops=[(H.ADD_STACK, 0, 0.)], geoid=0,
               range=1.)
    d = sc.oracle(drv.library)
    pos, idx = d.position([45.5], [2.5], [-10.], 0)
    q = d.step(pos)
    assert idx[0] == 0 and q["index"][0][0] == 0
    assert abs(q["elevation"][0][1] - 10. - q["altitude"][0]) < 1e-8


def test_marching_terminates(drv, small_stack):
    """ref: tests/test-turtle.c:873-883."""
    d = Scene(stacks=[small_stack], ops=[(H.ADD_STACK, 0, 0.)], range=100.).oracle(drv.library)
    pos, _ = d.position([45.5], [2.5], [0.5], 0)
    dirs = d.ecef_from_horizontal([45.5], [2.5], [30.], [20.])
    res, steps, _ = d.trace(pos, dirs, H.rule(5000.))
    assert res["status"][0] == 0 and 10 < steps < 100000 and res["altitude"][0] >= 5000.
