"""Residency planning: which tiles a device plan holds, how
they get there (device-side ingestion of tiles that are not on the host) and that the
answers do not depend on either."""
import math
import os
import shutil

import numpy as np
import pytest

import turtle_b200 as tb
from oracle import harness as H
from tests.common import Scene
from turtle_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
IO = os.path.join(HERE, "golden", "io")


def test_residency_from_rays_box():
    """Host side: the box holds every point of every ray up to where the rule stops it."""
    rng = np.random.default_rng(5)
    n = 200
    lat, lon = rng.uniform(44., 46., n), rng.uniform(2., 4., n)
    pos = np.array([tb.ecef_from_geodetic(la, lo, 500.) for la, lo in zip(lat, lon)])
    dirs = synth.random_unit(n, 7)
    rule = tb.trace_rule(9000., length_max=3e4)
    box = tb.residency_from_rays(pos, dirs, rule, step=500., margin=0.01)
    assert box.latitude_min <= lat.min() and box.latitude_max >= lat.max()
    assert box.longitude_min <= lon.min() and box.longitude_max >= lon.max()
    # 30 km of path is at most ~0.27 deg of latitude and ~0.39 deg of longitude at 46 N
    assert box.latitude_min > 44. - 0.3 and box.latitude_max < 46. + 0.3
    assert box.longitude_min > 2. - 0.45 and box.longitude_max < 4. + 0.45
    for t in (0., 1e4, 2e4, 3e4):
        for p, d in zip(pos[:50], dirs[:50]):
            la, lo, al = tb.ecef_to_geodetic(p + d * t)
            if al < 9000.:
                assert box.latitude_min <= la <= box.latitude_max
                assert box.longitude_min <= lo <= box.longitude_max
    # upward rays are cut by the altitude cap long before the length cap
    up = np.array([tb.ecef_from_horizontal(45., 3., 0., 80.)])
    p0 = np.array([tb.ecef_from_geodetic(45., 3., 0.)])
    box = tb.residency_from_rays(p0, up, tb.trace_rule(9000., length_max=1e7))
    assert 45. <= box.latitude_max < 45.05
    # a ray over the date line opens the longitude bounds
    p1 = np.array([tb.ecef_from_geodetic(10., 179.99, 100.)])
    east = np.array([tb.ecef_from_horizontal(10., 179.99, 90., 1.)])
    box = tb.residency_from_rays(p1, east, tb.trace_rule(9000., length_max=5e4))
    assert math.isnan(box.longitude_min) and math.isnan(box.longitude_max)
    assert not math.isnan(box.latitude_min)


def _fan(stepper, lat, lon, n_az, n_el):
    origin, _ = stepper.position(lat, lon, 1.0, 0)
    dirs = synth.fan_directions(lat, lon, n_az, n_el)
    return np.repeat(origin[None], len(dirs), 0), dirs


@pytest.mark.gpu
def test_device_ingestion_equals_host_load(small_stack):
    """A plan whose tiles are decoded ON THE DEVICE from the files (the stack was never
    loaded on the host) and a plan uploaded from host-resident tiles hold the same nodes:
    byte-identical traces."""
    sc = Scene(stacks=[small_stack], ops=[(H.ADD_STACK, 0, 0.)], range=0.)
    rule = tb.trace_rule(6000., max_steps=20000)
    a_stepper, _, _ = sc.product()
    plan_a = a_stepper.freeze(0)
    ra = plan_a.residency()
    assert ra["tiles_resident"] == 3 and ra["tiles_ingested"] == 3 and ra["tiles_skipped"] == 0
    b_stepper, _, b_stacks = sc.product()
    b_stacks[0].load()  # host-resident tiles: uploaded as they are
    plan_b = b_stepper.freeze(0)
    rb = plan_b.residency()
    assert rb["tiles_resident"] == 3 and rb["tiles_ingested"] == 0
    pos, dirs = _fan(a_stepper, 45.4, 2.6, 256, 128)
    assert plan_a.trace(pos, dirs, rule).tobytes() == plan_b.trace(pos, dirs, rule).tobytes()


@pytest.mark.gpu
def test_region_from_rays_same_answers(small_stack):
    """Rays that stay within one tile: the plan made for their box holds that tile only,
    and answers like the plan of the whole stack."""
    sc = Scene(stacks=[small_stack], ops=[(H.ADD_STACK, 0, 0.)], range=0.)
    stepper, _, _ = sc.product()
    pos, dirs = _fan(stepper, 45.5, 2.5, 128, 64)
    rule = tb.trace_rule(3000., length_max=2e4, max_steps=20000)
    full = stepper.freeze(0)
    box = tb.residency_from_rays(pos, dirs, rule, margin=0.001)
    assert box.latitude_min > 45.2 and box.latitude_max < 45.8
    part = stepper.freeze(0, region=box)
    r = part.residency()
    assert r["tiles_resident"] == 1 and r["tiles_skipped"] == 2
    assert part.bytes < full.bytes
    assert part.trace(pos, dirs, rule).tobytes() == full.trace(pos, dirs, rule).tobytes()
    # an explicit box that keeps nothing: every ray is outside at once
    none = stepper.freeze(0, region=(60., 61., None, None))
    assert none.residency()["tiles_resident"] == 0
    assert (none.trace(pos[:64], dirs[:64], rule)["status"] == tb.api.TRACE_DOMAIN).all()
    # a plan that does not fit its budget is refused, with the size it would need
    with pytest.raises(tb.TurtleError, match="residency plan needs"):
        stepper.freeze(0, memory_limit=1 << 20)


@pytest.mark.gpu
def test_mixed_format_stack_on_device(tmp_path):
    """GeoTIFF and PNG tiles side by side, ingested on the device (row order and byte
    order differ per format): the batched query answers like the scalar host path."""
    d = tmp_path / "tiles"
    d.mkdir()
    shutil.copy(os.path.join(IO, "n44e003.tif"), str(d / "n44e003.tif"))
    src = tb.Map(path=os.path.join(IO, "n44e003.tif"))
    info, _ = src.meta()
    z = np.array([[src.node(ix, iy)[2] for ix in range(info.nx)] for iy in range(info.ny)])
    tb.Map(info.nx, info.ny, (4., 5.), (44., 45.), (-32767., 32768.), None,
           z[::-1, ::-1].copy()).dump(str(d / "n44e004.png"))
    stack = tb.Stack(str(d))
    stepper = tb.Stepper(range=0.)
    stepper.add_stack(stack, 0.)
    plan = stepper.freeze(0)
    assert plan.residency()["tiles_ingested"] == 2
    rng = np.random.default_rng(9)
    n = 5000
    lat, lon = rng.uniform(43.9, 45.1, n), rng.uniform(2.9, 5.1, n)
    pos = tb.ecef_from_geodetic_batch(lat, lon, rng.uniform(-100., 9000., n)) \
        if tb.device_count() else None
    got = plan.step(pos.copy())
    for i in range(0, n, 7):
        want = stepper.step(pos[i].copy())
        assert got["index"][i, 0] == want["index"][0]
        if want["index"][0] >= 0:
            # latitude / longitude differ by a few ulp between the device and glibc
            # (DESIGN.md section 5): the interpolated ground moves by ~1e-9 m
            np.testing.assert_allclose(got["elevation"][i], want["elevation"], rtol=0,
                                       atol=1e-6)
            assert got["altitude"][i] == want["altitude"]
