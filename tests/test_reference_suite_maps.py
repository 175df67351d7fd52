"""The reference's own map / stack / client tests, restated against this library with the
same fixtures (PNG files written by turtle_map_dump) and the same assertions:
tests/test-turtle.c:66-143 (fixtures), :412-512 (test_map), :628-692 (test_stack),
:697-775 (test_client). `stack->tiles.size` of the reference is turtle_stack_tiles_loaded
here (the type is opaque)."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

import turtle_b200 as tb
from turtle_b200._lib import lib

X0, Y0, Z0, Z1, NX, NY = 496000., 5067000., 0., 1000., 201, 201
LOCKER = C.CFUNCTYPE(C.c_int)
NOTHING = LOCKER(lambda: 0)


def check(rc):
    tb.api._check(rc)


@pytest.fixture(scope="module")
def data(tmp_path_factory):
    root = tmp_path_factory.mktemp("reference_suite")
    # setup_map_data, test-turtle.c:66-96
    vals = np.where((np.arange(NX * NY) % 2) == 0, Z0, Z1).reshape(NX, NY).T  # k runs over (i, j)
    m = tb.Map(NX, NY, (X0 - 1000, X0 + 1000), (Y0 - 1000, Y0 + 1000), (Z0, Z1), "UTM 31N", vals)
    m.dump(str(root / "map.png"))
    # setup_stack_data, :99-143: four flat 1201 x 1201 tiles as PNG
    topo = root / "topography"
    topo.mkdir()
    for name, lat, lon in (("45N_002E", 45, 2), ("46N_002E", 46, 2), ("45N_003E", 45, 3),
                           ("46N_003E", 46, 3)):
        t = tb.Map(1201, 1201, (lon, lon + 1), (lat, lat + 1), (0., 1.), None,
                   np.zeros((1201, 1201)))
        t.dump(str(topo / (name + ".png")))
    return root


def test_map(data):
    m = tb.Map(path=str(data / "map.png"))
    info, projection = m.meta()
    assert (info.nx, info.ny) == (NX, NY)
    assert (info.x[0], info.x[1]) == (X0 - 1000, X0 + 1000)
    assert (info.y[0], info.y[1]) == (Y0 - 1000, Y0 + 1000)
    assert (info.z[0], info.z[1]) == (Z0, Z1)
    assert projection == "UTM 31N"
    k = 0
    for i in range(NX):
        for j in range(NY):
            assert m.node(i, j)[2] == (Z0 if k % 2 == 0 else Z1)
            k += 1
    m.elevation(X0, Y0)
    m.elevation(X0 + 0.5, Y0 + 0.5)
    assert m.elevation(X0 - 1000.5, Y0)[1] == 0
    for ix in range(0, NX, 7):  # re-fill with some dummy data (:455-470)
        for iy in range(0, NY, 7):
            x, y, _ = m.node(ix, iy)
            r = math.hypot(x - X0, y - Y0)
            check(lib.turtle_map_fill(m.handle, ix, iy, 0. if r >= 1e3 else 1e3 - r))
    # wrong maps (:473-507): codes and the message format with the reference's file names
    for path, code, pattern in (
            ("nothing", 2, r"\{ turtle_map_load \[#[0-9]*\], src/turtle/io.c:[0-9]* \} "
                           r"no valid format for file `nothing'"),
            ("nothing.png", 10, r"\{ turtle_map_load \[#[0-9]*\], src/turtle/io/png16.c:[0-9]* \} "
                               r"could not open file `nothing.png'")):
        with pytest.raises(tb.TurtleError) as e:
            tb.Map(path=path)
        assert e.value.code == code and re.match(pattern, str(e.value))
    with pytest.raises(tb.TurtleError) as e:
        tb.Map(NX, NY, (0, 1), (0, 1), (0, 1), "nothing")
    assert e.value.code == 4
    with pytest.raises(tb.TurtleError) as e:
        tb.Map(0, NY, (0, 1), (0, 1), (0, 1), "UTM 31N")
    assert e.value.code == 6


def test_stack(data):
    path = str(data / "topography").encode()
    stack = C.c_void_p()
    check(lib.turtle_stack_create(C.byref(stack), path, 3, None, None))
    size = lambda: lib.turtle_stack_tiles_loaded(stack)  # noqa: E731
    z, inside = C.c_double(), C.c_int()
    assert size() == 0
    for (lat, lon), n in (((45.5, 3.5), 1), ((45.0, 3.5), 1), ((46.5, 3.5), 2), ((45.0, 3.5), 2)):
        check(lib.turtle_stack_elevation(stack, lat, lon, C.byref(z), None))
        assert z.value == 0 and size() == n
    check(lib.turtle_stack_elevation(stack, 45.5, 4.5, C.byref(z), C.byref(inside)))
    assert inside.value == 0 and size() == 2
    for (lat, lon), n in (((45.5, 2.5), 3), ((46.5, 2.5), 3)):
        check(lib.turtle_stack_elevation(stack, lat, lon, C.byref(z), None))
        assert z.value == 0 and size() == n
    check(lib.turtle_stack_clear(stack))
    assert size() == 0
    check(lib.turtle_stack_elevation(stack, 45.5, 2.5, C.byref(z), None))
    assert z.value == 0 and size() == 1
    check(lib.turtle_stack_load(stack))
    assert size() == 3
    check(lib.turtle_stack_clear(stack))
    assert size() == 0
    check(lib.turtle_stack_load(stack))
    assert size() == 3
    check(lib.turtle_stack_load(stack))
    assert size() == 3
    lib.turtle_stack_destroy(C.byref(stack))
    check(lib.turtle_stack_create(C.byref(stack), path, 0, None, None))
    check(lib.turtle_stack_load(stack))
    assert size() == 4
    lib.turtle_stack_destroy(C.byref(stack))


def test_client(data):
    path = str(data / "topography").encode()
    stack, client = C.c_void_p(), C.c_void_p()
    check(lib.turtle_stack_create(C.byref(stack), path, 1, NOTHING, NOTHING))
    check(lib.turtle_client_create(C.byref(client), stack))
    z, inside = C.c_double(), C.c_int()
    for lat, lon in ((45.5, 3.5), (45.0, 3.5), (46.5, 3.5), (45.0, 3.5)):
        check(lib.turtle_client_elevation(client, lat, lon, C.byref(z), None))
        assert z.value == 0
    for _ in range(2):
        check(lib.turtle_client_elevation(client, 45.5, 4.5, C.byref(z), C.byref(inside)))
        assert inside.value == 0
    for lat, lon in ((45.5, 2.5), (46.5, 2.5)):
        check(lib.turtle_client_elevation(client, lat, lon, C.byref(z), None))
        assert z.value == 0
    check(lib.turtle_client_clear(client))
    check(lib.turtle_client_elevation(client, 45.5, 3.5, C.byref(z), None))
    assert z.value == 0
    check(lib.turtle_client_destroy(C.byref(client)))
    lib.turtle_stack_destroy(C.byref(stack))
    # false cases (:738-770)
    with pytest.raises(tb.TurtleError) as e:
        check(lib.turtle_client_create(C.byref(client), None))
    assert e.value.code == 1 and re.match(
        r"\{ turtle_client_create \[#[0-9]*\], src/turtle/client.c:[0-9]* \} invalid null stack",
        str(e.value))
    check(lib.turtle_stack_create(C.byref(stack), path, 0, None, None))
    with pytest.raises(tb.TurtleError) as e:
        check(lib.turtle_client_create(C.byref(client), stack))
    assert e.value.code == 1 and "stack has no lock" in str(e.value)
    lib.turtle_stack_destroy(C.byref(stack))
    check(lib.turtle_stack_create(C.byref(stack), path, 1, NOTHING, NOTHING))
    check(lib.turtle_client_create(C.byref(client), stack))
    with pytest.raises(tb.TurtleError) as e:
        check(lib.turtle_client_elevation(client, 45.5, 4.5, C.byref(z), None))
    assert e.value.code == 10 and re.match(
        r"\{ turtle_client_elevation \[#[0-9]*\], src/turtle/stack.c:[0-9]* \} missing elevation "
        r"data in `.*topography'", str(e.value))
    check(lib.turtle_client_destroy(C.byref(client)))
    lib.turtle_stack_destroy(C.byref(stack))


def test_projection(data):
    """tests/test-turtle.c:516-578."""
    m = tb.Map(11, 11, (45, 46), (3, 4), (-1, 1), None)
    assert lib.turtle_map_projection(m.handle) is None
    m = tb.Map(path=str(data / "map.png"))
    lib.turtle_projection_name.restype = C.c_char_p
    assert lib.turtle_projection_name(C.c_void_p(lib.turtle_map_projection(m.handle))) == b"UTM 31N"
    projection = C.c_void_p()
    check(lib.turtle_projection_create(C.byref(projection), None))
    assert lib.turtle_projection_name(projection) is None
    for tag in ("Lambert I", "Lambert II", "Lambert IIe", "Lambert III", "Lambert IV",
                "Lambert 93", "UTM 31N", "UTM 3.0N", "UTM 31S", "UTM 3.0S"):
        check(lib.turtle_projection_configure(projection, tag.encode()))
        assert lib.turtle_projection_name(projection) == tag.encode()
        x, y, la, lo = (C.c_double() for _ in range(4))
        check(lib.turtle_projection_project(projection, 45.5, 3.5, C.byref(x), C.byref(y)))
        check(lib.turtle_projection_unproject(projection, x, y, C.byref(la), C.byref(lo)))
        assert abs(la.value - 45.5) <= 1e-8 and abs(lo.value - 3.5) <= 1e-8
    lib.turtle_projection_destroy(C.byref(projection))
    with pytest.raises(tb.TurtleError) as e:
        check(lib.turtle_projection_create(C.byref(projection), b"nothing"))
    assert e.value.code == 4 and re.match(
        r"\{ turtle_projection_create \[#[0-9]*\], src/turtle/projection.c:[0-9]* \} "
        r"invalid projection `nothing'", str(e.value))


def test_strfunc():
    """tests/test-turtle.c:1216-1269: every API function has a name, a stranger has none."""
    lib.turtle_error_function.restype = C.c_char_p
    names = """turtle_client_clear turtle_client_create turtle_client_destroy
        turtle_client_elevation turtle_ecef_from_geodetic turtle_ecef_from_horizontal
        turtle_ecef_to_geodetic turtle_ecef_to_horizontal turtle_error_function
        turtle_error_handler_get turtle_error_handler_set turtle_map_create turtle_map_destroy
        turtle_map_dump turtle_map_elevation turtle_map_fill turtle_map_load turtle_map_meta
        turtle_map_node turtle_map_projection turtle_projection_configure
        turtle_projection_create turtle_projection_destroy turtle_projection_name
        turtle_projection_project turtle_projection_unproject turtle_stack_clear
        turtle_stack_create turtle_stack_destroy turtle_stack_elevation turtle_stack_load
        turtle_stepper_add_flat turtle_stepper_add_layer turtle_stepper_add_map
        turtle_stepper_add_stack turtle_stepper_create turtle_stepper_destroy
        turtle_stepper_geoid_get turtle_stepper_geoid_set turtle_stepper_range_get
        turtle_stepper_range_set turtle_stepper_position turtle_stepper_step""".split()
    raw = C.CDLL(tb._lib.LIB_PATH)
    for name in names:
        fn = C.cast(getattr(raw, name), C.c_void_p)
        assert lib.turtle_error_function(fn) == name.encode(), name
    assert lib.turtle_error_function(C.cast(NOTHING, C.c_void_p)) is None
