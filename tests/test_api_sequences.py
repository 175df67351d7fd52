"""Random SEQUENCES of stepper calls -- add_layer / add_flat / add_map / add_stack in any
order (also after steps), geoid on and off, range / slope / resolution changes, reset,
turtle_stepper_position, steps and queries -- made on the reference and on the product,
compared call by call: same return codes, same outputs to the bit. What the stepper
remembers between calls (the last sample, stepper.c:703-756; the local approximations; what
resets them, stepper.c:617-651; what does not: add_*, slope, resolution) is part of the
drop-in contract.

Not compared: latitude / longitude / altitude of a step on a stepper WITHOUT data -- the
reference returns the uninitialised `last.geographic` of its malloc (stepper.c:548-569)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import harness as H
from turtle_b200 import synth

pytestmark = pytest.mark.skipif(not os.path.exists(H.REF), reason="oracle/_ref not built")
D, P = C.c_double, C.c_void_p


class MapInfo(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("x", D * 2), ("y", D * 2), ("z", D * 2),
                ("encoding", C.c_char_p)]


def bind(path):
    lib = C.CDLL(path)
    for name in ("range", "slope", "resolution"):
        getter, setter = (getattr(lib, "turtle_stepper_%s_%s" % (name, w)) for w in ("get", "set"))
        getter.restype, getter.argtypes = D, [P]
        setter.restype, setter.argtypes = None, [P, D]
    lib.turtle_stepper_geoid_set.restype, lib.turtle_stepper_geoid_set.argtypes = None, [P, P]
    lib.turtle_stepper_geoid_get.restype, lib.turtle_stepper_geoid_get.argtypes = P, [P]
    lib.turtle_stepper_reset.restype, lib.turtle_stepper_reset.argtypes = None, [P]
    lib.turtle_stepper_create.argtypes = [C.POINTER(P)]
    lib.turtle_stepper_destroy.argtypes = [C.POINTER(P)]
    lib.turtle_stepper_add_layer.argtypes = [P]
    lib.turtle_stepper_add_flat.argtypes = [P, D]
    lib.turtle_stepper_add_map.argtypes = [P, P, D]
    lib.turtle_stepper_add_stack.argtypes = [P, P, D]
    lib.turtle_stepper_step.argtypes = ([P, C.POINTER(D), C.POINTER(D)] + [C.POINTER(D)] * 5 +
                                        [C.POINTER(C.c_int)])
    lib.turtle_stepper_position.argtypes = [P, D, D, D, C.c_int, C.POINTER(D),
                                            C.POINTER(C.c_int)]
    lib.turtle_map_create.argtypes = [C.POINTER(P), C.POINTER(MapInfo), C.c_char_p]
    lib.turtle_map_fill.argtypes = [P, C.c_int, C.c_int, D]
    lib.turtle_stack_create.argtypes = [C.POINTER(P), C.c_char_p, C.c_int, P, P]
    lib.turtle_ecef_from_geodetic.restype = None
    lib.turtle_ecef_from_geodetic.argtypes = [D, D, D, C.POINTER(D)]
    return lib


@pytest.fixture(scope="module")
def worlds(tmp_path_factory):
    """Per library: three maps (geodetic, UTM 31N, a 1 degree geoid) and a 2 x 2 tile stack,
    shared by all the sequences (steppers do not modify them)."""
    tiles = str(tmp_path_factory.mktemp("api_sequences"))
    synth.write_hgt_stack(tiles, 45, 2, 2, 2, n=1201)
    rng = np.random.default_rng(77)
    shapes = [(31, 27, (2.4, 3.1), (45.2, 45.9), (0., 3000.), None),
              (31, 27, (486000., 506000.), (5057000., 5077000.), (0., 3000.), b"UTM 31N"),
              (73, 37, (0., 360.), (-90., 90.), (-100., 100.), None)]
    values = [rng.uniform(0.5 * z[0], 0.5 * z[1], (ny, nx)) for nx, ny, _, _, z, _ in shapes]
    out = {}
    for key, path in (("reference", H.REF), ("product", H.PRODUCT)):
        lib = bind(path)
        maps = []
        for (nx, ny, x, y, z, tag), v in zip(shapes, values):
            info = MapInfo(nx, ny, (D * 2)(*x), (D * 2)(*y), (D * 2)(*z), None)
            m = P()
            assert lib.turtle_map_create(C.byref(m), C.byref(info), tag) == 0
            for iy in range(ny):
                for ix in range(nx):
                    assert lib.turtle_map_fill(m, ix, iy, float(v[iy, ix])) == 0
            maps.append(m)
        stack = P()
        assert lib.turtle_stack_create(C.byref(stack), tiles.encode(), 0, None, None) == 0
        out[key] = (lib, maps, stack)
    yield out
    if os.path.realpath(H.PRODUCT):  # hand the product's one error handler back to the binding
        from turtle_b200 import api
        api.install_handler()


def call(lib, maps, stack, stepper, position, op, a):
    """One call of the sequence; returns everything it answered."""
    if op == 0:
        return lib.turtle_stepper_add_layer(stepper)
    if op == 1:
        return lib.turtle_stepper_add_flat(stepper, a["offset"])
    if op == 2:
        return lib.turtle_stepper_add_map(stepper, maps[a["map"]], a["offset"])
    if op == 3:
        return lib.turtle_stepper_add_stack(stepper, stack, a["offset"])
    if op == 4:
        lib.turtle_stepper_geoid_set(stepper, maps[2] if a["which"] else None)
        return bool(lib.turtle_stepper_geoid_get(stepper))
    if op in (5, 6, 7):
        name = ("range", "slope", "resolution")[op - 5]
        value = ((0., 1., 10., -1.), (0.4, 1., 0.05, 2.), (1e-2, 1., 1e-3, 10.))[op - 5][a["which"]]
        getattr(lib, "turtle_stepper_%s_set" % name)(stepper, value)
        return getattr(lib, "turtle_stepper_%s_get" % name)(stepper)
    if op == 8:
        lib.turtle_stepper_reset(stepper)
        return 0
    if op == 9:
        p, index = (D * 3)(), C.c_int(-7)
        rc = lib.turtle_stepper_position(stepper, a["latitude"], a["longitude"], a["offset"],
                                         a["layer"], p, C.byref(index))
        if (rc == 0) and a["take"]:
            position[:] = list(p)
        return (rc, tuple(p), index.value) if rc == 0 else (rc,)
    lat, lon, alt, step = D(-1), D(-1), D(-1), D(-1)
    elevation, index = (D * 2)(-1, -1), (C.c_int * 2)(-9, -9)
    direction = None if a["query"] else (D * 3)(*a["direction"])
    rc = lib.turtle_stepper_step(stepper, position, direction, C.byref(lat), C.byref(lon),
                                 C.byref(alt), elevation, C.byref(step), index)
    geographic = (lat.value, lon.value, alt.value) if a["has_data"] else ()
    return (rc, tuple(position)) + geographic + (tuple(elevation), step.value, tuple(index))


@pytest.mark.parametrize("seed", range(60))
def test_call_by_call(worlds, seed):
    rng = np.random.default_rng(seed)
    steppers, positions = {}, {}
    latitude, longitude = rng.uniform(45.1, 46.9), rng.uniform(2.1, 3.9)
    for key, (lib, maps, stack) in worlds.items():
        lib.turtle_error_handler_set(None)  # errors come back as return codes
        steppers[key] = P()
        assert lib.turtle_stepper_create(C.byref(steppers[key])) == 0
        positions[key] = (D * 3)()
        lib.turtle_ecef_from_geodetic(latitude, longitude, rng.uniform(-100, 3000),
                                      positions[key])
    positions["product"][:] = list(positions["reference"])
    n_data, history = 0, []
    try:
        for _ in range(int(rng.integers(5, 60))):
            op = int(rng.integers(14))
            direction = rng.standard_normal(3)
            a = dict(offset=float(rng.uniform(-100, 500)), map=int(rng.integers(3)),
                     layer=int(rng.integers(-1, 4)), which=int(rng.integers(4)),
                     take=bool(rng.random() < 0.5), query=bool(rng.random() < 0.2),
                     direction=direction / np.linalg.norm(direction),
                     latitude=latitude + 0.01 * int(rng.integers(4)), longitude=longitude,
                     has_data=n_data > 0)
            answers = {key: call(lib, maps, stack, steppers[key], positions[key], op, a)
                       for key, (lib, maps, stack) in worlds.items()}
            history.append(op)
            assert repr(answers["product"]) == repr(answers["reference"]), history
            if (op in (1, 2, 3)) and (answers["reference"] == 0):
                n_data += 1
    finally:
        for key, (lib, maps, stack) in worlds.items():
            lib.turtle_stepper_destroy(C.byref(steppers[key]))
