"""Host-side object life cycle of the scalar turtle.h calls (no GPU): the cases the
reference handles that a residency-minded implementation gets wrong easily.

 * a directory whose only tile has a null span makes an EMPTY stack (stack.c:162-163);
 * the scalar stepper resolves stack tiles on demand and honours the `size` given to
   turtle_stack_create (stack.c:413-449) instead of loading the whole directory;
 * a stack may be destroyed before the stepper that used it.
"""
import ctypes as C
import os

import numpy as np

import turtle_b200 as tb
from oracle import harness as H
from turtle_b200 import synth
from turtle_b200._lib import lib


def test_stack_of_null_span_tiles_is_empty(tmp_path):
    d = tmp_path / "grd"
    d.mkdir()
    # header "y0 y1 x0 x1 dy dx" (grd.c:75-86): one row of two nodes, no latitude span
    (d / "one.grd").write_text("0 0 0 1 1 1\n0.5 1.5\n")
    st = tb.Stack(str(d))
    z, inside = st.elevation(0., 0.)
    assert inside == 0 and z == 0.
    assert lib.turtle_stack_tiles_loaded(st.handle) == 0


def test_scalar_stepper_loads_tiles_on_demand(small_stack):
    """size = 1: at most one tile of the 3-tile stack is ever resident on the host, and the
    answers are those of the reference stepping through the same rays."""
    st = tb.Stack(small_stack, size=1)
    s = tb.Stepper(range=0.)
    s.add_flat(0.)
    s.add_stack(st, 0.)
    ora = H.Driver(H.best_oracle())
    ost = ora.stack_create(small_stack)
    ora.geometry([(H.ADD_FLAT, 0, 0.), (H.ADD_STACK, ost, 0.)], range=0.)
    rng = np.random.default_rng(5)
    n = 200
    pos = ora.ecef_from_geodetic(rng.uniform(45.1, 46.9, n), rng.uniform(2.1, 3.9, n),
                                 rng.uniform(10., 3000., n))
    dirs = synth.random_unit(n, 9)
    want = ora.step(pos, dirs)
    most = 0
    for i in range(n):
        p = (C.c_double * 3)(*pos[i])
        d = (C.c_double * 3)(*dirs[i])
        alt, step = C.c_double(), C.c_double()
        idx = (C.c_int * 2)()
        tb.api._check(lib.turtle_stepper_step(s.handle, p, d, None, None, C.byref(alt), None,
                                              C.byref(step), idx))
        most = max(most, lib.turtle_stack_tiles_loaded(st.handle))
        assert step.value == want["step"][i] and alt.value == want["altitude"][i]
        assert list(idx) == list(want["index"][i]) and list(p) == list(want["position"][i])
    assert most == 1


def test_stack_destroyed_before_its_stepper(small_stack):
    st = tb.Stack(small_stack)
    s = tb.Stepper(range=0.)
    s.add_stack(st, 0.)
    pos, idx = s.position(45.5, 2.5, 1., 0)
    assert idx == 0
    lib.turtle_stack_destroy(C.byref(st._p))  # the stepper still refers to it
    del s                                      # ... and must not touch it when it goes
