"""GPU: one turtle_stepper_step for n independent particles (turtle_stepper_step_batch)
and ray origins (turtle_stepper_position_batch) against the oracle."""
import os

import numpy as np
import pytest

import turtle_b200 as tb
from oracle import harness as H
from tests.common import ulp_distance
from tests.test_gpu_trace import c3
from turtle_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


@pytest.mark.parametrize("rg,geoid", [(0., -1), (10., 1)])
def test_single_steps_golden(small_stack, rg, geoid):
    """Fresh stepper per particle (states = NULL): the committed reference outputs."""
    stepper, maps, stacks = c3(small_stack, rg, geoid).product()
    plan = stepper.freeze(0)
    key = "c3_r%d" % int(rg)
    got = plan.step(GOLD[key + "_pos"].copy(), GOLD[key + "_dir"])
    want = {f: GOLD[key + "_step_" + f] for f in got}
    same_medium = (got["index"] == want["index"]).all(1)
    assert (~same_medium).sum() <= 1
    assert np.array_equal(got["altitude"][same_medium], want["altitude"][same_medium]) or \
        np.abs(got["altitude"] - want["altitude"])[same_medium].max() < 1e-6
    assert np.abs(got["step"] - want["step"])[same_medium].max() < 1e-6
    assert np.abs(got["position"] - want["position"])[same_medium].max() < 1e-6
    assert ulp_distance(got["latitude"], want["latitude"])[same_medium].max() <= 64
    fin = same_medium & (want["index"][:, 0] >= 0)
    e_w, e_g = want["elevation"][fin], got["elevation"][fin]
    assert np.array_equal(np.abs(e_w) > 1e300, np.abs(e_g) > 1e300)  # +-DBL_MAX sentinels
    small = np.abs(e_w) < 1e300
    assert np.abs(e_w[small] - e_g[small]).max() < 1e-6
    # query mode: no direction, position untouched, step = slope * distance
    q = plan.step(GOLD[key + "_pos"].copy(), None)
    assert np.array_equal(q["position"], GOLD[key + "_pos"])
    out = q["index"][:, 0] < 0
    assert (q["step"][out] == 0.).all() and (q["elevation"][out] == 0.).all()


def test_particle_walk_with_states(small_stack):
    """Config-4 shape: many consecutive steps with a fresh direction each, stepper state
    (last sample + local approximation) kept per particle on the device."""
    sc = c3(small_stack, 1., -1)
    ora = sc.oracle(locked=True)
    stepper, maps, stacks = sc.product()
    plan = stepper.freeze(0)
    rng = np.random.default_rng(31)
    n, k = 4000, 12
    la, lo = rng.uniform(45.2, 45.9, n), rng.uniform(2.2, 2.9, n)
    ground, idx = ora.position(la, lo, rng.uniform(-50, 50, n), 0)
    dirs = synth.random_unit(n * k, 4).reshape(k, n, 3)
    want = ora.walk(ground, dirs, threads=os.cpu_count())
    states = plan.states(n)
    pos = ground.copy()
    bad = np.zeros(n, dtype=bool)
    for j in range(k):
        out = plan.step(pos, dirs[j], states=states)
        pos = out["position"]
        bad |= (out["index"] != want["index"][j]).any(1)
        ok = ~bad
        assert np.abs(out["step"] - want["step"][j])[ok].max() < 1e-5
        assert np.abs(out["altitude"] - want["altitude"][j])[ok].max() < 1e-5
    assert bad.sum() <= max(2, n // 500)
    assert np.abs(pos - want["position"])[~bad].max() < 1e-4


def test_position_batch(small_stack):
    sc = c3(small_stack, 0., 1)
    ora = sc.oracle()
    stepper, maps, stacks = sc.product()
    plan = stepper.freeze(0)
    rng = np.random.default_rng(32)
    n = 20000
    la, lo, h = rng.uniform(44.9, 47.1, n), rng.uniform(1.9, 4.1, n), rng.uniform(-5, 50, n)
    for layer in (0, 1):
        wp, wi = ora.position(la, lo, h, layer)
        gp, gi = plan.position(la, lo, h, layer)
        assert np.array_equal(wi, gi)
        assert np.abs(wp - gp).max() < 1e-8  # sin / cos of from_geodetic: a few ulp of 6.4e6 m
    assert (wi == -1).any()
    with pytest.raises(tb.TurtleError):
        plan.position(la, lo, h, 5)
