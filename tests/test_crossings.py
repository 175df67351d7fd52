"""Medium changes along every ray (turtle_stepper_trace_crossings):
where each medium begins, as the bisection of turtle_stepper_step locates it
(stepper.c:832-864). The oracle side is the canonical ray loop of the pthread driver over
the unmodified reference; the product side is the trace kernel."""
import os

import numpy as np
import pytest

import turtle_b200 as tb
from oracle import harness as H
from tests.test_gpu_trace import c3
from turtle_b200 import synth

K = 6


def _rays(ora, n, seed):
    rng = np.random.default_rng(seed)
    pos = ora.ecef_from_geodetic(rng.uniform(45.0, 46.9, n), rng.uniform(2.0, 3.9, n),
                                 rng.uniform(-300, 4000, n))
    return pos, synth.random_unit(n, seed + 1)


def _consistent(res, cross):
    """What must hold whatever made the records."""
    n_valid = np.minimum(res["n_changes"], K)
    for i in range(len(res)):
        c = cross[i, :n_valid[i]]
        assert (np.diff(c["length"]) >= 0).all() and (c["length"] <= res["total"][i] + 1e-9).all()
        assert (c["from"] != c["to"]).all()
        assert (c["to"][:-1] == c["from"][1:]).all()     # the chain of media is continuous
        if n_valid[i] == res["n_changes"][i] and n_valid[i] > 0:
            assert c["to"][-1] == res["index"][i, 0]     # ... and ends in the final medium
        assert (cross[i, n_valid[i]:]["length"] == 0).all()  # unused slots stay untouched


def test_oracle_crossings_reference_equals_restatement(small_stack):
    sc = c3(small_stack, 0., -1)
    ref, port = sc.oracle(H.best_oracle()), sc.oracle(H.PORT)
    pos, dirs = _rays(ref, 3000, 31)
    rule = H.rule(9000., length_max=5e4, max_steps=20000)
    ra, ca = ref.trace_crossings(pos, dirs, rule, K)
    rb, cb = port.trace_crossings(pos, dirs, rule, K)
    assert ra.tobytes() == rb.tobytes() and ca.tobytes() == cb.tobytes()
    _consistent(ra, ca)
    assert (ra["n_changes"] > 1).sum() > 100 and (ra["n_changes"] > K).sum() >= 0
    # the plain trace is the same trace
    assert ref.trace(pos, dirs, rule)[0].tobytes() == ra.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("rg,geoid", [(0., -1), (10., 1)])
def test_crossings_vs_oracle(small_stack, rg, geoid):
    sc = c3(small_stack, rg, geoid)
    ora = sc.oracle(locked=True)
    n = 20000 + 13
    pos, dirs = _rays(ora, n, 77)
    want, wcross = ora.trace_crossings(pos, dirs, H.rule(9000., length_max=5e4, max_steps=20000),
                                       K, threads=os.cpu_count())
    stepper, maps, stacks = sc.product()
    plan = stepper.freeze(0)
    rule = tb.trace_rule(9000., length_max=5e4, max_steps=20000)
    got, gcross = plan.trace_crossings(pos, dirs, rule, K)
    _consistent(got, gcross)
    # the records are those of the plain trace
    assert plan.trace(pos, dirs, rule).tobytes() == got.tobytes()
    same = (want["n_steps"] == got["n_steps"]) & (want["medium_hash"] == got["medium_hash"]) & \
        (want["status"] == got["status"])
    assert (~same).sum() <= max(3, n // 2000)            # grazing rays
    w, g = wcross[same], gcross[same]
    assert np.array_equal(w["from"], g["from"]) and np.array_equal(w["to"], g["to"])
    # boundaries are bisection-located: 1e-9 relative / 1 mm
    tol = np.maximum(1e-3, 1e-9 * np.abs(w["length"]))
    assert (np.abs(w["length"] - g["length"]) <= tol).all()
    assert (want["n_changes"] > 0).sum() > n // 10
    # capacity 0: no buffer needed, same records
    got0, none = plan.trace_crossings(pos[:100], dirs[:100], rule, 0)
    assert got0.tobytes() == got[:100].tobytes() and none.shape == (100, 0)
