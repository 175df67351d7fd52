"""GPU: whole-ray traces (turtle_stepper_trace_batch) against the oracle on the same
rays.

Parity protocol (oracle/parity.py): the per-ray DISCRETE outcome (step
count, stop status, final layer / data index, number and sequence of media) is compared
exactly; rays for which it differs are "grazing" rays and are counted (bound below).
EVERY continuous field north_star names -- all length[m], total, exit position, altitude
-- is compared for every ray: quantities that end on a bisected boundary against 1e-9
relative / 1 mm, quantities of rays stopped by an altitude / length / step threshold
against a small multiple of the reference's own rounding-noise floor (the reference vs
the reference compiled with FMA contraction, same rays, same test).
"""
import os

import numpy as np
import pytest

import turtle_b200 as tb
from oracle import harness as H
from oracle import parity as P
from tests.common import Scene, compare_traces, geoid_map, lambert_map, utm_map
from turtle_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


def noise_floor(scene, ref, pos, dirs, rule, locked=False):
    """The reference against ITSELF built with FMA contraction on the same rays
    (oracle/_ref/libturtle_ref_fma.so): the rounding-noise floor of every field."""
    if not H.available(H.REF_FMA) or H.best_oracle() != H.REF:
        return None
    fma, _, _ = scene.oracle(library=H.REF_FMA, locked=locked).trace(
        pos, dirs, rule, threads=os.cpu_count() if locked else 1)
    return P.report(ref, fma)


def check(ref, got, max_grazing, floor=None):
    """Every field north_star names, none dropped (oracle/parity.py): the discrete outcome
    is exact but for counted grazing rays; boundary-located quantities (lengths in media
    the ray has left, everything about rays that left the data) are held to 1 mm / 1e-9;
    threshold-exit quantities are held to a multiple of the reference's own FMA noise."""
    rep = P.report(ref, got)
    assert rep["discrete_mismatch"] <= max_grazing, P.table(rep, floor)
    assert rep["located_rays_over"] <= max_grazing, P.table(rep, floor)
    if floor is not None:
        bad = P.against_floor(rep, floor)
        assert not bad, "\n".join(bad) + "\n" + P.table(rep, floor)
    return rep


def c1(n_map=201, rg=0.):
    return Scene(maps=[utm_map(n=n_map)], ops=[(H.ADD_FLAT, 0, -100.), (H.ADD_LAYER, 0, 0.),
                                              (H.ADD_MAP, 0, 0.)], range=rg)


def c3(stack_dir, rg, geoid):
    lm = lambert_map(H.Driver(H.best_oracle()), n=201)
    return Scene(maps=[lm, geoid_map()], stacks=[stack_dir],
                 ops=[(H.ADD_FLAT, 0, 0.), (H.ADD_STACK, 0, 0.), (H.ADD_MAP, 0, 0.),
                      (H.ADD_LAYER, 0, 0.), (H.ADD_STACK, 0, 500.), (H.ADD_MAP, 0, 600.)],
                 geoid=geoid, range=rg)


@pytest.mark.parametrize("rg", [0., 10.])
def test_golden_fan_through_utm_map(rg):
    """Committed reference results (tests/golden), config-1 shape."""
    stepper, maps, stacks = c1(rg=rg).product()
    key = "c1_r%d" % int(rg)
    got = stepper.freeze(0).trace(GOLD[key + "_pos"], GOLD[key + "_dir"], tb.trace_rule(3100.))
    floor = noise_floor(c1(rg=rg), GOLD[key + "_res"], GOLD[key + "_pos"], GOLD[key + "_dir"],
                        H.rule(3100.))
    check(GOLD[key + "_res"], got, max_grazing=1, floor=floor)


@pytest.mark.parametrize("rg,geoid", [(0., -1), (10., 1)])
def test_golden_layered_geometry(small_stack, rg, geoid):
    """Committed reference results, config-3 shape: flat / stack / Lambert map, two
    layers, geoid, local approximation on; origins inside, outside and below ground."""
    stepper, maps, stacks = c3(small_stack, rg, geoid).product()
    key = "c3_r%d" % int(rg)
    got = stepper.freeze(0).trace(GOLD[key + "_pos"], GOLD[key + "_dir"],
                                  tb.trace_rule(9000., length_max=5e4))
    floor = noise_floor(c3(small_stack, rg, geoid), GOLD[key + "_res"], GOLD[key + "_pos"],
                        GOLD[key + "_dir"], H.rule(9000., length_max=5e4))
    check(GOLD[key + "_res"], got, max_grazing=2, floor=floor)


@pytest.mark.parametrize("rg,geoid", [(0., -1), (1., 1), (100., -1)])
def test_random_rays_vs_oracle(small_stack, rg, geoid):
    sc = c3(small_stack, rg, geoid)
    ora = sc.oracle(locked=True)
    rng = np.random.default_rng(21)
    n = 30000 + 17  # ragged: not a multiple of the warp size
    pos = ora.ecef_from_geodetic(rng.uniform(44.9, 47.1, n), rng.uniform(1.9, 4.1, n),
                                 rng.uniform(-500, 5000, n))
    dirs = synth.random_unit(n, 5)
    want, steps, _ = ora.trace(pos, dirs, H.rule(9000., length_max=1e5, max_steps=20000),
                               threads=os.cpu_count())
    stepper, maps, stacks = sc.product()
    plan = stepper.freeze(0)
    got = plan.trace(pos, dirs, tb.trace_rule(9000., length_max=1e5, max_steps=20000))
    floor = noise_floor(sc, want, pos, dirs, H.rule(9000., length_max=1e5, max_steps=20000),
                        locked=True)
    rep = check(want, got, max_grazing=max(3, n // 2000), floor=floor)
    c = plan.counters()
    assert c["rays"] == n and abs(c["steps"] - steps) <= 50 * max(1, rep["discrete_mismatch"])
    assert (want["status"] == tb.api.TRACE_LENGTH).any()    # some ran into the length cap
    assert (want["n_changes"] > 0).sum() > n // 10          # boundaries were crossed


def test_edge_cases(small_stack):
    sc = c3(small_stack, 0., -1)
    ora = sc.oracle()
    stepper, maps, stacks = sc.product()
    plan = stepper.freeze(0)
    rule = tb.trace_rule(9000., length_max=1e5)
    # empty batch
    assert len(plan.trace(np.zeros((0, 3)), np.zeros((0, 3)), rule)) == 0
    # one ray; NaN / inf inputs are flagged, never traced
    pos = ora.ecef_from_geodetic([45.5, 45.5, 45.5], [2.5, 2.5, 2.5], [100., 100., 100.])
    dirs = synth.random_unit(3, 1)
    pos[1, 0] = np.nan
    dirs[2, 1] = np.inf
    got = plan.trace(pos, dirs, rule)
    want, _, _ = ora.trace(pos, dirs, H.rule(9000., length_max=1e5))
    assert list(got["status"][1:]) == [tb.api.TRACE_INVALID] * 2
    assert list(want["status"][1:]) == [4, 4] and got["n_steps"][0] == want["n_steps"][0]
    # step cap and length cap
    pos = ora.ecef_from_geodetic([45.5] * 64, [2.5] * 64, [4000.] * 64)
    dirs = synth.random_unit(64, 2)
    for r_o, r_g in ((H.rule(1e9, max_steps=7), tb.trace_rule(1e9, max_steps=7)),
                     (H.rule(1e9, length_max=500.), tb.trace_rule(1e9, length_max=500.))):
        want, _, _ = ora.trace(pos, dirs, r_o)
        got = plan.trace(pos, dirs, r_g)
        assert np.array_equal(want["status"], got["status"])
        assert np.array_equal(want["n_steps"], got["n_steps"])
    # invalid rule is an argument error through the handler
    with pytest.raises(tb.TurtleError):
        plan.trace(pos, dirs, tb.trace_rule(1e9, max_steps=0))


def test_determinism_and_chunking(small_stack):
    """Size-independent properties: results do not depend on lane scheduling, on the
    chunking of the host pipeline, nor on the host / device entry point."""
    import torch
    sc = Scene(stacks=[small_stack], ops=[(H.ADD_STACK, 0, 0.)], range=0.)
    stepper, maps, stacks = sc.product()
    plan = stepper.freeze(0)
    n = (1 << 20) + (1 << 19) + 333  # > 1 pipeline chunk, ragged
    origin, _ = stepper.position(45.4, 2.6, 1.0, 0)
    dirs = synth.fan_directions(45.4, 2.6, 2048, 1024)[:n]
    pos = np.repeat(origin[None], n, 0)
    rule = tb.trace_rule(6000., max_steps=20000)
    a = plan.trace(pos, dirs, rule)
    b = plan.trace(pos, dirs, rule)
    assert a.tobytes() == b.tobytes()
    plan.launch_set(2, 64)  # another grid: same answers
    c = plan.trace(pos, dirs, rule)
    assert a.tobytes() == c.tobytes()
    plan.schedule_set(1)    # longest-expected-first queue order: same answers
    c = plan.trace(pos, dirs, rule)
    assert a.tobytes() == c.tobytes()
    plan.schedule_set(0)
    os.environ["TURTLE_B200_STREAM_MAX_RAYS"] = str(1 << 20)  # two streamed rounds
    c = plan.trace(pos, dirs, rule)
    del os.environ["TURTLE_B200_STREAM_MAX_RAYS"]
    assert a.tobytes() == c.tobytes() and plan.counters()["rays"] == n
    plan.pipeline_set(1)    # one kernel per chunk instead of the streamed single kernel
    c = plan.trace(pos, dirs, rule)
    assert a.tobytes() == c.tobytes()
    plan.pipeline_set(0)
    plan.specialise_set(0)  # generic list-walking kernel instead of the single-stack one
    c = plan.trace(pos, dirs, rule)
    assert a.tobytes() == c.tobytes()
    plan.specialise_set(1)
    d_res = torch.empty((n, 96), dtype=torch.uint8, device="cuda:0")
    plan.trace_device(n, torch.from_numpy(pos).cuda(), torch.from_numpy(dirs).cuda(), rule, d_res)
    torch.cuda.synchronize()
    assert d_res.cpu().numpy().tobytes() == a.tobytes()
    # the node gathers of the single-stack kernel: cell-packed tiles (one 8-byte load per
    # sample), a shared-memory window staged by the bulk copy engine -- same answers
    d_pos, d_dir = torch.from_numpy(pos).cuda(), torch.from_numpy(dirs).cuda()
    for mode in (1, 2):
        plan.gather_set(mode, 45.4, 2.6)
        d_res.zero_()
        plan.trace_device(n, d_pos, d_dir, rule, d_res)
        torch.cuda.synchronize()
        assert d_res.cpu().numpy().tobytes() == a.tobytes(), mode
        if mode == 2:  # the station sits in the window: a good part of the samples hit it
            c = plan.counters(sync=True)
            assert 0 < c["window_hits"] <= c["samples"]
    plan.gather_set(0)
    # a permutation of the rays permutes the results (no cross-ray state)
    perm = np.random.default_rng(3).permutation(n)
    p = plan.trace(pos[perm], dirs[perm], rule)
    assert p.tobytes() == a[perm].tobytes()
    # physics: rock length is bounded by the total length, all rays terminate
    assert (a["length"].sum(1) <= a["total"] * (1 + 1e-12) + 1e-9).all()
    assert (a["status"] <= tb.api.TRACE_STEPS).all() and (a["status"] != tb.api.TRACE_STEPS).mean() > 0.99
