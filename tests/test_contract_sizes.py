"""GPU: every BASELINE.json configuration at contract shapes -- the 2001 x 2001 UTM map, the
3 x 3 stack of 3601 x 3601 tiles, the 20 000 x 20 000 map -- with at least 1 Mi rays /
particles / points per case, against the reference's CPU path on a strided sample of the same
inputs. The parity protocol is the one of tests/test_gpu_trace.py (oracle/parity.py): discrete
outcome exact but for counted grazing rays, boundary-located quantities to 1 mm / 1e-9,
threshold-exit quantities against the reference's own FMA noise floor.

(tests/test_gpu_trace.py holds the same kernels on small shapes, where the reference can
trace every ray; bench.py and tools/bench_configs.py time the full sizes.)"""
import os
import sys

import numpy as np
import pytest

import turtle_b200 as tb
from oracle import harness as H
from oracle import parity as P
from turtle_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu
N = 1 << 20
CORES = os.cpu_count() or 1


@pytest.fixture(scope="module")
def configs():
    import bench_configs  # tools/: the scenes and ray sets of the configurations
    return bench_configs


def trace_and_compare(scene, pos, dirs, rule_g, rule_o, stride, max_grazing):
    stepper, maps, stacks = scene.product()
    plan = stepper.freeze(0)
    got = plan.trace(pos, dirs, rule_g)
    ref, _, _ = scene.oracle(locked=True).trace(pos[::stride], dirs[::stride], rule_o, threads=CORES)
    rep = P.report(ref, got[::stride])
    floor = None
    if H.available(H.REF_FMA) and H.best_oracle() == H.REF:
        fma, _, _ = scene.oracle(library=H.REF_FMA, locked=True).trace(
            pos[::stride], dirs[::stride], rule_o, threads=CORES)
        floor = P.report(ref, fma)
    assert rep["discrete_mismatch"] <= max_grazing, P.table(rep, floor)
    assert rep["located_rays_over"] <= max_grazing, P.table(rep, floor)
    if floor is not None:
        bad = P.against_floor(rep, floor)
        assert not bad, "\n".join(bad) + "\n" + P.table(rep, floor)
    assert (got["status"] <= tb.api.TRACE_STEPS).all()
    return got, rep


@pytest.mark.parametrize("rg", [0., 10.])
def test_c1_fan_through_the_2001_utm_map(configs, rg):
    scene, pos, dirs = configs.c1_inputs(N, rg)
    got, rep = trace_and_compare(scene, pos, dirs, tb.trace_rule(3100.), H.rule(3100.), 16, 2)
    assert rep["rays"] == N // 16 and (got["n_steps"] > 0).all()


def test_c2_fan_through_the_3x3_srtmgl1_stack(configs):
    import bench
    bench.make_stack()
    scene = configs.Scene(stacks=[bench.stack_dir()], ops=[(H.ADD_STACK, 0, 0.)], range=0.)
    stepper, _, _ = scene.product()
    origin, _ = stepper.position(bench.DET_LAT, bench.DET_LON, bench.DET_HEIGHT, 0)
    dirs = bench.fan_subsample(0, 1, (bench.N_AZ * bench.N_EL) // N, bench.N_AZ * bench.N_EL)
    pos = np.repeat(origin[None], len(dirs), 0)
    rule_g = tb.trace_rule(bench.ALTITUDE_MAX, max_steps=bench.MAX_STEPS)
    rule_o = H.rule(bench.ALTITUDE_MAX, max_steps=bench.MAX_STEPS)
    got, rep = trace_and_compare(scene, pos, dirs, rule_g, rule_o, 8, 2)
    assert rep["rays_domain_exit"] > 0 and rep["rays_threshold_exit"] > 0
    # the rock length, the number a muography caller keeps: boundary-located for every ray
    assert rep["located"]["length0"]["over"] == 0


@pytest.mark.parametrize("rg", [0., 10.])
def test_c3_random_rays_through_the_layered_geometry(configs, rg):
    scene, pos, dirs = configs.c3_inputs(N, rg)
    rule_g, rule_o = tb.trace_rule(9000., length_max=1e5), H.rule(9000., length_max=1e5)
    got, rep = trace_and_compare(scene, pos, dirs, rule_g, rule_o, 16, 3)
    # (the flat layer holds every ray: they end on the altitude or the length rule)
    assert (got["status"] == tb.api.TRACE_LENGTH).any() and (got["status"] == tb.api.TRACE_ALTITUDE).any()
    assert (got["n_changes"] > 0).mean() > 0.3


@pytest.mark.parametrize("rg", [0., 1.])
def test_c4_particles_random_walk(configs, rg):
    """1 Mi particles x 12 turtle_stepper_step with device-resident states against the
    reference walking every 32-nd particle through the same directions."""
    import torch
    scene = configs.layered_scene(rg)
    stepper, _, _ = scene.product()
    plan = stepper.freeze(0)
    n, k, stride = N, 12, 32
    lat = 45.5 + 2. * synth.random_uniform(n, 0xC4, 0)
    lon = 2.5 + 2. * synth.random_uniform(n, 0xC4, 1)
    h = -50. + 100. * synth.random_uniform(n, 0xC4, 2)
    origin, idx = plan.position(lat, lon, h, 0)
    assert (idx >= 0).all()
    states = plan.states(n)
    d_pos = torch.from_numpy(origin).cuda()
    d_step = torch.empty(n, dtype=torch.float64, device="cuda")
    d_alt = torch.empty(n, dtype=torch.float64, device="cuda")
    d_idx = torch.empty((n, 2), dtype=torch.int32, device="cuda")
    dirs, steps, alts, idxs = [], [], [], []
    for j in range(k):
        d = synth.random_unit(n, 0xC400 + j)
        plan.step_device(n, d_pos, torch.from_numpy(d).cuda(), states=states, altitude=d_alt,
                         step=d_step, index=d_idx)
        torch.cuda.synchronize()
        dirs.append(d[::stride])
        steps.append(d_step[::stride].cpu().numpy())
        alts.append(d_alt[::stride].cpu().numpy())
        idxs.append(d_idx[::stride].cpu().numpy())
    want = scene.oracle(locked=True).walk(origin[::stride], np.stack(dirs), threads=CORES)
    same = np.logical_and.accumulate((np.stack(idxs) == want["index"]).all(2), 0)
    assert (~same[-1]).sum() <= 2                      # particles off the reference's track
    assert np.abs(np.stack(steps) - want["step"])[same].max() < 1e-3
    assert np.abs(np.stack(alts) - want["altitude"])[same].max() < 1e-3
    assert (np.stack(steps) == want["step"])[same].mean() > 0.2  # many are bit-identical


def test_c5_queries_on_the_20k_map(configs):
    """16 Mi points -> turtle_ecef_to_geodetic_batch, turtle_map_elevation_batch and the
    fused kernel on the 20 000 x 20 000 map. Altitude is bit-exact, latitude / longitude
    within 4 ulp of the reference; elevations against the bilinear expression of
    map.c:229-277 evaluated (numpy, same operation order) at the REFERENCE's coordinates."""
    import torch
    from turtle_b200._lib import lib
    mp, row, col, (lon0, lat0, box) = configs.c5_map(20000)
    n = 1 << 24
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0xC5)
    u = torch.rand((3, n), generator=gen, device="cuda", dtype=torch.float64)
    la = (lat0 - 0.05 + (box + 0.1) * u[0]).contiguous()
    lo = (lon0 - 0.05 + (box + 0.1) * u[1]).contiguous()
    al = (5000. * u[2]).contiguous()
    d_ecef = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    ptr = lambda t: t.data_ptr()  # noqa: E731
    tb.api._check(lib.turtle_ecef_from_geodetic_batch_device(n, ptr(la), ptr(lo), ptr(al),
                                                             ptr(d_ecef), None))
    d_lat, d_lon, d_alt = (torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(3))
    d_z = torch.zeros(n, dtype=torch.float64, device="cuda")
    d_in = torch.zeros(n, dtype=torch.int32, device="cuda")
    tb.api._check(lib.turtle_ecef_to_geodetic_batch_device(n, ptr(d_ecef), ptr(d_lat), ptr(d_lon),
                                                           ptr(d_alt), None))
    tb.api._check(lib.turtle_map_elevation_batch_device(mp.handle, n, ptr(d_lon), ptr(d_lat),
                                                        ptr(d_z), ptr(d_in), None))
    torch.cuda.synchronize()
    stride = 64
    ecef = d_ecef[::stride].cpu().numpy()
    wla, wlo, wal = H.Driver(H.best_oracle()).ecef_to_geodetic(ecef)
    gla, glo, gal = (t[::stride].cpu().numpy() for t in (d_lat, d_lon, d_alt))
    assert np.array_equal(wal, gal)
    assert np.abs(wla.view(np.int64) - gla.view(np.int64)).max() <= 4
    assert np.abs(wlo.view(np.int64) - glo.view(np.int64)).max() <= 4
    # map.c:242-273 at the reference's coordinates
    nn = 20000
    dx = box / (nn - 1)
    hx, hy = (wlo - lon0) / dx, (wla - lat0) / dx
    inside = (hx >= 0) & (hx <= nn - 1) & (hy >= 0) & (hy <= nn - 1)
    ix, iy = np.minimum(hx.astype(np.int64), nn - 2), np.minimum(hy.astype(np.int64), nn - 2)
    ix, iy = np.where(inside, ix, 0), np.where(inside, iy, 0)
    fx, fy = hx - ix, hy - iy
    node = lambda i, j: np.rint(row[i] + col[j])  # noqa: E731  (z0 = 0, dz = 1: exact)
    want = (node(ix, iy) * (1 - fx) * (1 - fy) + node(ix, iy + 1) * (1 - fx) * fy +
            node(ix + 1, iy) * fx * (1 - fy) + node(ix + 1, iy + 1) * fx * fy)
    gz, gin = d_z[::stride].cpu().numpy(), d_in[::stride].cpu().numpy()
    edge = (np.abs(hx - np.rint(hx)) < 1e-6) | (np.abs(hy - np.rint(hy)) < 1e-6) | \
        (hx < 1e-6) | (hy < 1e-6) | (hx > nn - 1 - 1e-6) | (hy > nn - 1 - 1e-6)
    assert np.array_equal(gin[~edge], inside[~edge].astype(np.int32))
    both = inside & (gin == 1) & ~edge
    assert both.mean() > 0.9 and np.abs(gz[both] - want[both]).max() < 1e-6
    # the fused kernel returns the bits of the two-pass result
    z2 = d_z.clone()
    d_z.zero_()
    tb.api._check(lib.turtle_map_elevation_ecef_batch_device(
        mp.handle, n, ptr(d_ecef), ptr(d_lat), ptr(d_lon), ptr(d_alt), ptr(d_z), ptr(d_in), None))
    torch.cuda.synchronize()
    assert bool((d_z == z2).all())
