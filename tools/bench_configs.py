"""Measure the other BASELINE.json configurations (bench.py covers configs[1] = C2):

  c1  1 Mi straight rays from one station through a 2001 x 2001 UTM map + flat bottom
  c3  random rays through a layered stepper: Lambert map over a 3x3 tile stack over a
      flat geoid, local approximation on (range 10 m)
  c4  particles x 100 random-walk steps (turtle_stepper_step_batch with device states)
  c5  batched turtle_ecef_to_geodetic + turtle_map_elevation on a 20k x 20k map

    python tools/bench_configs.py --config c1 [--rays N] [--steps K]

Each run prints ONE JSON line: throughput on the GPU (CUDA events on the launching
stream, inputs resident in HBM), the reference's CPU path on all host cores over a
strided sample of the same inputs, and the parity of that sample. SURVEY.md section 8d
defines the workloads. Development / evidence tool: results are kept under profiles/.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench as B  # noqa: E402
import turtle_b200 as tb  # noqa: E402
from oracle import harness as H  # noqa: E402
from oracle import parity as P  # noqa: E402
from tests.common import Scene  # noqa: E402
from turtle_b200 import synth  # noqa: E402

RANK = int(os.environ.get("RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))
DEV = "cuda:%d" % LOCAL


def job_max(ms):
    """Whole-job time of a sharded measurement: the slowest rank (device time, max over ranks)."""
    if WORLD == 1:
        return ms, [ms]
    import torch.distributed as dist
    t = torch.tensor([ms], device=DEV, dtype=torch.float64)
    every = [torch.zeros_like(t) for _ in range(WORLD)]
    dist.all_gather(every, t)
    return max(float(x.item()) for x in every), [round(float(x.item()), 3) for x in every]


def emit(line):
    """ONE JSON line per job, from rank 0."""
    if RANK == 0:
        print(json.dumps(line), flush=True)


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if WORLD > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def trace_config(name, scene, pos, dirs, rule_o, rule_g, args, ops_per_sample, note, n_total=0):
    stepper, maps, stacks = scene.product()
    plan = stepper.freeze(LOCAL)
    plan.schedule_set(args.schedule)
    n_job = n_total or len(pos)
    if (WORLD > 1) and (len(pos) == n_job):
        # rays sharded with a stride (spreads the heavy tail), DEM replicated
        pos, dirs = np.ascontiguousarray(pos[RANK::WORLD]), np.ascontiguousarray(dirs[RANK::WORLD])
    n = len(pos)
    d_pos, d_dir = torch.from_numpy(pos).to(DEV), torch.from_numpy(dirs).to(DEV)
    d_res = torch.empty((n, 96), dtype=torch.uint8, device=DEV)
    if args.max_steps:  # profiling: bound the lone-lane tail of the launch
        rule_g = tb.trace_rule(rule_g.altitude_max, length_max=rule_g.length_max,
                               max_steps=args.max_steps)
        rule_o = H.rule(rule_o.altitude_max, length_max=rule_o.length_max,
                        max_steps=args.max_steps)
    ms_rank = timed(lambda: plan.trace_device(n, d_pos, d_dir, rule_g, d_res), args.steps, args.warmup)
    ms, per_rank = job_max(ms_rank)
    c = plan.counters(sync=True)
    piped = None
    if args.inflight > 1:
        # Steady state with several batches in flight on one plan: a launch ends when its
        # slowest ray does, and the SMs its persistent CTAs have left are taken by the next
        # batch (the CPU hides long rays behind each other the same way,
        # examples/example-pthread.c:66-114). One stream and one result array per batch in
        # flight; the device-pointer calls are asynchronous.
        k = args.inflight
        streams = [torch.cuda.Stream() for _ in range(k)]
        outs = [torch.empty((n, 96), dtype=torch.uint8, device=DEV) for _ in range(k)]
        batches = max(2 * k, args.steps * k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for b in range(batches):
            st = streams[b % k]
            plan.trace_device(n, d_pos, d_dir, rule_g, outs[b % k], stream=st.cuda_stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        same = all(bool((o == d_res).all()) for o in outs)
        piped = {"batches_in_flight": k, "batches": batches, "ms_per_batch": 1e3 * wall / batches,
                 "Mrays_per_s": n * batches / wall / 1e6, "results_identical": same,
                 "single_launch_ms": ms}
    if args.no_cpu:
        emit({"config": name, "n_gpus": WORLD, "rays": n_job, "ms_per_step": ms,
              "Mrays_per_s": n_job / ms / 1e3, "per_rank_ms": per_rank,
              "samples": c["samples"], "steps": c["steps"], "rebuilds": c["rebuilds"],
              "Gsamples_per_s": c["samples"] / ms_rank / 1e6, "pipelined": piped})
        return
    if RANK != 0:
        return
    got = d_res.cpu().numpy().view(tb.TRACE_RESULT).reshape(n)
    # reference on all cores, strided sample
    stride = max(1, n // args.cpu_rays)
    ora = scene.oracle(locked=True)
    want, steps, seconds = ora.trace(pos[::stride], dirs[::stride], rule_o,
                                     threads=os.cpu_count())
    rep = P.report(want, got[::stride])
    floor = None
    if H.available(H.REF_FMA) and H.best_oracle() == H.REF:
        fma, _, _ = scene.oracle(library=H.REF_FMA, locked=True).trace(
            pos[::stride], dirs[::stride], rule_o, threads=os.cpu_count())
        floor = P.report(want, fma)
    sys.stderr.write(P.table(rep, floor) + "\n")
    dfma = tb.dfma_peak(3)
    # executed FP64-pipe lane slots per sample: the committed ncu capture of this kernel on
    # this configuration (tools/kernel_profiles.py), refused when the kernel has changed
    lla = "_lla" if scene.range > 0 else ""
    prof, prof_src = B.kernel_profile("%s_trace%s_proj" % (name, lla))
    ops_per_sample = prof["fp64_lane_slots_per_sample"] if prof else None
    achieved = ops_per_sample * c["samples"] / (ms_rank * 1e-3) / 1e12 if prof else None
    line = {
        "config": name, "workload": note, "n_gpus": WORLD, "rays": n_job, "rays_per_gpu": n,
        "schedule": args.schedule, "ms_per_step": ms, "per_rank_ms": per_rank,
        "Mrays_per_s": n_job / ms / 1e3, "ns_per_step": ms_rank * 1e6 / max(c["steps"], 1),
        "steps_per_ray": c["steps"] / n, "samples_per_step": c["samples"] / max(c["steps"], 1),
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": dfma / 1e3,
                     "unit": "Tinst/s (FP64 pipe; FMA = 1)",
                     "frac": achieved / (dfma / 1e3) if prof else None,
                     "ops_per_sample": ops_per_sample, "ops_source": prof_src,
                     "active_threads_per_inst": prof["active_threads_per_inst"] if prof else None,
                     "issue_ceiling": prof["issue_ceiling"] if prof else None,
                     "dram_bytes_per_sample": prof["dram_bytes_per_sample"] if prof else None},
        "cpu_baseline": {"Mrays_per_s": len(want) / seconds / 1e6, "cores": os.cpu_count(),
                         "ns_per_step": 1e9 * seconds / max(steps, 1), "kind": "reference"
                         if H.best_oracle() == H.REF else "port",
                         "sample": "every %d-th ray (%d rays)" % (stride, len(want))},
        "rebuilds_per_sample": c.get("rebuilds", 0) / max(c["samples"], 1), "pipelined": piped,
        "parity": rep, "noise_floor": floor,
        "parity_vs_floor": P.against_floor(rep, floor) if floor else None,
        "status_counts": np.bincount(got["status"], minlength=5).tolist(),
        "plan_bytes": plan.bytes,
    }
    line["speedup_vs_cpu"] = line["Mrays_per_s"] / line["cpu_baseline"]["Mrays_per_s"]
    print(json.dumps(line), flush=True)


def c1_inputs(n, rg):
    """Scene and rays of configuration 1: 2001 x 2001 UTM map over a flat
    bottom, n-ray fan from the map centre + 1 m."""
    n_map = 2001
    x = (486000., 506000.)
    y = (5057000., 5077000.)
    vals = synth.fbm_grid(np.arange(n_map) / 3., np.arange(n_map) / 3.) * 3000.
    mp = dict(nx=n_map, ny=n_map, x=x, y=y, z=(0., 3000.), projection="UTM 31N", values=vals)
    scene = Scene(maps=[mp], ops=[(H.ADD_FLAT, 0, -100.), (H.ADD_LAYER, 0, 0.),
                                  (H.ADD_MAP, 0, 0.)], range=rg)
    ora = scene.oracle()
    lat, lon = ora.project("UTM 31N", [0.5 * (x[0] + x[1])], [0.5 * (y[0] + y[1])], inverse=True)
    origin, idx = ora.position(lat, lon, [1.0], 1)
    az, el = synth.golden_fan(n)
    dirs = synth.np_from_horizontal(np.full(n, lat[0]), np.full(n, lon[0]), az, el)
    return scene, np.repeat(origin, n, 0), dirs


def c1(args):
    n = args.rays or (1 << 20)
    scene, pos, dirs = c1_inputs(n, args.range)
    trace_config("c1", scene, pos, dirs, H.rule(3100.), tb.trace_rule(3100.), args, None,
                 "1 Mi-ray fan from the centre of a 2001x2001 UTM 31N map (10 m pitch) over a "
                 "flat bottom at -100 m, range %g, stop alt > 3100 m | 1e5 steps" % args.range)


def layered_scene(rg):
    B.make_stack()
    d0 = H.Driver(H.best_oracle())
    n_map = 2001
    cx, cy = d0.project("Lambert 93", [46.5], [3.5])
    half = 5. * (n_map - 1) / 2
    x = (cx[0] - half, cx[0] + half)
    y = (cy[0] - half, cy[0] + half)
    X, Y = np.meshgrid(np.linspace(x[0], x[1], n_map), np.linspace(y[0], y[1], n_map))
    la, lo = d0.project("Lambert 93", X.ravel(), Y.ravel(), inverse=True)
    vals = np.rint(synth.fbm_points((lo - B.STACK_LON0) * 3600., (la - B.STACK_LAT0) * 3600.)
                   * 3000.) + 0.
    mp = dict(nx=n_map, ny=n_map, x=x, y=y, z=(0., 6553.5), projection="Lambert 93", values=vals)
    return Scene(maps=[mp], stacks=[B.stack_dir()],
                 ops=[(H.ADD_FLAT, 0, 0.), (H.ADD_STACK, 0, 0.), (H.ADD_MAP, 0, 0.)], range=rg)


def c3_inputs(n, rg, rank=0, world=1):
    """Scene and rays of configuration 3: flat / 3 x 3 stack / Lambert map in one layer,
    n random rays (origins in the stack box + 0.1 deg, isotropic directions); with world > 1
    only the strided shard rank, rank + world, ... of the n rays is generated."""
    scene = layered_scene(rg)
    idx = np.arange(rank, n, world, dtype=np.uint64)
    lat = B.STACK_LAT0 - 0.1 + (B.STACK_N + 0.2) * synth.random_uniform(n, 0xC3, 0, idx)
    lon = B.STACK_LON0 - 0.1 + (B.STACK_N + 0.2) * synth.random_uniform(n, 0xC3, 1, idx)
    alt = -500. + 5500. * synth.random_uniform(n, 0xC3, 2, idx)
    return scene, synth.np_ecef_from_geodetic(lat, lon, alt), synth.random_unit(n, 0xC3, idx)


def c3(args):
    n = args.rays or (1 << 26)  # BASELINE.json configs[2]: 64 Mi rays
    scene, pos, dirs = c3_inputs(n, 10. if args.range is None else args.range, RANK, WORLD)
    trace_config("c3", scene, pos, dirs, H.rule(9000., length_max=1e5),
                 tb.trace_rule(9000., length_max=1e5), args, None,
                 "random rays (origins in the stack bbox + 0.1 deg, alt -500..5000 m, isotropic) "
                 "through flat(0) / 3x3 SRTMGL1 stack / 2001x2001 Lambert-93 map (5 m), range "
                 "%g, stop leaves data | alt > 9000 m | path > 100 km | 1e5 steps" % scene.range,
                 n_total=n)


def c4(args):
    scene = layered_scene(1. if args.range is None else args.range)
    stepper, maps, stacks = scene.product()
    plan = stepper.freeze(LOCAL)
    n_job = args.rays or (1 << 23)
    k = args.walk
    lat = (45.5 + 2. * synth.random_uniform(n_job, 0xC4, 0))[RANK::WORLD]
    lon = (2.5 + 2. * synth.random_uniform(n_job, 0xC4, 1))[RANK::WORLD]
    h = (-50. + 100. * synth.random_uniform(n_job, 0xC4, 2))[RANK::WORLD]
    n = len(lat)  # particles of this rank (strided shard)
    origin, idx = plan.position(lat, lon, h, 0)
    assert (idx >= 0).all()
    d_pos0 = torch.from_numpy(origin).to(DEV)
    gen = torch.Generator(device=DEV)
    gen.manual_seed(0xC4 + RANK)

    def directions():
        u = torch.rand((n, 2), generator=gen, device=DEV, dtype=torch.float64)
        cz = 2 * u[:, 0] - 1
        sz = torch.sqrt(torch.clamp(1 - cz * cz, min=0))
        ph = 2 * np.pi * u[:, 1]
        return torch.stack([sz * torch.cos(ph), sz * torch.sin(ph), cz], 1).contiguous()

    stride = max(1, n // args.cpu_rays)
    states = plan.states(n)
    kk = max(1, min(args.multi, k))  # steps per launch (1: turtle_stepper_step_batch)
    assert k % kk == 0
    d_step = torch.empty((kk, n), dtype=torch.float64, device=DEV)
    d_alt = torch.empty((kk, n), dtype=torch.float64, device=DEV)
    d_idx = torch.empty((kk, n, 2), dtype=torch.int32, device=DEV)
    total_ms, sample_dirs, got_step, got_alt, got_idx = 0., [], [], [], []
    rebuilds = 0
    for rep in range(2):  # pass 0 = warm-up, pass 1 = timed
        states.reset()
        d_pos = d_pos0.clone()
        gen.manual_seed(0xC4 + RANK)
        total_ms = 0.
        if WORLD > 1:
            torch.distributed.barrier()
        sample_dirs, got_step, got_alt, got_idx = [], [], [], []
        for j in range(k // kk):
            d_dir = torch.stack([directions() for _ in range(kk)], 0).contiguous()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if kk == 1:
                plan.step_device(n, d_pos, d_dir, states=states, altitude=d_alt, step=d_step,
                                 index=d_idx)
            else:
                plan.walk_device(n, kk, d_pos, d_dir, states=states, altitude=d_alt,
                                 step=d_step, index=d_idx)
            e1.record()
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
            rebuilds += plan.counters(sync=True)["rebuilds"] if rep == 1 else 0
            if rep == 1 and not args.no_cpu:
                for i in range(kk):
                    sample_dirs.append(d_dir[i, ::stride].cpu().numpy())
                    got_step.append(d_step[i, ::stride].cpu().numpy())
                    got_alt.append(d_alt[i, ::stride].cpu().numpy())
                    got_idx.append(d_idx[i, ::stride].cpu().numpy())
    ms_rank = total_ms
    total_ms, per_rank = job_max(ms_rank)
    if args.no_cpu:
        emit({"config": "c4", "n_gpus": WORLD, "particles": n_job, "walk_steps": k,
              "steps_per_launch": kk, "range": scene.range, "ms_total": total_ms, "per_rank_ms": per_rank,
              "Msteps_per_s": n_job * k / total_ms / 1e3, "rebuilds_per_step": rebuilds / (n * k)})
        return
    if RANK != 0:
        return
    ora = scene.oracle(locked=True)
    want = ora.walk(origin[::stride], np.stack(sample_dirs), threads=os.cpu_count())
    gs, ga, gi = np.stack(got_step), np.stack(got_alt), np.stack(got_idx)
    same_idx = (gi == want["index"]).all(2)
    ok = np.logical_and.accumulate(same_idx, 0)  # particle still on the oracle's track
    m = origin[::stride].shape[0]
    line = {
        "config": "c4", "workload": "%d particles x %d turtle_stepper_step, fresh isotropic "
        "direction each step, start within +-50 m of the ground; geometry of c3, range %g, "
        "per-particle stepper state on the device" % (n_job, k, scene.range),
        "n_gpus": WORLD, "particles": n_job, "walk_steps": k, "steps_per_launch": kk,
        "ms_total": total_ms, "per_rank_ms": per_rank,
        "Msteps_per_s": n_job * k / total_ms / 1e3, "ns_per_step": total_ms * 1e6 / (n_job * k),
        "state_bytes_per_particle": int(states.bytes_per_particle),
        "rebuilds_per_step": rebuilds / (n * k),
        "cpu_baseline": {"Msteps_per_s": m * k / want["seconds"] / 1e6, "cores": os.cpu_count(),
                         "ns_per_step": 1e9 * want["seconds"] / (m * k),
                         "sample": "every %d-th particle (%d particles)" % (stride, m)},
        "parity": {"particles": int(m), "off_track_particles": int((~ok[-1]).sum()),
                   "max_step_abs_diff_m": float(np.abs(gs - want["step"])[ok].max()),
                   "max_altitude_abs_diff_m": float(np.abs(ga - want["altitude"])[ok].max()),
                   "bit_identical_steps_frac": float((gs == want["step"])[ok].mean())},
    }
    line["speedup_vs_cpu"] = line["Msteps_per_s"] / line["cpu_baseline"]["Msteps_per_s"]
    print(json.dumps(line), flush=True)


def c5_map(n_map):
    """The geodetic map of configuration 5: n_map x n_map nodes over a 5.5 x 5.5 degree box,
    a cheap SEPARABLE synthetic terrain (node (ix, iy) = rint(row[ix] + col[iy]), z0 = 0,
    dz = 1) so that the host fill stays tractable and every node is known in closed form.
    -> (map, row, col, (lon0, lat0, box))"""
    import ctypes as C
    from turtle_b200._lib import lib
    box = 5.5
    lat0, lon0 = 44., 1.
    gx = np.arange(n_map, dtype=np.float64)
    row = 2500. + 1000. * np.sin(gx * 0.013) + 300. * np.sin(gx * 0.171)
    col = 400. * np.cos(gx * 0.007) + 100. * np.sin(gx * 0.31)
    mp = tb.Map(n_map, n_map, (lon0, lon0 + box), (lat0, lat0 + box), (0., 65535.), None)
    chunk = 1000
    for j0 in range(0, n_map, chunk):
        j1 = min(n_map, j0 + chunk)
        block = np.ascontiguousarray(np.rint(row[None, :] + col[j0:j1, None]))
        tb.api._check(lib.turtle_map_fill_rows(mp.handle, j0, j1 - j0,
                                               block.ctypes.data_as(C.c_void_p)))
    return mp, row, col, (lon0, lat0, box)


def c5(args):
    import ctypes as C
    from turtle_b200._lib import lib
    n_map = args.map_nodes
    t0 = time.time()
    mp, row, col, (lon0, lat0, box) = c5_map(n_map)
    fill_s = time.time() - t0
    lib.turtle_map_gather_set(mp.handle, args.gather)  # 1: cell-packed second copy
    n_job = args.rays or (1 << 30)
    n = n_job // WORLD  # a contiguous chunk of the points per rank
    gen = torch.Generator(device=DEV)
    gen.manual_seed(0xC5 + RANK)
    d_ecef = torch.empty((n, 3), dtype=torch.float64, device=DEV)
    piece = 1 << 26
    for i0 in range(0, n, piece):
        m = min(piece, n - i0)
        u = torch.rand((3, m), generator=gen, device=DEV, dtype=torch.float64)
        la = (lat0 - 0.05 + (box + 0.1) * u[0]).contiguous()
        lo = (lon0 - 0.05 + (box + 0.1) * u[1]).contiguous()
        al = (5000. * u[2]).contiguous()
        tb.api._check(lib.turtle_ecef_from_geodetic_batch_device(
            m, la.data_ptr(), lo.data_ptr(), al.data_ptr(), d_ecef[i0:].data_ptr(), None))
        torch.cuda.synchronize()
    d_lat = torch.empty(n, dtype=torch.float64, device=DEV)
    d_lon = torch.empty(n, dtype=torch.float64, device=DEV)
    d_alt = torch.empty(n, dtype=torch.float64, device=DEV)
    d_z = torch.zeros(n, dtype=torch.float64, device=DEV)
    d_in = torch.zeros(n, dtype=torch.int32, device=DEV)
    P = lambda t: t.data_ptr()  # noqa: E731
    ms_geo = timed(lambda: tb.api._check(lib.turtle_ecef_to_geodetic_batch_device(
        n, P(d_ecef), P(d_lat), P(d_lon), P(d_alt), None)), args.steps, args.warmup)
    ms_map = timed(lambda: tb.api._check(lib.turtle_map_elevation_batch_device(
        mp.handle, n, P(d_lon), P(d_lat), P(d_z), P(d_in), None)), args.steps, args.warmup)
    z_two = d_z[::4096].cpu().numpy().copy()
    ms_fused = timed(lambda: tb.api._check(lib.turtle_map_elevation_ecef_batch_device(
        mp.handle, n, P(d_ecef), P(d_lat), P(d_lon), P(d_alt), P(d_z), P(d_in), None)), args.steps,
        args.warmup)
    (ms_geo, _), (ms_map, _), (ms_fused, per_rank) = job_max(ms_geo), job_max(ms_map), job_max(ms_fused)
    if args.no_cpu:
        emit({"config": "c5", "gather": args.gather, "n_gpus": WORLD, "points": n_job, "ms_to_geodetic": ms_geo,
              "ms_map_elevation": ms_map, "ms_fused": ms_fused, "per_rank_ms_fused": per_rank,
              "Gpoints_per_s": {"to_geodetic": n_job / ms_geo / 1e6, "map_elevation": n_job / ms_map / 1e6,
                                "fused": n_job / ms_fused / 1e6}})
        return
    if RANK != 0:
        return
    # oracle on a strided sample
    stride = max(1, n // args.cpu_rays)
    ecef_s = d_ecef[::stride].cpu().numpy()
    ora = H.Driver(H.best_oracle())
    t0 = time.time()
    wla, wlo, wal = ora.ecef_to_geodetic(ecef_s)
    t_geo = time.time() - t0
    gla, glo, gal = d_lat[::stride].cpu().numpy(), d_lon[::stride].cpu().numpy(), \
        d_alt[::stride].cpu().numpy()
    gz, gin = d_z[::stride].cpu().numpy(), d_in[::stride].cpu().numpy()
    # elevation of the oracle's own lat / lon through the product's scalar host call
    # (bit-identical to the reference on maps, tests/test_oracle.py)
    hz = np.zeros(len(wla))
    hin = np.zeros(len(wla), dtype=np.int32)
    zz, ii = C.c_double(), C.c_int()
    for i in range(min(len(wla), 200000)):
        lib.turtle_map_elevation(mp.handle, wlo[i], wla[i], C.byref(zz), C.byref(ii))
        hz[i], hin[i] = zz.value, ii.value
    m = min(len(wla), 200000)
    both = (hin[:m] == 1) & (gin[:m] == 1)
    hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) \
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.
    dfma = tb.dfma_peak(3)

    def roof(ms, nbytes, profile_name):
        """Both roofs of a query kernel: algorithmic bytes per point against the measured HBM
        bandwidth, executed FP64-pipe lane slots per point (committed ncu capture, refused
        when stale) against the measured DFMA peak; plus the DRAM bytes actually moved."""
        out = {"GB_per_s": n * nbytes / ms / 1e6, "hbm_frac": n * nbytes / ms / 1e6 / hbm,
               "algorithmic_bytes_per_point": nbytes}
        prof, src = B.kernel_profile(profile_name)
        out["profile"] = src
        if prof is not None:
            ops = prof["fp64_lane_slots_per_sample"]
            out.update({"fp64_lane_slots_per_point": ops,
                        "fp64_Tinst_per_s": n * ops / ms / 1e9,
                        "fp64_frac": n * ops / ms / 1e9 / (dfma / 1e3),
                        "dram_bytes_per_point": prof["dram_bytes_per_sample"],
                        "bound": "fp64" if n * ops / ms / 1e9 / (dfma / 1e3) >
                        n * nbytes / ms / 1e6 / hbm else "hbm"})
        return out
    packed = "_packed" if args.gather == 1 else ""
    line = {
        "config": "c5", "workload": "%d ECEF points (uniform over the map box + 0.05 deg, alt "
        "0-5000 m) -> turtle_ecef_to_geodetic_batch, turtle_map_elevation_batch and the fused "
        "kernel on a %dx%d uint16 geodetic map (%.0f MB)" % (n_job, n_map, n_map, n_map * n_map * 2 / 1e6),
        "gather": args.gather, "n_gpus": WORLD, "points": n_job, "points_per_gpu": n,
        "map_fill_seconds": fill_s,
        "to_geodetic": dict(ms=ms_geo, Gpoints_per_s=n_job / ms_geo / 1e6,
                            **roof(ms_geo, 48, "c5_to_geodetic")),
        "map_elevation": dict(ms=ms_map, Gpoints_per_s=n_job / ms_map / 1e6,
                              **roof(ms_map, 36, "c5_map_elevation" + packed)),
        "fused": dict(ms=ms_fused, Gpoints_per_s=n_job / ms_fused / 1e6,
                      **roof(ms_fused, 68, "c5_map_elevation_ecef" + packed)),
        "peaks": {"hbm_GB_per_s": hbm, "fp64_Tinst_per_s": dfma / 1e3},
        "cpu_baseline": {"to_geodetic_Mpoints_per_s_1core": len(wla) / t_geo / 1e6,
                         "sample": "every %d-th point (%d points), 1 thread" % (stride, len(wla))},
        "parity": {"points": int(len(wla)), "altitude_bit_exact": bool(np.array_equal(wal, gal)),
                   "lat_max_ulp": int(np.abs(wla.view(np.int64) - gla.view(np.int64)).max()),
                   "lon_max_ulp": int(np.abs(wlo.view(np.int64) - glo.view(np.int64)).max()),
                   "inside_flips": int((hin[:m] != gin[:m]).sum()),
                   "elevation_max_abs_diff_m": float(np.abs(hz[:m] - gz[:m])[both].max()),
                   "fused_equals_two_pass": bool(np.array_equal(z_two, d_z[::4096].cpu().numpy()))},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["c1", "c3", "c4", "c5"])
    ap.add_argument("--rays", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--range", type=float, default=None)
    ap.add_argument("--walk", type=int, default=100)
    ap.add_argument("--multi", type=int, default=1,
                    help="c4: steps per launch (> 1: turtle_stepper_walk_batch)")
    ap.add_argument("--cpu-rays", type=int, default=1 << 18)
    ap.add_argument("--map-nodes", type=int, default=20000)
    ap.add_argument("--schedule", type=int, default=0)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--l2-fetch", type=int, default=0,
                    help="experiment: cudaLimitMaxL2FetchGranularity in bytes (32, 64, 128)")
    ap.add_argument("--gather", type=int, default=0,
                    help="c5: 1 = elevation queries gather from the cell-packed copy of the map")
    ap.add_argument("--inflight", type=int, default=0,
                    help="also measure the steady state with this many batches in flight")
    ap.add_argument("--max-steps", type=int, default=0,
                    help="profiling: stop rays after this many steps (bounds the launch tail)")
    ap.add_argument("--no-cpu", action="store_true", help="GPU part only (ncu captures)")
    args = ap.parse_args()
    if args.config == "c1" and args.range is None:
        args.range = 0.
    torch.cuda.set_device(LOCAL)
    if args.l2_fetch:
        # experiment: the granularity at which an L2 miss fetches from HBM (a device limit,
        # default 64 B on this GPU): random 2 x 2 gathers use 8 of the bytes they fetch
        from cuda import cudart
        torch.zeros(1, device=DEV)
        err, = cudart.cudaDeviceSetLimit(cudart.cudaLimit.cudaLimitMaxL2FetchGranularity, args.l2_fetch)
        err2, got = cudart.cudaDeviceGetLimit(cudart.cudaLimit.cudaLimitMaxL2FetchGranularity)
        sys.stderr.write("cudaLimitMaxL2FetchGranularity: set %d -> %s, now %s\n" % (args.l2_fetch, err, got))
    if WORLD > 1:  # one process per GPU (torchrun): --rays is the size of the WHOLE job
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/turtle_b200_nccl.%h.%p.log")
        torch.distributed.init_process_group("nccl", device_id=torch.device(DEV))
        if RANK == 0:
            B.make_stack()
        torch.distributed.barrier()
    {"c1": c1, "c3": c3, "c4": c4, "c5": c5}[args.config](args)
    if WORLD > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
