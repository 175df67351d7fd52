"""Development probe (run under gpurun): first contact of the CUDA path with the oracle.
Prints ulp statistics of the device transforms vs the oracle, trace parity on a
small layered geometry, and a first timing. Not part of the product or the tests."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import turtle_b200 as tb
from turtle_b200 import synth
from oracle import harness as H

def ulps(a, b):
    ia = a.view(np.int64); ib = b.view(np.int64)
    return np.abs(ia - ib)

def main():
    import functools; global print; print = functools.partial(print, flush=True)
    print("devices", tb.device_count())
    print("dfma peak Gop/s", tb.dfma_peak(3))
    lib = H.best_oracle(); print("oracle", lib)
    ora = H.Driver(lib)
    rng = np.random.default_rng(1)
    n = 1000000
    lat = rng.uniform(-90, 90, n); lon = rng.uniform(-180, 180, n); alt = rng.uniform(-500, 9000, n)
    e_o = ora.ecef_from_geodetic(lat, lon, alt)
    e_g = tb.ecef_from_geodetic_batch(lat, lon, alt)
    u = ulps(e_o, e_g); print("from_geodetic ulp max", u.max(), "mismatch frac", (u > 0).mean())
    g_o = ora.ecef_to_geodetic(e_o); g_g = tb.ecef_to_geodetic_batch(e_o)
    for k, nm in enumerate(("lat", "lon", "alt")):
        u = ulps(g_o[k], g_g[k]); print("to_geodetic", nm, "ulp max", u.max(), "mismatch frac", (u > 0).mean())
    # layered geometry on a 1201 stack
    d0 = '/tmp/stack1201'
    synth.write_hgt_stack(d0, 45, 2, 2, 2, n=1201, skip=((46, 3),))
    st = ora.stack_create(d0, locked=True)
    nx = ny = 401
    cx, cy = ora.project("Lambert 93", [45.6], [2.7])
    x = (cx[0] - 5000., cx[0] + 5000.); y = (cy[0] - 5000., cy[0] + 5000.)
    X, Y = np.meshgrid(np.linspace(x[0], x[1], nx), np.linspace(y[0], y[1], ny))
    la, lo = ora.project("Lambert 93", X.ravel(), Y.ravel(), inverse=True)
    vals = np.rint(synth.fbm_points((lo - 2) * 1200, (la - 45) * 1200) * 3000.) + 0.0
    m = ora.map_create(nx, ny, x, y, (0., 6553.5), "Lambert 93", vals)
    ops = [(H.ADD_FLAT, 0, 0.), (H.ADD_STACK, st, 0.), (H.ADD_MAP, m, 0.), (H.ADD_LAYER, 0, 0),
           (H.ADD_STACK, st, 500.), (H.ADD_MAP, m, 600.)]
    gmap = tb.Map(nx, ny, x, y, (0., 6553.5), "Lambert 93", vals.reshape(ny, nx))
    gstack = tb.Stack(d0)
    n = 20000
    olat = rng.uniform(44.9, 47.1, n); olon = rng.uniform(1.9, 4.1, n); oalt = rng.uniform(-500, 5000, n)
    pos = ora.ecef_from_geodetic(olat, olon, oalt)
    dirs = synth.random_unit(n, 7)
    for rg in (0., 10.):
        ora.geometry(ops, range=rg)
        ro, steps, sec = ora.trace(pos, dirs, H.rule(9000., length_max=1e5), threads=os.cpu_count())
        s = tb.Stepper(range=rg)
        s.add_flat(0.); s.add_stack(gstack, 0.); s.add_map(gmap, 0.); s.add_layer(); s.add_stack(gstack, 500.); s.add_map(gmap, 600.)
        plan = s.freeze(0)
        t = time.time(); rgp = plan.trace(pos, dirs, tb.trace_rule(9000., length_max=1e5)); dt = time.time() - t
        c = plan.counters()
        print("range", rg, "oracle steps", steps, "gpu", c, "wall %.3f" % dt)
        for f in ("n_steps", "status", "index", "medium_hash", "n_changes"):
            print("   ", f, "mismatch", int((ro[f] != rgp[f]).reshape(n, -1).any(1).sum()))
        same = (ro["n_steps"] == rgp["n_steps"]) & (ro["medium_hash"] == rgp["medium_hash"])
        dp = np.abs(ro["position"] - rgp["position"]).max(1)
        print("    pos diff max (same-steps rays) %.3e m, all %.3e" % (dp[same].max(), dp.max()),
              "len diff max %.3e" % np.abs(ro["length"] - rgp["length"])[same].max(),
              "bit-identical rays", int((ro.tobytes() == rgp.tobytes())), int(sum(ro[i].tobytes() == rgp[i].tobytes() for i in range(n))))
    # throughput on this small geometry with many rays
    n = 2000000
    olat = rng.uniform(45.0, 46.0, n); olon = rng.uniform(2.0, 3.0, n); oalt = rng.uniform(0, 4000, n)
    pos = synth.np_ecef_from_geodetic(olat, olon, oalt); dirs = synth.random_unit(n, 9)
    s = tb.Stepper(range=0.); s.add_stack(gstack, 0.); plan = s.freeze(0)
    for it in range(3):
        t = time.time(); r = plan.trace(pos, dirs, tb.trace_rule(9000., length_max=1e5)); dt = time.time() - t
        c = plan.counters()
        print("stack-only 2M rays: wall %.3f s kernel_ms %.1f steps %d samples %d -> %.2f Mrays/s, %.3f ns/step (kernel)" % (
            dt, c["kernel_ms"], c["steps"], c["samples"], n / dt / 1e6, c["kernel_ms"] * 1e6 / max(c["steps"], 1)))

if __name__ == "__main__":
    main()
