"""Development probe: what do the longest rays of config c3 look like?"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, turtle_b200 as tb
import bench as B
from tools.bench_configs import layered_scene
from turtle_b200 import synth
scene = layered_scene(0.)
stepper, maps, stacks = scene.product()
plan = stepper.freeze(0)
n = 1 << 22
lat = B.STACK_LAT0 - 0.1 + (B.STACK_N + 0.2) * synth.random_uniform(n, 0xC3, 0)
lon = B.STACK_LON0 - 0.1 + (B.STACK_N + 0.2) * synth.random_uniform(n, 0xC3, 1)
alt = -500. + 5500. * synth.random_uniform(n, 0xC3, 2)
pos = synth.np_ecef_from_geodetic(lat, lon, alt); dirs = synth.random_unit(n, 0xC3)
r = plan.trace(pos, dirs, tb.trace_rule(9000., length_max=1e5))
o = np.argsort(-r["n_steps"])[:25]
print("steps percentiles", np.percentile(r["n_steps"], [50, 90, 99, 99.9, 99.99, 100]))
for i in o:
    print(i, "steps", r["n_steps"][i], "changes", r["n_changes"][i], "status", r["status"][i], "idx", r["index"][i], "len", np.round(r["length"][i], 1), "alt0 %.1f" % alt[i], "final alt %.3f" % r["altitude"][i])
