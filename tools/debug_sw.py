import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import turtle_b200 as tb
from oracle import harness as H
from turtle_b200 import synth
from tests.test_scenes_extra import scenes, rays
d1 = '/tmp/dbg_sw'; d2 = '/tmp/dbg_mixed'
synth.write_hgt_stack(d1, -35, -72, 2, 2, n=1201)
synth.write_hgt_stack(d2, 45, 2, 1, 1, n=1201)
scene, box = scenes(d1, d2)["south_west"]
ora = scene.oracle(locked=True)
pos, dirs = rays(ora, *box, 20000, 5)
stepper, maps, stacks = scene.product()
plan = stepper.freeze(0)
# single query samples
q_o = ora.step(pos)
q_g = plan.step(pos.copy(), None)
for f in ("latitude", "longitude", "altitude", "step"):
    d = np.abs(q_o[f] - q_g[f]); print(f, "max abs diff", d.max(), "argmax", d.argmax())
print("index mismatch", (q_o["index"] != q_g["index"]).any(1).sum())
e = np.abs(q_o["elevation"] - q_g["elevation"]); e[np.abs(q_o["elevation"]) > 1e300] = 0
print("elevation max diff", e.max(), "rows with diff>1e-6:", (e.max(1) > 1e-6).sum())
bad = np.where(e.max(1) > 1e-6)[0][:10]
for i in bad:
    la, lo, al = q_o["latitude"][i], q_o["longitude"][i], q_o["altitude"][i]
    print(i, la, lo, al, q_o["elevation"][i], q_g["elevation"][i], q_o["index"][i], q_g["index"][i])
want, steps, _ = ora.trace(pos, dirs, H.rule(6000., length_max=5e4, max_steps=20000), threads=os.cpu_count())
got = plan.trace(pos, dirs, tb.trace_rule(6000., length_max=5e4, max_steps=20000))
m = np.where((want["n_steps"] != got["n_steps"]) | (want["status"] != got["status"]) | (want["medium_hash"] != got["medium_hash"]))[0]
print("mismatching rays", len(m))
for i in m[:12]:
    print(i, "steps", want["n_steps"][i], got["n_steps"][i], "status", want["status"][i], got["status"][i], "changes", want["n_changes"][i], got["n_changes"][i], "idx", want["index"][i], got["index"][i], "len", np.round(want["length"][i], 3), np.round(got["length"][i], 3))
