"""North-star item 2, measured: how the trace kernel of the bench fan (C2) fetches the four
nodes of a sample -- four 16-bit loads (default), ONE 8-byte load from a cell-packed copy of
the tiles, or a shared-memory window of the station's surroundings staged by the bulk copy
engine (cp.async.bulk + mbarrier). Prints one JSON line per mode: kernel time (CUDA events,
16 Mi-ray fan, rays resident), window hit fraction, byte-identity of the records.

    python tools/gather_modes.py [--rays N] [--steps K] [--mode M]   (--mode: one mode only,
                                                                      for an ncu capture)
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench as B  # noqa: E402
import turtle_b200 as tb  # noqa: E402

NAMES = {0: "global 16-bit loads", 1: "cell-packed tiles, one 8-byte load",
         2: "shared-memory window (cp.async.bulk) + global"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=B.N_AZ * B.N_EL)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--mode", type=int, default=-1)
    args = ap.parse_args()
    B.make_stack()
    stepper = tb.Stepper(range=0., slope=0.4, resolution=1e-2)
    stepper.add_stack(tb.Stack(B.stack_dir()), 0.)
    plan = stepper.freeze(0)
    rule = tb.trace_rule(B.ALTITUDE_MAX, max_steps=B.MAX_STEPS)
    n = args.rays
    total = B.N_AZ * B.N_EL
    _, _, dirs = B.fan(0, 1, 0, n) if n == total else (0, 0, B.fan_subsample(0, 1, total // n, total)[:n])
    origin, _ = stepper.position(B.DET_LAT, B.DET_LON, B.DET_HEIGHT, 0)
    d_pos = torch.from_numpy(np.repeat(origin[None], n, 0)).cuda()
    d_dir = torch.from_numpy(dirs).cuda()
    d_res = torch.empty((n, 96), dtype=torch.uint8, device="cuda")
    reference = None
    for mode in ([args.mode] if args.mode >= 0 else [0, 1, 2]):
        plan.gather_set(mode, B.DET_LAT, B.DET_LON)
        for _ in range(3):
            plan.trace_device(n, d_pos, d_dir, rule, d_res)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            plan.trace_device(n, d_pos, d_dir, rule, d_res)
        e1.record()
        torch.cuda.synchronize()
        c = plan.counters(sync=True)
        got = d_res.cpu()
        if reference is None:
            reference = got
        print(json.dumps({"gather": mode, "how": NAMES[mode], "rays": n,
                          "kernel_ms": e0.elapsed_time(e1) / args.steps,
                          "samples": c["samples"], "window_hit_fraction": c["window_hits"] / c["samples"],
                          "plan_bytes": plan.bytes,
                          "records_identical_to_first_mode": bool((got == reference).all())}),
              flush=True)


if __name__ == "__main__":
    main()
