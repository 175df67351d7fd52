"""Development probe (run under gpurun): the host-pointer path of the bench workload,
streamed (one persistent kernel fed / drained by the copy engines) against chunked (one
kernel per chunk), next to the device-resident kernel time."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench as B  # noqa: E402
import turtle_b200 as tb  # noqa: E402
from turtle_b200 import synth  # noqa: E402


def main():
    B.make_stack()
    stack = tb.Stack(B.stack_dir())
    stepper = tb.Stepper(range=0., slope=0.4, resolution=1e-2)
    stepper.add_stack(stack, 0.)
    plan = stepper.freeze(0)
    print(json.dumps(dict(residency=plan.residency())), flush=True)
    rule = tb.trace_rule(B.ALTITUDE_MAX, max_steps=B.MAX_STEPS)
    n = B.N_AZ * B.N_EL
    dirs = synth.fan_directions(B.DET_LAT, B.DET_LON, B.N_AZ, B.N_EL, bundle=32)
    origin, _ = stepper.position(B.DET_LAT, B.DET_LON, B.DET_HEIGHT, 0)
    h_pos = torch.empty((n, 3), dtype=torch.float64, pin_memory=True)
    h_pos.numpy()[:] = origin
    h_dir = torch.empty((n, 3), dtype=torch.float64, pin_memory=True)
    h_dir.numpy()[:] = dirs
    h_res = torch.empty((n, 96), dtype=torch.uint8, pin_memory=True)
    res_np = h_res.numpy().view(tb.TRACE_RESULT).reshape(n)
    d_pos, d_dir = h_pos.cuda(), h_dir.cuda()
    d_res = torch.empty((n, 96), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        plan.trace_device(n, d_pos, d_dir, rule, d_res)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        plan.trace_device(n, d_pos, d_dir, rule, d_res)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps(dict(device_ms=round(e0.elapsed_time(e1) / 3, 2))), flush=True)
    for mode in (0, 1):
        plan.pipeline_set(mode)
        for _ in range(2):
            plan.trace(h_pos.numpy(), h_dir.numpy(), rule, results=res_np)
        t0 = time.perf_counter()
        for _ in range(3):
            plan.trace(h_pos.numpy(), h_dir.numpy(), rule, results=res_np)
        ms = (time.perf_counter() - t0) / 3 * 1e3
        same = bool((torch.from_numpy(h_res.numpy()) == d_res.cpu()).all())
        print(json.dumps(dict(pipeline=mode, host_ms=round(ms, 2), mrays=round(n / ms / 1e3, 1),
                              same=same, counters=plan.counters())), flush=True)


if __name__ == "__main__":
    main()
