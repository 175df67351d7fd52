"""Development sweep (run under gpurun): launch geometry / register budget of the trace
kernel on the bench workload, device path and host-pointer path."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import turtle_b200 as tb
import bench as B

def main():
    order = os.environ.get("SWEEP_ORDER", "el")
    B.make_stack()
    stack = tb.Stack(B.stack_dir())
    stepper = tb.Stepper(range=0., slope=0.4, resolution=1e-2)
    stepper.add_stack(stack, 0.)
    plan = stepper.freeze(0)
    rule = tb.trace_rule(B.ALTITUDE_MAX, max_steps=B.MAX_STEPS)
    n = B.N_AZ * B.N_EL
    from turtle_b200 import synth
    bundle = {"el": 1, "band": 32, "band8": 8, "band128": 128}.get(order, 1)
    lat, lon = B.DET_LAT, B.DET_LON
    dirs = synth.fan_directions(lat, lon, B.N_AZ, B.N_EL, bundle=bundle)
    if order == "az":  # azimuth-major
        dirs = dirs.reshape(B.N_EL, B.N_AZ, 3).transpose(1, 0, 2).reshape(n, 3).copy()
    origin, _ = stepper.position(lat, lon, B.DET_HEIGHT, 0)
    h_pos = torch.empty((n, 3), dtype=torch.float64, pin_memory=True); h_pos.numpy()[:] = origin
    h_dir = torch.empty((n, 3), dtype=torch.float64, pin_memory=True); h_dir.numpy()[:] = dirs
    h_res = torch.empty((n, 96), dtype=torch.uint8, pin_memory=True)
    d_pos, d_dir = h_pos.cuda(), h_dir.cuda()
    d_res = torch.empty((n, 96), dtype=torch.uint8, device="cuda")
    res_np = h_res.numpy().view(tb.TRACE_RESULT).reshape(n)
    configs = [(6, 128, 0, 0), (6, 128, 0, 1), (8, 128, 0, 1)]
    if os.environ.get("SWEEP_CONFIGS"):
        configs = [tuple(int(v) for v in c.split(",")) for c in os.environ["SWEEP_CONFIGS"].split(";")]
    for (c, t, sched, special) in configs:
        plan.launch_set(c, t)
        plan.schedule_set(sched)
        plan.specialise_set(special)
        for _ in range(2):
            plan.trace_device(n, d_pos, d_dir, rule, d_res)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            plan.trace_device(n, d_pos, d_dir, rule, d_res)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        if os.environ.get("SWEEP_SKIP_HOST"):
            print(json.dumps(dict(order=order, special=special, ctas_per_sm=c, device_ms=round(ms, 2))), flush=True)
            continue
        plan.trace(h_pos.numpy(), h_dir.numpy(), rule, results=res_np)
        t0 = time.perf_counter()
        for _ in range(2):
            plan.trace(h_pos.numpy(), h_dir.numpy(), rule, results=res_np)
        host_ms = (time.perf_counter() - t0) / 2 * 1e3
        c2 = plan.counters()
        print(json.dumps(dict(order=order, schedule=sched, special=special, ctas_per_sm=c, threads=t, device_ms=round(ms, 2), mrays=round(n / ms / 1e3, 2),
                              host_ms=round(host_ms, 2), host_mrays=round(n / host_ms / 1e3, 2),
                              host_kernel_ms_sum=round(c2["kernel_ms"], 1))), flush=True)

if __name__ == "__main__":
    main()
