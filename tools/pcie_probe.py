"""Development probe (run under gpurun): pinned host<->device copy bandwidth of the box,
one direction at a time and both at once -- the floor of the host-pointer (e2e) path."""
import json
import time

import torch


def main():
    n = 1 << 30
    h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}
    for name in ("h2d", "d2h", "both"):
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if name in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_a.copy_(h_a, non_blocking=True)
            if name in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_b.copy_(d_b, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        out[name + "_GBps_each"] = round(n / dt / 1e9, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
