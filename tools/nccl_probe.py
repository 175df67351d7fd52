"""Development probe: how fast does the result gather go between ranks on this box?"""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 16777216
x = torch.full((n, 96), rank, dtype=torch.uint8, device=dev)
out = [torch.empty_like(x) for _ in range(world)] if rank == 0 else None
big = torch.empty((world, n, 96), dtype=torch.uint8, device=dev)
def t(fn, name, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print("%-28s %.2f ms  (%.1f GB/s per sender)" % (name, e0.elapsed_time(e1) / reps, x.numel() / (e0.elapsed_time(e1) / reps) / 1e6), flush=True)
t(lambda: dist.gather(x, out, dst=0), "gather(list)")
t(lambda: dist.all_gather_into_tensor(big.view(-1), x.view(-1)), "all_gather_into_tensor")
def sr():
    if rank == 0:
        for r in range(1, world): dist.recv(big[r], src=r)
    else:
        dist.send(x, dst=0)
t(sr, "send/recv")
dist.destroy_process_group()
