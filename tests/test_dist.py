"""World-size-2 check of the multi-GPU plumbing on CPU (gloo): shard the rays, trace each
shard independently (here with the C oracle standing in for a rank's GPU), gather the
records on rank 0 and compare with the single-process answer."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import harness as H
from tests.common import Scene, utm_map
from turtle_b200 import synth
from turtle_b200.dist import gather_records, shard_bounds


def _scene():
    return Scene(maps=[utm_map(n=101)], ops=[(H.ADD_FLAT, 0, -100.), (H.ADD_LAYER, 0, 0.),
                                             (H.ADD_MAP, 0, 0.)], range=0.)


def _rays(n):
    d = _scene().oracle(H.PORT)
    mp_ = _scene().maps[0]
    cx, cy = 0.5 * (mp_["x"][0] + mp_["x"][1]), 0.5 * (mp_["y"][0] + mp_["y"][1])
    lat, lon = d.project("UTM 31N", [cx], [cy], inverse=True)
    pos, _ = d.position(lat, lon, [1.0], 1)
    az, el = synth.golden_fan(n)
    dirs = d.ecef_from_horizontal(np.full(n, lat[0]), np.full(n, lon[0]), az, el)
    return np.repeat(pos, n, 0), dirs


def _worker(rank, size, port, n, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    pos, dirs = _rays(n)
    first, last = shard_bounds(n, rank, size)
    res, _, _ = _scene().oracle(H.PORT).trace(pos[first:last], dirs[first:last], H.rule(3100.))
    local = torch.from_numpy(res.view(np.uint8).reshape(-1, 96).copy())
    full = gather_records(local, dst=0)
    if rank == 0:
        np.save(out_path, full.numpy())
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    for n in (0, 1, 7, 64, 1000003):
        for size in (1, 2, 3, 8):
            b = [shard_bounds(n, r, size) for r in range(size)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(size - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_trace_gather_world2(tmp_path):
    n = 301  # ragged: 151 + 150
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, 29531 + os.getpid() % 1000, n, out), nprocs=2, join=True)
    got = np.load(out)
    pos, dirs = _rays(n)
    want, _, _ = _scene().oracle(H.PORT).trace(pos, dirs, H.rule(3100.))
    assert got.shape == (n, 96)
    assert got.tobytes() == want.tobytes()
