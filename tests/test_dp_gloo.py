"""world_size-2 gloo tests (CPU) of the data-parallel plumbing used by bench.py --gpus N."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

from quantized_training import dp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        units = dp.shard_indices(11, world, rank)
        # slowest rank defines the time; units add up
        t = dp.reduce_max(1.0 + rank)
        (n_units, nll) = dp.reduce_sum([len(units), sum(0.5 * u for u in units)])
        torch.manual_seed(0)
        model = nn.Linear(4, 3)
        model.weight.requires_grad_(False)                      # frozen weight: its grad must not be exchanged
        x = torch.full((2, 4), float(rank + 1))
        model(x).sum().backward()
        nb = dp.allreduce_grads_(model.parameters())
        # overlapped reducer: buckets fire from grad hooks during backward, finish() before the optimizer step
        torch.manual_seed(1)
        net = nn.Sequential(nn.Linear(4, 8), nn.ReLU(), nn.Linear(8, 8), nn.ReLU(), nn.Linear(8, 2))
        net[0].weight.requires_grad_(False)
        red = dp.GradReducer(net.parameters(), bucket_bytes=64)      # tiny buckets: several per backward
        sums = []
        for step in range(2):
            red.zero_grad()
            xin = torch.full((3, 4), float(rank + 1 + step))
            net(xin).square().sum().backward()
            nbk = red.finish()
            sums.append([float(p.grad.sum()) for p in net.parameters() if p.requires_grad])
        # reference: the same gradients computed locally for both ranks' inputs, averaged
        want = []
        for step in range(2):
            acc = None
            for r in range(world):
                net.zero_grad(set_to_none=True)
                net(torch.full((3, 4), float(r + 1 + step))).square().sum().backward()
                g = [float(p.grad.sum()) for p in net.parameters() if p.requires_grad]
                acc = g if acc is None else [a + b for a, b in zip(acc, g)]
            want.append([a / world for a in acc])
        out[rank] = dict(units=units, t=t, n_units=n_units, nll=nll, bias_grad=model.bias.grad.tolist(), buckets=nb,
                         weight_grad=model.weight.grad, reducer_sums=sums, reducer_want=want, reducer_buckets=nbk)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert sorted(r0["units"] + r1["units"]) == list(range(11)) and not set(r0["units"]) & set(r1["units"])
    assert abs(len(r0["units"]) - len(r1["units"])) <= 1
    assert r0["t"] == r1["t"] == 2.0
    assert r0["n_units"] == r1["n_units"] == 11 and r0["nll"] == r1["nll"] == 0.5 * sum(range(11))
    assert r0["bias_grad"] == r1["bias_grad"] == [2.0, 2.0, 2.0]   # each rank's bias grad is 2 (two rows); mean is 2
    assert r0["buckets"] == 1 and r0["weight_grad"] is None
    assert r0["reducer_buckets"] > 1
    for r in (r0, r1):
        for got, want in zip(r["reducer_sums"], r["reducer_want"]):
            assert len(got) == 5 and all(abs(a - b) <= 1e-4 * max(1.0, abs(b)) for a, b in zip(got, want)), (got, want)


def test_single_process_is_a_no_op():
    assert dp.reduce_max(3.5) == 3.5 and dp.reduce_sum([1, 2]) == [1.0, 2.0]
    assert dp.shard_indices(5, 1, 0) == [0, 1, 2, 3, 4]


def test_sliding_windows_match_the_reference_recipe():
    w = dp.windows_for(2500, 1024, 512)
    assert w[0] == (0, 1024, 1024) and w[1] == (512, 1536, 512) and w[-1][1] == 2500
    assert sum(t for _, _, t in w) == 2500                       # every token is scored exactly once
    assert dp.windows_for(1000, 1024, 512) == [(0, 1000, 1000)]
