"""World-size-2 gloo tests (CPU) of the data-parallel host logic: ray sharding and the flat gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from iron_b200.parallel import allreduce_gradients, collect_params
    torch.manual_seed(0)                              # replicated weights
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    extra = torch.nn.Linear(2, 2)                     # a module that gets no gradient on rank 1
    params = collect_params([net, extra])
    x = torch.full((4, 5), float(rank + 1))
    net(x).sum().backward()
    if rank == 0:
        extra(torch.ones(1, 2)).sum().backward()
    local = [None if p.grad is None else p.grad.clone() for p in params]
    n = allreduce_gradients(params, world)
    out[rank] = (n, [p.grad.clone() for p in params], local)
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_world2():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        (n0, g0, l0), (n1, g1, l1) = out[0], out[1]
    assert n0 == n1 == sum(t.numel() for t in g0)
    for a, b, la, lb in zip(g0, g1, l0, l1):
        assert torch.equal(a, b)                      # every rank ends with the same averaged gradient
        za = la if la is not None else torch.zeros_like(a)
        zb = lb if lb is not None else torch.zeros_like(a)
        assert torch.allclose(a, (za + zb) / 2, atol=1e-6)


def _flat_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from iron_b200.parallel import allreduce_flat
    # what GraphedStage2Step(flat_grads=True, grad_scale=1 / world) leaves behind: one bucket, pre-scaled, .grad = views
    local = torch.arange(10, dtype=torch.float32) * (rank + 1)
    flat = local / world
    views = [flat[0:4].view(2, 2), flat[4:10]]
    n = allreduce_flat(flat, world)
    out[rank] = (n, flat.clone(), [v.clone() for v in views])
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_world2():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_flat_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        (n0, f0, v0), (n1, f1, v1) = out[0], out[1]
    assert n0 == n1 == 10 and torch.equal(f0, f1)
    assert torch.allclose(f0, torch.arange(10, dtype=torch.float32) * 1.5)      # mean of x1 and x2
    assert torch.equal(v0[0], f0[0:4].view(2, 2)) and torch.equal(v0[1], f0[4:10])   # the views saw the in-place exchange


def test_shard_range_partitions_exactly():
    from iron_b200.parallel import crop_for_rank, shard_range
    for n in (0, 1, 7, 4096, 65537):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                a, b = shard_range(n, r, world)
                assert 0 <= a <= b <= n and (b - a) in (n // world, n // world + 1)
                cover += list(range(a, b))
            assert cover == list(range(n))
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
    corners = {crop_for_rank(r, 64) for r in range(8)}
    assert len(corners) == 8 and crop_for_rank(0, 64) == (224, 224)
    assert all(0 <= x <= 448 and 0 <= y <= 448 for x, y in corners)
    near = {crop_for_rank(r, 64, stride=8) for r in range(8)}                  # the bench's equal-work windows
    assert len(near) == 8 and all(abs(x - 224) <= 8 and abs(y - 224) <= 8 for x, y in near)


def test_bench_view_sharding_orbits_the_fixture_camera(monkeypatch):
    """bench.py's default weak-scaling inputs: rank r renders the centre crop of the fixture view orbited by r x 45 degrees
    about the y-axis -- same distance and elevation, rank 0 = the canonical view; 'windows' keeps one view."""
    import numpy as np
    import bench
    from oracle import iron_oracle as O
    monkeypatch.delenv("IRONB_BENCH_SHARD", raising=False)
    monkeypatch.delenv("IRONB_BENCH_CROP_STRIDE", raising=False)
    W0 = np.array(O.FIXTURE_W2C, dtype=np.float64).reshape(4, 4)
    assert np.array_equal(bench.view_w2c(0), W0)
    c0 = np.linalg.inv(W0)[:3, 3]
    centres = []
    for r in range(8):
        W = bench.view_w2c(r)
        c = np.linalg.inv(W)[:3, 3]
        centres.append(c)
        assert abs(np.linalg.norm(c) - np.linalg.norm(c0)) < 1e-9 and abs(c[1] - c0[1]) < 1e-9      # same distance, same height
        assert np.allclose(W[:3, :3] @ W[:3, :3].T, np.eye(3), atol=1e-12)                            # still a rotation
        assert bench.crop_corner(r, 64) == (224, 224)
    assert min(np.linalg.norm(centres[i] - centres[j]) for i in range(8) for j in range(i)) > 1.0     # eight distinct views
    monkeypatch.setenv("IRONB_BENCH_SHARD", "windows")
    assert np.array_equal(bench.view_w2c(3), W0)
    assert len({bench.crop_corner(r, 64) for r in range(8)}) == 8
