"""Batch sharding + the single gather, on CPU with gloo (world_size 2 and 3): results are rank-count invariant."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ipoke_b200.parallel import global_noise, shard_bounds, sharded_sample


def test_shard_bounds_cover_batch():
    for n in (0, 1, 5, 64, 257):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_global_noise_is_rank_count_invariant():
    a = global_noise(8, 32, seed=42)
    torch.manual_seed(42)
    assert torch.equal(a, torch.randn(8, 32, 8, 8))          # CPU default-generator draw (second_stage_video.py:300)


def _fake_compute(z, cond, x0, length):
    # stand-in for the native sampler: deterministic per-sample function so the gathered order can be checked
    b = z.shape[0]
    base = z.flatten(1).sum(1) + cond.flatten(1).sum(1) * 2 + x0.flatten(1).sum(1) * 3
    return base.view(b, 1, 1, 1, 1).expand(b, length, 3, 4, 4).contiguous()


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = global_noise(B, 4, seed=7)
    g = torch.Generator().manual_seed(1)
    cond = torch.randn((B, 2, 8, 8), generator=g)
    x0 = torch.randn((B, 3, 4, 4), generator=g)
    out = sharded_sample(_fake_compute, z, cond, x0, 2)
    local = sharded_sample(_fake_compute, z, cond, x0, 2, gather=False)
    lo, hi = shard_bounds(B, world, rank)
    assert local.shape[0] == hi - lo
    if rank == 0:
        q.put(out)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,B", [(2, 6), (2, 5), (3, 7)])
def test_sharded_sample_gather_matches_single_process(world, B):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    z = global_noise(B, 4, seed=7)
    g = torch.Generator().manual_seed(1)
    cond = torch.randn((B, 2, 8, 8), generator=g)
    x0 = torch.randn((B, 3, 4, 4), generator=g)
    ref = _fake_compute(z, cond, x0, 2)
    assert torch.equal(out, ref)


# ---- training: the sharded optimizer step (reduce the flat gradient, update the local shard, all-gather the parameters) ----
def _train_worker(rank, world, port, n, q):
    from ipoke_b200.train import shard_range, sharded_update
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    padded, lo, hi = shard_range(n, world, rank)
    g0 = torch.Generator().manual_seed(3)
    params = torch.zeros(padded)
    params[:n] = torch.randn(n, generator=g0)                              # identical on every rank
    grads = torch.zeros(padded)
    grads[:n] = torch.randn(n, generator=torch.Generator().manual_seed(100 + rank))      # per-rank gradient of the local batch

    def sgd(p, g):                                                         # stand-in for ipk_adam_step: p -= lr * mean gradient
        p -= 0.1 * g / world

    sharded_update(params, grads, lo, hi, world, sgd)
    q.put((rank, params[:n].clone()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 10), (2, 7), (3, 11)])
def test_sharded_update_matches_single_process_mean_gradient(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get() for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    ref = torch.randn(n, generator=torch.Generator().manual_seed(3))
    mean_g = sum(torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)) / world
    ref = ref - 0.1 * mean_g
    for r in range(world):
        assert torch.allclose(got[r], ref, atol=1e-6), r                   # the global-batch update ...
        assert torch.equal(got[r], got[0])                                 # ... and bit-identical parameters on every rank
