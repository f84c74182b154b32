"""N>1 host logic on CPU: world_size-2 gloo processes exercise the column sharding and the
differentiable loss all-reduce (the CUDA kernels themselves are covered by -m gpu tests; the oracle
stands in for them here so the composition rule 'shards + global count + SUM == full batch' is checked)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from di_hpc_b200.sharding import all_reduce_losses, all_reduce_moments, shard_columns
        from oracle import oracle as orc
        T, B = 12, 37
        g = np.random.default_rng(7)
        value = g.standard_normal((T + 1, B), dtype=np.float32)
        reward = g.standard_normal((T, B), dtype=np.float32)
        b0, b1 = shard_columns(B, rank, world)
        # GAE: no collective, shards concatenate to the full result
        adv_local = orc.gae_forward(np.ascontiguousarray(value[:, b0:b1]), np.ascontiguousarray(reward[:, b0:b1]))
        parts = [None] * world
        dist.all_gather_object(parts, (b0, b1, adv_local))
        full = np.concatenate([p[2] for p in sorted(parts, key=lambda p: p[0])], axis=1)
        ok_gae = np.array_equal(full, orc.gae_forward(value, reward))
        # a mean-type loss: local sum / GLOBAL count, then differentiable all-reduce(SUM)
        x = torch.from_numpy(value[:-1, b0:b1].copy()).requires_grad_(True)
        local = (x * x).sum().reshape(1) / (T * B)
        (glob, ) = all_reduce_losses([local])
        glob.sum().backward()
        want = float((value[:-1] ** 2).mean())
        ok_loss = abs(float(glob.item()) - want) < 1e-6
        ok_grad = np.allclose(x.grad.numpy(), 2 * value[:-1, b0:b1] / (T * B), atol=1e-7)
        # advantage moments [sum, sum sq, count] of the shards all-reduce to the whole batch's statistics
        a64 = adv_local.astype(np.float64)
        mom = torch.tensor([a64.sum(), np.square(a64).sum(), float(a64.size)], dtype=torch.float64)
        all_reduce_moments(mom)
        s1, s2, n = mom.tolist()
        mean, sd = s1 / n, np.sqrt((s2 - s1 * s1 / n) / (n - 1))
        want_st = orc.adv_stats(orc.gae_forward(value, reward))
        ok_mom = n == T * B and abs(mean - want_st[0]) < 1e-6 and abs(sd + 1e-8 - want_st[1]) < 1e-6
        q.put((rank, ok_gae, ok_loss and ok_mom, ok_grad, (b0, b1)))
    finally:
        dist.destroy_process_group()


def test_shard_columns_cover_and_align():
    sys.path.insert(0, ROOT)
    from di_hpc_b200.sharding import shard_columns
    for B in (1, 5, 37, 64, 65536, 524288, 1000003):
        for W in (1, 2, 3, 4, 8):
            rs = [shard_columns(B, r, W) for r in range(W)]
            assert rs[0][0] == 0 and rs[-1][1] == B
            for (a0, a1), (c0, c1) in zip(rs, rs[1:]):
                assert a1 == c0 and a0 <= a1
            assert all(r[0] % 4 == 0 for r in rs if r[1] > r[0])  # non-empty shards start 16B-aligned


@pytest.mark.timeout(120)
def test_two_rank_gloo_composition():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, ok_gae, ok_loss, ok_grad, rng_ in res:
        assert ok_gae and ok_loss and ok_grad, (rank, ok_gae, ok_loss, ok_grad, rng_)
