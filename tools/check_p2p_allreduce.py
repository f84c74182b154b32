"""torchrun script (N >= 2 GPUs of one box): the NVLink peer-memory scalar all-reduce (csrc/p2p.cu) against NCCL --
values, bit-identity across ranks, 2000 back-to-back calls (epoch parity / overwrite safety), CUDA-graph replay, latency.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/check_p2p_allreduce.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from di_hpc_b200.sharding import P2PScalarAllReduce, all_reduce_losses  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    comm = P2PScalarAllReduce()
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    ok = True
    # values vs NCCL, bit-identity across ranks, every n in 1..16
    for n in range(1, 17):
        x = torch.randn(n, device="cuda", generator=g)
        want = x.clone()
        dist.all_reduce(want)
        got = comm(x.clone())
        allr = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(allr, got)
        same = all(torch.equal(a, allr[0]) for a in allr)
        close = torch.allclose(got, want, rtol=1e-6, atol=1e-6)
        ok = ok and same and close
    # many calls back to back with rank-dependent skew (a sleeping rank must not lose words)
    acc_ok = True
    for i in range(2000):
        x = torch.full((3, ), float(rank + 1) * (i % 7 + 1), device="cuda")
        if i % 97 == rank:
            torch.cuda._sleep(2000000)
        y = comm(x)
        if i % 250 == 0:
            acc_ok = acc_ok and bool((y == (i % 7 + 1) * world * (world + 1) / 2).all().item())
    # differentiable wrapper
    l = torch.ones(1, device="cuda", requires_grad=True) * (rank + 1)
    (tot, ) = all_reduce_losses([l], comm=comm)
    tot.backward() if l.is_leaf else None
    wrap_ok = float(tot.item()) == world * (world + 1) / 2
    # CUDA graph: capture one call, replay with new data
    buf = torch.zeros(4, device="cuda")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        comm(buf)  # warm-up on the capture stream
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        comm(buf)
    graph_ok = True
    for i in range(5):
        buf.fill_(float(rank + i))
        torch.cuda.synchronize()
        dist.barrier()
        graph.replay()
        torch.cuda.synchronize()
        graph_ok = graph_ok and bool((buf == sum(r + i for r in range(world))).all().item())
    # latency: queued calls between events
    x = torch.ones(1, device="cuda")

    def timeit(fn, n=200):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n * 1e3

    p2p_us = timeit(lambda: comm(x))
    nccl_us = timeit(lambda: dist.all_reduce(x))
    res = torch.tensor([float(ok), float(acc_ok), float(wrap_ok), float(graph_ok), p2p_us, nccl_us], device="cuda")
    mins = res.clone()
    dist.all_reduce(mins, op=dist.ReduceOp.MIN)
    maxs = res.clone()
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"world": world, "values_and_bit_identity": bool(mins[0]), "2000_skewed_calls": bool(mins[1]),
                          "autograd_wrapper": bool(mins[2]), "graph_replay": bool(mins[3]),
                          "p2p_us_per_call_max_over_ranks": float(maxs[4]), "nccl_us_per_call_max_over_ranks": float(maxs[5])}),
              flush=True)
    comm.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
