"""Why is the queued (back-to-back) GAE module loop slower than isolated calls?  Prints CPU time per iteration,
GPU time per iteration and allocator counters for a few loop shapes."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hpc_rll.rl_utils.gae import GAE  # noqa: E402

T, B = 1024, 65536
v = torch.randn(T + 1, B, device="cuda", requires_grad=True)
r = torch.randn(T, B, device="cuda", requires_grad=True)
g = torch.randn(T, B, device="cuda")
m = GAE(T, B)


def loop(name, body, iters=20):
    for _ in range(3):
        body()
    torch.cuda.synchronize()
    s0 = torch.cuda.memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        body()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    s1 = torch.cuda.memory_stats()
    print("%-28s cpu %.3f ms/iter  gpu %.3f ms/iter  cudaMalloc %d  retries %d" %
          (name, (t1 - t0) * 1e3 / iters, e0.elapsed_time(e1) / iters,
           s1["segment.all.allocated"] - s0["segment.all.allocated"],
           s1["num_alloc_retries"] - s0["num_alloc_retries"]))


def fwd_only():
    with torch.no_grad():
        m(v, r)


def fwd_bwd():
    torch.autograd.grad(m(v, r), [v, r], grad_outputs=g)


def fwd_bwd_keep():
    global KEEP
    KEEP = torch.autograd.grad(m(v, r), [v, r], grad_outputs=g)


loop("fwd only (no_grad)", fwd_only)
loop("fwd+bwd, results dropped", fwd_bwd)
loop("fwd+bwd, results kept", fwd_bwd_keep)
adv = torch.empty(T, B, device="cuda")
gv = torch.empty(T + 1, B, device="cuda")
gr = torch.empty(T, B, device="cuda")
from di_hpc_b200 import _abi  # noqa: E402
L = _abi.lib()
st = torch.cuda.current_stream().cuda_stream


def raw():
    L.hpc_rll_gae_forward(v.data_ptr(), r.data_ptr(), adv.data_ptr(), T, B, 0.99, 0.97, st)
    L.hpc_rll_gae_backward(g.data_ptr(), gv.data_ptr(), gr.data_ptr(), T, B, 0.99, 0.97, st)


loop("C ABI, preallocated", raw)
