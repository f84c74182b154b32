"""Write-only HBM bandwidth reference for the dense scatter kernels (backward of the n-step ops): cudaMemsetAsync
(`Tensor.zero_`) and a plain fill kernel on buffers of the sizes the backward passes write, next to the scatter itself."""
import json
import sys
import os
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timeit(fn, iters=30):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    out = []
    for mb in (128, 428, 1024, 2147):
        n = mb * 1000 * 1000 // 4
        x = torch.empty(n, device="cuda")
        t0 = timeit(lambda: x.zero_())
        t1 = timeit(lambda: x.fill_(1.5))
        y = torch.empty(n, device="cuda")
        t2 = timeit(lambda: y.copy_(x))
        out.append(dict(mb=mb, memset_ms=t0, memset_gbs=mb / t0, fill_ms=t1, fill_gbs=mb / t1, copy_ms=t2,
                        copy_gbs=2 * mb / t2))
        print(json.dumps(out[-1]))
    from di_hpc_b200 import _abi
    from hpc_rll.rl_utils.td import DistNStepTD
    B, N, A, T = 262144, 8, 51, 5
    d0 = torch.softmax(torch.randn(B, N, A, device="cuda"), -1).requires_grad_(True)
    d1 = torch.softmax(torch.randn(B, N, A, device="cuda"), -1)
    a = torch.randint(0, N, (B,), device="cuda")
    r = torch.randn(T, B, device="cuda")
    d = torch.zeros(B, device="cuda")
    m = DistNStepTD(T, B, N, A)
    loss = m(d0, d1, a, a, r, d, None, 0.99, -10.0, 10.0)[0]
    one = torch.ones_like(loss)
    t = timeit(lambda: torch.autograd.grad(loss, [d0], grad_outputs=one, retain_graph=True))
    print(json.dumps(dict(op="dist backward (scatter 428 MB)", ms=t, write_gbs=B * N * A * 4 / t / 1e6)))


if __name__ == "__main__":
    main()
