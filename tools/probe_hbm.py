"""HBM ceilings for different read:write mixes (torch ops, CUDA events): context for the roofline fractions."""
import json

import torch


def t(fn, it=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it


n = 1024 * 65536
x = torch.randn(n, device="cuda")
y = torch.empty_like(x)
z = torch.empty_like(x)
out = {}
out["copy_1r1w_gbs"] = 8 * n / t(lambda: y.copy_(x)) / 1e6
out["fill_0r1w_gbs"] = 4 * n / t(lambda: y.zero_()) / 1e6
out["sum_1r0w_gbs"] = 4 * n / t(lambda: x.sum()) / 1e6
out["add_2r1w_gbs"] = 12 * n / t(lambda: torch.add(x, y, out=z)) / 1e6
# 1 read : 2 writes, the mix of the GAE backward (grad_adv in; grad_value, grad_reward out): frexp is one
# TensorIterator kernel with two outputs (fp32 mantissa + int32 exponent)
mant, expo = torch.empty_like(x), torch.empty(n, dtype=torch.int32, device="cuda")
out["frexp_1r2w_gbs"] = 12 * n / t(lambda: torch.frexp(x, out=(mant, expo))) / 1e6
print(json.dumps(out))
