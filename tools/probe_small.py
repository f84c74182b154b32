"""Module-level latency of GAE on the reference's small test shape, column-scan vs T-split."""
import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from di_hpc_b200 import _abi
from hpc_rll.rl_utils.gae import GAE
T, B = 1024, 64
v = torch.randn(T + 1, B, device="cuda", requires_grad=True)
r = torch.randn(T, B, device="cuda", requires_grad=True)
g = torch.randn(T, B, device="cuda")
m = GAE(T, B)
for cfg in (2, 13, 20):
    _abi.set_config(0, cfg)
    for _ in range(5):
        o = m(v, r); torch.autograd.grad(o, [v, r], grad_outputs=g)
    torch.cuda.synchronize()
    tf = tb = 0.0
    for _ in range(50):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); o = m(v, r); e[1].record(); torch.autograd.grad(o, [v, r], grad_outputs=g); e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    print("cfg", cfg, "fwd %.1f us bwd %.1f us" % (tf / 50 * 1e3, tb / 50 * 1e3))
