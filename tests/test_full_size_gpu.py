"""-m gpu parity at BASELINE.json's FULL sizes (C2: V-trace/UPGO T=512, B=32768, N=16; C3: QR-DQN/IQN
B=1,048,576, tau=tau'=64, N=8, nstep=5) through size-independent properties: every column / sample of
these ops is independent except for the final mean, so a random SUBSET of columns (samples) pushed through
the oracle must reproduce the corresponding slice of the full-size CUDA result -- per-sample errors exactly,
gradients up to the known 1/count ratio -- and the scalar loss must equal the mean of the per-sample terms.
Inputs are generated on the device (seeded); only the subset and the small outputs travel to the host."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests._gpu import host, need_cuda, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def close(got, want, what, tol=TOL):
    e = rel_err(got, want)
    assert e <= tol, "%s: rel err %.3e" % (what, e)


def gen(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


def nstep_inputs(g, T, B, N):
    return dict(action=torch.randint(0, N, (B, ), device="cuda", generator=g),
                next_n_action=torch.randint(0, N, (B, ), device="cuda", generator=g),
                reward=torch.randn(T, B, device="cuda", generator=g),
                done=(torch.rand(B, device="cuda", generator=g) < 0.1).float(),
                weight=torch.rand(B, device="cuda", generator=g))


def test_qrdqn_c3_subset_vs_oracle():
    need_cuda()
    from hpc_rll.rl_utils.td import QRDQNNStepTDError
    tau, T, B, N, S = 64, 5, 1 << 20, 8, 384
    g = gen(31)
    inp = nstep_inputs(g, T, B, N)
    q = torch.randn(B, N, tau, device="cuda", generator=g).requires_grad_(True)
    nq = torch.randn(B, N, tau, device="cuda", generator=g)
    loss, td = QRDQNNStepTDError(tau, T, B, N)(q, nq, inp["action"], inp["next_n_action"], inp["reward"], inp["done"],
                                               0.99, inp["weight"], None)
    loss.sum().backward()
    idx = torch.randperm(B, device="cuda", generator=g)[:S].sort().values
    o = orc.qrdqn_nstep_td(host(q[idx]), host(nq[idx]), host(inp["action"][idx]), host(inp["next_n_action"][idx]),
                           host(inp["reward"][:, idx]), host(inp["done"][idx]), host(inp["weight"][idx]), None, 0.99, 1.0)
    close(host(td[idx]), o["td_error_per_sample"], "td_error_per_sample")
    close(host(q.grad[idx]) * (B / S), o["grad_q"], "grad_q")
    want = float((td.double() * inp["weight"].double()).mean().item())  # td.py:513-517: loss = mean(td * weight)
    assert abs(float(loss.item()) - want) <= 1e-6 * max(1.0, abs(want))
    # off-action slots receive exactly zero everywhere
    sel = torch.zeros(B, N, dtype=torch.bool, device="cuda")
    sel[torch.arange(B, device="cuda"), inp["action"]] = True
    assert int((q.grad != 0).any(dim=2).logical_and(~sel).sum().item()) == 0


def test_iqn_c3_subset_vs_oracle():
    need_cuda()
    from hpc_rll.rl_utils.td import IQNNStepTDError
    tau, tau_p, T, B, N, S = 64, 64, 5, 1 << 20, 8, 384
    g = gen(32)
    inp = nstep_inputs(g, T, B, N)
    q = torch.randn(tau, B, N, device="cuda", generator=g).requires_grad_(True)
    nq = torch.randn(tau_p, B, N, device="cuda", generator=g)
    rq = torch.rand(tau, B, device="cuda", generator=g)
    loss, td = IQNNStepTDError(tau, tau_p, T, B, N)(q, nq, inp["action"], inp["next_n_action"], inp["reward"],
                                                    inp["done"], rq, 0.99, 1.0, inp["weight"], None)
    loss.sum().backward()
    idx = torch.randperm(B, device="cuda", generator=g)[:S].sort().values
    o = orc.iqn_nstep_td(host(q[:, idx]), host(nq[:, idx]), host(inp["action"][idx]), host(inp["next_n_action"][idx]),
                         host(inp["reward"][:, idx]), host(inp["done"][idx]), host(rq[:, idx]), host(inp["weight"][idx]),
                         None, 0.99, 1.0, 1.0)
    close(host(td[idx]), o["td_error_per_sample"], "td_error_per_sample")
    close(host(q.grad[:, idx]) * (B / S), o["grad_q"], "grad_q")
    want = float((td.double() * inp["weight"].double()).mean().item())
    assert abs(float(loss.item()) - want) <= 1e-6 * max(1.0, abs(want))


def test_vtrace_upgo_c2_column_subset_vs_oracle():
    need_cuda()
    from hpc_rll.rl_utils.upgo import UPGO
    from hpc_rll.rl_utils.vtrace import VTrace
    T, B, N, S = 512, 32768, 16, 96
    g = gen(33)
    tgt = torch.randn(T, B, N, device="cuda", generator=g).requires_grad_(True)
    beh = torch.randn(T, B, N, device="cuda", generator=g)
    action = torch.randint(0, N, (T, B), device="cuda", generator=g)
    value = torch.randn(T + 1, B, device="cuda", generator=g).requires_grad_(True)
    reward = torch.randn(T, B, device="cuda", generator=g)
    weight = torch.rand(T, B, device="cuda", generator=g)
    cols = torch.randperm(B, device="cuda", generator=g)[:S].sort().values
    coef = [1.0, 0.5, -0.25]

    l = VTrace(T, B, N)(tgt, beh, action, value, reward, weight)
    (coef[0] * l.policy_loss + coef[1] * l.value_loss + coef[2] * l.entropy_loss).sum().backward()
    o = orc.vtrace(host(tgt[:, cols]), host(beh[:, cols]), host(action[:, cols]), host(value[:, cols]),
                   host(reward[:, cols]), host(weight[:, cols]), coef=coef)
    close(host(tgt.grad[:, cols]) * (B / S), o["grad_target_output"], "vtrace grad_target_output")
    close(host(value.grad[:, cols]) * (B / S), o["grad_value"], "vtrace grad_value")

    tgt.grad = None
    rhos = torch.rand(T, B, device="cuda", generator=g) * 2
    loss = UPGO(T, B, N)(tgt, rhos, action, reward, value.detach())
    loss.sum().backward()
    o = orc.upgo(host(tgt[:, cols]), host(rhos[:, cols]), host(action[:, cols]), host(reward[:, cols]),
                 host(value[:, cols]))
    close(host(tgt.grad[:, cols]) * (B / S), o["grad_target_output"], "upgo grad_target_output")


def test_vtrace_upgo_c2_full_losses_vs_oracle():
    """The whole C2 problem (T=512, B=32768, N=16: 16.8 M steps) through the C oracle: the three V-trace losses
    and the UPGO loss at full size -- the column-subset test above only sees gradients -- plus the complete
    gradient tensors.  ~15 s of host time (oracle is C + OpenMP)."""
    need_cuda()
    from hpc_rll.rl_utils.upgo import UPGO
    from hpc_rll.rl_utils.vtrace import VTrace
    T, B, N = 512, 32768, 16
    g = gen(35)
    tgt = torch.randn(T, B, N, device="cuda", generator=g).requires_grad_(True)
    beh = torch.randn(T, B, N, device="cuda", generator=g)
    action = torch.randint(0, N, (T, B), device="cuda", generator=g)
    value = torch.randn(T + 1, B, device="cuda", generator=g).requires_grad_(True)
    reward = torch.randn(T, B, device="cuda", generator=g)
    weight = torch.rand(T, B, device="cuda", generator=g)
    coef = [1.0, 0.5, -0.25]
    l = VTrace(T, B, N)(tgt, beh, action, value, reward, weight)
    (coef[0] * l.policy_loss + coef[1] * l.value_loss + coef[2] * l.entropy_loss).sum().backward()
    o = orc.vtrace(host(tgt), host(beh), host(action), host(value), host(reward), host(weight), coef=coef)
    for got, name in zip(l, ("policy_loss", "value_loss", "entropy_loss")):
        want = float(o[name])
        assert abs(float(got.item()) - want) <= TOL * max(1.0, abs(want)), (name, float(got.item()), want)
    n = float(T * B)  # gradients carry 1/(T*B): compare them at O(1) scale, or the norm-relative bar is vacuous
    close(host(tgt.grad * n), o["grad_target_output"] * np.float32(n), "vtrace grad_target_output (full)")
    close(host(value.grad * n), o["grad_value"] * np.float32(n), "vtrace grad_value (full)")
    del o
    tgt.grad = None
    rhos = torch.rand(T, B, device="cuda", generator=g) * 2
    loss = UPGO(T, B, N)(tgt, rhos, action, reward, value.detach())
    loss.sum().backward()
    o = orc.upgo(host(tgt), host(rhos), host(action), host(reward), host(value))
    want = float(o["loss"])
    assert abs(float(loss.item()) - want) <= TOL * max(1.0, abs(want)), (float(loss.item()), want)
    close(host(tgt.grad * n), o["grad_target_output"] * np.float32(n), "upgo grad_target_output (full)")
