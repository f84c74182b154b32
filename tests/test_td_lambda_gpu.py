"""TD(lambda) parity on the GPU vs the oracle and the origin-generated golden fixtures.
ret / grad are element-wise chains in the oracle's fp32 order; the loss is an fp64-accumulated sum:
tolerance 1e-5 norm-relative everywhere (north_star), observed ~1e-7."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests._golden import Case, grad_err, names, rel_err
from tests._gpu import dev, host, need_cuda, rng

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _run(value, reward, weight, gamma, lam, coef):
    from hpc_rll.rl_utils.td import TDLambda
    v = dev(value).requires_grad_(True)
    r = dev(reward)
    w = None if weight is None else dev(weight)
    loss = TDLambda(*reward.shape)(v, r, w, gamma, lam)
    assert loss.shape == (1, )
    (loss * coef).sum().backward()
    torch.cuda.synchronize()
    return float(loss.item()), host(v.grad)


@pytest.mark.parametrize("T,B,use_w", [(1024, 64, True), (128, 128, False), (1, 3, True), (37, 5, True),
                                         (100, 260, True), (33, 4100, False), (16, 33000, True), (7, 1, False)])
def test_td_lambda_vs_oracle(T, B, use_w):
    need_cuda()
    g = rng(T * 7919 + B)
    value = g.standard_normal((T + 1, B), dtype=np.float32)
    reward = g.standard_normal((T, B), dtype=np.float32)
    weight = g.random((T, B), dtype=np.float32) if use_w else None
    loss, gv = _run(value, reward, weight, 0.9, 0.8, 1.7)
    o = orc.td_lambda(value, reward, weight, 0.9, 0.8, 1.7)
    o64 = orc.td_lambda(value.astype(np.float64), reward.astype(np.float64),
                        None if weight is None else weight.astype(np.float64), 0.9, 0.8, 1.7)
    assert rel_err(loss, o["loss"]) <= TOL and rel_err(loss, o64["loss"]) <= TOL
    assert grad_err(gv, o["grad_value"]) <= TOL
    assert np.all(gv[-1] == 0)


@pytest.mark.parametrize("name", names("tdlambda"))
def test_td_lambda_vs_golden(name):
    need_cuda()
    c = Case(name)
    loss, gv = _run(c.inp("value"), c.inp("reward"), c.inp("weight"), c.attr("gamma"), c.attr("lambda_"),
                    c.attr("coef_loss"))
    assert rel_err(loss, c.out("loss", 32)) <= TOL and rel_err(loss, c.out("loss", 64)) <= TOL
    assert grad_err(gv, c.grad("value", 32)) <= TOL and grad_err(gv, c.grad("value", 64)) <= TOL


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 21, 99])  # 21 = single-launch T-split with look-back
def test_td_lambda_configs_agree(cfg):
    need_cuda()
    from di_hpc_b200 import _abi
    g = rng(cfg)
    T, B = 77, 1300
    value = g.standard_normal((T + 1, B), dtype=np.float32)
    reward = g.standard_normal((T, B), dtype=np.float32)
    weight = g.random((T, B), dtype=np.float32)
    try:
        _abi.set_config(_abi.OP_TD_LAMBDA, cfg)
        loss, gv = _run(value, reward, weight, 0.99, 0.95, 1.0)
        loss2, gv2 = _run(value, reward, None, 0.99, 0.95, 1.0)
    finally:
        _abi.set_config(_abi.OP_TD_LAMBDA, -1)
    o = orc.td_lambda(value, reward, weight, 0.99, 0.95, 1.0)
    o2 = orc.td_lambda(value, reward, None, 0.99, 0.95, 1.0)
    assert rel_err(loss, o["loss"]) <= TOL and grad_err(gv, o["grad_value"]) <= TOL
    assert rel_err(loss2, o2["loss"]) <= TOL and grad_err(gv2, o2["grad_value"]) <= TOL


def test_td_lambda_deterministic_and_sharded():
    """Run-to-run bit reproducibility (fixed-order reduction) and shard composition via global_B."""
    need_cuda()
    from hpc_rll.rl_utils.td import TDLambda
    g = rng(3)
    T, B = 64, 2048
    value = dev(g.standard_normal((T + 1, B), dtype=np.float32))
    reward = dev(g.standard_normal((T, B), dtype=np.float32))
    m = TDLambda(T, B)
    a = m(value, reward)
    b = m(value, reward)
    assert torch.equal(a, b)
    half = TDLambda(T, B // 2)
    half.global_B = B
    parts = half(value[:, :B // 2].contiguous(), reward[:, :B // 2].contiguous()) + \
        half(value[:, B // 2:].contiguous(), reward[:, B // 2:].contiguous())
    assert rel_err(host(parts), host(a)) <= 1e-6
