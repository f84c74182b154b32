"""-m gpu parity of the fused chain GAE -> advantage moments -> normalised PPO:
hpc_rll_gae_forward_moments + hpc_rll_adv_stats + hpc_rll_ppo_forward_norm vs the oracle restatement of
origin.gae -> (adv - mean) / (std + 1e-8) -> origin.ppo_error, and vs the origin-generated fixtures."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests._golden import Case, names
from tests._gpu import dev, host, need_cuda, rel_err, rng

pytestmark = pytest.mark.gpu
TOL = 1e-5


def close(got, want, what):
    e = rel_err(got, want)
    assert e <= TOL, "%s: rel err %.3e" % (what, e)


def run_chain(inp, hp, coef):
    from hpc_rll.rl_utils.gae import gae_with_adv_stats
    from hpc_rll.rl_utils.ppo import PPO
    adv, stats = gae_with_adv_stats(dev(inp["value"]), dev(inp["reward"]), hp["gamma"], hp["lambda_"])
    ln = dev(inp["logits_new"]).requires_grad_(True)
    vn = dev(inp["value_new"]).requires_grad_(True)
    w = None if inp.get("weight") is None else dev(inp["weight"])
    R, N = inp["logits_new"].shape
    loss, info = PPO(R, N)(ln, dev(inp["logits_old"]), dev(inp["action"]), vn, dev(inp["value_old"]), adv.reshape(-1),
                           dev(inp["return_"]), w, hp["clip_ratio"], hp["use_value_clip"], hp["dual_clip"],
                           adv_stats=stats)
    (coef[0] * loss.policy_loss + coef[1] * loss.value_loss + coef[2] * loss.entropy_loss).sum().backward()
    torch.cuda.synchronize()
    return dict(adv=host(adv), stats=host(stats), losses=[float(x.item()) for x in loss] + list(info),
                grad_logits_new=host(ln.grad), grad_value_new=host(vn.grad))


def chain_inputs(g, T, B, N, use_w):
    R = T * B
    lo = g.standard_normal((R, N)).astype(np.float32)
    return dict(value=g.standard_normal((T + 1, B)).astype(np.float32),
                reward=g.standard_normal((T, B)).astype(np.float32),
                logits_new=(lo + 0.3 * g.standard_normal((R, N))).astype(np.float32), logits_old=lo,
                action=g.integers(0, N, (R, )).astype(np.int64), value_new=g.standard_normal(R).astype(np.float32),
                value_old=g.standard_normal(R).astype(np.float32), return_=g.standard_normal(R).astype(np.float32),
                weight=g.random(R).astype(np.float32) if use_w else None)


@pytest.mark.parametrize("T,B,N,use_w,dual", [(64, 512, 16, True, None), (33, 40, 6, False, 3.0), (128, 1024, 18, True, None),
                                             (5, 7, 3, True, 2.0), (1, 4, 40, False, None), (700, 36, 8, False, None),
                                             (16, 4100, 4, True, None), (12, 33, 300, False, None)])
def test_chain_vs_oracle(T, B, N, use_w, dual):
    need_cuda()
    inp = chain_inputs(rng(T * 7 + B * 3 + N), T, B, N, use_w)
    hp = dict(gamma=0.99, lambda_=0.97, clip_ratio=0.2, use_value_clip=True, dual_clip=dual)
    coef = [1.0, 0.5, -0.01]
    r = run_chain(inp, hp, coef)
    o = orc.gae_norm_ppo(inp["value"], inp["reward"], inp["logits_new"], inp["logits_old"], inp["action"],
                         inp["value_new"], inp["value_old"], inp["return_"], inp["weight"], hp["gamma"], hp["lambda_"],
                         hp["clip_ratio"], hp["use_value_clip"], hp["dual_clip"], coef)
    assert np.array_equal(r["adv"], o["adv"]), "GAE forward must stay bit-exact with moments on"
    # moments are accumulated in fp64 (GPU: over fp32 runs of 16 rows): the statistics agree to ~an fp32 ulp
    assert abs(r["stats"][0] - o["adv_mean"]) <= 1e-6 * max(1.0, abs(o["adv_mean"]))
    assert abs(r["stats"][1] - o["adv_denom"]) <= 1e-6 * o["adv_denom"]
    for k, nm in enumerate(("policy_loss", "value_loss", "entropy_loss", "approx_kl")):
        close(r["losses"][k], o[nm], nm)
    # clipfrac counts threshold crossings: a ratio within an ulp of 1 +- clip may fall either side
    assert abs(r["losses"][4] - o["clipfrac"]) <= max(1e-5, 2.0 / (T * B))
    close(r["grad_logits_new"], o["grad_logits_new"], "grad_logits_new")
    close(r["grad_value_new"], o["grad_value_new"], "grad_value_new")


@pytest.mark.parametrize("name", names("gaeppo"))
def test_chain_vs_golden(name):
    need_cuda()
    c = Case(name)
    inp = {k: c.inp(k) for k in ("value", "reward", "logits_new", "logits_old", "action", "value_new", "value_old",
                                 "return_", "weight")}
    hp = dict(gamma=c.attr("gamma"), lambda_=c.attr("lambda_"), clip_ratio=c.attr("clip_ratio"),
              use_value_clip=bool(c.attr("use_value_clip")), dual_clip=c.attr("dual_clip"))
    coef = [c.attr("coef_policy"), c.attr("coef_value"), c.attr("coef_entropy")]
    r = run_chain(inp, hp, coef)
    assert np.array_equal(r["adv"], c.out("adv", 32))
    for prec in (32, 64):
        close(r["stats"][0], c.out("adv_mean", prec), "adv_mean")
        close(r["stats"][1], c.out("adv_denom", prec), "adv_denom")
        for k, nm in enumerate(("policy_loss", "value_loss", "entropy_loss")):
            close(r["losses"][k], c.out(nm, prec), nm)
        close(r["grad_logits_new"], c.grad("logits_new", prec), "grad_logits_new")
        close(r["grad_value_new"], c.grad("value_new", prec), "grad_value_new")
    close(r["losses"][3], c.out("approx_kl", 32), "approx_kl")
    close(r["losses"][4], c.out("clipfrac", 32), "clipfrac")


def test_stats_equal_unfused_normalisation():
    """Handing adv_stats to PPO == normalising adv first with the same two scalars (bit-identical losses)."""
    need_cuda()
    from hpc_rll.rl_utils.gae import gae_with_adv_stats
    from hpc_rll.rl_utils.ppo import PPO
    inp = chain_inputs(rng(5), 32, 256, 16, True)
    adv, stats = gae_with_adv_stats(dev(inp["value"]), dev(inp["reward"]))
    args = [dev(inp[k]) for k in ("logits_new", "logits_old", "action", "value_new", "value_old")]
    tail = [dev(inp["return_"]), dev(inp["weight"])]
    fused, _ = PPO(1, 1)(*args, adv.reshape(-1), *tail, adv_stats=stats)
    pre = ((adv - stats[0]) / stats[1]).reshape(-1)
    plain, _ = PPO(1, 1)(*args, pre, *tail)
    for a, b in zip(fused, plain):
        assert torch.equal(a, b)


def test_moments_sharded_equals_whole():
    """Moments of column shards add up to the moments of the whole batch (what all_reduce_moments sums)."""
    need_cuda()
    from di_hpc_b200 import _abi
    from di_hpc_b200.sharding import shard_columns
    g = rng(9)
    T, B = 48, 1000
    value, reward = g.standard_normal((T + 1, B)).astype(np.float32), g.standard_normal((T, B)).astype(np.float32)

    def moments(v, r):
        v, r = dev(v), dev(r)
        t, b = r.shape
        adv = torch.empty_like(r)
        m = torch.zeros(3, dtype=torch.float64, device="cuda")
        ws = _abi.workspace(_abi.OP_GAE_MOMENTS, t, b, 0, r.device)
        _abi.check(_abi.lib().hpc_rll_gae_forward_moments(_abi.ptr(v), _abi.ptr(r), _abi.ptr(adv), _abi.ptr(m), t, b,
                                                          0.99, 0.97, _abi.ptr(ws), ws.numel(), _abi.stream_of(r)),
                   "moments")
        return host(m)[:2], host(adv)

    whole, adv = moments(value, reward)
    parts = np.zeros(2)
    for rank in range(3):
        b0, b1 = shard_columns(B, rank, 3)
        m, _ = moments(value[:, b0:b1], reward[:, b0:b1])
        parts += m
    a64 = adv.astype(np.float64)
    # runs of 16 rows are summed in fp32 before they are folded into the fp64 accumulators
    assert abs(whole[0] - a64.sum()) <= 2e-7 * np.abs(a64).sum()
    assert abs(whole[1] - np.square(a64).sum()) <= 2e-7 * np.square(a64).sum()
    assert np.allclose(parts, whole, rtol=1e-12, atol=1e-9)


def test_argument_errors():
    need_cuda()
    from hpc_rll.rl_utils.gae import gae_with_adv_stats
    from hpc_rll.rl_utils.ppo import PPO
    with pytest.raises(ValueError):
        gae_with_adv_stats(torch.zeros(4, 3, device="cuda"), torch.zeros(4, 3, device="cuda"))
    inp = chain_inputs(rng(1), 2, 4, 4, False)
    args = [dev(inp[k]) for k in ("logits_new", "logits_old", "action", "value_new", "value_old")]
    with pytest.raises(ValueError):
        PPO(8, 4)(*args, torch.zeros(8, device="cuda"), dev(inp["return_"]), adv_stats=torch.zeros(3, device="cuda"))
