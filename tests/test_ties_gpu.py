"""Exact-tie semantics on the DEVICE ("no GPU test hits a tie/boundary").

The fixtures `*_tie_*.npz` are produced by running hpc_rll.origin (tests/golden/make_golden.py, main_ties) on
inputs built so that the comparisons with tie rules are hit exactly:

  QR-DQN  origin/td.py:512-515   Huber `|err| < 1` (strict) and the quantile side `err <= 0` -- integer errors
  IQN     origin/td.py:433,443   `|err| <= kappa` (inclusive) and `err < 0` (strict)            -- integer errors
  UPGO    origin/upgo.py:36      `r[t+1] + v[t+2] >= v[t+1]` with equality on ~1/3 of the steps
  PPO     origin/ppo.py:62-76    ratio == 1 (torch.min tie), adv == 0 (dual-clip torch.max tie), value-clip tie
                                 (ret-v)^2 == (ret-v_clip)^2 where autograd gives v_new HALF the gradient

Every quantity in these cases is a short dyadic rational, so the only roundings are the final scalings: the
CUDA path must agree with origin-fp32 to 1e-6 (ten times tighter than the general bar) and the gradient's
sign/zero structure must be IDENTICAL -- a flipped tie changes a gradient entry by O(1/B), far above that.
"""
import numpy as np
import pytest

from tests._golden import Case, names, rel_err
from tests._gpu import need_cuda
from tests.test_nstep_gpu import run_iqn, run_qrdqn
from tests.test_policy_ops_gpu import run_ppo, run_upgo

pytestmark = pytest.mark.gpu
TIE_TOL = 1e-6


def tight(got, want, what):
    e = rel_err(got, want)
    assert e <= TIE_TOL, "%s: rel err %.3e on an exact-tie fixture" % (what, e)


def same_structure(got, want, what):
    """identical sign pattern wherever origin's entry is not rounding dust: origin scales every pair term by 1/(B*tau)
    BEFORE summing, so a sum that cancels exactly in exact arithmetic can come out as ~1e-10 instead of 0 there"""
    got, want = np.asarray(got), np.asarray(want)
    big = np.abs(want) > 1e-6 * max(1.0, float(np.abs(want).max()))
    assert np.array_equal(np.sign(got)[big], np.sign(want)[big]), "%s: sign structure differs on a tie fixture" % what
    assert np.all(np.abs(got[~big]) <= 1e-6 * max(1.0, float(np.abs(want).max()))), "%s: spurious entries" % what


def tie_names(prefix):
    out = [n for n in names(prefix) if "_tie_" in n]
    assert out, "no tie fixtures for %s (run tests/golden/make_golden.py ties)" % prefix
    return out


@pytest.mark.parametrize("name", tie_names("qrdqn"))
def test_qrdqn_ties(name):
    need_cuda()
    c = Case(name)
    inp = {k: c.inp(k) for k in ("q", "next_n_q", "action", "next_n_action", "reward", "done", "weight",
                                 "value_gamma")}
    # the fixture really contains the ties it is named after
    tgt = inp["next_n_q"][np.arange(len(inp["action"])), inp["next_n_action"]]
    cur = inp["q"][np.arange(len(inp["action"])), inp["action"]]
    R = inp["reward"].sum(0)[:, None]
    err = (R + tgt * (1 - inp["done"])[:, None])[:, None, :] - cur[:, :, None]
    assert (err == 0).mean() > 0.05 and (np.abs(err) == 1).mean() > 0.05
    loss, td, gq = run_qrdqn(inp, c.attr("gamma"), c.attr("coef_loss"))
    tight(loss, c.out("loss", 32), "loss")
    tight(td, c.out("td_error_per_sample", 32), "td_error_per_sample")
    tight(gq, c.grad("q", 32), "grad_q")
    same_structure(gq, c.grad("q", 32), "grad_q")
    assert np.array_equal(td, c.out("td_error_per_sample", 32)), "integer-error td sums are exact in fp32"


@pytest.mark.parametrize("name", tie_names("iqn"))
def test_iqn_ties(name):
    need_cuda()
    c = Case(name)
    inp = {k: c.inp(k) for k in ("q", "next_n_q", "action", "next_n_action", "reward", "done", "replay_quantiles",
                                 "weight", "value_gamma")}
    loss, td, gq = run_iqn(inp, c.attr("gamma"), c.attr("kappa"), c.attr("coef_loss"))
    tight(loss, c.out("loss", 32), "loss")
    tight(td, c.out("td_error_per_sample", 32), "td_error_per_sample")
    tight(gq, c.grad("q", 32), "grad_q")
    same_structure(gq, c.grad("q", 32), "grad_q")


@pytest.mark.parametrize("name", tie_names("upgo"))
def test_upgo_ties(name):
    need_cuda()
    c = Case(name)
    inp = {k: c.inp(k) for k in ("target_output", "rhos", "action", "rewards", "bootstrap_values")}
    r, v = inp["rewards"], inp["bootstrap_values"]
    assert (r[1:] + v[2:] == v[1:-1]).mean() > 0.1  # the `>=` boundary is exercised
    loss, gt = run_upgo(inp, c.attr("coef_loss"))
    tight(loss, c.out("loss", 32), "loss")
    tight(gt, c.grad("target_output", 32), "grad_target_output")
    # a flipped tie changes the return of that step by an integer, i.e. the advantage that scales the whole
    # logits row (and its sign / exact zeros, since every advantage here is an integer multiple of rho)
    same_structure(gt, c.grad("target_output", 32), "grad_target_output")


@pytest.mark.parametrize("name", tie_names("ppo"))
def test_ppo_ties(name):
    need_cuda()
    c = Case(name)
    inp = {k: c.inp(k) for k in ("logits_new", "logits_old", "action", "value_new", "value_old", "adv", "return_",
                                 "weight")}
    assert np.array_equal(inp["logits_new"], inp["logits_old"]) and (inp["adv"] == 0).any()
    coef = [c.attr("coef_policy"), c.attr("coef_value"), c.attr("coef_entropy")]
    outs, gl, gv = run_ppo(inp, c.attr("clip_ratio"), bool(c.attr("use_value_clip")), c.attr("dual_clip"), coef)
    for k, nm in enumerate(("policy_loss", "value_loss", "entropy_loss")):
        tight(outs[k], c.out(nm, 32), nm)
    tight(gl, c.grad("logits_new", 32), "grad_logits_new")
    tight(gv, c.grad("value_new", 32), "grad_value_new")
    assert outs[4] == 0.0 and float(c.out("clipfrac", 32)) == 0.0  # ratio == 1 is inside [1-eps, 1+eps]
    assert abs(outs[3]) <= 1e-7  # approx_kl of identical policies
    if bool(c.attr("use_value_clip")):
        # the value-clip tie samples (every other one): autograd's torch.max tie hands v_new HALF of
        # d/dv 0.5*(ret-v)^2 = (v-ret) = 0.125, i.e. 0.5 * 0.125 / B, bit for bit
        B = len(inp["adv"])
        want = np.float32(0.5 * 0.125 * coef[1]) / np.float32(B)
        assert np.array_equal(gv[::2], np.full_like(gv[::2], want)), (gv[::2], want)
        assert np.array_equal(gv[::2], c.grad("value_new", 32)[::2])
