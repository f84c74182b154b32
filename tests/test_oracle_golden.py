"""Pin oracle/oracle.c to the reference: every origin-generated golden fixture (fp32 and fp64)
must be reproduced by the C restatement.  fp64: <=1e-12 (same algorithm, different summation
order only).  fp32: <=2e-6 norm-relative; GAE forward is element-wise and must be BIT-EXACT."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests._golden import Case, names, rel_err

TOL = {32: 2e-6, 64: 1e-12}
DT = {32: np.float32, 64: np.float64}


def check(got, want, prec, what):
    e = rel_err(got, want)
    assert e <= TOL[prec], "%s: rel err %.3e > %.1e" % (what, e, TOL[prec])


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", names("gae"))
def test_gae(name, prec):
    c = Case(name)
    dt = DT[prec]
    g, l = c.attr("gamma"), c.attr("lambda_")
    adv = orc.gae_forward(c.inp("value", dt), c.inp("reward", dt), g, l)
    if prec == 32:
        assert np.array_equal(adv, c.out("adv", 32)), "GAE forward must be bit-exact vs origin fp32"
    check(adv, c.out("adv", prec), prec, "adv")
    gr = orc.gae_backward(c.inp("grad_adv", dt), g, l)
    check(gr["value"], c.grad("value", prec), prec, "grad_value")
    check(gr["reward"], c.grad("reward", prec), prec, "grad_reward")


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", names("tdlambda"))
def test_td_lambda(name, prec):
    c = Case(name)
    dt = DT[prec]
    r = orc.td_lambda(c.inp("value", dt), c.inp("reward", dt), c.inp("weight", dt), c.attr("gamma"),
                      c.attr("lambda_"), c.attr("coef_loss"))
    check(r["loss"], c.out("loss", prec), prec, "loss")
    check(r["ret"], c.out("ret", prec), prec, "ret")
    check(r["grad_value"], c.grad("value", prec), prec, "grad_value")
    assert np.all(r["grad_value"][-1] == 0)


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", names("vtrace"))
def test_vtrace(name, prec):
    c = Case(name)
    dt = DT[prec]
    coef = [c.attr("coef_policy"), c.attr("coef_value"), c.attr("coef_entropy")]
    r = orc.vtrace(c.inp("target_output", dt), c.inp("behaviour_output", dt), c.inp("action"), c.inp("value", dt),
                   c.inp("reward", dt), c.inp("weight", dt), c.attr("gamma"), c.attr("lambda_"),
                   c.attr("rho_clip_ratio"), c.attr("c_clip_ratio"), c.attr("rho_pg_clip_ratio"), coef)
    for k in ("policy_loss", "value_loss", "entropy_loss"):
        check(r[k], c.out(k, prec), prec, k)
    check(r["grad_target_output"], c.grad("target_output", prec), prec, "grad_target_output")
    check(r["grad_value"], c.grad("value", prec), prec, "grad_value")


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", names("upgo"))
def test_upgo(name, prec):
    c = Case(name)
    dt = DT[prec]
    r = orc.upgo(c.inp("target_output", dt), c.inp("rhos", dt), c.inp("action"), c.inp("rewards", dt),
                 c.inp("bootstrap_values", dt), c.attr("coef_loss"))
    check(r["loss"], c.out("loss", prec), prec, "loss")
    check(r["ret"], c.out("ret", prec), prec, "ret")
    check(r["grad_target_output"], c.grad("target_output", prec), prec, "grad_target_output")


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", names("ppo"))
def test_ppo(name, prec):
    c = Case(name)
    dt = DT[prec]
    coef = [c.attr("coef_policy"), c.attr("coef_value"), c.attr("coef_entropy")]
    r = orc.ppo(c.inp("logits_new", dt), c.inp("logits_old", dt), c.inp("action"), c.inp("value_new", dt),
                c.inp("value_old", dt), c.inp("adv", dt), c.inp("return_", dt), c.inp("weight", dt),
                c.attr("clip_ratio"), bool(c.attr("use_value_clip")), c.attr("dual_clip"), coef)
    for k in ("policy_loss", "value_loss", "entropy_loss", "approx_kl", "clipfrac"):
        # approx_kl / clipfrac are python floats taken from an fp32 mean in origin
        check(r[k], c.out(k, prec), 32 if k in ("approx_kl", "clipfrac") else prec, k)
    check(r["grad_logits_new"], c.grad("logits_new", prec), prec, "grad_logits_new")
    check(r["grad_value_new"], c.grad("value_new", prec), prec, "grad_value_new")


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", names("gaeppo"))
def test_gae_norm_ppo_chain(name, prec):
    """origin.gae -> (adv - mean) / (std + 1e-8) -> origin.ppo_error."""
    c = Case(name)
    dt = DT[prec]
    coef = [c.attr("coef_policy"), c.attr("coef_value"), c.attr("coef_entropy")]
    r = orc.gae_norm_ppo(c.inp("value", dt), c.inp("reward", dt), c.inp("logits_new", dt), c.inp("logits_old", dt),
                         c.inp("action"), c.inp("value_new", dt), c.inp("value_old", dt), c.inp("return_", dt),
                         c.inp("weight", dt), c.attr("gamma"), c.attr("lambda_"), c.attr("clip_ratio"),
                         bool(c.attr("use_value_clip")), c.attr("dual_clip"), coef)
    check(r["adv"], c.out("adv", prec), prec, "adv")
    check(r["adv_mean"], c.out("adv_mean", prec), prec, "adv_mean")
    check(r["adv_denom"], c.out("adv_denom", prec), prec, "adv_denom")
    for k in ("policy_loss", "value_loss", "entropy_loss", "approx_kl", "clipfrac"):
        check(r[k], c.out(k, prec), 32 if k in ("approx_kl", "clipfrac") else prec, k)
    check(r["grad_logits_new"], c.grad("logits_new", prec), prec, "grad_logits_new")
    check(r["grad_value_new"], c.grad("value_new", prec), prec, "grad_value_new")


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", names("qnstep"))
def test_q_nstep(name, prec):
    c = Case(name)
    dt = DT[prec]
    r = orc.q_nstep_td(c.inp("q", dt), c.inp("next_n_q", dt), c.inp("action"), c.inp("next_n_action"),
                       c.inp("reward", dt), c.inp("done", dt), c.inp("weight", dt), c.attr("gamma"),
                       "rescale" in name, c.attr("coef_loss"))
    check(r["loss"], c.out("loss", prec), prec, "loss")
    check(r["td_error_per_sample"], c.out("td_error_per_sample", prec), prec, "td_err")
    gq = c.grad("q", prec)
    check(r["grad_q"], gq, prec, "grad_q")
    assert np.array_equal(r["grad_q"] != 0, gq != 0), "gradient zero-pattern (action gather) must be exact"


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", names("dist"))
def test_dist_nstep(name, prec):
    c = Case(name)
    dt = DT[prec]
    r = orc.dist_nstep_td(c.inp("dist", dt), c.inp("next_n_dist", dt), c.inp("action"), c.inp("next_n_action"),
                          c.inp("reward", dt), c.inp("done", dt), c.inp("weight", dt), c.attr("gamma"),
                          c.attr("v_min"), c.attr("v_max"), c.attr("coef_loss"))
    check(r["loss"], c.out("loss", prec), prec, "loss")
    check(r["td_error_per_sample"], c.out("td_error_per_sample", prec), prec, "td_err")
    check(r["grad_dist"], c.grad("dist", prec), prec, "grad_dist")


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", names("qrdqn"))
def test_qrdqn(name, prec):
    c = Case(name)
    dt = DT[prec]
    r = orc.qrdqn_nstep_td(c.inp("q", dt), c.inp("next_n_q", dt), c.inp("action"), c.inp("next_n_action"),
                           c.inp("reward", dt), c.inp("done", dt), c.inp("weight", dt), c.inp("value_gamma", dt),
                           c.attr("gamma"), c.attr("coef_loss"))
    check(r["loss"], c.out("loss", prec), prec, "loss")
    check(r["td_error_per_sample"], c.out("td_error_per_sample", prec), prec, "td_err")
    check(r["grad_q"], c.grad("q", prec), prec, "grad_q")


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", names("iqn"))
def test_iqn(name, prec):
    c = Case(name)
    dt = DT[prec]
    r = orc.iqn_nstep_td(c.inp("q", dt), c.inp("next_n_q", dt), c.inp("action"), c.inp("next_n_action"),
                         c.inp("reward", dt), c.inp("done", dt), c.inp("replay_quantiles", dt),
                         c.inp("weight", dt), c.inp("value_gamma", dt), c.attr("gamma"), c.attr("kappa"),
                         c.attr("coef_loss"))
    check(r["loss"], c.out("loss", prec), prec, "loss")
    check(r["td_error_per_sample"], c.out("td_error_per_sample", prec), prec, "td_err")
    check(r["grad_q"], c.grad("q", prec), prec, "grad_q")
