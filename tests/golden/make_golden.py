#!/usr/bin/env python
"""Generate the committed golden fixtures (tests/golden/*.npz) by RUNNING THE REFERENCE.

Runs ``/root/reference/hpc_rll/origin`` (imported read-only via ``_ref_origin``; nothing is
copied) on seeded inputs, in fp32 and again in fp64 (``torch.set_default_dtype(float64)``,
because origin/td.py creates ``torch.ones(nstep)`` / ``linspace`` in the default dtype), and
stores inputs, every output and the autograd gradient of every differentiable input.

The reference ships no golden vectors or asserting tests for this path (its tests only print
an error figure, /root/reference/tests/testbase.py:8-11), so these origin-generated vectors are
the pin for ``oracle/`` -- see DESIGN.md "Oracle".  Re-run only in the build container:

    python tests/golden/make_golden.py

Scalar losses are combined as  L = sum_k coef_k * loss_k  with distinct coefficients so the
fixture pins the gradient path of every loss term separately.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_origin  # noqa: E402

O = _ref_origin.load()
F32, F64 = torch.float32, torch.float64


def _np(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def _run(fn, inputs, diff, dtype):
    """inputs: dict name->tensor (fp32/int64/None).  diff: names needing grad.  Returns (outs, grads)."""
    torch.set_default_dtype(dtype)
    try:
        cast = {}
        for k, v in inputs.items():
            if isinstance(v, torch.Tensor) and v.is_floating_point():
                cast[k] = v.detach().to(dtype).clone().requires_grad_(k in diff)
            else:
                cast[k] = v
        outs, total = fn(cast)
        grads = {}
        if diff:
            gs = torch.autograd.grad(total, [cast[k] for k in diff], allow_unused=True)
            for k, g in zip(diff, gs):
                grads[k] = torch.zeros_like(cast[k]) if g is None else g
        return {k: _np(v) for k, v in outs.items()}, {k: _np(v) for k, v in grads.items()}
    finally:
        torch.set_default_dtype(F32)


def save_case(name, fn, inputs, diff, attrs):
    out32, grad32 = _run(fn, inputs, diff, F32)
    out64, grad64 = _run(fn, inputs, diff, F64)
    blob = {}
    for k, v in inputs.items():
        if v is not None:
            blob["in_" + k] = _np(v)
    for k, v in out32.items():
        blob["out32_" + k] = v
    for k, v in out64.items():
        blob["out64_" + k] = v
    for k, v in grad32.items():
        blob["grad32_" + k] = v
    for k, v in grad64.items():
        blob["grad64_" + k] = v
    for k, v in attrs.items():
        blob["attr_" + k] = np.asarray(-1.0 if v is None else v, dtype=np.float64)
        if v is None:
            blob["attrnone_" + k] = np.asarray(1)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **blob)
    return path


def gen(seed):
    g = torch.Generator().manual_seed(seed)
    return g


# ----------------------------------------------------------------------------- gae
def case_gae(name, T, B, gamma, lam, seed):
    g = gen(seed)
    inputs = dict(value=torch.randn(T + 1, B, generator=g), reward=torch.randn(T, B, generator=g),
                  grad_adv=torch.randn(T, B, generator=g))

    def fn(c):
        adv = O.gae.gae(O.gae.gae_data(c["value"], c["reward"]), gamma, lam)
        return dict(adv=adv), (adv * c["grad_adv"]).sum()

    save_case(name, fn, inputs, ["value", "reward"], dict(gamma=gamma, lambda_=lam))


# ----------------------------------------------------------------------------- td_lambda
def case_tdlambda(name, T, B, gamma, lam, use_weight, seed):
    g = gen(seed)
    inputs = dict(value=torch.randn(T + 1, B, generator=g), reward=torch.randn(T, B, generator=g),
                  weight=torch.rand(T, B, generator=g) if use_weight else None)

    def fn(c):
        loss = O.td.td_lambda_error(O.td.td_lambda_data(c["value"], c["reward"], c["weight"]), gamma, lam)
        with torch.no_grad():
            ret = O.td.generalized_lambda_returns(c["value"], c["reward"], gamma, lam)
        return dict(loss=loss, ret=ret), 1.7 * loss

    save_case(name, fn, inputs, ["value"], dict(gamma=gamma, lambda_=lam, coef_loss=1.7))


# ----------------------------------------------------------------------------- vtrace
def case_vtrace(name, T, B, N, use_weight, hp, seed):
    g = gen(seed)
    inputs = dict(target_output=torch.randn(T, B, N, generator=g) * 1.5,
                  behaviour_output=torch.randn(T, B, N, generator=g) * 1.5,
                  action=torch.randint(0, N, (T, B), generator=g),
                  value=torch.randn(T + 1, B, generator=g), reward=torch.randn(T, B, generator=g),
                  weight=torch.rand(T, B, generator=g) if use_weight else None)
    coef = (1.0, 0.5, -0.25)

    def fn(c):
        l = O.vtrace.vtrace_error(O.vtrace.vtrace_data(c["target_output"], c["behaviour_output"], c["action"],
                                                       c["value"], c["reward"], c["weight"]), **hp)
        return (dict(policy_loss=l.policy_loss, value_loss=l.value_loss, entropy_loss=l.entropy_loss),
                coef[0] * l.policy_loss + coef[1] * l.value_loss + coef[2] * l.entropy_loss)

    attrs = dict(hp)
    attrs.update(coef_policy=coef[0], coef_value=coef[1], coef_entropy=coef[2])
    save_case(name, fn, inputs, ["target_output", "value"], attrs)


# ----------------------------------------------------------------------------- upgo
def case_upgo(name, T, B, N, seed):
    g = gen(seed)
    inputs = dict(target_output=torch.randn(T, B, N, generator=g) * 1.5,
                  rhos=torch.rand(T, B, generator=g) * 2,
                  action=torch.randint(0, N, (T, B), generator=g),
                  rewards=torch.randn(T, B, generator=g),
                  bootstrap_values=torch.randn(T + 1, B, generator=g))

    def fn(c):
        loss = O.upgo.upgo_loss(c["target_output"], c["rhos"], c["action"], c["rewards"], c["bootstrap_values"])
        with torch.no_grad():
            ret = O.upgo.upgo_returns(c["rewards"], c["bootstrap_values"])
        return dict(loss=loss, ret=ret), -0.7 * loss

    save_case(name, fn, inputs, ["target_output"], dict(coef_loss=-0.7))


# ----------------------------------------------------------------------------- ppo
def case_ppo(name, B, N, use_weight, clip_ratio, use_value_clip, dual_clip, seed):
    g = gen(seed)
    logit_old = torch.randn(B, N, generator=g)
    inputs = dict(logits_new=logit_old + 0.3 * torch.randn(B, N, generator=g), logits_old=logit_old,
                  action=torch.randint(0, N, (B,), generator=g),
                  value_new=torch.randn(B, generator=g), value_old=torch.randn(B, generator=g),
                  adv=torch.randn(B, generator=g), return_=torch.randn(B, generator=g),
                  weight=torch.rand(B, generator=g) if use_weight else None)
    coef = (1.0, 0.5, -0.01)

    def fn(c):
        l, info = O.ppo.ppo_error(O.ppo.ppo_data(c["logits_new"], c["logits_old"], c["action"], c["value_new"],
                                                 c["value_old"], c["adv"], c["return_"], c["weight"]),
                                  clip_ratio, use_value_clip, dual_clip)
        outs = dict(policy_loss=l.policy_loss, value_loss=l.value_loss, entropy_loss=l.entropy_loss,
                    approx_kl=torch.tensor(info.approx_kl), clipfrac=torch.tensor(info.clipfrac))
        return outs, coef[0] * l.policy_loss + coef[1] * l.value_loss + coef[2] * l.entropy_loss

    save_case(name, fn, inputs, ["logits_new", "value_new"],
              dict(clip_ratio=clip_ratio, use_value_clip=float(use_value_clip), dual_clip=dual_clip,
                   coef_policy=coef[0], coef_value=coef[1], coef_entropy=coef[2]))


# ----------------------------------------------------------------------------- gae -> normalise -> ppo
def case_gae_ppo(name, T, B, N, use_weight, clip_ratio, use_value_clip, dual_clip, seed):
    """The chain SURVEY.md 8(f)3 names: origin.gae -> (adv - mean) / (std + 1e-8) (the normalisation
    origin/ppo.py:43-47 describes) -> origin.ppo_error on the flattened (T*B,) batch."""
    g = gen(seed)
    R = T * B
    logit_old = torch.randn(R, N, generator=g)
    inputs = dict(value=torch.randn(T + 1, B, generator=g), reward=torch.randn(T, B, generator=g),
                  logits_new=logit_old + 0.3 * torch.randn(R, N, generator=g), logits_old=logit_old,
                  action=torch.randint(0, N, (R,), generator=g),
                  value_new=torch.randn(R, generator=g), value_old=torch.randn(R, generator=g),
                  return_=torch.randn(R, generator=g),
                  weight=torch.rand(R, generator=g) if use_weight else None)
    coef = (1.0, 0.5, -0.01)
    gamma, lam = 0.99, 0.97

    def fn(c):
        with torch.no_grad():
            adv = O.gae.gae(O.gae.gae_data(c["value"], c["reward"]), gamma, lam)
            mean, denom = adv.mean(), adv.std() + 1e-8
            advn = ((adv - mean) / denom).reshape(-1)
        l, info = O.ppo.ppo_error(O.ppo.ppo_data(c["logits_new"], c["logits_old"], c["action"], c["value_new"],
                                                 c["value_old"], advn, c["return_"], c["weight"]),
                                  clip_ratio, use_value_clip, dual_clip)
        outs = dict(adv=adv, adv_mean=mean, adv_denom=denom, policy_loss=l.policy_loss, value_loss=l.value_loss,
                    entropy_loss=l.entropy_loss, approx_kl=torch.tensor(info.approx_kl),
                    clipfrac=torch.tensor(info.clipfrac))
        return outs, coef[0] * l.policy_loss + coef[1] * l.value_loss + coef[2] * l.entropy_loss

    save_case(name, fn, inputs, ["logits_new", "value_new"],
              dict(gamma=gamma, lambda_=lam, clip_ratio=clip_ratio, use_value_clip=float(use_value_clip),
                   dual_clip=dual_clip, coef_policy=coef[0], coef_value=coef[1], coef_entropy=coef[2]))


# ----------------------------------------------------------------------------- q n-step (+rescale)
def _nstep_inputs(g, T, B, N, use_weight):
    return dict(q=torch.randn(B, N, generator=g), next_n_q=torch.randn(B, N, generator=g),
                action=torch.randint(0, N, (B,), generator=g), next_n_action=torch.randint(0, N, (B,), generator=g),
                reward=torch.randn(T, B, generator=g),
                done=(torch.rand(B, generator=g) < 0.3).float(),
                weight=torch.rand(B, generator=g) if use_weight else None)


def case_qnstep(name, T, B, N, gamma, use_weight, rescale, seed):
    g = gen(seed)
    inputs = _nstep_inputs(g, T, B, N, use_weight)
    if rescale:
        inputs["q"] = inputs["q"] * 3
        inputs["next_n_q"] = inputs["next_n_q"] * 3
    f = O.td.q_nstep_td_error_with_rescale if rescale else O.td.q_nstep_td_error

    def fn(c):
        loss, td = f(O.td.q_nstep_td_data(c["q"], c["next_n_q"], c["action"], c["next_n_action"], c["reward"],
                                          c["done"], c["weight"]), gamma, T)
        return dict(loss=loss, td_error_per_sample=td), 1.3 * loss

    save_case(name, fn, inputs, ["q"], dict(gamma=gamma, coef_loss=1.3))


# ----------------------------------------------------------------------------- dist n-step (C51)
def case_dist(name, T, B, N, n_atom, gamma, v_min, v_max, use_weight, seed):
    g = gen(seed)
    inputs = dict(dist=torch.softmax(torch.randn(B, N, n_atom, generator=g), -1),
                  next_n_dist=torch.softmax(torch.randn(B, N, n_atom, generator=g), -1),
                  action=torch.randint(0, N, (B,), generator=g), next_n_action=torch.randint(0, N, (B,), generator=g),
                  reward=torch.randn(T, B, generator=g) * 2,
                  done=(torch.rand(B, generator=g) < 0.3).float(),
                  weight=torch.rand(B, generator=g) if use_weight else None)

    def fn(c):
        loss, td = O.td.dist_nstep_td_error(
            O.td.dist_nstep_td_data(c["dist"], c["next_n_dist"], c["action"], c["next_n_action"], c["reward"],
                                    c["done"], c["weight"]), gamma, v_min, v_max, n_atom, T)
        return dict(loss=loss, td_error_per_sample=td), 0.9 * loss

    save_case(name, fn, inputs, ["dist"], dict(gamma=gamma, v_min=v_min, v_max=v_max, n_atom=n_atom, coef_loss=0.9))


# ----------------------------------------------------------------------------- qrdqn
def case_qrdqn(name, tau, T, B, N, gamma, use_weight, use_vg, seed):
    g = gen(seed)
    inputs = dict(q=torch.randn(B, N, tau, generator=g), next_n_q=torch.randn(B, N, tau, generator=g),
                  action=torch.randint(0, N, (B,), generator=g), next_n_action=torch.randint(0, N, (B,), generator=g),
                  reward=torch.randn(T, B, generator=g), done=(torch.rand(B, generator=g) < 0.3).float(),
                  weight=torch.rand(B, generator=g) if use_weight else None,
                  value_gamma=torch.rand(B, generator=g) if use_vg else None)

    def fn(c):
        # `tau` is the INTEGER quantile count, as in /root/reference/tests/test_qrdqn_nstep_td_error.py:57
        loss, td = O.td.qrdqn_nstep_td_error(
            O.td.qrdqn_nstep_td_data(c["q"], c["next_n_q"], c["action"], c["next_n_action"], c["reward"], c["done"],
                                     tau, c["weight"]), gamma, T, c["value_gamma"])
        return dict(loss=loss, td_error_per_sample=td), 1.1 * loss

    save_case(name, fn, inputs, ["q"], dict(gamma=gamma, tau=tau, coef_loss=1.1))


# ----------------------------------------------------------------------------- iqn
def case_iqn(name, tau, tau_p, T, B, N, gamma, kappa, use_weight, use_vg, seed):
    g = gen(seed)
    inputs = dict(q=torch.randn(tau, B, N, generator=g), next_n_q=torch.randn(tau_p, B, N, generator=g),
                  action=torch.randint(0, N, (B,), generator=g), next_n_action=torch.randint(0, N, (B,), generator=g),
                  reward=torch.randn(T, B, generator=g), done=(torch.rand(B, generator=g) < 0.3).float(),
                  replay_quantiles=torch.rand(tau, B, generator=g),
                  weight=torch.rand(B, generator=g) if use_weight else None,
                  value_gamma=torch.rand(B, generator=g) if use_vg else None)

    def fn(c):
        loss, td = O.td.iqn_nstep_td_error(
            O.td.iqn_nstep_td_data(c["q"], c["next_n_q"], c["action"], c["next_n_action"], c["reward"], c["done"],
                                   c["replay_quantiles"], c["weight"]), gamma, T, kappa, c["value_gamma"])
        return dict(loss=loss, td_error_per_sample=td), 0.8 * loss

    save_case(name, fn, inputs, ["q"], dict(gamma=gamma, kappa=kappa, coef_loss=0.8))


# ----------------------------------------------------------------------------- padding
def case_padding(name, ndim, n, lo_hi, value, group, seed):
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_origin_padding",
                                                  os.path.join(_ref_origin.REF_ROOT, "hpc_rll", "origin", "padding.py"))
    P = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(P)
    rng = np.random.default_rng(seed)
    shapes = [tuple(int(rng.integers(lo, hi)) for lo, hi in lo_hi) for _ in range(n)]
    g = gen(seed)
    data = [torch.randn(*s, generator=g) for s in shapes]
    pad = {1: P.Padding1D, 2: P.Padding2D, 3: P.Padding3D}[ndim]
    new_x, mask, ori_shapes = pad(data, value=value)
    blob = {"n": np.asarray(n), "ndim": np.asarray(ndim), "value": np.asarray(value), "group": np.asarray(group),
            "shapes": np.asarray(shapes, dtype=np.int64), "new_x": _np(new_x), "mask": _np(mask)}
    for i, d in enumerate(data):
        blob["x%d" % i] = _np(d)
    sorted_shapes = sorted(shapes, key=lambda t: int(np.prod(t)))
    sorted_data = sorted(data, key=lambda t: int(np.prod(t.shape)))
    _, positions = P.oracle_split_group(sorted_data, group)
    blob["sorted_shapes"] = np.asarray(sorted_shapes, dtype=np.int64)
    blob["positions"] = np.asarray(positions, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)



# ----------------------------------------------------------------------------- exact-tie fixtures
# Inputs constructed so that the comparisons with tie semantics are hit EXACTLY:
#   origin/td.py:512-515 (QR-DQN Huber `<1`, `le(0.)`), origin/td.py:433,443 (IQN `abs()<=kappa`, `err<0`),
#   origin/upgo.py:36 (`>=`), origin/ppo.py:62-76 (torch.min / torch.max / clamp).
def _int_tensor(g, shape, lo, hi):
    return torch.randint(lo, hi + 1, shape, generator=g).float()


def case_qrdqn_tie(name, tau, T, B, N, use_weight, seed):
    """integer-valued q / next_n_q / reward, gamma = 1, value_gamma = None (= gamma**T = 1): every pairwise error
    is an integer, so err == 0 (the `le(0.)` side) and |err| == 1 (the Huber `<1` boundary) occur in bulk."""
    g = gen(seed)
    inputs = dict(q=_int_tensor(g, (B, N, tau), -2, 2), next_n_q=_int_tensor(g, (B, N, tau), -2, 2),
                  action=torch.randint(0, N, (B,), generator=g), next_n_action=torch.randint(0, N, (B,), generator=g),
                  reward=_int_tensor(g, (T, B), -1, 1), done=(torch.rand(B, generator=g) < 0.3).float(),
                  weight=_int_tensor(g, (B,), 1, 3) if use_weight else None, value_gamma=None)

    def fn(c):
        loss, td = O.td.qrdqn_nstep_td_error(
            O.td.qrdqn_nstep_td_data(c["q"], c["next_n_q"], c["action"], c["next_n_action"], c["reward"], c["done"],
                                     tau, c["weight"]), 1.0, T, c["value_gamma"])
        return dict(loss=loss, td_error_per_sample=td), 1.0 * loss

    save_case(name, fn, inputs, ["q"], dict(gamma=1.0, tau=tau, coef_loss=1.0))


def case_iqn_tie(name, tau, tau_p, T, B, N, kappa, seed):
    """integer-valued inputs, gamma = 1: errors are integers, so |err| == kappa (inclusive quadratic branch,
    origin/td.py:433) and err == 0 (`err < 0` is false there, origin/td.py:443) are both hit; replay quantiles
    are multiples of 1/4 so the quantile weights are exact too."""
    g = gen(seed)
    inputs = dict(q=_int_tensor(g, (tau, B, N), -2, 2), next_n_q=_int_tensor(g, (tau_p, B, N), -2, 2),
                  action=torch.randint(0, N, (B,), generator=g), next_n_action=torch.randint(0, N, (B,), generator=g),
                  reward=_int_tensor(g, (T, B), -1, 1), done=(torch.rand(B, generator=g) < 0.3).float(),
                  replay_quantiles=_int_tensor(g, (tau, B), 0, 4) / 4.0, weight=None, value_gamma=None)

    def fn(c):
        loss, td = O.td.iqn_nstep_td_error(
            O.td.iqn_nstep_td_data(c["q"], c["next_n_q"], c["action"], c["next_n_action"], c["reward"], c["done"],
                                   c["replay_quantiles"], c["weight"]), 1.0, T, kappa, c["value_gamma"])
        return dict(loss=loss, td_error_per_sample=td), 1.0 * loss

    save_case(name, fn, inputs, ["q"], dict(gamma=1.0, kappa=kappa, coef_loss=1.0))


def case_upgo_tie(name, T, B, N, seed):
    """integer rewards / bootstrap values: r[t+1] + v[t+2] == v[t+1] happens on about a third of the steps, where
    origin's `>=` (upgo.py:36) keeps following the trajectory; rhos are powers of two."""
    g = gen(seed)
    inputs = dict(target_output=torch.randn(T, B, N, generator=g) * 1.5,
                  rhos=torch.tensor([0.5, 1.0, 2.0])[torch.randint(0, 3, (T, B), generator=g)],
                  action=torch.randint(0, N, (T, B), generator=g), rewards=_int_tensor(g, (T, B), -1, 1),
                  bootstrap_values=_int_tensor(g, (T + 1, B), -1, 1))

    def fn(c):
        loss = O.upgo.upgo_loss(c["target_output"], c["rhos"], c["action"], c["rewards"], c["bootstrap_values"])
        with torch.no_grad():
            ret = O.upgo.upgo_returns(c["rewards"], c["bootstrap_values"])
        return dict(loss=loss, ret=ret), 1.0 * loss

    save_case(name, fn, inputs, ["target_output"], dict(coef_loss=1.0))


def case_ppo_tie(name, B, N, clip_ratio, use_value_clip, dual_clip, seed):
    """Every PPO tie at once (origin/ppo.py:62-76):
      * logits_new == logits_old            -> ratio == 1 exactly, surr1 == surr2 (torch.min tie)
      * adv == 0 on a third of the samples  -> surr1 == surr2 == dual_clip*adv == 0 (torch.max tie of the dual clip)
      * value clip: v_new=.5, v_old=0, eps=.25, ret=.375 -> (ret-v)^2 == (ret-v_clip)^2 == 1/64 (torch.max tie:
        autograd splits the gradient, half of it reaches v_new) on every other sample."""
    g = gen(seed)
    logit = torch.randn(B, N, generator=g)
    adv = _int_tensor(g, (B,), -1, 1)
    vn = torch.randn(B, generator=g)
    vo = torch.randn(B, generator=g)
    ret = torch.randn(B, generator=g)
    vn[::2], vo[::2], ret[::2] = 0.5, 0.0, 0.375
    inputs = dict(logits_new=logit.clone(), logits_old=logit, action=torch.randint(0, N, (B,), generator=g),
                  value_new=vn, value_old=vo, adv=adv, return_=ret, weight=None)
    coef = (1.0, 1.0, -0.5)

    def fn(c):
        l, info = O.ppo.ppo_error(O.ppo.ppo_data(c["logits_new"], c["logits_old"], c["action"], c["value_new"],
                                                 c["value_old"], c["adv"], c["return_"], c["weight"]),
                                  clip_ratio, use_value_clip, dual_clip)
        outs = dict(policy_loss=l.policy_loss, value_loss=l.value_loss, entropy_loss=l.entropy_loss,
                    approx_kl=torch.tensor(info.approx_kl), clipfrac=torch.tensor(info.clipfrac))
        return outs, coef[0] * l.policy_loss + coef[1] * l.value_loss + coef[2] * l.entropy_loss

    save_case(name, fn, inputs, ["logits_new", "value_new"],
              dict(clip_ratio=clip_ratio, use_value_clip=float(use_value_clip), dual_clip=dual_clip,
                   coef_policy=coef[0], coef_value=coef[1], coef_entropy=coef[2]))


def main_ties():
    s = 4321
    case_qrdqn_tie("qrdqn_tie_tau16_t3_b12_n3", 16, 3, 12, 3, False, s + 1)
    case_qrdqn_tie("qrdqn_tie_tau64_t5_b6_n2_w", 64, 5, 6, 2, True, s + 2)
    case_iqn_tie("iqn_tie_tau16_16_t3_b12_n3_k1", 16, 16, 3, 12, 3, 1.0, s + 3)
    case_iqn_tie("iqn_tie_tau33_20_t2_b7_n2_k2", 33, 20, 2, 7, 2, 2.0, s + 4)
    case_upgo_tie("upgo_tie_t24_b9_n5", 24, 9, 5, s + 5)
    case_upgo_tie("upgo_tie_t7_b33_n16", 7, 33, 16, s + 6)
    case_ppo_tie("ppo_tie_b24_n6_vclip", 24, 6, 0.25, True, None, s + 7)
    case_ppo_tie("ppo_tie_b24_n16_vclip_dual", 24, 16, 0.25, True, 2.0, s + 8)
    case_ppo_tie("ppo_tie_b10_n3_noclip_dual", 10, 3, 0.25, False, 3.0, s + 9)


def main_gae_ppo():
    s = 1234
    case_gae_ppo("gaeppo_t16_b8_n6", 16, 8, 6, False, 0.2, True, None, s + 100)
    case_gae_ppo("gaeppo_t12_b5_n16_w_dual", 12, 5, 16, True, 0.2, True, 3.0, s + 101)
    case_gae_ppo("gaeppo_t7_b3_n18_w", 7, 3, 18, True, 0.1, False, None, s + 102)


def main():
    if sys.argv[1:] == ["gaeppo"]:  # add the chain fixtures without re-zipping the others
        return main_gae_ppo()
    if sys.argv[1:] == ["ties"]:  # add the exact-tie fixtures without re-zipping the others
        return main_ties()
    for f in os.listdir(HERE):
        if f.endswith(".npz"):
            os.remove(os.path.join(HERE, f))
    main_gae_ppo()
    main_ties()
    s = 1234
    case_gae("gae_t16_b8", 16, 8, 0.99, 0.97, s + 1)
    case_gae("gae_t1_b5", 1, 5, 0.99, 0.97, s + 2)
    case_gae("gae_t37_b3", 37, 3, 0.9, 0.5, s + 3)
    case_gae("gae_t64_b1", 64, 1, 1.0, 1.0, s + 4)
    case_gae("gae_t128_b12", 128, 12, 0.99, 0.97, s + 5)

    case_tdlambda("tdlambda_t16_b8_w", 16, 8, 0.9, 0.8, True, s + 10)
    case_tdlambda("tdlambda_t16_b8", 16, 8, 0.9, 0.8, False, s + 11)
    case_tdlambda("tdlambda_t1_b3", 1, 3, 0.99, 0.95, True, s + 12)
    case_tdlambda("tdlambda_t50_b5", 50, 5, 0.99, 1.0, True, s + 13)

    hp = dict(gamma=0.99, lambda_=0.95, rho_clip_ratio=1.0, c_clip_ratio=1.0, rho_pg_clip_ratio=1.0)
    hp2 = dict(gamma=0.9, lambda_=0.8, rho_clip_ratio=1.5, c_clip_ratio=0.9, rho_pg_clip_ratio=2.0)
    case_vtrace("vtrace_t12_b6_n5", 12, 6, 5, False, hp, s + 20)
    case_vtrace("vtrace_t12_b6_n16_w", 12, 6, 16, True, hp2, s + 21)
    case_vtrace("vtrace_t1_b2_n1", 1, 2, 1, True, hp, s + 22)
    case_vtrace("vtrace_t9_b3_n40", 9, 3, 40, False, hp2, s + 23)

    case_upgo("upgo_t12_b6_n5", 12, 6, 5, s + 30)
    case_upgo("upgo_t1_b4_n3", 1, 4, 3, s + 31)
    case_upgo("upgo_t2_b4_n16", 2, 4, 16, s + 32)
    case_upgo("upgo_t20_b3_n33", 20, 3, 33, s + 33)

    case_ppo("ppo_b32_n6", 32, 6, False, 0.2, True, None, s + 40)
    case_ppo("ppo_b32_n6_w_dual", 32, 6, True, 0.2, True, 3.0, s + 41)
    case_ppo("ppo_b17_n16_noclip", 17, 16, True, 0.1, False, None, s + 42)
    case_ppo("ppo_b5_n1", 5, 1, False, 0.2, True, 2.0, s + 43)
    case_ppo("ppo_b9_n37_dual", 9, 37, False, 0.3, False, 1.5, s + 44)

    case_qnstep("qnstep_t5_b16_n6", 5, 16, 6, 0.95, False, False, s + 50)
    case_qnstep("qnstep_t1_b7_n3_w", 1, 7, 3, 0.99, True, False, s + 51)
    case_qnstep("qnstep_t30_b4_n1_w", 30, 4, 1, 0.95, True, False, s + 52)
    case_qnstep("qnstep_rescale_t5_b16_n6", 5, 16, 6, 0.95, False, True, s + 53)
    case_qnstep("qnstep_rescale_t3_b9_n4_w", 3, 9, 4, 0.99, True, True, s + 54)

    case_dist("dist_t5_b8_n4_a51", 5, 8, 4, 51, 0.95, -10.0, 10.0, False, s + 60)
    case_dist("dist_t3_b6_n3_a21_w", 3, 6, 3, 21, 0.99, -5.0, 5.0, True, s + 61)
    case_dist("dist_t1_b4_n2_a5_w", 1, 4, 2, 5, 0.9, 0.0, 4.0, True, s + 62)

    case_qrdqn("qrdqn_tau8_t5_b6_n4", 8, 5, 6, 4, 0.95, False, False, s + 70)
    case_qrdqn("qrdqn_tau39_t10_b5_n3_w_vg", 39, 10, 5, 3, 0.95, True, True, s + 71)
    case_qrdqn("qrdqn_tau64_t5_b3_n2_w", 64, 5, 3, 2, 0.99, True, False, s + 72)
    case_qrdqn("qrdqn_tau1_t1_b4_n1", 1, 1, 4, 1, 0.9, False, False, s + 73)

    case_iqn("iqn_tau8_9_t5_b6_n4", 8, 9, 5, 6, 4, 0.95, 1.0, False, False, s + 80)
    case_iqn("iqn_tau33_34_t10_b5_n3_w_vg", 33, 34, 10, 5, 3, 0.95, 0.9, True, True, s + 81)
    case_iqn("iqn_tau64_64_t5_b3_n2_w", 64, 64, 5, 3, 2, 0.99, 1.0, True, False, s + 82)
    case_iqn("iqn_tau1_1_t1_b4_n1", 1, 1, 1, 4, 1, 0.9, 0.5, False, False, s + 83)
    case_padding("padding_1d_n64", 1, 64, [(32, 128)], 0, 4, s + 90)       # tests/test_padding.py:10-11
    case_padding("padding_1d_n7_v3", 1, 7, [(1, 9)], 3, 3, s + 91)
    case_padding("padding_2d_n12", 2, 12, [(6, 14), (4, 11)], 0, 4, s + 92)   # (test_padding.py:12 ranges, scaled down)
    case_padding("padding_3d_n8", 3, 8, [(3, 7), (3, 6), (4, 9)], 0, 3, s + 93)  # (test_padding.py:13, scaled down)
    case_padding("padding_3d_n5_small", 3, 5, [(1, 4), (1, 5), (1, 6)], -1, 2, s + 94)
    n = len([f for f in os.listdir(HERE) if f.endswith(".npz")])
    tot = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz"))
    print("wrote %d fixtures, %.1f KB" % (n, tot / 1024))


if __name__ == "__main__":
    main()
