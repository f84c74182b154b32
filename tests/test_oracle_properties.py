"""Independent checks of the oracle's closed-form adjoints and structure (CPU, fp64): the golden fixtures pin
oracle/oracle.c to origin's autograd; here every gradient is ALSO compared with central finite differences of the
oracle's own forward, and the structural identities the GPU tests rely on at full size are verified exactly."""
import numpy as np
import pytest

from oracle import oracle as orc

F = np.float64


def rng(seed):
    return np.random.default_rng(seed)


def fd_check(f, x, grad, n_probe, g, eps=1e-6, tol=2e-6):
    """central differences of the scalar f() w.r.t. a few entries of x (perturbed in place) vs grad"""
    flat, gflat = x.reshape(-1), np.asarray(grad).reshape(-1)
    for i in g.choice(flat.size, size=min(n_probe, flat.size), replace=False):
        old = flat[i]
        flat[i] = old + eps
        fp = f()
        flat[i] = old - eps
        fm = f()
        flat[i] = old
        num = (fp - fm) / (2 * eps)
        assert abs(num - gflat[i]) <= tol * max(1.0, abs(num)), (i, num, gflat[i])


def test_gae_is_linear_and_adjoint_identity():
    g = rng(1)
    T, B = 37, 11
    v, r, G = g.standard_normal((T + 1, B)), g.standard_normal((T, B)), g.standard_normal((T, B))
    v2, r2 = g.standard_normal((T + 1, B)), g.standard_normal((T, B))
    a1, a2 = orc.gae_forward(v, r, 0.97, 0.9), orc.gae_forward(v2, r2, 0.97, 0.9)
    a12 = orc.gae_forward(v + 2.5 * v2, r + 2.5 * r2, 0.97, 0.9)
    assert np.allclose(a12, a1 + 2.5 * a2, rtol=1e-12, atol=1e-12)
    gb = orc.gae_backward(G, 0.97, 0.9)
    lhs = float((a1 * G).sum())
    rhs = float((v * gb["value"]).sum() + (r * gb["reward"]).sum())
    assert abs(lhs - rhs) <= 1e-11 * max(1.0, abs(lhs))


def test_td_lambda_gradient_fd():
    g = rng(2)
    T, B = 9, 5
    v, r, w = g.standard_normal((T + 1, B)), g.standard_normal((T, B)), g.random((T, B))
    o = orc.td_lambda(v, r, w, 0.9, 0.8, 1.0)
    # the lambda-return is a constant for the gradient (origin: no_grad), so differentiate with ret frozen
    ret = o["ret"]

    def f():
        return 0.5 * float((w * (ret - v[:-1]) ** 2).mean())

    fd_check(f, v, o["grad_value"], 12, g)


def test_vtrace_gradients_fd():
    g = rng(3)
    T, B, N = 5, 4, 6
    t, b = g.standard_normal((T, B, N)), g.standard_normal((T, B, N))
    a = g.integers(0, N, (T, B)).astype(np.int64)
    v, r, w = g.standard_normal((T + 1, B)), g.standard_normal((T, B)), g.random((T, B))
    coef = [1.0, 0.5, -0.25]
    hp = dict(gamma=0.9, lambda_=0.8, rho_clip_ratio=1.5, c_clip_ratio=0.9, rho_pg_clip_ratio=2.0)
    o = orc.vtrace(t, b, a, v, r, w, coef=coef, **hp)

    # V-trace targets are stop-gradient in origin (vtrace.py:63-79): freeze them at the base point
    def total(tt, vv, frozen):
        logp = tt - np.log(np.exp(tt - tt.max(-1, keepdims=True)).sum(-1, keepdims=True)) - tt.max(-1, keepdims=True)
        p = np.exp(logp)
        lp_a = np.take_along_axis(logp, a[..., None], -1)[..., 0]
        ent = -(p * logp).sum(-1)
        pol = -(lp_a * frozen["adv"] * w).mean()
        val = 0.5 * (((vv[:-1] - frozen["ret"]) ** 2) * w).mean()
        return coef[0] * pol + coef[1] * val + coef[2] * (ent * w).mean()

    # recover the frozen quantities from the oracle's gradient of the value head: d val / d v = w (v - ret) / n
    n = T * B
    gv_val = o["grad_value"][:-1] / coef[1]
    ret = v[:-1] - gv_val * n / np.where(w == 0, 1, w)
    # policy advantage: d pol / d logit_a-part is not separable from the entropy part, so take it from a
    # second oracle call with the entropy coefficient off
    o_pol = orc.vtrace(t, b, a, v, r, w, coef=[1.0, 0.0, 0.0], **hp)
    logp = t - t.max(-1, keepdims=True)
    logp = logp - np.log(np.exp(logp).sum(-1, keepdims=True))
    p = np.exp(logp)
    onehot = np.zeros_like(p)
    np.put_along_axis(onehot, a[..., None], 1.0, -1)
    # grad = -(adv w / n) (onehot - p)  ->  read adv off the action's own slot
    ga = np.take_along_axis(o_pol["grad_target_output"], a[..., None], -1)[..., 0]
    pa = np.take_along_axis(p, a[..., None], -1)[..., 0]
    adv = -ga * n / (np.where(w == 0, 1, w) * (1 - pa))
    frozen = dict(adv=adv, ret=ret)
    fd_check(lambda: total(t, v, frozen), t, o["grad_target_output"], 10, g)
    fd_check(lambda: total(t, v, frozen), v, o["grad_value"], 8, g)


def test_upgo_gradient_fd():
    g = rng(4)
    T, B, N = 6, 3, 5
    t = g.standard_normal((T, B, N))
    rho, a = g.random((T, B)) * 2, g.integers(0, N, (T, B)).astype(np.int64)
    rew, bv = g.standard_normal((T, B)), g.standard_normal((T + 1, B))
    o = orc.upgo(t, rho, a, rew, bv, 1.0)
    adv = rho * (o["ret"] - bv[:-1])  # upgo.py:57-60: returns and advantages carry no gradient

    def f():
        logp = t - t.max(-1, keepdims=True)
        logp = logp - np.log(np.exp(logp).sum(-1, keepdims=True))
        return float(-(np.take_along_axis(logp, a[..., None], -1)[..., 0] * adv).mean())

    assert abs(f() - o["loss"]) <= 1e-12 * max(1.0, abs(o["loss"]))
    fd_check(f, t, o["grad_target_output"], 12, g)


def test_ppo_gradients_fd():
    g = rng(5)
    B, N = 9, 5
    lo = g.standard_normal((B, N))
    ln = lo + 0.3 * g.standard_normal((B, N))
    a = g.integers(0, N, (B, )).astype(np.int64)
    vn, vo, adv, ret, w = (g.standard_normal(B) for _ in range(5))
    w = np.abs(w)
    coef = [1.0, 0.5, -0.01]

    def f():
        o = orc.ppo(ln, lo, a, vn, vo, adv, ret, w, 0.2, True, 3.0, coef, want_grad=False)
        return coef[0] * o["policy_loss"] + coef[1] * o["value_loss"] + coef[2] * o["entropy_loss"]

    o = orc.ppo(ln, lo, a, vn, vo, adv, ret, w, 0.2, True, 3.0, coef)
    fd_check(f, ln, o["grad_logits_new"], 15, g)
    fd_check(f, vn, o["grad_value_new"], 9, g)


@pytest.mark.parametrize("rescale", [False, True])
def test_q_nstep_gradient_fd(rescale):
    g = rng(6)
    T, B, N = 4, 7, 5
    q, nq = g.standard_normal((B, N)), g.standard_normal((B, N))
    a, an = g.integers(0, N, (B, )).astype(np.int64), g.integers(0, N, (B, )).astype(np.int64)
    r, d, w = g.standard_normal((T, B)), (g.random(B) < 0.3).astype(F), g.random(B)
    o = orc.q_nstep_td(q, nq, a, an, r, d, w, 0.95, rescale, 1.0)
    fd_check(lambda: float(orc.q_nstep_td(q, nq, a, an, r, d, w, 0.95, rescale, 1.0, want_grad=False)["loss"]), q,
             o["grad_q"], 15, g)


def test_dist_nstep_gradient_fd():
    g = rng(7)
    T, B, N, A = 3, 5, 3, 11
    e = np.exp(g.standard_normal((B, N, A)))
    dist = e / e.sum(-1, keepdims=True)
    e2 = np.exp(g.standard_normal((B, N, A)))
    nd = e2 / e2.sum(-1, keepdims=True)
    a, an = g.integers(0, N, (B, )).astype(np.int64), g.integers(0, N, (B, )).astype(np.int64)
    r, d, w = g.standard_normal((T, B)), (g.random(B) < 0.3).astype(F), g.random(B)
    args = (dist, nd, a, an, r, d, w, 0.95, -3.0, 3.0)
    o = orc.dist_nstep_td(*args, 1.0)
    fd_check(lambda: float(orc.dist_nstep_td(*args, 1.0, want_grad=False)["loss"]), dist, o["grad_dist"], 20, g,
             eps=1e-7, tol=5e-5)


def test_qrdqn_iqn_gradients_fd():
    g = rng(8)
    tau, tau_p, T, B, N = 7, 6, 3, 4, 3
    a, an = g.integers(0, N, (B, )).astype(np.int64), g.integers(0, N, (B, )).astype(np.int64)
    r, d, w = g.standard_normal((T, B)), (g.random(B) < 0.3).astype(F), g.random(B)
    q, nq = g.standard_normal((B, N, tau)), g.standard_normal((B, N, tau))
    o = orc.qrdqn_nstep_td(q, nq, a, an, r, d, w, None, 0.95, 1.0)
    fd_check(lambda: float(orc.qrdqn_nstep_td(q, nq, a, an, r, d, w, None, 0.95, 1.0, want_grad=False)["loss"]), q,
             o["grad_q"], 20, g)
    qi, nqi, rq = g.standard_normal((tau, B, N)), g.standard_normal((tau_p, B, N)), g.random((tau, B))
    o = orc.iqn_nstep_td(qi, nqi, a, an, r, d, rq, w, None, 0.95, 0.9, 1.0)
    fd_check(lambda: float(orc.iqn_nstep_td(qi, nqi, a, an, r, d, rq, w, None, 0.95, 0.9, 1.0,
                                            want_grad=False)["loss"]), qi, o["grad_q"], 20, g)


def test_losses_compose_over_batch_shards():
    """mean-type losses: sum_k (B_k / B) * loss(shard k) == loss(whole) -- the rule behind global_B + all-reduce"""
    g = rng(9)
    T, B, N = 6, 10, 4
    t, b = g.standard_normal((T, B, N)), g.standard_normal((T, B, N))
    a = g.integers(0, N, (T, B)).astype(np.int64)
    v, r, w = g.standard_normal((T + 1, B)), g.standard_normal((T, B)), g.random((T, B))
    whole = orc.vtrace(t, b, a, v, r, w, want_grad=False)
    parts = [(0, 4), (4, 10)]
    for key in ("policy_loss", "value_loss", "entropy_loss"):
        s = sum((b1 - b0) / B * orc.vtrace(np.ascontiguousarray(t[:, b0:b1]), np.ascontiguousarray(b[:, b0:b1]),
                                           np.ascontiguousarray(a[:, b0:b1]), np.ascontiguousarray(v[:, b0:b1]),
                                           np.ascontiguousarray(r[:, b0:b1]), np.ascontiguousarray(w[:, b0:b1]),
                                           want_grad=False)[key] for b0, b1 in parts)
        assert abs(s - whole[key]) <= 1e-12 * max(1.0, abs(whole[key]))
