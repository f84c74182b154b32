"""-m gpu: the reference's native boundary on the B200 kernels.

(a) `test_names_*`: each of the 19 hot-path `hpc_rl_utils` names called directly with tensor lists in the reference
    launchers' positional order (src/rl_utils/*.cu `index++`; buffers shaped as the reference modules register them),
    results compared with the oracle.
(b) `test_reference_wrappers_*`: the UNMODIFIED reference wrappers (staged copy of hpc_rll/rl_utils/*.py) running on the
    shim, forward + autograd backward, compared with the oracle -- what a DI-hpc user gets by swapping the extension.
"""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests import _refwrap
from tests._gpu import dev, host, need_cuda, rel_err, rng

pytestmark = pytest.mark.gpu
TOL = 1e-5


def close(got, want, what):
    e = rel_err(got, want)
    assert e <= TOL, "%s: rel err %.3e" % (what, e)


def z(*s):
    return torch.zeros(*s, device="cuda")


def one():
    return torch.ones(1, device="cuda")


def scan_inputs(g, T, B):
    return (g.standard_normal((T + 1, B), dtype=np.float32), g.standard_normal((T, B), dtype=np.float32),
            g.random((T, B), dtype=np.float32))


def nstep_inputs(g, T, B, N):
    return dict(action=g.integers(0, N, (B, )).astype(np.int64), next_n_action=g.integers(0, N, (B, )).astype(np.int64),
                reward=g.standard_normal((T, B), dtype=np.float32), done=(g.random(B) < 0.3).astype(np.float32),
                weight=g.random(B).astype(np.float32))


# ------------------------------------------------------------------------------------------------ (a) direct calls
def test_names_gae_and_td_lambda():
    need_cuda()
    import hpc_rl_utils as H
    T, B = 33, 260
    value, reward, weight = scan_inputs(rng(1), T, B)
    adv = z(T, B)
    H.GaeForward([dev(value), dev(reward)], [adv], 0.99, 0.97)
    close(host(adv), orc.gae_forward(value, reward, np.float32(0.99), np.float32(0.97)), "adv")
    loss, gbuf, gval = z(1), z(T, B), z(T + 1, B)
    H.TdLambdaForward([dev(value), dev(reward), dev(weight)], [loss, gbuf], 0.9, 0.8)
    H.TdLambdaBackward([one() * 1.7, gbuf], [gval])
    o = orc.td_lambda(value, reward, weight, np.float32(0.9), np.float32(0.8), 1.7)
    close(float(loss.item()), o["loss"], "loss")
    close(host(gval) * T * B, o["grad_value"] * T * B, "grad_value")
    # the reference module's default weight is ones(B) (rl_utils/td.py:160): broadcast, not read out of bounds
    H.TdLambdaForward([dev(value), dev(reward), torch.ones(B, device="cuda")], [loss, gbuf], 0.9, 0.8)
    close(float(loss.item()), orc.td_lambda(value, reward, None, np.float32(0.9), np.float32(0.8))["loss"], "loss (B,)")


@pytest.mark.parametrize("rescale", [False, True])
def test_names_q_nstep(rescale):
    need_cuda()
    import hpc_rl_utils as H
    T, B, N = 5, 300, 6
    g = rng(2 + rescale)
    inp = nstep_inputs(g, T, B, N)
    q, nq = g.standard_normal((B, N), dtype=np.float32), g.standard_normal((B, N), dtype=np.float32)
    td, loss, gbuf, gq = z(B), z(1), z(B), z(B, N)
    fwd, bwd = (H.QNStepTdRescaleForward, H.QNStepTdRescaleBackward) if rescale else (H.QNStepTdForward, H.QNStepTdBackward)
    fwd([dev(q), dev(nq), dev(inp["action"]), dev(inp["next_n_action"]), dev(inp["reward"]), dev(inp["done"]),
         dev(inp["weight"])], [td, loss, gbuf], 0.95)
    bwd([one() * 1.3, gbuf, dev(inp["action"])], [gq])
    o = orc.q_nstep_td(q, nq, inp["action"], inp["next_n_action"], inp["reward"], inp["done"], inp["weight"],
                       np.float32(0.95), rescale, 1.3)
    close(float(loss.item()), o["loss"], "loss")
    close(host(td), o["td_error_per_sample"], "td")
    close(host(gq) * B, o["grad_q"] * B, "grad_q")
    assert np.array_equal(host(gq) != 0, o["grad_q"] != 0)


def test_names_dist_nstep():
    need_cuda()
    import hpc_rl_utils as H
    T, B, N, A = 3, 200, 4, 51
    g = rng(4)
    inp = nstep_inputs(g, T, B, N)
    sm = lambda x: (np.exp(x) / np.exp(x).sum(-1, keepdims=True)).astype(np.float32)  # noqa: E731
    dist, ndist = sm(g.standard_normal((B, N, A))), sm(g.standard_normal((B, N, A)))
    td, loss, buf, gd = z(B), z(1), z(B + B * A), z(B, N, A)
    H.DistNStepTdForward([dev(dist), dev(ndist), dev(inp["action"]), dev(inp["next_n_action"]), dev(inp["reward"]),
                          dev(inp["done"]), dev(inp["weight"])], [td, loss, buf], 0.95, -10.0, 10.0)
    H.DistNStepTdBackward([one() * 0.9, buf, dev(inp["action"])], [gd])
    o = orc.dist_nstep_td(dist, ndist, inp["action"], inp["next_n_action"], inp["reward"], inp["done"], inp["weight"],
                          np.float32(0.95), -10.0, 10.0, 0.9)
    close(float(loss.item()), o["loss"], "loss")
    close(host(td), o["td_error_per_sample"], "td")
    close(host(gd) * B, o["grad_dist"] * B, "grad_dist")


def test_names_qrdqn_and_iqn():
    need_cuda()
    import hpc_rl_utils as H
    tau, tau_p, T, B, N = 39, 34, 10, 89, 7
    g = rng(5)
    inp = nstep_inputs(g, T, B, N)
    vg = g.random(B).astype(np.float32)
    q, nq = g.standard_normal((B, N, tau), dtype=np.float32), g.standard_normal((B, N, tau), dtype=np.float32)
    loss, td, gbuf, gq = z(1), z(B), z(B, tau), z(B, N, tau)  # grad_buf (B,tau): the reference's own (too small) shape
    H.QRDQNNStepTDErrorForward([dev(q), dev(nq), dev(inp["action"]), dev(inp["next_n_action"]), dev(inp["reward"]),
                                dev(inp["done"]), dev(inp["weight"]), dev(vg)],
                               [loss, td, z(B, tau, tau), z(B, tau, tau), gbuf], 0.95)
    H.QRDQNNStepTDErrorBackward([one() * 1.1, gbuf, dev(inp["weight"]), dev(inp["action"])], [gq])
    o = orc.qrdqn_nstep_td(q, nq, inp["action"], inp["next_n_action"], inp["reward"], inp["done"], inp["weight"], vg,
                           np.float32(0.95), 1.1)
    close(float(loss.item()), o["loss"], "qrdqn loss")
    close(host(td), o["td_error_per_sample"], "qrdqn td")
    close(host(gq) * B, o["grad_q"] * B, "qrdqn grad_q")
    qi, nqi = g.standard_normal((tau, B, N), dtype=np.float32), g.standard_normal((tau_p, B, N), dtype=np.float32)
    rq = g.random((tau, B)).astype(np.float32)
    gbuf, gqi = z(B, tau_p, tau), z(tau, B, N)
    H.IQNNStepTDErrorForward([dev(qi), dev(nqi), dev(inp["action"]), dev(inp["next_n_action"]), dev(inp["reward"]),
                              dev(inp["done"]), dev(rq), dev(inp["weight"]), dev(vg)],
                             [loss, td, z(B, tau_p, tau), z(B, tau_p, tau), gbuf], 0.95, 0.9)
    H.IQNNStepTDErrorBackward([one() * 0.8, gbuf, dev(inp["weight"]), dev(inp["action"])], [gqi])
    o = orc.iqn_nstep_td(qi, nqi, inp["action"], inp["next_n_action"], inp["reward"], inp["done"], rq, inp["weight"], vg,
                         np.float32(0.95), np.float32(0.9), 0.8)
    close(float(loss.item()), o["loss"], "iqn loss")
    close(host(td), o["td_error_per_sample"], "iqn td")
    close(host(gqi) * B, o["grad_q"] * B, "iqn grad_q")


def test_names_upgo_vtrace_ppo():
    need_cuda()
    import hpc_rl_utils as H
    T, B, N = 12, 70, 18
    g = rng(6)
    value, reward, weight = scan_inputs(g, T, B)
    tgt = (g.standard_normal((T, B, N)) * 1.5).astype(np.float32)
    beh = (g.standard_normal((T, B, N)) * 1.5).astype(np.float32)
    act = g.integers(0, N, (T, B)).astype(np.int64)
    rhos = (g.random((T, B)) * 2).astype(np.float32)
    # ---- upgo
    advb, metric, loss, gbuf, gt = z(T, B), z(T, B), z(1), z(T, B, N), z(T, B, N)
    H.UpgoForward([dev(tgt), dev(rhos), dev(act), dev(reward), dev(value)], [advb, metric, loss, gbuf])
    H.UpgoBackward([one() * -0.7, gbuf, advb], [gt])
    o = orc.upgo(tgt, rhos, act, reward, value, -0.7)
    close(float(loss.item()), o["loss"], "upgo loss")
    close(host(gt) * T * B, o["grad_target_output"] * T * B, "upgo grad")
    # ---- vtrace: 12 forward outputs in vtrace.cu:26-37 order, 11 backward inputs in vtrace.cu:93-103 order
    outs = [z(T, B), z(T, B), z(T, B, N), z(T, B, N), z(T, B, N), z(T, B), z(T, B), z(T, B), z(T, B), z(1), z(1), z(1)]
    H.VTraceForward([dev(tgt), dev(beh), dev(act), dev(value), dev(reward), dev(weight)], outs, 0.9, 0.8, 1.5, 0.9, 2.0)
    gv, gtt = z(T + 1, B), z(T, B, N)
    coef = [1.0, 0.5, -0.25]
    H.VTraceBackward([one() * coef[0], one() * coef[1], one() * coef[2], dev(value), dev(act), dev(weight), outs[7],
                      outs[8], outs[2], outs[3], outs[4]], [gv, gtt])
    o = orc.vtrace(tgt, beh, act, value, reward, weight, np.float32(0.9), np.float32(0.8), np.float32(1.5),
                   np.float32(0.9), np.float32(2.0), coef)
    for k, nm in enumerate(("policy_loss", "value_loss", "entropy_loss")):
        close(float(outs[9 + k].item()), o[nm], "vtrace " + nm)
    close(host(gtt) * T * B, o["grad_target_output"] * T * B, "vtrace grad_target_output")
    close(host(gv) * T * B, o["grad_value"] * T * B, "vtrace grad_value")
    # ---- ppo: 14 outputs in ppo.cu:26-39 order, 9 backward inputs in ppo.cu:80-88 order
    Bp = 257
    lo = g.standard_normal((Bp, N)).astype(np.float32)
    ln = (lo + 0.3 * g.standard_normal((Bp, N))).astype(np.float32)
    a = g.integers(0, N, (Bp, )).astype(np.int64)
    vn, vo, adv, ret, w = (g.standard_normal(Bp).astype(np.float32) for _ in range(5))
    w = np.abs(w)
    outs = [z(Bp), z(Bp), z(Bp, N), z(Bp, N), z(Bp, N), z(Bp), z(Bp), z(Bp), z(Bp), z(1), z(1), z(1), z(1), z(1)]
    for dual_ref, dual in ((3.0, 3.0), (0.0, None)):  # the reference wrapper encodes None as 0.0 (rl_utils/ppo.py:136)
        H.PPOForward([dev(ln), dev(lo), dev(a), dev(vn), dev(vo), dev(adv), dev(ret), dev(w)], outs, True, 0.2, dual_ref)
        gval, glog = z(Bp), z(Bp, N)
        c3 = [1.0, 0.5, -0.01]
        H.PPOBackward([one() * c3[0], one() * c3[1], one() * c3[2], outs[6], outs[7], outs[8], outs[2], outs[3], outs[4]],
                      [gval, glog])
        o = orc.ppo(ln, lo, a, vn, vo, adv, ret, w, np.float32(0.2), True, dual, c3)
        for k, nm in enumerate(("policy_loss", "value_loss", "entropy_loss", "approx_kl")):
            close(float(outs[9 + k].item()), o[nm], "ppo " + nm)
        assert abs(float(outs[13].item()) - o["clipfrac"]) <= 2.0 / Bp
        close(host(glog) * Bp, o["grad_logits_new"] * Bp, "ppo grad_logits_new")
        close(host(gval) * Bp, o["grad_value_new"] * Bp, "ppo grad_value_new")


# ------------------------------------------------------------------------------------------------ (b) reference wrappers
def _ref(mod):
    need_cuda()
    m = _refwrap.load(mod)
    if m is None:
        pytest.skip("reference wrappers not staged (tools/stage_ref_wrappers.sh)")
    return m


def test_reference_wrappers_gae_tdlambda():
    g = rng(11)
    T, B = 1024, 64  # tests/test_gae.py:10-11, tests/test_tdlambda.py:10-11
    value, reward, weight = scan_inputs(g, T, B)
    adv = _ref("gae").GAE(T, B).cuda()(dev(value), dev(reward))
    close(host(adv), orc.gae_forward(value, reward, np.float32(0.99), np.float32(0.97)), "adv")
    v = dev(value).requires_grad_(True)
    loss = _ref("td").TDLambda(T, B).cuda()(v, dev(reward), dev(weight))
    loss.backward()
    o = orc.td_lambda(value, reward, weight, np.float32(0.9), np.float32(0.8))
    close(float(loss.item()), o["loss"], "loss")
    close(host(v.grad) * T * B, o["grad_value"] * T * B, "grad_value")


def test_reference_wrappers_vtrace_upgo_ppo():
    g = rng(12)
    T, B, N = 64, 48, 16
    value, reward, weight = scan_inputs(g, T, B)
    tgt = (g.standard_normal((T, B, N)) * 1.5).astype(np.float32)
    beh = (g.standard_normal((T, B, N)) * 1.5).astype(np.float32)
    act = g.integers(0, N, (T, B)).astype(np.int64)
    t, v = dev(tgt).requires_grad_(True), dev(value).requires_grad_(True)
    l = _ref("vtrace").VTrace(T, B, N).cuda()(t, dev(beh), dev(act), v, dev(reward))  # weight None -> module ones
    coef = [1.0, 0.5, -0.25]
    (coef[0] * l.policy_loss + coef[1] * l.value_loss + coef[2] * l.entropy_loss).sum().backward()
    o = orc.vtrace(tgt, beh, act, value, reward, None, coef=coef)
    for got, nm in zip(l, ("policy_loss", "value_loss", "entropy_loss")):
        close(float(got.item()), o[nm], nm)
    close(host(t.grad) * T * B, o["grad_target_output"] * T * B, "vtrace grad_target_output")
    close(host(v.grad) * T * B, o["grad_value"] * T * B, "vtrace grad_value")
    rhos = (g.random((T, B)) * 2).astype(np.float32)
    t2 = dev(tgt).requires_grad_(True)
    loss = _ref("upgo").UPGO(T, B, N).cuda()(t2, dev(rhos), dev(act), dev(reward), dev(value))
    loss.sum().backward()
    o = orc.upgo(tgt, rhos, act, reward, value)
    close(float(loss.item()), o["loss"], "upgo loss")
    close(host(t2.grad) * T * B, o["grad_target_output"] * T * B, "upgo grad")
    Bp = 128  # tests/test_ppo.py:11-12
    lo = g.standard_normal((Bp, N)).astype(np.float32)
    ln = (lo + 0.3 * g.standard_normal((Bp, N))).astype(np.float32)
    a = g.integers(0, N, (Bp, )).astype(np.int64)
    vn, vo, adv, ret = (g.standard_normal(Bp).astype(np.float32) for _ in range(4))
    lnt, vnt = dev(ln).requires_grad_(True), dev(vn).requires_grad_(True)
    loss, info = _ref("ppo").PPO(Bp, N).cuda()(lnt, dev(lo), dev(a), vnt, dev(vo), dev(adv), dev(ret), None, 0.2, True, None)
    (loss.policy_loss + 0.5 * loss.value_loss - 0.01 * loss.entropy_loss).sum().backward()
    o = orc.ppo(ln, lo, a, vn, vo, adv, ret, None, 0.2, True, None, [1.0, 0.5, -0.01])
    for got, nm in zip(loss, ("policy_loss", "value_loss", "entropy_loss")):
        close(float(got.item()), o[nm], "ppo " + nm)
    close(info.approx_kl, o["approx_kl"], "approx_kl")
    close(host(lnt.grad) * Bp, o["grad_logits_new"] * Bp, "ppo grad_logits_new")
    close(host(vnt.grad) * Bp, o["grad_value_new"] * Bp, "ppo grad_value_new")


def test_reference_wrappers_nstep_family():
    g = rng(13)
    td = _ref("td")
    T, B, N = 10, 89, 7
    inp = nstep_inputs(g, T, B, N)
    dv = {k: dev(v) for k, v in inp.items()}
    for rescale, cls in ((False, td.QNStepTD), (True, td.QNStepTDRescale)):
        q, nq = g.standard_normal((B, N), dtype=np.float32), g.standard_normal((B, N), dtype=np.float32)
        qt = dev(q).requires_grad_(True)
        loss, tde = cls(T, B, N).cuda()(qt, dev(nq), dv["action"], dv["next_n_action"], dv["reward"], dv["done"],
                                        dv["weight"], 0.95)
        loss.sum().backward()
        o = orc.q_nstep_td(q, nq, inp["action"], inp["next_n_action"], inp["reward"], inp["done"], inp["weight"],
                           np.float32(0.95), rescale)
        close(float(loss.item()), o["loss"], "q loss")
        close(host(tde), o["td_error_per_sample"], "q td")
        close(host(qt.grad) * B, o["grad_q"] * B, "grad_q")
    A = 51
    sm = lambda x: (np.exp(x) / np.exp(x).sum(-1, keepdims=True)).astype(np.float32)  # noqa: E731
    dist, ndist = sm(g.standard_normal((B, N, A))), sm(g.standard_normal((B, N, A)))
    dt = dev(dist).requires_grad_(True)
    loss, tde = td.DistNStepTD(T, B, N, A).cuda()(dt, dev(ndist), dv["action"], dv["next_n_action"], dv["reward"],
                                                  dv["done"], dv["weight"], 0.95, -10.0, 10.0)
    loss.sum().backward()
    o = orc.dist_nstep_td(dist, ndist, inp["action"], inp["next_n_action"], inp["reward"], inp["done"], inp["weight"],
                          np.float32(0.95), -10.0, 10.0)
    close(float(loss.item()), o["loss"], "dist loss")
    close(host(dt.grad) * B, o["grad_dist"] * B, "grad_dist")
    tau, tau_p = 39, 34  # tests/test_qrdqn_nstep_td_error.py:10-14
    q, nq = g.standard_normal((B, N, tau), dtype=np.float32), g.standard_normal((B, N, tau), dtype=np.float32)
    qt = dev(q).requires_grad_(True)
    loss, tde = td.QRDQNNStepTDError(tau, T, B, N).cuda()(qt, dev(nq), dv["action"], dv["next_n_action"], dv["reward"],
                                                          dv["done"], 0.95, dv["weight"])
    loss.sum().backward()
    o = orc.qrdqn_nstep_td(q, nq, inp["action"], inp["next_n_action"], inp["reward"], inp["done"], inp["weight"], None,
                           np.float32(0.95))
    close(float(loss.item()), o["loss"], "qrdqn loss")
    close(host(tde), o["td_error_per_sample"], "qrdqn td")
    close(host(qt.grad) * B, o["grad_q"] * B, "qrdqn grad_q")
    qi, nqi = g.standard_normal((tau, B, N), dtype=np.float32), g.standard_normal((tau_p, B, N), dtype=np.float32)
    rq = g.random((tau, B)).astype(np.float32)
    qit = dev(qi).requires_grad_(True)
    loss, tde = td.IQNNStepTDError(tau, tau_p, T, B, N).cuda()(qit, dev(nqi), dv["action"], dv["next_n_action"],
                                                               dv["reward"], dv["done"], dev(rq), 0.95, 0.9, dv["weight"])
    loss.sum().backward()
    o = orc.iqn_nstep_td(qi, nqi, inp["action"], inp["next_n_action"], inp["reward"], inp["done"], rq, inp["weight"], None,
                         np.float32(0.95), np.float32(0.9))
    close(float(loss.item()), o["loss"], "iqn loss")
    close(host(tde), o["td_error_per_sample"], "iqn td")
    close(host(qit.grad) * B, o["grad_q"] * B, "iqn grad_q")
