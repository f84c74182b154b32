"""Per-op throughput of every op of the path at the BASELINE.json configs (C1..C3) through the drop-in
modules (autograd forward + backward), CUDA-event timed, with the algorithmic-byte roofline fraction
(SURVEY.md 8d formulas).  One JSON line per op; a summary goes to profiles/.

    python tools/bench_ops.py [--quick] [--ops gae,vtrace,...]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from di_hpc_b200 import _abi  # noqa: E402
from hpc_rll.rl_utils.gae import GAE  # noqa: E402
from hpc_rll.rl_utils.ppo import PPO  # noqa: E402
from hpc_rll.rl_utils.td import (DistNStepTD, IQNNStepTDError, QNStepTD, QNStepTDRescale, QRDQNNStepTDError,  # noqa: E402
                                 TDLambda)
from hpc_rll.rl_utils.upgo import UPGO  # noqa: E402
from hpc_rll.rl_utils.vtrace import VTrace  # noqa: E402

DEV = "cuda"
ONE = None  # ones(1) on the device, set in main()
LAST_PIPELINED = 0.0


def peak():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"])
    except Exception:
        return 6650.0


def timed(fwd, bwd, iters, warm=3):
    """median ms of fwd and bwd separately (events on the current stream)."""
    for _ in range(warm):
        out = fwd()
        bwd(out)
    torch.cuda.synchronize()
    tf, tb = [], []
    for _ in range(iters):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        n0 = _abi.launch_count()
        e[0].record()
        out = fwd()
        e[1].record()
        bwd(out)
        e[2].record()
        torch.cuda.synchronize()
        launches = _abi.launch_count() - n0
        tf.append(e[0].elapsed_time(e[1]))
        tb.append(e[1].elapsed_time(e[2]))
    tf.sort()
    tb.sort()
    # sustained rate: `iters` fwd+bwd pairs queued back to back between two events, so the host-side launch
    # work of call i+1 overlaps the kernels of call i (what a training loop sees); ops whose API returns Python
    # scalars (PPO info) still synchronise once per call
    global LAST_PIPELINED
    best = None
    for _ in range(3):  # best of three bursts (the first one after a sync-per-call phase runs at idle clocks)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            bwd(fwd())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        best = ms if best is None else min(best, ms)
    LAST_PIPELINED = best
    return tf[len(tf) // 2], tb[len(tb) // 2], launches


def rnd(*shape):
    return torch.randn(*shape, device=DEV)


def bench_gae(T, B, iters):
    v, r, g = rnd(T + 1, B).requires_grad_(True), rnd(T, B).requires_grad_(True), rnd(T, B)
    m = GAE(T, B)
    f, b, L = timed(lambda: m(v, r), lambda o: torch.autograd.grad(o, [v, r], grad_outputs=g), iters)
    return dict(op="gae", shape=dict(T=T, B=B), units=T * B, unit="steps", alg_bytes=24 * T * B + 8 * B, fwd_ms=f,
                bwd_ms=b, launches=L)


def bench_td_lambda(T, B, iters, use_w=True):
    v, r = rnd(T + 1, B).requires_grad_(True), rnd(T, B)
    w = torch.rand(T, B, device=DEV) if use_w else None
    m = TDLambda(T, B)
    f, b, L = timed(lambda: m(v, r, w), lambda o: torch.autograd.grad(o, [v], grad_outputs=ONE), iters)
    return dict(op="td_lambda" + ("_w" if use_w else ""), shape=dict(T=T, B=B), units=T * B, unit="steps",
                alg_bytes=(24 if use_w else 20) * T * B, fwd_ms=f, bwd_ms=b, launches=L)


def bench_vtrace(T, B, N, iters, use_w=False):
    t = rnd(T, B, N).requires_grad_(True)
    bh = rnd(T, B, N)
    a = torch.randint(0, N, (T, B), device=DEV)
    v, r = rnd(T + 1, B).requires_grad_(True), rnd(T, B)
    w = torch.rand(T, B, device=DEV) if use_w else None
    m = VTrace(T, B, N)

    def fwd():
        l = m(t, bh, a, v, r, w)
        return l.policy_loss + l.value_loss + l.entropy_loss

    f, b, L = timed(fwd, lambda o: torch.autograd.grad(o, [t, v], grad_outputs=ONE), iters)
    return dict(op="vtrace", shape=dict(T=T, B=B, N=N), units=T * B, unit="steps",
                alg_bytes=(16 * N + 52 + (8 if use_w else 0)) * T * B, fwd_ms=f, bwd_ms=b, launches=L)


def bench_upgo(T, B, N, iters):
    t = rnd(T, B, N).requires_grad_(True)
    rho = torch.rand(T, B, device=DEV) * 2
    a = torch.randint(0, N, (T, B), device=DEV)
    r, v = rnd(T, B), rnd(T + 1, B)
    m = UPGO(T, B, N)
    f, b, L = timed(lambda: m(t, rho, a, r, v), lambda o: torch.autograd.grad(o, [t], grad_outputs=ONE), iters)
    return dict(op="upgo", shape=dict(T=T, B=B, N=N), units=T * B, unit="steps", alg_bytes=(12 * N + 36) * T * B,
                fwd_ms=f, bwd_ms=b, launches=L)


def bench_ppo(B, N, iters):
    lo = rnd(B, N)
    ln = (lo + 0.3 * rnd(B, N)).requires_grad_(True)
    a = torch.randint(0, N, (B, ), device=DEV)
    vn = rnd(B).requires_grad_(True)
    vo, adv, ret = rnd(B), rnd(B), rnd(B)
    m = PPO(B, N)
    m.lazy_info = True  # approx_kl / clipfrac as LazyScalars: no blocking D2H copy per call

    def fwd():
        l, _ = m(ln, lo, a, vn, vo, adv, ret)
        return l.policy_loss + l.value_loss + l.entropy_loss

    f, b, L = timed(fwd, lambda o: torch.autograd.grad(o, [ln, vn], grad_outputs=ONE), iters)
    return dict(op="ppo", shape=dict(B=B, N=N), units=B, unit="samples", alg_bytes=(16 * N + 50) * B, fwd_ms=f,
                bwd_ms=b, launches=L)


def bench_chain(T, B, N, iters, fused):
    """GAE -> (adv - mean) / (std + 1e-8) -> PPO on the flattened (T*B,) batch.  fused: moments
    ride in the GAE scan and the normalisation happens inside the PPO kernel; unfused: same kernels with the
    normalisation done by three PyTorch passes over adv (mean, std, normalise)."""
    from hpc_rll.rl_utils.gae import gae_with_adv_stats
    R = T * B
    v, r = rnd(T + 1, B), rnd(T, B)
    lo = rnd(R, N)
    ln = (lo + 0.3 * rnd(R, N)).requires_grad_(True)
    a = torch.randint(0, N, (R, ), device=DEV)
    vn = rnd(R).requires_grad_(True)
    vo, ret = rnd(R), rnd(R)
    g, m = GAE(T, B), PPO(R, N)

    def fwd():
        if fused:
            adv, st = gae_with_adv_stats(v, r)
            l, _ = m(ln, lo, a, vn, vo, adv.reshape(-1), ret, adv_stats=st)
        else:
            with torch.no_grad():
                adv = g(v, r)
                adv = ((adv - adv.mean()) / (adv.std() + 1e-8)).reshape(-1)
            l, _ = m(ln, lo, a, vn, vo, adv, ret)
        return l.policy_loss + l.value_loss + l.entropy_loss

    f, b, L = timed(fwd, lambda o: torch.autograd.grad(o, [ln, vn], grad_outputs=ONE), iters)
    return dict(op="gae_norm_ppo_" + ("fused" if fused else "unfused"), shape=dict(T=T, B=B, N=N), units=R,
                unit="steps", alg_bytes=(12 + 16 * N + 50) * R, fwd_ms=f, bwd_ms=b, launches=L)


def nstep_common(T, B, N):
    return (torch.randint(0, N, (B, ), device=DEV), torch.randint(0, N, (B, ), device=DEV), rnd(T, B),
            (torch.rand(B, device=DEV) < 0.1).float())


def bench_q(T, B, N, iters, rescale=False):
    q, nq = rnd(B, N).requires_grad_(True), rnd(B, N)
    a, an, r, d = nstep_common(T, B, N)
    m = (QNStepTDRescale if rescale else QNStepTD)(T, B, N)
    f, b, L = timed(lambda: m(q, nq, a, an, r, d, None, 0.99)[0],
                    lambda o: torch.autograd.grad(o, [q], grad_outputs=ONE), iters)
    return dict(op="q_nstep_td" + ("_rescale" if rescale else ""), shape=dict(T=T, B=B, N=N), units=B, unit="samples",
                alg_bytes=(4 * T + 4 * N + 100) * B, fwd_ms=f, bwd_ms=b, launches=L)


def bench_dist(T, B, N, n_atom, iters):
    d0 = torch.softmax(rnd(B, N, n_atom), -1).requires_grad_(True)
    d1 = torch.softmax(rnd(B, N, n_atom), -1)
    a, an, r, d = nstep_common(T, B, N)
    m = DistNStepTD(T, B, N, n_atom)
    f, b, L = timed(lambda: m(d0, d1, a, an, r, d, None, 0.99, -10.0, 10.0)[0],
                    lambda o: torch.autograd.grad(o, [d0], grad_outputs=ONE), iters)
    return dict(op="dist_nstep_td", shape=dict(T=T, B=B, N=N, n_atom=n_atom), units=B, unit="samples",
                alg_bytes=(8 * n_atom + 4 * T + 30 + 4 * N * n_atom + 4 * n_atom) * B, fwd_ms=f, bwd_ms=b, launches=L)


def bench_qrdqn(tau, T, B, N, iters):
    q, nq = rnd(B, N, tau).requires_grad_(True), rnd(B, N, tau)
    a, an, r, d = nstep_common(T, B, N)
    m = QRDQNNStepTDError(tau, T, B, N)
    f, b, L = timed(lambda: m(q, nq, a, an, r, d, 0.99)[0], lambda o: torch.autograd.grad(o, [q], grad_outputs=ONE),
                    iters)
    return dict(op="qrdqn_nstep_td", shape=dict(tau=tau, T=T, B=B, N=N), units=B, unit="samples",
                alg_bytes=(8 * tau + 4 * N * tau + 4 * T + 32) * B, pair_flops=tau * tau * 12 * B, fwd_ms=f, bwd_ms=b,
                launches=L)


def bench_iqn(tau, tau_p, T, B, N, iters):
    q, nq = rnd(tau, B, N).requires_grad_(True), rnd(tau_p, B, N)
    a, an, r, d = nstep_common(T, B, N)
    rq = torch.rand(tau, B, device=DEV)
    m = IQNNStepTDError(tau, tau_p, T, B, N)
    f, b, L = timed(lambda: m(q, nq, a, an, r, d, rq, 0.99, 1.0)[0],
                    lambda o: torch.autograd.grad(o, [q], grad_outputs=ONE), iters)
    # gathers touch one 32 B sector per (quantile, sample): min(4N, 32) B each
    sect = max(32, 4) if N <= 8 else 32
    return dict(op="iqn_nstep_td", shape=dict(tau=tau, tau_p=tau_p, T=T, B=B, N=N), units=B, unit="samples",
                alg_bytes=(sect * (tau + tau_p) + 4 * tau + 4 * N * tau + 4 * T + 50) * B,
                pair_flops=tau * tau_p * 13 * B, fwd_ms=f, bwd_ms=b, launches=L)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--ops", default="all")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    global ONE
    ONE = torch.ones(1, device=DEV)
    it = 5 if args.quick else 20
    sel = None if args.ops == "all" else set(args.ops.split(","))
    pk = peak()
    plan = [
        ("gae", lambda: bench_gae(1024, 65536, it)),
        ("gae_small", lambda: bench_gae(1024, 64, it)),
        ("td_lambda", lambda: bench_td_lambda(1024, 65536, it)),
        ("td_lambda_now", lambda: bench_td_lambda(1024, 65536, it, use_w=False)),
        ("vtrace", lambda: bench_vtrace(512, 32768, 16, it)),
        ("vtrace128", lambda: bench_vtrace(512, 4096, 128, it)),
        ("vtrace6", lambda: bench_vtrace(512, 32768, 6, it)),
        ("vtrace18", lambda: bench_vtrace(512, 32768, 18, it)),
        ("upgo6", lambda: bench_upgo(512, 32768, 6, it)),
        ("ppo6", lambda: bench_ppo(1 << 22, 6, it)),
        ("upgo", lambda: bench_upgo(512, 32768, 16, it)),
        ("upgo128", lambda: bench_upgo(512, 4096, 128, it)),
        ("ppo", lambda: bench_ppo(1 << 22, 16, it)),
        ("chain", lambda: bench_chain(1024, 16384, 16, it, True)),
        ("chain_unfused", lambda: bench_chain(1024, 16384, 16, it, False)),
        ("q", lambda: bench_q(5, 1 << 22, 8, it)),
        ("q_rescale", lambda: bench_q(5, 1 << 22, 8, it, rescale=True)),
        ("dist", lambda: bench_dist(5, 1 << 18, 8, 51, it)),
        ("qrdqn", lambda: bench_qrdqn(64, 5, 1 << 20, 8, it)),
        ("iqn", lambda: bench_iqn(64, 64, 5, 1 << 20, 8, it)),
    ]
    lines = []
    for name, fn in plan:
        if sel is not None and name not in sel:
            continue
        rec = fn()
        tot = rec["fwd_ms"] + rec["bwd_ms"]
        rec["total_ms"] = tot
        rec["throughput"] = rec["units"] / (tot * 1e-3)
        rec["alg_gbs"] = rec["alg_bytes"] / (tot * 1e-3) / 1e9
        rec["hbm_frac_of_measured_peak"] = rec["alg_gbs"] / pk
        rec["pipelined_ms"] = LAST_PIPELINED
        rec["pipelined_hbm_frac"] = rec["alg_bytes"] / (LAST_PIPELINED * 1e-3) / 1e9 / pk
        if "pair_flops" in rec:
            rec["pair_tflops"] = rec["pair_flops"] / (rec["fwd_ms"] * 1e-3) / 1e12
        print(json.dumps(rec), flush=True)
        lines.append(rec)
        torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as f:
            for r in lines:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
