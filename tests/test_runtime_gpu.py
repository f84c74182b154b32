"""Runtime behaviour a drop-in must get right on the GPU: side streams, CUDA-graph capture/replay,
non-contiguous inputs, 64-bit indexing (> 2^31 elements) and shard composition through global_B."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests._golden import rel_err
from tests._gpu import dev, host, need_cuda, rng

pytestmark = pytest.mark.gpu


def test_side_stream_and_noncontiguous_inputs():
    need_cuda()
    from hpc_rll.rl_utils.gae import GAE
    from hpc_rll.rl_utils.td import TDLambda
    g = rng(1)
    T, B = 40, 516
    value = g.standard_normal((T + 1, B), dtype=np.float32)
    reward = g.standard_normal((T, B), dtype=np.float32)
    s = torch.cuda.Stream()
    vt = dev(value.T.copy()).T  # (T+1,B) view with B-major strides: not contiguous
    assert not vt.is_contiguous()
    with torch.cuda.stream(s):
        adv = GAE(T, B)(vt, dev(reward))
        loss = TDLambda(T, B)(vt, dev(reward))
    s.synchronize()
    assert rel_err(host(adv), orc.gae_forward(value, reward)) <= 2e-6  # B <= 4096: T-split path (re-associated)
    assert rel_err(float(loss.item()), orc.td_lambda(value, reward)["loss"]) <= 1e-5


def test_cuda_graph_capture_and_replay():
    need_cuda()
    from di_hpc_b200 import _abi
    g = rng(2)
    T, B = 64, 2048
    value = dev(g.standard_normal((T + 1, B), dtype=np.float32))
    reward = dev(g.standard_normal((T, B), dtype=np.float32))
    gadv = dev(g.standard_normal((T, B), dtype=np.float32))
    adv, gv, gr = torch.empty_like(reward), torch.empty_like(value), torch.empty_like(reward)
    loss, gbuf = torch.empty(1, device="cuda"), torch.empty_like(reward)
    ws = _abi.workspace(_abi.OP_TD_LAMBDA, T, B, 0, "cuda")
    L = _abi.lib()

    def step(st):
        _abi.check(L.hpc_rll_gae_forward(value.data_ptr(), reward.data_ptr(), adv.data_ptr(), T, B, 0.99, 0.97, st), "f")
        _abi.check(L.hpc_rll_gae_backward(gadv.data_ptr(), gv.data_ptr(), gr.data_ptr(), T, B, 0.99, 0.97, st), "b")
        _abi.check(L.hpc_rll_td_lambda_forward(value.data_ptr(), reward.data_ptr(), None, loss.data_ptr(),
                                               gbuf.data_ptr(), T, B, 0.9, 0.8, 0, ws.data_ptr(), ws.numel(), st), "t")

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        step(side.cuda_stream)  # warm-up: uploads the (T, lambda) table, opts kernels into large smem
    side.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        step(torch.cuda.current_stream().cuda_stream)
    # new data, replay
    value.copy_(dev(g.standard_normal((T + 1, B), dtype=np.float32)))
    reward.copy_(dev(g.standard_normal((T, B), dtype=np.float32)))
    torch.cuda.synchronize()
    graph.replay()
    torch.cuda.synchronize()
    # B = 2048: the small-batch T-split kernels (epoch-tagged look-back scratch must survive graph replay)
    assert rel_err(host(adv), orc.gae_forward(host(value), host(reward))) <= 2e-6
    ob = orc.gae_backward(host(gadv))
    assert rel_err(host(gv), ob["value"]) <= 2e-6 and rel_err(host(gr), ob["reward"]) <= 2e-6
    for _ in range(3):  # replay again with new data: stale aggregates of the previous replay must not be taken
        value.copy_(dev(g.standard_normal((T + 1, B), dtype=np.float32)))
        gadv.copy_(dev(g.standard_normal((T, B), dtype=np.float32)))
        torch.cuda.synchronize()
        graph.replay()
        torch.cuda.synchronize()
        assert rel_err(host(adv), orc.gae_forward(host(value), host(reward))) <= 2e-6
        assert rel_err(host(gv), orc.gae_backward(host(gadv))["value"]) <= 2e-6
    assert rel_err(float(loss.item()), orc.td_lambda(host(value), host(reward))["loss"]) <= 1e-5


def test_vtrace_int64_indexing_and_shard_composition():
    """T*B*N = 2^31 + ... elements: the full batch must equal the sum of its two half-batch shards run
    with global_B (exercises 64-bit row offsets and the global normalisation used for data parallelism)."""
    need_cuda()
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs ~45 GB of device memory")
    from hpc_rll.rl_utils.vtrace import VTrace
    T, B, N = 512, 32768, 128
    assert T * B * N >= 2**31
    gen = torch.Generator(device="cuda").manual_seed(5)
    tgt = torch.randn(T, B, N, device="cuda", generator=gen)
    beh = torch.randn(T, B, N, device="cuda", generator=gen)
    act = torch.randint(0, N, (T, B), device="cuda", generator=gen)
    val = torch.randn(T + 1, B, device="cuda", generator=gen)
    rew = torch.randn(T, B, device="cuda", generator=gen)

    def run(sl, global_B):
        t = tgt[:, sl].contiguous().requires_grad_(True)
        v = val[:, sl].contiguous().requires_grad_(True)
        m = VTrace(T, t.shape[1], N)
        m.global_B = global_B
        l = m(t, beh[:, sl].contiguous(), act[:, sl].contiguous(), v, rew[:, sl].contiguous())
        tot = l.policy_loss + 0.5 * l.value_loss - 0.01 * l.entropy_loss
        gt, gvv = torch.autograd.grad(tot, [t, v], grad_outputs=torch.ones(1, device="cuda"))
        return torch.stack([l.policy_loss, l.value_loss, l.entropy_loss]).flatten().double(), gt, gvv

    full_l, full_gt, full_gv = run(slice(0, B), 0)
    h = B // 2
    a_l, a_gt, a_gv = run(slice(0, h), B)
    b_l, b_gt, b_gv = run(slice(h, B), B)
    assert torch.allclose(a_l + b_l, full_l, rtol=1e-5, atol=1e-7)
    # the last rows of the tensor live beyond the 2^31-element mark
    assert torch.equal(full_gt[:, :h], a_gt) and torch.equal(full_gt[:, h:], b_gt)
    assert torch.equal(full_gv[:, :h], a_gv) and torch.equal(full_gv[:, h:], b_gv)
    assert float(full_gt[-1, -1].abs().sum()) > 0


def test_empty_and_degenerate_sizes():
    need_cuda()
    from di_hpc_b200 import _abi
    L = _abi.lib()
    st = torch.cuda.current_stream().cuda_stream
    x = torch.zeros(8, device="cuda")
    # GAE with T == 0 or B == 0 is a no-op (backward zero-fills the single value row)
    assert L.hpc_rll_gae_forward(x.data_ptr(), x.data_ptr(), x.data_ptr(), 0, 8, 0.99, 0.97, st) == 0
    assert L.hpc_rll_gae_forward(x.data_ptr(), x.data_ptr(), x.data_ptr(), 4, 0, 0.99, 0.97, st) == 0
    gv = torch.ones(8, device="cuda")
    assert L.hpc_rll_gae_backward(None, gv.data_ptr(), None, 0, 8, 0.99, 0.97, st) == 0
    torch.cuda.synchronize()
    assert float(gv.abs().sum()) == 0.0
    # loss ops reject empty batches with a message instead of dividing by zero
    ws = _abi.workspace(_abi.OP_TD_LAMBDA, 4, 8, 0, "cuda")
    rc = L.hpc_rll_td_lambda_forward(x.data_ptr(), x.data_ptr(), None, x.data_ptr(), x.data_ptr(), 0, 8, 0.9, 0.8, 0,
                                     ws.data_ptr(), ws.numel(), st)
    assert rc == 1 and b"positive" in L.hpc_rll_last_error()


def test_make_graphed_callables_small_batch():
    """The module-level answer to launch latency at small batches: PyTorch's own `torch.cuda.make_graphed_callables`
    captures forward AND autograd backward of a drop-in module into CUDA graphs.  It captures on an internal stream, so
    this exercises the look-back scratch adoption during capture (csrc/scan_lookback.cu) and the epoch re-arming across
    replays, on the reference's own test shape (tests/test_gae.py:10-11, tests/test_tdlambda.py:10-11)."""
    need_cuda()
    from hpc_rll.rl_utils.gae import GAE
    from hpc_rll.rl_utils.td import TDLambda
    g = rng(9)
    T, B = 1024, 64
    sv = torch.randn(T + 1, B, device="cuda", requires_grad=True)
    sr = torch.randn(T, B, device="cuda", requires_grad=True)
    ggae = torch.cuda.make_graphed_callables(GAE(T, B), (sv, sr))
    gtd = torch.cuda.make_graphed_callables(TDLambda(T, B), (sv.detach().clone().requires_grad_(True), sr.detach().clone()))
    for _ in range(3):
        value = g.standard_normal((T + 1, B), dtype=np.float32)
        reward = g.standard_normal((T, B), dtype=np.float32)
        gadv = g.standard_normal((T, B), dtype=np.float32)
        v, r = dev(value).requires_grad_(True), dev(reward).requires_grad_(True)
        adv = ggae(v, r)
        adv.backward(dev(gadv))
        torch.cuda.synchronize()
        ob = orc.gae_backward(gadv)
        assert rel_err(host(adv), orc.gae_forward(value, reward)) <= 2e-6
        assert rel_err(host(v.grad), ob["value"]) <= 2e-6 and rel_err(host(r.grad), ob["reward"]) <= 2e-6
        v2 = dev(value).requires_grad_(True)
        loss = gtd(v2, dev(reward))
        loss.backward()
        torch.cuda.synchronize()
        o = orc.td_lambda(value, reward)
        assert rel_err(float(loss.item()), o["loss"]) <= 1e-5
        assert float(np.max(np.abs(host(v2.grad) - o["grad_value"])) / np.max(np.abs(o["grad_value"]))) <= 1e-5
