"""The dense backward scatter of the n-step ops (csrc/nstep.cu: `scatter_rows_kernel`, and from 1 MB of output on the
shared-memory-image + bulk-store kernel `scatter_rows_bulk_kernel`) against a plain torch scatter, through the C ABI.
One case per kernel instantiation the launcher can pick, each with a tail that is not a multiple of the image rows.
The product g * buf is one fp32 multiply, so the comparison is bit-exact."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")


def reference(buf, action, g, R, N, L, period):
    out = torch.zeros(R, N, L, device=buf.device)
    rows = torch.arange(R, device=buf.device)
    out[rows, action[rows % period]] = g * buf.view(R, L)
    return out


def run(kind, B, N, L, tau=1, offset=0):
    from di_hpc_b200 import _abi
    lib = _abi.lib()
    gen = torch.Generator(device="cuda").manual_seed(B * 7 + N * 3 + L + tau)
    action = torch.randint(0, N, (B,), device="cuda", generator=gen)
    g = torch.tensor([0.731], device="cuda")
    R = tau * B if kind == "iqn" else B
    buf = torch.randn(R * L, device="cuda", generator=gen)
    store = torch.full((R * N * L + 8,), float("nan"), device="cuda")
    out = store[offset:offset + R * N * L]  # offset != 0: base not 16-byte aligned -> the per-thread-store kernel
    s = _abi.stream_of(buf)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    if kind == "q":
        rc = lib.hpc_rll_q_nstep_td_backward(p(g), p(buf), p(action), p(out), B, N, s)
    elif kind == "dist":
        rc = lib.hpc_rll_dist_nstep_td_backward(p(g), p(buf), p(action), p(out), B, N, L, s)
    elif kind == "qrdqn":
        rc = lib.hpc_rll_qrdqn_nstep_td_backward(p(g), p(buf), p(action), p(out), L, B, N, s)
    else:
        rc = lib.hpc_rll_iqn_nstep_td_backward(p(g), p(buf), p(action), p(out), tau, B, N, s)
    _abi.check(rc, kind)
    torch.cuda.synchronize()
    want = reference(buf, action, g, R, N, L, B)
    got = out.view(R, N, L)
    assert torch.equal(got, want), "%s B=%d N=%d L=%d tau=%d: %d elements differ" % (
        kind, B, N, L, tau, int((got != want).sum()))
    assert torch.isnan(store[:offset]).all() and torch.isnan(store[offset + R * N * L:]).all()  # nothing outside


# (kind, B, N, L, tau): every shape writes > 1 MB, so the aligned runs take the bulk-store kernel
CASES = [
    ("q", 70001, 4, 1, 1),       # lane per row, 8 rows per lane (16-byte rows)
    ("q", 40003, 8, 1, 1),       # 4 rows per lane
    ("q", 20001, 16, 1, 1),      # 2 rows per lane
    ("q", 9001, 40, 1, 1),       # 1 row per lane
    ("q", 3001, 128, 1, 1),      # 1 row per lane, 512-byte rows (16 KB images)
    ("q", 3001, 200, 1, 1),      # rows too long for the lane-per-row images -> per-thread stores
    ("iqn", 5003, 8, 1, 13),     # R = tau * B rows, action index r % B wraps inside images
    ("iqn", 37, 8, 1, 1024),     # period shorter than an image: several wraps per image
    ("dist", 9001, 4, 7, 1),     # lane per row, L > 1 (rows of buf read at store time)
    ("dist", 5001, 4, 16, 1),
    ("dist", 4001, 8, 21, 1),    # warp per row, <= 32 elements per row of buf
    ("dist", 3001, 8, 51, 1),    # <= 64 (C51)
    ("qrdqn", 2003, 8, 64, 1),   # <= 64, 16-byte aligned rows of buf
    ("qrdqn", 1501, 6, 100, 1),  # <= 128
    ("qrdqn", 1203, 5, 130, 1),  # any length
    ("qrdqn", 803, 4, 512, 1),   # 8 KB rows: image too large -> per-thread stores
]


@pytest.mark.parametrize("kind,B,N,L,tau", CASES)
def test_scatter_matches_torch(kind, B, N, L, tau):
    need_cuda()
    run(kind, B, N, L, tau)


@pytest.mark.parametrize("kind,B,N,L,tau", [CASES[1], CASES[11], CASES[12]])
def test_scatter_unaligned_output(kind, B, N, L, tau):
    need_cuda()
    run(kind, B, N, L, tau, offset=1)


def test_scatter_small_outputs_use_thread_stores():
    need_cuda()
    for kind, B, N, L, tau in [("q", 5, 3, 1, 1), ("dist", 33, 5, 51, 1), ("qrdqn", 17, 3, 32, 1), ("iqn", 6, 4, 1, 5)]:
        run(kind, B, N, L, tau)
