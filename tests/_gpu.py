"""Helpers for the -m gpu parity tests (CUDA path through the C ABI vs the oracle)."""
import numpy as np
import pytest
import torch

from tests._golden import grad_err, rel_err  # noqa: F401


def need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def dev(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def host(t):
    return t.detach().cpu().numpy()


def rng(seed):
    return np.random.default_rng(seed)
