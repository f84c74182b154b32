"""GAE -- drop-in for /root/reference/hpc_rll/rl_utils/gae.py (class GAE, GAEFunction).

Same constructor ``GAE(T, B)`` and ``forward(value, reward, gamma=0.99, lambda_=0.97)`` as the
reference (gae.py:20-61).  Differences, all deliberate (DESIGN.md "Deviations"):
  * ``GAEFunction.backward`` is a real adjoint (the reference returns None for every input,
    gae.py:16-18, while hpc_rll/origin/gae.py is differentiable);
  * the result is a fresh tensor sized from the inputs, not the module-owned ``adv`` buffer returned
    by alias (gae.py:39,61) -- T and B of the constructor are kept only for API compatibility;
  * kernels run on PyTorch's current stream, with argument checks.
"""
import torch

from .. import _abi, _ext


class GAEFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, value, reward, gamma, lambda_):
        value = _abi.require_f32_cuda("value", value)
        reward = _abi.require_f32_cuda("reward", reward)
        T, B = reward.shape
        if value.shape != (T + 1, B):
            raise ValueError("value must be (T+1, B)=(%d, %d), got %s" % (T + 1, B, tuple(value.shape)))
        adv = torch.empty_like(reward)
        with _abi.on_device(reward.device):
            _abi.check(
                _abi.lib().hpc_rll_gae_forward(_abi.ptr(value), _abi.ptr(reward), _abi.ptr(adv), T, B, float(gamma),
                                               float(lambda_), _abi.stream_of(reward)), "hpc_rll_gae_forward")
        ctx.gamma, ctx.lambda_, ctx.T, ctx.B = float(gamma), float(lambda_), T, B
        return adv

    @staticmethod
    def backward(ctx, grad_adv):
        T, B = ctx.T, ctx.B
        grad_adv = _abi.require_f32_cuda("grad_adv", grad_adv)
        grad_value = torch.empty((T + 1, B), dtype=torch.float32, device=grad_adv.device)
        grad_reward = torch.empty((T, B), dtype=torch.float32, device=grad_adv.device)
        with _abi.on_device(grad_adv.device):
            _abi.check(
                _abi.lib().hpc_rll_gae_backward(_abi.ptr(grad_adv), _abi.ptr(grad_value), _abi.ptr(grad_reward), T, B,
                                                ctx.gamma, ctx.lambda_, _abi.stream_of(grad_adv)),
                "hpc_rll_gae_backward")
        return grad_value, grad_reward, None, None


def gae_with_adv_stats(value, reward, gamma: float = 0.99, lambda_: float = 0.97, group=None):
    """
    Overview:
        GAE forward fused with the statistics of the advantage normalisation
        ``(adv - adv.mean()) / (adv.std() + 1e-8)`` that precedes ``ppo_error``
        (hpc_rll/origin/ppo.py:43-47): the scan kernel accumulates sum(adv) and sum(adv^2) in fp64 while it
        writes ``adv``, so mean/std cost no extra pass over HBM.  Hand ``adv_stats`` to ``PPO.forward(...,
        adv_stats=adv_stats)`` and the normalisation itself happens inside the PPO kernel.  No autograd
        (advantages are targets; the reference's GAE has no backward at all, rl_utils/gae.py:16-18).
    Arguments:
        - value (:obj:`torch.FloatTensor`): :math:`(T + 1, B)`; reward :math:`(T, B)`
        - group: ``torch.distributed`` group over which the batch axis is sharded (moments are all-reduced)
    Returns:
        - adv (:obj:`torch.FloatTensor`): :math:`(T, B)`, NOT normalised
        - adv_stats (:obj:`torch.FloatTensor`): :math:`(2,)` = ``[mean, std + 1e-8]`` of the (global) advantages
    """
    from ..sharding import all_reduce_moments
    value = _abi.require_f32_cuda("value", value.detach())
    reward = _abi.require_f32_cuda("reward", reward.detach())
    T, B = reward.shape
    if value.shape != (T + 1, B):
        raise ValueError("value must be (T+1, B)=(%d, %d), got %s" % (T + 1, B, tuple(value.shape)))
    if T == 0 or B == 0:
        raise ValueError("gae_with_adv_stats needs a non-empty batch")
    dev = reward.device
    adv = torch.empty_like(reward)
    moments = torch.empty(3, dtype=torch.float64, device=dev)
    moments[2] = T * B
    stats = torch.empty(2, dtype=torch.float32, device=dev)
    ws = _abi.workspace(_abi.OP_GAE_MOMENTS, T, B, 0, dev)
    with _abi.on_device(dev):
        _abi.check(
            _abi.lib().hpc_rll_gae_forward_moments(_abi.ptr(value), _abi.ptr(reward), _abi.ptr(adv),
                                                   _abi.ptr(moments), T, B, float(gamma), float(lambda_),
                                                   _abi.ptr(ws), ws.numel(), _abi.stream_of(reward)),
            "hpc_rll_gae_forward_moments")
        all_reduce_moments(moments, group)
        _abi.check(_abi.lib().hpc_rll_adv_stats(_abi.ptr(moments), 0, _abi.ptr(stats), _abi.stream_of(reward)),
                   "hpc_rll_adv_stats")
    return adv, stats


class GAE(torch.nn.Module):
    """
    Overview:
        Generalized Advantage Estimator (arXiv:1506.02438), normalised variant of
        hpc_rll/origin/gae.py:28-37.

    Interface:
        __init__, forward
    """

    def __init__(self, T, B):
        r"""
        Arguments:
            - T (:obj:`int`): trajectory length
            - B (:obj:`int`): batch size
        """
        super().__init__()
        self.T, self.B = T, B

    def forward(self, value, reward, gamma: float = 0.99, lambda_: float = 0.97) -> torch.FloatTensor:
        """
        Arguments:
            - value (:obj:`torch.FloatTensor`): :math:`(T + 1, B)`
            - reward (:obj:`torch.FloatTensor`): :math:`(T, B)`
            - gamma (:obj:`float`): discount factor, defaults to 0.99
            - lambda_ (:obj:`float`): GAE lambda, defaults to 0.97
        Returns:
            - adv (:obj:`torch.FloatTensor`): :math:`(T, B)`
        """
        assert (value.is_cuda)
        assert (reward.is_cuda)
        fast = _ext.fast()
        if fast is not None:  # C++ autograd function (csrc_torch/fast_ops.cpp); GAEFunction is its ctypes-bound twin
            return fast.gae(value, reward, float(gamma), float(lambda_))
        return GAEFunction.apply(value, reward, gamma, lambda_)
