"""V-trace -- drop-in for /root/reference/hpc_rll/rl_utils/vtrace.py (VTrace, VtraceFunction,
hpc_vtrace_loss).  Same constructor ``VTrace(T, B, N)`` and forward signature (vtrace.py:83-133).
The three losses are shape-(1,) tensors as in the reference (vtrace.py:77-79)."""
from collections import namedtuple

import torch

from .. import _abi, _ext

hpc_vtrace_loss = namedtuple('hpc_vtrace_loss', ['policy_loss', 'value_loss', 'entropy_loss'])


class VtraceFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, target_output, behaviour_output, action, value, reward, weight, gamma, lambda_, rho_clip_ratio,
                c_clip_ratio, rho_pg_clip_ratio, global_B):
        target_output = _abi.require_f32_cuda("target_output", target_output)
        behaviour_output = _abi.require_f32_cuda("behaviour_output", behaviour_output)
        action = _abi.require_i64_cuda("action", action)
        value = _abi.require_f32_cuda("value", value)
        reward = _abi.require_f32_cuda("reward", reward)
        T, B, N = target_output.shape
        if behaviour_output.shape != (T, B, N) or action.shape != (T, B) or reward.shape != (T, B) \
                or value.shape != (T + 1, B):
            raise ValueError("vtrace: inconsistent shapes")
        if weight is not None:
            weight = _abi.require_f32_cuda("weight", weight)
            if weight.shape != (T, B):
                raise ValueError("weight must be (T, B)")
        dev = reward.device
        losses = torch.empty(3, dtype=torch.float32, device=dev)
        pg_coef = torch.empty((T, B), dtype=torch.float32, device=dev)
        gv_buf = torch.empty((T, B), dtype=torch.float32, device=dev)
        ws = _abi.workspace(_abi.OP_VTRACE, T, B, N, dev)
        with _abi.on_device(dev):
            _abi.check(
                _abi.lib().hpc_rll_vtrace_forward(
                    _abi.ptr(target_output), _abi.ptr(behaviour_output), _abi.ptr(action), _abi.ptr(value),
                    _abi.ptr(reward), _abi.ptr(weight), _abi.ptr(losses), _abi.ptr(pg_coef), _abi.ptr(gv_buf), T, B, N,
                    float(gamma), float(lambda_), float(rho_clip_ratio), float(c_clip_ratio),
                    float(rho_pg_clip_ratio), int(global_B), _abi.ptr(ws), ws.numel(), _abi.stream_of(reward)),
                "hpc_rll_vtrace_forward")
        ctx.save_for_backward(target_output, action, weight, pg_coef, gv_buf)
        ctx.global_B = int(global_B)
        return losses[0:1], losses[1:2], losses[2:3]

    @staticmethod
    def backward(ctx, grad_pg_loss, grad_value_loss, grad_entropy_loss):
        target_output, action, weight, pg_coef, gv_buf = ctx.saved_tensors
        T, B, N = target_output.shape
        g_pg = _abi.grad_scalar(grad_pg_loss, pg_coef)
        g_v = _abi.grad_scalar(grad_value_loss, pg_coef)
        g_e = _abi.grad_scalar(grad_entropy_loss, pg_coef)
        grad_target = torch.empty_like(target_output)
        grad_value = torch.empty((T + 1, B), dtype=torch.float32, device=pg_coef.device)
        with _abi.on_device(pg_coef.device):
            _abi.check(
                _abi.lib().hpc_rll_vtrace_backward(_abi.ptr(g_pg), _abi.ptr(g_v), _abi.ptr(g_e),
                                                   _abi.ptr(target_output), _abi.ptr(action), _abi.ptr(weight),
                                                   _abi.ptr(pg_coef), _abi.ptr(gv_buf), _abi.ptr(grad_target),
                                                   _abi.ptr(grad_value), T, B, N, ctx.global_B,
                                                   _abi.stream_of(pg_coef)), "hpc_rll_vtrace_backward")
        return grad_target, None, None, grad_value, None, None, None, None, None, None, None, None


class VTrace(torch.nn.Module):
    """
    Overview:
        V-trace losses (IMPALA, arXiv:1802.01561), hpc_rll/origin/vtrace.py:24-79.

    Interface:
        __init__, forward
    """

    def __init__(self, T, B, N):
        r"""
        Arguments:
            - T (:obj:`int`): trajectory length
            - B (:obj:`int`): batch size
            - N (:obj:`int`): number of output
        """
        super().__init__()
        self.T, self.B, self.N = T, B, N
        self.global_B = 0

    def forward(self, target_output, behaviour_output, action, value, reward, weight=None, gamma: float = 0.99,
                lambda_: float = 0.95, rho_clip_ratio: float = 1.0, c_clip_ratio: float = 1.0,
                rho_pg_clip_ratio: float = 1.0):
        """
        Arguments:
            - target_output (:obj:`torch.Tensor`): :math:`(T, B, N)` logits of the current policy
            - behaviour_output (:obj:`torch.Tensor`): :math:`(T, B, N)` logits of the behaviour policy
            - action (:obj:`torch.Tensor`): :math:`(T, B)` int64
            - value (:obj:`torch.Tensor`): :math:`(T + 1, B)`
            - reward (:obj:`torch.Tensor`): :math:`(T, B)`
            - weight (:obj:`torch.Tensor` or None): :math:`(T, B)`
        Returns:
            - trace_loss (:obj:`hpc_vtrace_loss`): policy_loss, value_loss, entropy_loss -- shape (1,) each
        """
        assert (target_output.is_cuda)
        assert (behaviour_output.is_cuda)
        assert (action.is_cuda)
        assert (value.is_cuda)
        assert (reward.is_cuda)
        if weight is not None:
            assert (weight.is_cuda)
        fast = _ext.fast()
        if fast is not None:
            return hpc_vtrace_loss(*fast.vtrace(target_output, behaviour_output, action, value, reward, weight,
                                                float(gamma), float(lambda_), float(rho_clip_ratio),
                                                float(c_clip_ratio), float(rho_pg_clip_ratio), int(self.global_B)))
        pg_loss, value_loss, entropy_loss = VtraceFunction.apply(target_output, behaviour_output, action, value,
                                                                 reward, weight, gamma, lambda_, rho_clip_ratio,
                                                                 c_clip_ratio, rho_pg_clip_ratio, self.global_B)
        return hpc_vtrace_loss(pg_loss, value_loss, entropy_loss)
