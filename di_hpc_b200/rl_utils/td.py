"""TD-error family -- drop-in for /root/reference/hpc_rll/rl_utils/td.py.

Classes (same names, constructor arguments and forward signatures as the reference):
    DistNStepTD(T, B, N, n_atom)        td.py:33-108
    TDLambda(T, B)                      td.py:140-187
    QNStepTD(T, B, N)                   td.py:213-277
    QNStepTDRescale(T, B, N)            td.py:303-371
    IQNNStepTDError(tau, tauPrime, T, B, N)   td.py:397-485
    QRDQNNStepTDError(tau, T, B, N)     td.py:513-592

Common deviations from the reference wrapper (DESIGN.md "Deviations"): results are fresh tensors
(the reference returns module-owned buffers by alias), scratch lives in a per-call workspace rather
than in registered buffers that pollute ``state_dict``, shapes come from the inputs (constructor
sizes are kept for API compatibility), saved state goes through ``save_for_backward``.
Scalars keep the reference's shape ``(1,)`` (register_buffer('loss', torch.zeros(1)), td.py:161).
"""
from typing import Optional

import torch

from .. import _abi, _ext


def _nstep_common(q_like, action, next_n_action, reward, done, weight, B):
    """Shared argument normalisation of the n-step family (reference asserts is_cuda only)."""
    action = _abi.require_i64_cuda("action", action)
    next_n_action = _abi.require_i64_cuda("next_n_action", next_n_action)
    reward = _abi.require_f32_cuda("reward", reward)
    if done.dtype != torch.float32:
        # the reference kernels read `done` as float* (src/rl_utils/q_nstep_td.cu:39); accept bool too
        done = done.to(torch.float32)
    done = _abi.require_f32_cuda("done", done)
    if action.shape != (B, ) or next_n_action.shape != (B, ) or done.shape != (B, ) or reward.dim() != 2 \
            or reward.shape[1] != B:
        raise ValueError("n-step td: inconsistent shapes (B=%d)" % B)
    if weight is not None:
        weight = _abi.require_f32_cuda("weight", weight)
        if weight.shape != (B, ):
            raise ValueError("weight must be (B,)")
    return action, next_n_action, reward, done, weight


class TDLambdaFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, value, reward, weight, gamma, lambda_, global_B):
        value = _abi.require_f32_cuda("value", value)
        reward = _abi.require_f32_cuda("reward", reward)
        T, B = reward.shape
        if value.shape != (T + 1, B):
            raise ValueError("value must be (T+1, B)=(%d, %d), got %s" % (T + 1, B, tuple(value.shape)))
        if weight is not None:
            weight = _abi.require_f32_cuda("weight", weight)
            if weight.dim() == 1:
                # the reference documents weight as (B,) (td.py:172) though its kernel indexes (T,B)
                # (td_lambda_kernel.h:24); broadcast the documented form
                weight = weight.unsqueeze(0).expand(T, B).contiguous()
            if weight.shape != (T, B):
                raise ValueError("weight must be (T, B) or (B,), got %s" % (tuple(weight.shape),))
        loss = torch.empty(1, dtype=torch.float32, device=reward.device)
        grad_buf = torch.empty_like(reward)
        ws = _abi.workspace(_abi.OP_TD_LAMBDA, T, B, 0, reward.device)
        with _abi.on_device(reward.device):
            _abi.check(
                _abi.lib().hpc_rll_td_lambda_forward(_abi.ptr(value), _abi.ptr(reward), _abi.ptr(weight),
                                                     _abi.ptr(loss), _abi.ptr(grad_buf), T, B, float(gamma),
                                                     float(lambda_), int(global_B), _abi.ptr(ws), ws.numel(),
                                                     _abi.stream_of(reward)), "hpc_rll_td_lambda_forward")
        ctx.save_for_backward(grad_buf)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        (grad_buf, ) = ctx.saved_tensors
        T, B = grad_buf.shape
        g = _abi.grad_scalar(grad_loss, grad_buf)
        grad_value = torch.empty((T + 1, B), dtype=torch.float32, device=grad_buf.device)
        with _abi.on_device(grad_buf.device):
            _abi.check(
                _abi.lib().hpc_rll_td_lambda_backward(_abi.ptr(g), _abi.ptr(grad_buf), _abi.ptr(grad_value), T, B,
                                                      _abi.stream_of(grad_buf)), "hpc_rll_td_lambda_backward")
        return grad_value, None, None, None, None, None


class TDLambda(torch.nn.Module):
    """
    Overview:
        TD(lambda) loss with constant gamma and lambda (hpc_rll/origin/td.py:148-176).

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
        self.global_B = 0  # set >0 when B is one shard of a data-parallel batch (see di_hpc_b200.sharding)

    def forward(self, value, reward, weight=None, gamma: float = 0.9, lambda_: float = 0.8) -> torch.Tensor:
        """
        Arguments:
            - value (:obj:`torch.FloatTensor`): :math:`(T + 1, B)`
            - reward (:obj:`torch.FloatTensor`): :math:`(T, B)`
            - weight (:obj:`torch.FloatTensor` or None): :math:`(T, B)` (or :math:`(B,)`, broadcast)
            - gamma (:obj:`float`): discount factor, defaults to 0.9
            - lambda_ (:obj:`float`): lambda, defaults to 0.8
        Returns:
            - loss (:obj:`torch.Tensor`): shape :math:`(1,)`
        """
        assert (value.is_cuda)
        assert (reward.is_cuda)
        if weight is not None:
            assert (weight.is_cuda)
        fast = _ext.fast()
        if fast is not None:
            return fast.td_lambda(value, reward, weight, float(gamma), float(lambda_), int(self.global_B))
        return TDLambdaFunction.apply(value, reward, weight, gamma, lambda_, self.global_B)


# ------------------------------------------------------------------------------------------- q n-step
class _QNStepTDBase(torch.autograd.Function):
    RESCALE = 0

    @classmethod
    def _fwd(cls, ctx, q, next_n_q, action, next_n_action, reward, done, weight, gamma, global_B):
        q = _abi.require_f32_cuda("q", q)
        next_n_q = _abi.require_f32_cuda("next_n_q", next_n_q)
        B, N = q.shape
        if next_n_q.shape != (B, N):
            raise ValueError("next_n_q must match q")
        action, next_n_action, reward, done, weight = _nstep_common(q, action, next_n_action, reward, done, weight, B)
        T = reward.shape[0]
        dev = q.device
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        td_err = torch.empty(B, dtype=torch.float32, device=dev)
        grad_buf = torch.empty(B, dtype=torch.float32, device=dev)
        ws = _abi.workspace(_abi.OP_Q_NSTEP_TD, T, B, N, dev)
        with _abi.on_device(dev):
            _abi.check(
                _abi.lib().hpc_rll_q_nstep_td_forward(_abi.ptr(q), _abi.ptr(next_n_q), _abi.ptr(action),
                                                      _abi.ptr(next_n_action), _abi.ptr(reward), _abi.ptr(done),
                                                      _abi.ptr(weight), _abi.ptr(loss), _abi.ptr(td_err),
                                                      _abi.ptr(grad_buf), T, B, N, float(gamma), cls.RESCALE,
                                                      int(global_B), _abi.ptr(ws), ws.numel(), _abi.stream_of(q)),
                "hpc_rll_q_nstep_td_forward")
        ctx.save_for_backward(grad_buf, action)
        ctx.N = N
        ctx.mark_non_differentiable(td_err)
        return loss, td_err

    @staticmethod
    def _bwd(ctx, grad_loss):
        grad_buf, action = ctx.saved_tensors
        B, N = grad_buf.shape[0], ctx.N
        g = _abi.grad_scalar(grad_loss, grad_buf)
        grad_q = torch.empty((B, N), dtype=torch.float32, device=grad_buf.device)
        with _abi.on_device(grad_buf.device):
            _abi.check(
                _abi.lib().hpc_rll_q_nstep_td_backward(_abi.ptr(g), _abi.ptr(grad_buf), _abi.ptr(action),
                                                       _abi.ptr(grad_q), B, N, _abi.stream_of(grad_buf)),
                "hpc_rll_q_nstep_td_backward")
        return grad_q


class QNStepTDFunction(_QNStepTDBase):
    RESCALE = 0

    @staticmethod
    def forward(ctx, q, next_n_q, action, next_n_action, reward, done, weight, gamma, global_B):
        return QNStepTDFunction._fwd(ctx, q, next_n_q, action, next_n_action, reward, done, weight, gamma, global_B)

    @staticmethod
    def backward(ctx, grad_loss, grad_td_err):
        return (_QNStepTDBase._bwd(ctx, grad_loss), ) + (None, ) * 8


class QNStepTDRescaleFunction(_QNStepTDBase):
    RESCALE = 1

    @staticmethod
    def forward(ctx, q, next_n_q, action, next_n_action, reward, done, weight, gamma, global_B):
        return QNStepTDRescaleFunction._fwd(ctx, q, next_n_q, action, next_n_action, reward, done, weight, gamma,
                                            global_B)

    @staticmethod
    def backward(ctx, grad_loss, grad_td_err):
        return (_QNStepTDBase._bwd(ctx, grad_loss), ) + (None, ) * 8


class QNStepTD(torch.nn.Module):
    """
    Overview:
        Multistep (1 step or n step) td_error for q-learning based algorithm
        (hpc_rll/origin/td.py:252-291; criterion = MSELoss as in the reference, td.py:266-268).

    Interface:
        __init__, forward
    """
    _fn = QNStepTDFunction

    def __init__(self, T, B, N):
        r"""
        Arguments:
            - T (:obj:`int`): nstep
            - B (:obj:`int`): batch size
            - N (:obj:`int`): action dim
        """
        super().__init__()
        self.T, self.B, self.N = T, B, N
        self.global_B = 0

    def forward(self, q, next_n_q, action, next_n_action, reward, done, weight, gamma: float) -> torch.Tensor:
        """
        Arguments:
            - q, next_n_q (:obj:`torch.FloatTensor`): :math:`(B, N)`
            - action, next_n_action (:obj:`torch.LongTensor`): :math:`(B, )`
            - reward (:obj:`torch.FloatTensor`): :math:`(T, B)`
            - done (:obj:`torch.FloatTensor`): :math:`(B, )` 0/1
            - weight (:obj:`torch.FloatTensor` or None): :math:`(B, )`
            - gamma (:obj:`float`): discount factor
        Returns:
            - loss (:obj:`torch.Tensor`): shape :math:`(1,)`
            - td_error_per_sample (:obj:`torch.Tensor`): :math:`(B, )`
        """
        for t in (q, next_n_q, action, next_n_action, reward, done):
            assert (t.is_cuda)
        if weight is not None:
            assert (weight.is_cuda)
        fast = _ext.fast()
        if fast is not None:
            return tuple(fast.q_nstep_td(q, next_n_q, action, next_n_action, reward, done, weight, float(gamma),
                                         bool(self._fn.RESCALE), int(self.global_B)))
        return self._fn.apply(q, next_n_q, action, next_n_action, reward, done, weight, gamma, self.global_B)


class QNStepTDRescale(QNStepTD):
    """
    Overview:
        n-step td_error with value rescaling (hpc_rll/origin/td.py:294-340), same interface as QNStepTD.
    """
    _fn = QNStepTDRescaleFunction


# ------------------------------------------------------------------------------------------- C51
class DistNStepTDFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, dist, next_n_dist, action, next_n_action, reward, done, weight, gamma, v_min, v_max, global_B):
        dist = _abi.require_f32_cuda("dist", dist)
        next_n_dist = _abi.require_f32_cuda("next_n_dist", next_n_dist)
        B, N, n_atom = dist.shape
        if next_n_dist.shape != (B, N, n_atom):
            raise ValueError("next_n_dist must match dist")
        action, next_n_action, reward, done, weight = _nstep_common(dist, action, next_n_action, reward, done, weight,
                                                                    B)
        T = reward.shape[0]
        dev = dist.device
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        td_err = torch.empty(B, dtype=torch.float32, device=dev)
        grad_buf = torch.empty((B, n_atom), dtype=torch.float32, device=dev)
        ws = _abi.workspace(_abi.OP_DIST_NSTEP_TD, T, B, N, dev)
        with _abi.on_device(dev):
            _abi.check(
                _abi.lib().hpc_rll_dist_nstep_td_forward(_abi.ptr(dist), _abi.ptr(next_n_dist), _abi.ptr(action),
                                                         _abi.ptr(next_n_action), _abi.ptr(reward), _abi.ptr(done),
                                                         _abi.ptr(weight), _abi.ptr(loss), _abi.ptr(td_err),
                                                         _abi.ptr(grad_buf), T, B, N, n_atom, float(gamma),
                                                         float(v_min), float(v_max), int(global_B), _abi.ptr(ws),
                                                         ws.numel(), _abi.stream_of(dist)),
                "hpc_rll_dist_nstep_td_forward")
        ctx.save_for_backward(grad_buf, action)
        ctx.N = N
        ctx.mark_non_differentiable(td_err)
        return loss, td_err

    @staticmethod
    def backward(ctx, grad_loss, grad_td_err):
        grad_buf, action = ctx.saved_tensors
        B, n_atom = grad_buf.shape
        N = ctx.N
        g = _abi.grad_scalar(grad_loss, grad_buf)
        grad_dist = torch.empty((B, N, n_atom), dtype=torch.float32, device=grad_buf.device)
        with _abi.on_device(grad_buf.device):
            _abi.check(
                _abi.lib().hpc_rll_dist_nstep_td_backward(_abi.ptr(g), _abi.ptr(grad_buf), _abi.ptr(action),
                                                          _abi.ptr(grad_dist), B, N, n_atom,
                                                          _abi.stream_of(grad_buf)), "hpc_rll_dist_nstep_td_backward")
        return (grad_dist, ) + (None, ) * 10


class DistNStepTD(torch.nn.Module):
    """
    Overview:
        Multistep td_error for distributional (C51) q-learning (hpc_rll/origin/td.py:29-143).

    Interface:
        __init__, forward
    """

    def __init__(self, T, B, N, n_atom):
        r"""
        Arguments:
            - T (:obj:`int`): nstep
            - B (:obj:`int`): batch size
            - N (:obj:`int`): action dim
            - n_atom (:obj:`int`): number of atoms
        """
        super().__init__()
        self.T, self.B, self.N, self.n_atom = T, B, N, n_atom
        self.global_B = 0
        # the reference asserts dist[range, action] > 0 on the host every call (td.py:101-103), which forces
        # a device sync; keep the check available but off the hot path
        self.check_positive = False

    def forward(self, dist, next_n_dist, action, next_n_action, reward, done, weight, gamma: float, v_min: float,
                v_max: float) -> torch.Tensor:
        """
        Arguments:
            - dist, next_n_dist (:obj:`torch.FloatTensor`): :math:`(B, N, n_atom)`
            - action, next_n_action (:obj:`torch.LongTensor`): :math:`(B, )`
            - reward (:obj:`torch.FloatTensor`): :math:`(T, B)`
            - done (:obj:`torch.FloatTensor`): :math:`(B, )`
            - weight (:obj:`torch.FloatTensor` or None): :math:`(B, )`
            - gamma, v_min, v_max (:obj:`float`)
        Returns:
            - loss (:obj:`torch.Tensor`): shape :math:`(1,)`
            - td_error_per_sample (:obj:`torch.Tensor`): :math:`(B, )`
        """
        for t in (dist, next_n_dist, action, next_n_action, reward, done):
            assert (t.is_cuda)
        if weight is not None:
            assert (weight.is_cuda)
        if self.check_positive:
            batch_range = torch.arange(action.shape[0], device=action.device)
            assert (dist[batch_range, action] > 0.0).all(), ("dist act", dist[batch_range, action], "dist:", dist)
        fast = _ext.fast()
        if fast is not None:
            return tuple(fast.dist_nstep_td(dist, next_n_dist, action, next_n_action, reward, done, weight,
                                            float(gamma), float(v_min), float(v_max), int(self.global_B)))
        return DistNStepTDFunction.apply(dist, next_n_dist, action, next_n_action, reward, done, weight, gamma, v_min,
                                         v_max, self.global_B)


# ------------------------------------------------------------------------------------------- QR-DQN
class QRDQNNStepTDErrorFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, q, next_n_q, action, next_n_action, reward, done, weight, value_gamma, gamma, global_B):
        q = _abi.require_f32_cuda("q", q)
        next_n_q = _abi.require_f32_cuda("next_n_q", next_n_q)
        B, N, tau = q.shape
        if next_n_q.shape != (B, N, tau):
            raise ValueError("next_n_q must match q")
        action, next_n_action, reward, done, weight = _nstep_common(q, action, next_n_action, reward, done, weight, B)
        if value_gamma is not None:
            value_gamma = _abi.require_f32_cuda("value_gamma", value_gamma)
            if value_gamma.shape != (B, ):
                raise ValueError("value_gamma must be (B,)")
        T = reward.shape[0]
        dev = q.device
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        td_err = torch.empty(B, dtype=torch.float32, device=dev)
        grad_buf = torch.empty((B, tau), dtype=torch.float32, device=dev)
        ws = _abi.workspace(_abi.OP_QRDQN_NSTEP_TD, T, B, N, dev)
        with _abi.on_device(dev):
            _abi.check(
                _abi.lib().hpc_rll_qrdqn_nstep_td_forward(_abi.ptr(q), _abi.ptr(next_n_q), _abi.ptr(action),
                                                          _abi.ptr(next_n_action), _abi.ptr(reward), _abi.ptr(done),
                                                          _abi.ptr(weight), _abi.ptr(value_gamma), _abi.ptr(loss),
                                                          _abi.ptr(td_err), _abi.ptr(grad_buf), tau, T, B, N,
                                                          float(gamma), int(global_B), _abi.ptr(ws), ws.numel(),
                                                          _abi.stream_of(q)), "hpc_rll_qrdqn_nstep_td_forward")
        ctx.save_for_backward(grad_buf, action)
        ctx.N = N
        ctx.mark_non_differentiable(td_err)
        return loss, td_err

    @staticmethod
    def backward(ctx, grad_loss, grad_td_err):
        grad_buf, action = ctx.saved_tensors
        B, tau = grad_buf.shape
        N = ctx.N
        g = _abi.grad_scalar(grad_loss, grad_buf)
        grad_q = torch.empty((B, N, tau), dtype=torch.float32, device=grad_buf.device)
        with _abi.on_device(grad_buf.device):
            _abi.check(
                _abi.lib().hpc_rll_qrdqn_nstep_td_backward(_abi.ptr(g), _abi.ptr(grad_buf), _abi.ptr(action),
                                                           _abi.ptr(grad_q), tau, B, N, _abi.stream_of(grad_buf)),
                "hpc_rll_qrdqn_nstep_td_backward")
        return (grad_q, ) + (None, ) * 9


class QRDQNNStepTDError(torch.nn.Module):
    """
    Overview:
        Multistep td_error in QR-DQN (hpc_rll/origin/td.py:455-517).  As in the reference wrapper, the
        quantile weight uses the integer quantile COUNT ``tau`` (hpc_rll/rl_utils/td.py:538,
        qrdqn_nstep_td_error_kernel.h:60), not a tensor of fractions.

    Interface:
        __init__, forward
    """

    def __init__(self, tau, T, B, N):
        r"""
        Arguments:
            - tau (:obj:`int`): num of quantiles
            - T (:obj:`int`): nstep
            - B (:obj:`int`): batch size
            - N (:obj:`int`): action dim
        """
        super().__init__()
        self.tau, self.T, self.B, self.N = tau, T, B, N
        self.global_B = 0

    def forward(self, q, next_n_q, action, next_n_action, reward, done, gamma: float,
                weight: Optional[torch.Tensor] = None, value_gamma: Optional[torch.Tensor] = None) -> torch.Tensor:
        """
        Arguments:
            - q, next_n_q (:obj:`torch.FloatTensor`): :math:`(B, N, tau)`
            - action, next_n_action (:obj:`torch.LongTensor`): :math:`(B, )`
            - reward (:obj:`torch.FloatTensor`): :math:`(T, B)`
            - done (:obj:`torch.FloatTensor`): :math:`(B, )`
            - gamma (:obj:`float`)
            - weight, value_gamma (:obj:`torch.FloatTensor` or None): :math:`(B, )`
        Returns:
            - loss (:obj:`torch.Tensor`): shape :math:`(1,)`;  td_error_per_sample :math:`(B, )`
        """
        for t in (q, next_n_q, action, next_n_action, reward, done):
            assert (t.is_cuda)
        if weight is not None:
            assert (weight.is_cuda)
        if value_gamma is not None:
            assert (value_gamma.is_cuda)
        fast = _ext.fast()
        if fast is not None:
            return tuple(fast.qrdqn_nstep_td(q, next_n_q, action, next_n_action, reward, done, weight, value_gamma,
                                             float(gamma), int(self.global_B)))
        return QRDQNNStepTDErrorFunction.apply(q, next_n_q, action, next_n_action, reward, done, weight, value_gamma,
                                               gamma, self.global_B)


# ------------------------------------------------------------------------------------------- IQN
class IQNNStepTDErrorFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, q, next_n_q, action, next_n_action, reward, done, replay_quantiles, weight, value_gamma, gamma,
                kappa, global_B):
        q = _abi.require_f32_cuda("q", q)
        next_n_q = _abi.require_f32_cuda("next_n_q", next_n_q)
        replay_quantiles = _abi.require_f32_cuda("replay_quantiles", replay_quantiles)
        tau, B, N = q.shape
        tau_p = next_n_q.shape[0]
        if next_n_q.shape != (tau_p, B, N):
            raise ValueError("next_n_q must be (tau', B, N)")
        if replay_quantiles.numel() != tau * B:
            raise ValueError("replay_quantiles must hold tau*B values")
        action, next_n_action, reward, done, weight = _nstep_common(q, action, next_n_action, reward, done, weight, B)
        if value_gamma is not None:
            value_gamma = _abi.require_f32_cuda("value_gamma", value_gamma)
            if value_gamma.shape != (B, ):
                raise ValueError("value_gamma must be (B,)")
        T = reward.shape[0]
        dev = q.device
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        td_err = torch.empty(B, dtype=torch.float32, device=dev)
        grad_buf = torch.empty((tau, B), dtype=torch.float32, device=dev)
        ws = _abi.workspace(_abi.OP_IQN_NSTEP_TD, T, B, N, dev)
        with _abi.on_device(dev):
            _abi.check(
                _abi.lib().hpc_rll_iqn_nstep_td_forward(_abi.ptr(q), _abi.ptr(next_n_q), _abi.ptr(action),
                                                        _abi.ptr(next_n_action), _abi.ptr(reward), _abi.ptr(done),
                                                        _abi.ptr(replay_quantiles), _abi.ptr(weight),
                                                        _abi.ptr(value_gamma), _abi.ptr(loss), _abi.ptr(td_err),
                                                        _abi.ptr(grad_buf), tau, tau_p, T, B, N, float(gamma),
                                                        float(kappa), int(global_B), _abi.ptr(ws), ws.numel(),
                                                        _abi.stream_of(q)), "hpc_rll_iqn_nstep_td_forward")
        ctx.save_for_backward(grad_buf, action)
        ctx.N = N
        ctx.mark_non_differentiable(td_err)
        return loss, td_err

    @staticmethod
    def backward(ctx, grad_loss, grad_td_err):
        grad_buf, action = ctx.saved_tensors
        tau, B = grad_buf.shape
        N = ctx.N
        g = _abi.grad_scalar(grad_loss, grad_buf)
        grad_q = torch.empty((tau, B, N), dtype=torch.float32, device=grad_buf.device)
        with _abi.on_device(grad_buf.device):
            _abi.check(
                _abi.lib().hpc_rll_iqn_nstep_td_backward(_abi.ptr(g), _abi.ptr(grad_buf), _abi.ptr(action),
                                                         _abi.ptr(grad_q), tau, B, N, _abi.stream_of(grad_buf)),
                "hpc_rll_iqn_nstep_td_backward")
        return (grad_q, ) + (None, ) * 11


class IQNNStepTDError(torch.nn.Module):
    """
    Overview:
        Multistep td_error in IQN (arXiv:1806.06923), hpc_rll/origin/td.py:361-448.

    Interface:
        __init__, forward
    """

    def __init__(self, tau, tauPrime, T, B, N):
        r"""
        Arguments:
            - tau (:obj:`int`): num of quantiles
            - tauPrime (:obj:`int`): num of target quantiles
            - T (:obj:`int`): nstep
            - B (:obj:`int`): batch size
            - N (:obj:`int`): action dim
        """
        super().__init__()
        self.tau, self.tauPrime, self.T, self.B, self.N = tau, tauPrime, T, B, N
        self.global_B = 0

    def forward(self, q, next_n_q, action, next_n_action, reward, done, replay_quantiles, gamma: float,
                kappa: float = 1.0, weight: Optional[torch.Tensor] = None,
                value_gamma: Optional[torch.Tensor] = None) -> torch.Tensor:
        """
        Arguments:
            - q (:obj:`torch.FloatTensor`): :math:`(tau, B, N)`
            - next_n_q (:obj:`torch.FloatTensor`): :math:`(tau', B, N)`
            - action, next_n_action (:obj:`torch.LongTensor`): :math:`(B, )`
            - reward (:obj:`torch.FloatTensor`): :math:`(T, B)`
            - done (:obj:`torch.FloatTensor`): :math:`(B, )`
            - replay_quantiles (:obj:`torch.FloatTensor`): :math:`(tau, B)`
            - gamma, kappa (:obj:`float`)
            - weight, value_gamma (:obj:`torch.FloatTensor` or None): :math:`(B, )`
        Returns:
            - loss (:obj:`torch.Tensor`): shape :math:`(1,)`;  td_error_per_sample :math:`(B, )`
        """
        for t in (q, next_n_q, action, next_n_action, reward, done, replay_quantiles):
            assert (t.is_cuda)
        if weight is not None:
            assert (weight.is_cuda)
        if value_gamma is not None:
            assert (value_gamma.is_cuda)
        fast = _ext.fast()
        if fast is not None:
            return tuple(fast.iqn_nstep_td(q, next_n_q, action, next_n_action, reward, done, replay_quantiles, weight,
                                           value_gamma, float(gamma), float(kappa), int(self.global_B)))
        return IQNNStepTDErrorFunction.apply(q, next_n_q, action, next_n_action, reward, done, replay_quantiles,
                                             weight, value_gamma, gamma, kappa, self.global_B)
