"""PPO -- drop-in for /root/reference/hpc_rll/rl_utils/ppo.py (PPO, PPOFunction, hpc_ppo_loss,
hpc_ppo_info).  Same constructor ``PPO(B, N)`` and forward signature (ppo.py:90-148).  approx_kl and
clipfrac are returned as Python floats like the reference, but fetched with ONE device->host copy
(the reference issues two ``.item()`` syncs, ppo.py:148)."""
import os
from collections import namedtuple
from typing import Optional

import torch

from .. import _abi, _ext

hpc_ppo_loss = namedtuple('hpc_ppo_loss', ['policy_loss', 'value_loss', 'entropy_loss'])
hpc_ppo_info = namedtuple('hpc_ppo_info', ['approx_kl', 'clipfrac'])


class LazyScalar:
    """A device-computed scalar that is only brought to the host when somebody reads it.

    ``PPO.forward`` returns ``approx_kl`` / ``clipfrac`` as Python floats like the reference (rl_utils/ppo.py:148,
    two ``.item()`` syncs there, one D2H copy here) -- which still blocks the host once per call.  With
    ``PPO.lazy_info = True`` the two values come back as ``LazyScalar``: the 8-byte device->host copy is queued on the
    current stream into pinned memory next to an event, and the host waits for it only in ``float(x)`` / formatting /
    arithmetic / comparison.  A training loop that logs the info every k steps never stalls the launch queue."""
    __slots__ = ("_host", "_idx", "_event", "_value")

    def __init__(self, host, idx, event):
        self._host, self._idx, self._event, self._value = host, idx, event, None

    def item(self) -> float:
        if self._value is None:
            self._event.synchronize()
            self._value = float(self._host[self._idx])
        return self._value

    __float__ = item

    def ready(self) -> bool:
        return self._value is not None or self._event.query()

    def __repr__(self):
        return repr(self.item())

    def __format__(self, spec):
        return format(self.item(), spec)

    def __bool__(self):
        return bool(self.item())

    def __eq__(self, o):
        return self.item() == float(o)

    def __lt__(self, o):
        return self.item() < float(o)

    def __le__(self, o):
        return self.item() <= float(o)

    def __gt__(self, o):
        return self.item() > float(o)

    def __ge__(self, o):
        return self.item() >= float(o)

    def __hash__(self):
        return hash(self.item())

    def __add__(self, o):
        return self.item() + float(o)

    __radd__ = __add__

    def __sub__(self, o):
        return self.item() - float(o)

    def __rsub__(self, o):
        return float(o) - self.item()

    def __mul__(self, o):
        return self.item() * float(o)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self.item() / float(o)

    def __rtruediv__(self, o):
        return float(o) / self.item()

    def __neg__(self):
        return -self.item()

    def __abs__(self):
        return abs(self.item())


def _lazy_info(info):
    host = torch.empty(2, dtype=torch.float32, pin_memory=True)  # caching host allocator: no cudaHostAlloc per call
    host.copy_(info, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(info.device))
    return hpc_ppo_info(LazyScalar(host, 0, ev), LazyScalar(host, 1, ev))


class PPOFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, logits_new, logits_old, action, value_new, value_old, adv, return_, weight, clip_ratio,
                use_value_clip, dual_clip, global_B, adv_stats=None):
        logits_new = _abi.require_f32_cuda("logits_new", logits_new)
        logits_old = _abi.require_f32_cuda("logits_old", logits_old)
        action = _abi.require_i64_cuda("action", action)
        value_new = _abi.require_f32_cuda("value_new", value_new)
        value_old = _abi.require_f32_cuda("value_old", value_old)
        adv = _abi.require_f32_cuda("adv", adv)
        return_ = _abi.require_f32_cuda("return_", return_)
        B, N = logits_new.shape
        if logits_old.shape != (B, N) or any(t.shape != (B, ) for t in (action, value_new, value_old, adv, return_)):
            raise ValueError("ppo: inconsistent shapes")
        if weight is not None:
            weight = _abi.require_f32_cuda("weight", weight)
            if weight.shape != (B, ):
                raise ValueError("weight must be (B,)")
        if adv_stats is not None:
            adv_stats = _abi.require_f32_cuda("adv_stats", adv_stats)
            if adv_stats.shape != (2, ):
                raise ValueError("adv_stats must be (2,) = [mean, std + 1e-8]")
        dev = adv.device
        out = torch.empty(5, dtype=torch.float32, device=dev)
        pol_coef = torch.empty(B, dtype=torch.float32, device=dev)
        val_coef = torch.empty(B, dtype=torch.float32, device=dev)
        ws = _abi.workspace(_abi.OP_PPO, 0, B, N, dev)
        tail = (_abi.ptr(out), _abi.ptr(pol_coef), _abi.ptr(val_coef), B, N, float(clip_ratio),
                1 if use_value_clip else 0, -1.0 if dual_clip is None else float(dual_clip), int(global_B),
                _abi.ptr(ws), ws.numel(), _abi.stream_of(adv))
        head = (_abi.ptr(logits_new), _abi.ptr(logits_old), _abi.ptr(action), _abi.ptr(value_new),
                _abi.ptr(value_old), _abi.ptr(adv), _abi.ptr(return_), _abi.ptr(weight))
        with _abi.on_device(dev):
            if adv_stats is None:
                _abi.check(_abi.lib().hpc_rll_ppo_forward(*head, *tail), "hpc_rll_ppo_forward")
            else:
                _abi.check(_abi.lib().hpc_rll_ppo_forward_norm(*head, _abi.ptr(adv_stats), *tail),
                           "hpc_rll_ppo_forward_norm")
        ctx.save_for_backward(logits_new, action, weight, pol_coef, val_coef)
        ctx.global_B = int(global_B)
        info = out[3:5]
        ctx.mark_non_differentiable(info)
        return out[0:1], out[1:2], out[2:3], info

    @staticmethod
    def backward(ctx, grad_policy_loss, grad_value_loss, grad_entropy_loss, grad_info):
        logits_new, action, weight, pol_coef, val_coef = ctx.saved_tensors
        B, N = logits_new.shape
        g_p = _abi.grad_scalar(grad_policy_loss, pol_coef)
        g_v = _abi.grad_scalar(grad_value_loss, pol_coef)
        g_e = _abi.grad_scalar(grad_entropy_loss, pol_coef)
        grad_logits = torch.empty_like(logits_new)
        grad_value = torch.empty(B, dtype=torch.float32, device=pol_coef.device)
        with _abi.on_device(pol_coef.device):
            _abi.check(
                _abi.lib().hpc_rll_ppo_backward(_abi.ptr(g_p), _abi.ptr(g_v), _abi.ptr(g_e), _abi.ptr(logits_new),
                                                _abi.ptr(action), _abi.ptr(weight), _abi.ptr(pol_coef),
                                                _abi.ptr(val_coef), _abi.ptr(grad_logits), _abi.ptr(grad_value), B, N,
                                                ctx.global_B, _abi.stream_of(pol_coef)), "hpc_rll_ppo_backward")
        return grad_logits, None, None, grad_value, None, None, None, None, None, None, None, None, None


class PPO(torch.nn.Module):
    """
    Overview:
        Proximal Policy Optimization losses with value clip and dual clip (hpc_rll/origin/ppo.py:13-80).

    Interface:
        __init__, forward
    """

    def __init__(self, B, N):
        r"""
        Arguments:
            - B (:obj:`int`): batch size
            - N (:obj:`int`): number of output
        """
        super().__init__()
        self.B, self.N = B, N
        self.global_B = 0
        # False: approx_kl / clipfrac are Python floats as in the reference (one blocking D2H copy per call);
        # True: LazyScalar objects that synchronise only when read
        self.lazy_info = os.environ.get("HPC_RLL_PPO_LAZY_INFO", "0") == "1"

    def forward(self, logits_new, logits_old, action, value_new, value_old, adv, return_, weight=None,
                clip_ratio: float = 0.2, use_value_clip: bool = True, dual_clip: Optional[float] = None,
                *, adv_stats: Optional[torch.Tensor] = None):
        """
        Arguments:
            - logits_new, logits_old (:obj:`torch.FloatTensor`): :math:`(B, N)`
            - action (:obj:`torch.LongTensor`): :math:`(B, )`
            - value_new, value_old, adv, return_ (:obj:`torch.FloatTensor`): :math:`(B, )`
            - weight (:obj:`torch.FloatTensor` or :obj:`None`): :math:`(B, )`
            - clip_ratio (:obj:`float`): defaults to 0.2
            - use_value_clip (:obj:`bool`)
            - dual_clip (:obj:`float` or None): must be > 1.0 when given
            - adv_stats (:obj:`torch.FloatTensor` or None): extension -- :math:`(2,)` ``[mean, std + 1e-8]`` from
              ``gae_with_adv_stats``; ``adv`` is then normalised inside the kernel (origin/ppo.py:43-47 leaves
              that step to the caller)
        Returns:
            - ppo_loss (:obj:`hpc_ppo_loss`): shape-(1,) tensors
            - ppo_info (:obj:`hpc_ppo_info`): Python floats
        """
        assert (logits_new.is_cuda)
        assert (logits_old.is_cuda)
        assert (action.is_cuda)
        assert (value_new.is_cuda)
        assert (value_old.is_cuda)
        assert (adv.is_cuda)
        assert (return_.is_cuda)
        if weight is not None:
            assert (weight.is_cuda)
        assert dual_clip is None or dual_clip > 1.0, \
            "dual_clip value must be greater than 1.0, but get value: {}".format(dual_clip)
        fast = _ext.fast()
        if fast is not None:
            policy_loss, value_loss, entropy_loss, info = fast.ppo(
                logits_new, logits_old, action, value_new, value_old, adv, return_, weight, float(clip_ratio),
                bool(use_value_clip), -1.0 if dual_clip is None else float(dual_clip), int(self.global_B), adv_stats)
        else:
            policy_loss, value_loss, entropy_loss, info = PPOFunction.apply(logits_new, logits_old, action, value_new,
                                                                            value_old, adv, return_, weight,
                                                                            clip_ratio, use_value_clip, dual_clip,
                                                                            self.global_B, adv_stats)
        if self.lazy_info:
            return hpc_ppo_loss(policy_loss, value_loss, entropy_loss), _lazy_info(info)
        approx_kl, clipfrac = info.tolist()  # one device->host copy
        return hpc_ppo_loss(policy_loss, value_loss, entropy_loss), hpc_ppo_info(approx_kl, clipfrac)
