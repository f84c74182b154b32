"""Batch-axis sharding of the trajectory-return path across the GPUs of one box.

Every op of the path is independent per column/sample; the only cross-sample coupling is the final
``mean`` of the loss ops (and the 1/count in their gradients).  So the batch axis B shards trivially:
one process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch), rank r owns columns
``[r*B/W, (r+1)*B/W)``, kernels normalise by the GLOBAL element count (``module.global_B``) and the
only collective is one all-reduce(SUM) of the few loss scalars.  GAE has no collective at all.
The reference has no multi-GPU story; this module is the whole of ours.
"""
from typing import Sequence, Tuple

import torch
import torch.distributed as dist


def shard_columns(B: int, rank: int, world: int, align: int = 4) -> Tuple[int, int]:
    """Column range [b0, b1) of ``rank``.  Shard starts are multiples of ``align`` columns (16 bytes of
    fp32) so every shard of a (T,B) tensor can be described by a TMA tensor map."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world %d/%d" % (rank, world))
    per = -(-B // world)
    per = -(-per // align) * align
    b0 = min(B, rank * per)
    b1 = min(B, b0 + per)
    return b0, b1


def local_shard(x: torch.Tensor, rank: int, world: int, dim: int = -1) -> torch.Tensor:
    """Contiguous copy of this rank's column block of a global tensor (time-major tensors are
    B-innermost, so a column block of a global tensor is strided; real data-parallel jobs already hold
    their own contiguous shard and never call this)."""
    b0, b1 = shard_columns(x.shape[dim], rank, world)
    return x.narrow(dim, b0, b1 - b0).contiguous()


class _AllReduceSum(torch.autograd.Function):
    """loss_global = sum over ranks of loss_local; d loss_global / d loss_local = 1 on every rank."""

    @staticmethod
    def forward(ctx, x, group):
        y = x.clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


def all_reduce_losses(losses: Sequence[torch.Tensor], group=None):
    """Sum per-rank partial losses (already divided by the global count) into the global losses.
    Differentiable: gradients flow back to the local losses unchanged.  Packs the scalars into one
    tensor so a step costs a single latency-bound NCCL call."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(losses)
    flat = torch.cat([l.reshape(-1) for l in losses])
    red = _AllReduceSum.apply(flat, group)
    out, o = [], 0
    for l in losses:
        n = l.numel()
        out.append(red[o:o + n].reshape(l.shape))
        o += n
    return out


def all_reduce_moments(moments: torch.Tensor, group=None) -> torch.Tensor:
    """In-place all-reduce(SUM) of ``[sum(adv), sum(adv^2), count]`` (fp64) over the ranks that shard the
    batch: the element count rides in the same tensor, so the global mean/std need one latency-bound
    collective and no device->host read."""
    if moments.dtype != torch.float64 or moments.numel() != 3:
        raise TypeError("moments must be 3 float64 values [sum, sum of squares, count]")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(moments, op=dist.ReduceOp.SUM, group=group)
    return moments


def set_global_batch(module: torch.nn.Module, global_B: int) -> torch.nn.Module:
    """Tell a loss module that its batch is one shard of ``global_B`` columns/samples."""
    if not hasattr(module, "global_B"):
        raise TypeError("%s has no global_B (GAE needs none)" % type(module).__name__)
    module.global_B = int(global_B)
    return module
