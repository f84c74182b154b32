"""Data-parallel check of the loss ops on real GPUs over NCCL: every rank holds a contiguous
column shard, kernels normalise by the GLOBAL count (module.global_B), per-rank losses are summed with one
all_reduce, gradients need no collective.  Rank 0 compares against the single-GPU full-batch result.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multi_gpu.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from di_hpc_b200.sharding import all_reduce_losses, local_shard, set_global_batch, shard_columns  # noqa: E402
from hpc_rll.rl_utils.gae import GAE  # noqa: E402
from hpc_rll.rl_utils.td import TDLambda  # noqa: E402
from hpc_rll.rl_utils.vtrace import VTrace  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    T, B, N = 64, 4096, 8
    g = torch.Generator(device="cuda").manual_seed(7)  # same seed on every rank: identical global tensors
    tgt, beh = torch.randn(T, B, N, device="cuda", generator=g), torch.randn(T, B, N, device="cuda", generator=g)
    act = torch.randint(0, N, (T, B), device="cuda", generator=g)
    val, rew = torch.randn(T + 1, B, device="cuda", generator=g), torch.randn(T, B, device="cuda", generator=g)
    one = torch.ones(1, device="cuda")
    out = {}

    # ---- sharded run -------------------------------------------------------------------------------------
    b0, b1 = shard_columns(B, rank, world)
    t_l = local_shard(tgt, rank, world, dim=1).requires_grad_(True)
    v_l = local_shard(val, rank, world, dim=1).requires_grad_(True)
    m = set_global_batch(VTrace(T, b1 - b0, N), B)
    l = m(t_l, local_shard(beh, rank, world, dim=1), local_shard(act, rank, world, dim=1), v_l,
          local_shard(rew, rank, world, dim=1))
    pg, vl, ent = all_reduce_losses([l.policy_loss, l.value_loss, l.entropy_loss])
    gt_l, gv_l = torch.autograd.grad(pg + 0.5 * vl - 0.01 * ent, [t_l, v_l], grad_outputs=one)
    td = set_global_batch(TDLambda(T, b1 - b0), B)
    (tdl, ) = all_reduce_losses([td(v_l.detach().requires_grad_(True), local_shard(rew, rank, world, dim=1))])
    adv_l = GAE(T, b1 - b0)(v_l.detach(), local_shard(rew, rank, world, dim=1))  # no collective at all

    # ---- full-batch reference on every rank (cheap at this size) -----------------------------------------
    t_f, v_f = tgt.clone().requires_grad_(True), val.clone().requires_grad_(True)
    lf = VTrace(T, B, N)(t_f, beh, act, v_f, rew)
    gt_f, gv_f = torch.autograd.grad(lf.policy_loss + 0.5 * lf.value_loss - 0.01 * lf.entropy_loss, [t_f, v_f],
                                     grad_outputs=one)
    tdf = TDLambda(T, B)(val, rew)
    adv_f = GAE(T, B)(val, rew)

    def rel(a, b):
        return float((a.detach().double() - b.detach().double()).abs().max() / max(1.0, float(b.detach().double().abs().max())))

    out["vtrace_losses"] = max(rel(pg, lf.policy_loss), rel(vl, lf.value_loss), rel(ent, lf.entropy_loss))
    out["vtrace_grad_target"] = rel(gt_l, gt_f[:, b0:b1])
    out["vtrace_grad_value"] = rel(gv_l, gv_f[:, b0:b1])
    out["td_lambda_loss"] = rel(tdl, tdf)
    out["gae_shard_bitexact"] = bool(torch.equal(adv_l, adv_f[:, b0:b1]))
    worst = torch.tensor([max(v for v in out.values() if not isinstance(v, bool))], device="cuda")
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    ok = torch.tensor([1 if out["gae_shard_bitexact"] else 0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        res = dict(world=world, T=T, B=B, N=N, max_rel_err_over_ranks=float(worst.item()),
                   gae_bitexact_all_ranks=bool(ok.item()), rank0=out)
        print(json.dumps(res))
        assert worst.item() <= 1e-5 and ok.item() == 1
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
