"""Data-parallel context (one process per GPU, torch.distributed / NCCL over NVLink).

The reference is single-device (crank/bin/train.py:158-159); N-GPU training here is defined as
"N ranks x per-rank batch b == reference 1 device x batch N*b" (SURVEY.md section 8e):
  * gradients: one all-reduce (sum, pre-scaled by 1/N) of each sub-model's flat gradient bucket
    per `step_model`;
  * VQ EMA statistics [counts | per-code sums] are all-reduced BEFORE the EMA normalisation so the
    codebooks stay identical to the big-batch run (crank/net/module/vqvae2.py:316-330).
"""

import torch
import torch.distributed as dist

_state = {"enabled": False, "group": None}


def enable(group=None):
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _state["enabled"] = True
    _state["group"] = group


def disable():
    _state["enabled"] = False
    _state["group"] = None


def world_size():
    if _state["enabled"] and dist.is_initialized():
        return dist.get_world_size(_state["group"])
    return 1


def active():
    return world_size() > 1


def all_reduce_sum(t):
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_state["group"])
    return t


def stats_reducer():
    return all_reduce_sum if active() else None


def average_gradients(params):
    """All-reduce the gradients of `params` as ONE flat bucket and divide by the world size."""
    if not active():
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    all_reduce_sum(flat)
    flat.mul_(1.0 / world_size())
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off : off + n].view_as(g))
        off += n
