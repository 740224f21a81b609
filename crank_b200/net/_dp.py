"""Data-parallel context (one process per GPU, torch.distributed / NCCL over NVLink).

The reference is single-device (crank/bin/train.py:158-159); N-GPU training here is defined as
"N ranks x per-rank batch b == reference 1 device x batch N*b" (SURVEY.md section 8e):
  * gradients: one all-reduce (sum, pre-scaled by 1/N) of each sub-model's flat gradient bucket
    per `step_model`;
  * VQ EMA statistics [counts | per-code sums] are all-reduced BEFORE the EMA normalisation so the
    codebooks stay identical to the big-batch run (crank/net/module/vqvae2.py:316-330);
  * masked means (reconstruction / commitment / LSGAN / cross-entropy terms; the reference selects
    frames with masked_select, crank/net/trainer/trainer_vqvae.py:229-237): the big-batch value is
    sum_r(num_r) / sum_r(count_r), so each rank's local mean is weighted by
    world * count_r / sum_r(count_r) before backward.  With equal
    valid-frame counts the weight is exactly 1.  The per-step weights of the batch's masks / label
    tensors come from ONE small all-reduce (`begin_step`); anything else is reduced on the fly.
  * the loss values reported by `train()` are averaged over ranks (one all-reduce of the packed
    loss vector), i.e. they are the big-batch values.

Ranks must run the same batch shape (B, T): terms that average over all frames (the STFT trajectory loss,
un-masked LSGAN terms) are weighted 1.
"""

import torch
import torch.distributed as dist

_state = {"enabled": False, "group": None, "weights": {}}

# batch entries whose valid-element counts define the masked means of a step
MASK_KEYS = ("encoder_mask", "decoder_mask", "cycle_encoder_mask", "cycle_decoder_mask")
LABEL_KEYS = ("org_h", "cv_h")
IGNORE_INDEX = -100


def enable(group=None):
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _state["enabled"] = True
    _state["group"] = group
    _state["weights"] = {}
    from .. import ops

    ops.set_mean_weight_hook(mean_weight)


def disable():
    _state["enabled"] = False
    _state["group"] = None
    _state["weights"] = {}
    from .. import ops

    ops.set_mean_weight_hook(None)


def world_size():
    if _state["enabled"] and dist.is_initialized():
        return dist.get_world_size(_state["group"])
    return 1


def active():
    return world_size() > 1


def rank():
    if _state["enabled"] and dist.is_initialized():
        return dist.get_rank(_state["group"])
    return 0


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


def _weights_from_counts(local):
    """local: 1-D float tensor of this rank's valid-element counts -> world * local / global (0/0 -> 0)."""
    total = local.clone()
    all_reduce_sum(total)
    return torch.where(total > 0, local * float(world_size()) / total.clamp_min(1.0), torch.zeros_like(local))


def begin_step(batch):
    """Per-step weights of the batch's masks and label tensors with ONE all-reduce.  The tensors are
    kept referenced until the next call so that their addresses identify them during the step."""
    _state["weights"] = {}
    if not active():
        return
    tensors, counts = [], []
    for k in MASK_KEYS:
        t = batch.get(k)
        if isinstance(t, torch.Tensor):
            tensors.append(t)
            counts.append((t != 0).sum())
    for k in LABEL_KEYS:
        t = batch.get(k)
        if isinstance(t, torch.Tensor):
            tensors.append(t)
            counts.append((t != IGNORE_INDEX).sum())
    if not tensors:
        return
    w = _weights_from_counts(torch.stack(counts).float())
    for i, t in enumerate(tensors):
        _state["weights"][(t.data_ptr(), t.numel(), 0)] = (t, w[i])


def mean_weight(key_tensor, shift, count):
    """Weight of a masked mean over `count` local elements selected by `key_tensor` (mask or labels)."""
    if not active():
        return None
    hit = _state["weights"].get((key_tensor.data_ptr(), key_tensor.numel(), int(shift)))
    if hit is not None:
        return hit[1]
    # not one of the batch tensors (sliced labels of the causal mode, shifted masks): reduce its count
    # now -- every rank takes this path for the same call, so the collective is matched
    return _weights_from_counts(count.detach().reshape(1).float())[0]


def average_loss_vector(packed):
    """Rank-average of the packed per-step loss scalars (reporting only)."""
    if active():
        all_reduce_sum(packed)
        packed.mul_(1.0 / world_size())
    return packed
