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

import os

_state = {"enabled": False, "group": None, "weights": {}, "comm_stream": None, "pending": {},
          "overlap": os.environ.get("CRANK_B200_DP_OVERLAP", "1") != "0"}

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
    if torch.cuda.is_available():
        flush()
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
    """All-reduce the gradients of `params` IN PLACE and divide by the world size.  A network's gradient is one flat
    tensor per parameter pack (`theta.grad`, a few MB) plus a handful of small extras (embeddings): with NCCL they go
    out as ONE coalesced group call (ncclGroupStart/End: a single fused kernel, no cat / copy-back staging); other
    backends (gloo in the CPU tests) reduce them one by one."""
    if not active():
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    scale = 1.0 / world_size()
    coalesce = getattr(dist, "_coalescing_manager", None)
    if grads[0].is_cuda and len(grads) > 1 and coalesce is not None and dist.get_backend(_state["group"]) == "nccl":
        try:
            with coalesce(group=_state["group"], device=grads[0].device, async_ops=False):
                for g in grads:
                    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=_state["group"])
        except TypeError:            # (older signature): plain per-tensor calls
            for g in grads:
                all_reduce_sum(g)
    else:
        for g in grads:
            all_reduce_sum(g)
    torch._foreach_mul_(grads, scale)


# ---- overlap: collectives (and what depends only on them) run on a side stream ----------------------------------
# `run_async(keys, fn)` executes fn() -- all-reduce + clip + Adam of one sub-model, or the EMA-statistics all-reduce +
# codebook update of one quantiser -- on the communication stream, after everything the main stream has issued so
# far, and remembers a completion event under every key (a parameter / buffer).  Whoever reads such a tensor next calls
# `wait_for(...)` first (the packed networks do it where they fetch their effective weights, VQVAE2 / Quantizer at the
# top of forward).  So the discriminator's gradient all-reduce + update overlaps the generator forward that follows it,
# the classifier's overlaps the next step, and none of the eight 133 KB EMA-statistics all-reduces of an LSGAN step
# sits on the compute stream any more.  Inside a CUDA-graph capture, on CPU tensors, or with CRANK_B200_DP_OVERLAP=0
# everything runs inline (same results: only the stream changes).
def _overlap_ok(sample):
    return (active() and _state["overlap"] and sample is not None and sample.is_cuda
            and not torch.cuda.is_current_stream_capturing())


def run_async(keys, fn, tensors=()):
    keys = list(keys)
    sample = keys[0] if keys else None
    if not _overlap_ok(sample):
        fn()
        return
    if _state["comm_stream"] is None:
        _state["comm_stream"] = torch.cuda.Stream()
    cs = _state["comm_stream"]
    cs.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cs):
        fn()
        ev = torch.cuda.Event()
        ev.record(cs)
    for t in tensors:                       # allocator: these were allocated on the main stream, last used on `cs`
        if isinstance(t, torch.Tensor) and t.is_cuda:
            t.record_stream(cs)
    for k in keys:
        _state["pending"][id(k)] = (k, ev)


def wait_for(tensors):
    """Make the current stream wait for any pending side-stream update of these tensors."""
    pend = _state["pending"]
    if not pend:
        return
    for t in tensors:
        hit = pend.pop(id(t), None)
        if hit is not None:
            torch.cuda.current_stream().wait_event(hit[1])


def flush():
    """Wait (on the current stream) for every pending side-stream update: before checkpoints / state_dict reads."""
    pend = _state["pending"]
    for _, ev in list(pend.values()):
        torch.cuda.current_stream().wait_event(ev)
    pend.clear()


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
