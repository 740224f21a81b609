"""Fused Adam over flat parameter packs (crk_adam_step): torch.optim.Adam semantics
(crank/net/trainer/utils.py:43 -> Adam(lr) with default betas/eps, no weight decay / amsgrad).
Each network owns a handful of flat parameters (see parallel_wavegan.models), so a step is a
handful of launches instead of the reference's per-tensor loop over 200+ tiny tensors."""

import torch

from ... import ops


class FusedAdam(torch.optim.Optimizer):
    """capturable=True keeps the step count in device memory (crk_adam_step_dev) so that the optimizer step can be
    captured in a CUDA graph; the default passes it as a host integer (crk_adam_step)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, capturable=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.capturable = bool(capturable)

    def state_dict(self):
        # capturable mode advances the step count on the device (graph replays never run this Python): mirror it into the host
        # field the checkpoint carries, so that a resumed run continues with the right bias corrections
        for st in self.state.values():
            if "step_dev" in st:
                st["step"] = int(st["step_dev"].item())
        return super().state_dict()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                if self.capturable:
                    if "step_dev" not in st:        # (created outside any capture: the first steps run eagerly)
                        st["step_dev"] = torch.full((1,), int(st["step"]), dtype=torch.int64, device=p.device)
                    ops.adam_step_dev(p.data, g, st["exp_avg"], st["exp_avg_sq"], float(group["lr"]), b1, b2,
                                      group["eps"], st["step_dev"])
                else:
                    st["step"] += 1
                    ops.adam_step(p.data, g, st["exp_avg"], st["exp_avg_sq"], float(group["lr"]), b1, b2,
                                  group["eps"], st["step"])
                # the kernel wrote through the raw pointer: bump the autograd version so that the
                # weight-norm caches keyed on it (parallel_wavegan.models) are invalidated
                torch.autograd.graph.increment_version(p)
        return loss


class FusedRAdam(torch.optim.Optimizer):
    """torch_optimizer.RAdam as built at crank/net/trainer/utils.py:44-45 (crk_radam_step)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                st["step"] += 1
                ops.radam_step(p.data, g, st["exp_avg"], st["exp_avg_sq"], float(group["lr"]), b1, b2, group["eps"], st["step"])
                torch.autograd.graph.increment_version(p)
        return loss


class FusedLamb(torch.optim.Optimizer):
    """pytorch_lamb.Lamb as built at crank/net/trainer/utils.py:46-47 (crk_lamb_step).  The trust ratio is computed per
    parameter tensor of the REFERENCE: a flat pack carries its (offset, length) segments in `param._crk_segments`
    (set by parallel_wavegan.models), any other parameter is one segment."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    segs = getattr(p, "_crk_segments", None) or [(0, p.numel())]
                    st["seg_off"] = torch.tensor([o for o, _ in segs], dtype=torch.int64, device=p.device)
                    st["seg_len"] = torch.tensor([n for _, n in segs], dtype=torch.int64, device=p.device)
                    st["upd"] = torch.empty(p.numel(), dtype=torch.float32, device=p.device)
                    st["trust"] = torch.empty(len(segs), dtype=torch.float32, device=p.device)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                st["step"] += 1
                ops.lamb_step(p.data, g, st["exp_avg"], st["exp_avg_sq"], st["upd"], st["seg_off"], st["seg_len"], st["trust"],
                              float(group["lr"]), b1, b2, group["eps"])
                torch.autograd.graph.increment_version(p)
        return loss
