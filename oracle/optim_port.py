"""TEST INFRASTRUCTURE (oracle): restatement of the two third-party optimizers crank/net/trainer/utils.py:40-58 can build.

Both packages are absent from /root/reference and this image (tools/requirements.txt: torch_optimizer, pytorch_lamb;
un-pinned), so the algorithms are restated from their published sources -- torch_optimizer.RAdam (Liu et al., the
original implementation's arithmetic) and pytorch_lamb.Lamb (cybertronai: no bias correction, weight norm clamped to
[0, 10], trust ratio 1 when either norm is 0).  Parity: unpinned by the reference (it has no optimizer test); used as the
checker of crank_b200.net.trainer.optim.FusedRAdam / FusedLamb."""
import math

import torch


class RAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self):
        for group in self.param_groups:
            lr, (beta1, beta2), eps = group["lr"], group["betas"], group["eps"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                grad = p.grad.float()
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                exp_avg, exp_avg_sq = st["exp_avg"], st["exp_avg_sq"]
                exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
                exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)
                st["step"] += 1
                step = st["step"]
                beta2_t = beta2 ** step
                n_sma_max = 2 / (1 - beta2) - 1
                n_sma = n_sma_max - 2 * step * beta2_t / (1 - beta2_t)
                if n_sma >= 5:
                    step_size = lr * math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_sma_max - 4) * (n_sma - 2) / n_sma
                                               * n_sma_max / (n_sma_max - 2)) / (1 - beta1 ** step)
                    p.addcdiv_(exp_avg, exp_avg_sq.sqrt().add_(eps), value=-step_size)
                else:
                    step_size = lr / (1 - beta1 ** step)
                    p.add_(exp_avg, alpha=-step_size)


class Lamb(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self):
        for group in self.param_groups:
            lr, (beta1, beta2), eps = group["lr"], group["betas"], group["eps"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                grad = p.grad
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                exp_avg, exp_avg_sq = st["exp_avg"], st["exp_avg_sq"]
                st["step"] += 1
                exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)
                exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
                weight_norm = p.pow(2).sum().sqrt().clamp(0, 10)
                adam_step = exp_avg / exp_avg_sq.sqrt().add(eps)
                adam_norm = adam_step.pow(2).sum().sqrt()
                trust = 1.0 if (weight_norm == 0 or adam_norm == 0) else float(weight_norm / adam_norm)
                p.add_(adam_step, alpha=-lr * trust)
