"""TEST INFRASTRUCTURE ONLY -- torch-CPU stand-ins for the C-ABI ops, installed by monkeypatching inside tests.

Purpose: exercise the product's HOST logic (the four trainers' loss assembly and update order, VQVAE2 / Quantizer
orchestration incl. the list-mutation quirk, frozen / no-grad passes, flat-parameter packing and the reference-keyed
state dicts, FusedAdam plumbing) on a box without a GPU, for every recipe switch, against the oracle.  The kernels
themselves are shape-generic and are verified on the GPU (`-m gpu`); what differs between config variants is Python.

This is NOT a fallback: nothing in `crank_b200/` imports it, the product raises without CUDA, and the emulation is only
reachable through the `emulated_ops()` context manager below.  Each stand-in restates the math of the op it replaces
(the same definitions the kernels implement; citations in crank_b200/ops.py and include/crank_b200.h).
"""
import contextlib
import math

import torch
import torch.nn.functional as F


def _views(net, theta):
    out = {}
    for name, d in zip(net._names, net._descs):
        g = theta[d.g_off : d.g_off + d.cout].view(d.cout, 1, 1)
        v = theta[d.v_off : d.v_off + d.cout * d.cin * d.k].view(d.cout, d.cin, d.k)
        b = theta[d.b_off : d.b_off + d.cout] if d.b_off >= 0 else None
        out[name] = (v * (g / v.flatten(1).norm(dim=1).view(-1, 1, 1)), b)     # torch._weight_norm(v, g, dim=0)
    return out


def _act(kind, slope):
    return {0: lambda t: t, 1: F.relu, 2: lambda t: F.leaky_relu(t, slope)}[int(kind)]


class WavenetEmu:
    """crk_wavenet_fwd / _bwd (autograd supplies the backward)."""

    @staticmethod
    def apply(net, x, c, dropmul, theta, save_gates=True):
        cfg = net.cfg
        W = _views(net, theta)
        B, T, _ = x.shape
        h = x.transpose(1, 2)
        cond = c.transpose(1, 2) if c is not None else None
        first, head = _act(cfg.first_act, cfg.slope), _act(cfg.head_act, cfg.slope)
        w, b = W[net._names[0]]
        h = first(F.conv1d(h, w, b))
        skips = 0
        lps = cfg.layers // cfg.stacks
        k = cfg.kernel_size
        for l in range(cfg.layers):
            dil = 2 ** (l % lps)
            res = h
            xin = h if dropmul is None else h * dropmul[l].view(B, T, 64).transpose(1, 2)
            w, b = W[f"conv_layers.{l}.conv"]
            if cfg.causal:
                a = F.conv1d(xin, w, b, dilation=dil, padding=(k - 1) * dil)[:, :, :T]
            else:
                a = F.conv1d(xin, w, b, dilation=dil, padding=(k - 1) // 2 * dil)
            if cfg.aux_ch > 0:
                a = a + F.conv1d(cond, W[f"conv_layers.{l}.conv1x1_aux"][0], None)
            xa, xb = a.split(a.size(1) // 2, dim=1)
            z = torch.tanh(xa) * torch.sigmoid(xb)
            s = F.conv1d(z, *W[f"conv_layers.{l}.conv1x1_skip"])
            h = (F.conv1d(z, *W[f"conv_layers.{l}.conv1x1_out"]) + res) * math.sqrt(0.5)
            skips = skips + s
        y = head(skips * math.sqrt(1.0 / cfg.layers))
        y = head(F.conv1d(y, *W["last_conv_layers.1"]))
        y = F.conv1d(y, *W["last_conv_layers.3"])
        return y.transpose(1, 2)


class _ScaleGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, s):
        ctx.s = float(s)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.s, None


class ConvstackEmu:
    """crk_convstack_fwd / _bwd (ParallelWaveGANDiscriminator; dx_scale = gradient reversal)."""

    @staticmethod
    def apply(net, x, theta, grad_scale):
        cfg = net.cfg
        W = _views(net, theta)
        if float(grad_scale) != 1.0:
            x = _ScaleGrad.apply(x, grad_scale)
        h = x.transpose(1, 2)
        k = cfg.kernel_size
        for i in range(cfg.layers):
            last = i == cfg.layers - 1
            dil = 1 if (i == 0 or last) else (i if cfg.dilation_factor == 1 else cfg.dilation_factor ** i)
            w, b = W[f"conv_layers.{2 * i}"]
            h = F.conv1d(h, w, b, dilation=dil, padding=(k - 1) // 2 * dil)
            if not last:
                h = F.leaky_relu(h, cfg.slope)
        return h.transpose(1, 2)


class VQEmu:
    """crk_vq_argmin: reference distance expression, first-minimum ties, gather, straight-through."""

    @staticmethod
    def apply(x, W, opblob=None, variant=None):
        D = W.shape[1]
        flat = x.reshape(-1, D)
        Wd = W.detach()
        dist = torch.sum(Wd ** 2, dim=1) - 2 * torch.matmul(flat.detach(), Wd.T) + torch.sum(flat.detach() ** 2, dim=1, keepdim=True)
        idx = torch.argmin(dist, dim=1).view(x.shape[0], x.shape[1])
        e = F.embedding(idx, W)
        qx = x + (e - x).detach()
        return e, qx, idx


def vq_ema_update(x, idx, ema_size, ema_w, W, decay, eps, reduce_fn=None, runner=None, opblob=None):
    K, D = W.shape
    onehot = F.one_hot(idx.reshape(-1), K).float()
    stats = torch.cat([onehot.sum(0), (x.reshape(-1, D).T @ onehot).reshape(-1)])
    if reduce_fn is not None:
        reduce_fn(stats)
    counts, esum = stats[:K], stats[K:].view(D, K)
    ema_size.copy_(decay * ema_size + (1 - decay) * counts)
    ema_w.copy_(decay * ema_w + (1 - decay) * esum)
    n = ema_size.sum()
    ema_size.copy_((ema_size + eps) / (n + K * eps) * n)
    W.copy_((ema_w / ema_size.unsqueeze(0)).T)


class MaskedLossEmu:
    """crk_masked_loss_fwd / _bwd: (mean |x-y|, mean (x-y)^2, #selected elements); the product's wrapper
    `ops.masked_l1_mse` (with its data-parallel weighting hook) stays in place on top of this."""

    @staticmethod
    def apply(x, y, mask, shift):
        l1, mse, n = _masked_l1_mse(x, y, mask, shift)
        return l1, mse, n


class CrossEntropyEmu:
    @staticmethod
    def apply(logits, labels, ignore_index):
        return (F.cross_entropy(logits.float(), labels, ignore_index=ignore_index),
                (labels != ignore_index).sum().float())


def _masked_l1_mse(x, y, mask=None, shift=0):
    if shift > 0:
        x = x[:, shift:]
        y = y[:, :-shift] if isinstance(y, torch.Tensor) else y
        mask = mask[:, shift:] if mask is not None else None
    elif shift < 0:
        cs = -shift
        x = x[:, :-cs]
        y = y[:, cs:] if isinstance(y, torch.Tensor) else y
        mask = mask[:, :-cs] if mask is not None else None
    if not isinstance(y, torch.Tensor):
        y = torch.full_like(x, float(y))
    if mask is not None:
        x, y = x.masked_select(mask.bool()), y.masked_select(mask.bool())
    return F.l1_loss(x, y), F.mse_loss(x, y), torch.tensor(float(x.numel()))


class StftLossEmu:
    @staticmethod
    def apply(x, y, n_fft, hop, win):
        from oracle.crank_port import stft_mag

        xm, ym = stft_mag(x, n_fft, hop, win), stft_mag(y.detach(), n_fft, hop, win)
        return F.l1_loss(xm, ym), F.l1_loss(xm.log(), ym.log())


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step_count):
    """torch.optim.Adam single-tensor formula (crk_adam_step)."""
    m.lerp_(g, 1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step_count, 1 - beta2 ** step_count
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


def logmel(wav, window, mel_basis, n_fft, hop, eps=1e-10, mean=None, std=None, fused=None):
    """crk_logmel_fwd / crk_logmel_fused_fwd: frames start at m*hop (no centring here), |rFFT(window * frame)| . mel_basis -> log10.
    Differentiable in wav and window: also the stand-in of ops.logmel_learnable (crk_logmel_fused_bwd)."""
    wav = wav.float()
    frames = wav.unfold(-1, n_fft, hop) * window                        # (B, M, n_fft)
    mag = torch.fft.rfft(frames, dim=-1).abs()
    out = torch.log10(torch.clamp(mag @ mel_basis, min=eps))
    if mean is not None:
        out = (out - mean) / std
    return out


def radam_step(p, g, m, v, lr, beta1, beta2, eps, step_count):
    """crk_radam_step: torch_optimizer.RAdam on one flat tensor."""
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    beta2_t = beta2 ** step_count
    n_max = 2 / (1 - beta2) - 1
    n_sma = n_max - 2 * step_count * beta2_t / (1 - beta2_t)
    if n_sma >= 5:
        step_size = lr * math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_max - 4) * (n_sma - 2) / n_sma * n_max / (n_max - 2)) / (1 - beta1 ** step_count)
        p.addcdiv_(m, v.sqrt().add_(eps), value=-step_size)
    else:
        p.add_(m, alpha=-lr / (1 - beta1 ** step_count))


def lamb_step(p, g, m, v, upd, seg_off, seg_len, trust, lr, beta1, beta2, eps):
    """crk_lamb_step: pytorch_lamb.Lamb with one trust ratio per segment of the flat pack."""
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    upd.copy_((m / v.sqrt().add(eps)).reshape(-1))
    pf = p.view(-1)
    for i, (o, n) in enumerate(zip(seg_off.tolist(), seg_len.tolist())):
        wn = pf[o:o + n].pow(2).sum().sqrt().clamp(0, 10)
        an = upd[o:o + n].pow(2).sum().sqrt()
        t = 1.0 if (wn == 0 or an == 0) else float(wn / an)
        trust[i] = t
        pf[o:o + n].add_(upd[o:o + n], alpha=-lr * t)


def adam_step_dev(p, g, m, v, lr, beta1, beta2, eps, step_dev):
    """crk_adam_step_dev: the counter lives in a tensor and is incremented by the call."""
    step_dev += 1
    adam_step(p, g, m, v, lr, beta1, beta2, eps, int(step_dev.item()))


@contextlib.contextmanager
def emulated_ops():
    """Swap the product's op bindings for the CPU stand-ins; everything is restored on exit."""
    from crank_b200 import lib, ops
    from crank_b200.parallel_wavegan import models

    saved = [(models, "WavenetFn", models.WavenetFn), (models, "ConvstackFn", models.ConvstackFn),
             (ops, "VQFn", ops.VQFn), (ops, "vq_ema_update", ops.vq_ema_update),
             (ops, "MaskedLossFn", ops.MaskedLossFn), (ops, "CrossEntropyFn", ops.CrossEntropyFn),
             (ops, "StftLossFn", ops.StftLossFn), (ops, "adam_step", ops.adam_step),
             (ops, "adam_step_dev", ops.adam_step_dev), (ops, "logmel", ops.logmel),
             (ops, "logmel_learnable", ops.logmel_learnable), (ops, "radam_step", ops.radam_step),
             (ops, "lamb_step", ops.lamb_step), (lib, "require_cuda", lib.require_cuda)]
    models.WavenetFn, models.ConvstackFn = WavenetEmu, ConvstackEmu
    ops.VQFn, ops.vq_ema_update = VQEmu, vq_ema_update
    ops.MaskedLossFn, ops.CrossEntropyFn = MaskedLossEmu, CrossEntropyEmu
    ops.StftLossFn, ops.adam_step, ops.adam_step_dev = StftLossEmu, adam_step, adam_step_dev
    ops.logmel = logmel
    ops.logmel_learnable = logmel
    ops.radam_step, ops.lamb_step = radam_step, lamb_step
    lib.require_cuda = lambda *a, **k: None
    try:
        yield
    finally:
        for mod, name, val in saved:
            setattr(mod, name, val)
