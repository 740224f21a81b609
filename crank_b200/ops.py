"""torch.autograd bindings of the C-ABI kernels (argument checking + pointer extraction only).

Activations are channels-last (B, T, C) fp32 CUDA tensors; the reference's (B, C, T) layout only
exists at the public module boundaries (crank_b200.parallel_wavegan.models).
"""

import ctypes as C
import math

import torch

from . import lib as L

_f32 = torch.float32


def panel(t):
    """(tensor, ld) for a (B,T,C) fp32 tensor viewed as a (B*T, C) row panel with row stride ld."""
    if t is None:
        return None, 0
    if t.dtype != _f32:
        t = t.float()
    L.require_cuda(t)
    assert t.dim() == 3, "expected (B, T, C)"
    B, T, Cn = t.shape
    if t.stride(2) != 1 or (B > 1 and t.stride(0) != T * t.stride(1)) or t.stride(1) < Cn or (
        t.data_ptr() % 4
    ):
        t = t.contiguous()
    return t, t.stride(1)


def _empty(n, device):
    return torch.empty(max(int(n), 4), dtype=_f32, device=device)


# ---------------------------------------------------------------------------------------------
class WavenetFn(torch.autograd.Function):
    """One WaveNet stack (first 1x1, L gated residual blocks, head) -- crk_wavenet_fwd / _bwd."""

    @staticmethod
    def forward(ctx, net, x, c, dropmul, theta, save_gates=True):
        x, ldx = panel(x)
        c, ldc = panel(c)
        B, T, _ = x.shape
        cfg = net.cfg
        weff = net.effective_weights()
        y = torch.empty(B, T, cfg.out_ch, dtype=_f32, device=x.device)
        act = _empty(L.lib().crk_wavenet_act_floats(C.byref(cfg), B, T), x.device)
        # `save_gates` is decided by the CALLER (forward_cl): inside Function.forward grad mode is always off and
        # needs_input_grad mirrors requires_grad whatever the caller's grad mode, so neither can tell a
        # torch.no_grad() pass from a training pass.  False: inference entry, nothing is saved for backward.
        entry = "crk_wavenet_fwd" if save_gates else "crk_wavenet_infer"
        WavenetFn.last_entry = entry
        L.call(entry, C.byref(cfg), L.ptr(weff), L.ptr(x), ldx, L.ptr(c), ldc,
               L.ptr(dropmul), L.ptr(y), cfg.out_ch, L.ptr(act), B, T)
        ctx.net = net
        ctx.dims = (B, T, ldx, ldc)
        ctx.has_c = c is not None
        ctx.has_drop = dropmul is not None
        saved = [x, theta, weff, act]
        if c is not None:
            saved.append(c)
        if dropmul is not None:
            saved.append(dropmul)
        ctx.save_for_backward(*saved)
        return y

    @staticmethod
    def backward(ctx, dy):
        saved = list(ctx.saved_tensors)
        x, theta, weff, act = saved[:4]
        rest = saved[4:]
        c = rest.pop(0) if ctx.has_c else None
        dropmul = rest.pop(0) if ctx.has_drop else None
        B, T, ldx, ldc = ctx.dims
        cfg = ctx.net.cfg
        dy, lddy = panel(dy)
        need_dx = ctx.needs_input_grad[1]
        need_dc = ctx.has_c and ctx.needs_input_grad[2]
        dx = torch.empty(B, T, cfg.in_ch, dtype=_f32, device=x.device) if need_dx else None
        dc = torch.empty(B, T, cfg.aux_ch, dtype=_f32, device=x.device) if need_dc else None
        # parameters frozen at forward time (e.g. the discriminator inside the generator update): input
        # gradients only -- the library skips every weight-gradient kernel when gtheta is NULL
        gtheta = torch.empty_like(theta) if ctx.needs_input_grad[4] else None
        ws = _empty(L.lib().crk_wavenet_ws_floats(C.byref(cfg), B, T), x.device)
        L.call("crk_wavenet_bwd", C.byref(cfg), L.ptr(theta), L.ptr(weff), L.ptr(x), ldx,
               L.ptr(c), ldc, L.ptr(dropmul), L.ptr(act), L.ptr(dy), lddy,
               L.ptr(dx), cfg.in_ch, L.ptr(dc), max(cfg.aux_ch, 0), L.ptr(gtheta), L.ptr(ws), B, T)
        return None, dx, dc, None, gtheta, None


class ConvstackFn(torch.autograd.Function):
    """Plain Conv1d+LeakyReLU stack -- crk_convstack_fwd / _bwd.  dx_scale folds a gradient
    reversal layer (crank/net/module/spkradv.py:63-72) into the input gradient."""

    @staticmethod
    def forward(ctx, net, x, theta, dx_scale):
        x, ldx = panel(x)
        B, T, _ = x.shape
        cfg = net.cfg
        weff = net.effective_weights()
        y = torch.empty(B, T, cfg.out_ch, dtype=_f32, device=x.device)
        act = _empty(L.lib().crk_convstack_act_floats(C.byref(cfg), B, T), x.device)
        L.call("crk_convstack_fwd", C.byref(cfg), L.ptr(weff), L.ptr(x), ldx, L.ptr(y), cfg.out_ch,
               L.ptr(act), B, T)
        ctx.net = net
        ctx.dims = (B, T, ldx)
        ctx.dx_scale = float(dx_scale)
        ctx.save_for_backward(x, theta, weff, act)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, theta, weff, act = ctx.saved_tensors
        B, T, ldx = ctx.dims
        cfg = ctx.net.cfg
        dy, lddy = panel(dy)
        need_dx = ctx.needs_input_grad[1]
        dx = torch.empty(B, T, cfg.in_ch, dtype=_f32, device=x.device) if need_dx else None
        gtheta = torch.empty_like(theta) if ctx.needs_input_grad[2] else None
        ws = _empty(L.lib().crk_convstack_ws_floats(C.byref(cfg), B, T), x.device)
        L.call("crk_convstack_bwd", C.byref(cfg), L.ptr(theta), L.ptr(weff), L.ptr(x), ldx,
               L.ptr(act), L.ptr(dy), lddy, L.ptr(dx), cfg.in_ch, C.c_float(ctx.dx_scale),
               L.ptr(gtheta), L.ptr(ws), B, T)
        return None, dx, gtheta, None


# ---------------------------------------------------------------------------------------------
def vq_fast_ok(K, D):
    """the round-2 quantiser kernels (crk_vq_fast.cuh): tensor-core modes, D = 64, K in {128, 256, 384, 512}"""
    return L.get_precision() != "fp32" and D == 64 and K % 128 == 0 and 128 <= K <= 512


def vq_pack_operand(W, out=None):
    """codebook (K, D) -> operand blob of crk_vq_argmin_fast (raw fp32 codebook in the tensor-core layout | |w|^2)"""
    K, D = W.shape
    if out is None:
        out = torch.zeros(L.lib().crk_vq_op_floats(K, D), dtype=_f32, device=W.device)    # (pad rows stay 0)
    Wc = W.detach()
    L.call("crk_vq_pack_op", L.ptr(Wc if Wc.is_contiguous() else Wc.contiguous()), L.ptr(out), K, D)
    return out


class VQFn(torch.autograd.Function):
    """idx = argmin L2, e = W[idx], qx = x + (e - x) (straight-through).

    Tensor-core modes: crk_vq_argmin_fast (one TF32 pass + exact fp32 re-score; `opblob` = the caller's cached operand
    blob of W, packed here when None).  fp32 mode / other shapes: crk_vq_prepare + crk_vq_argmin (CUDA cores).
    `variant="tf32x3"` selects round 1's crk_vq_argmin_tc (kept for the cross-check tests)."""

    @staticmethod
    def forward(ctx, x, W, opblob=None, variant=None):
        x, ldx = panel(x)
        B, T, D = x.shape
        K = W.shape[0]
        dev = x.device
        Wc = W.detach()
        if not Wc.is_contiguous():
            Wc = Wc.contiguous()
        idx = torch.empty(B, T, dtype=torch.int64, device=dev)
        e = torch.empty(B, T, D, dtype=_f32, device=dev)
        qx = torch.empty(B, T, D, dtype=_f32, device=dev)
        if vq_fast_ok(K, D) and variant is None:
            if opblob is None:
                opblob = vq_pack_operand(Wc)
            L.call("crk_vq_argmin_fast", L.ptr(x), ldx, L.ptr(opblob), L.ptr(idx), L.ptr(e), D, L.ptr(qx), D, B * T, K, D)
        else:
            WT = torch.empty(D, K, dtype=_f32, device=dev)
            wn = torch.empty(K, dtype=_f32, device=dev)
            L.call("crk_vq_prepare", L.ptr(Wc), L.ptr(WT), L.ptr(wn), K, D)
            if variant == "tf32x3" and D == 64 and K % 128 == 0 and K <= 512:
                # round 1: 3xTF32 distance GEMM + exact fp32 re-score of near-ties
                blob = torch.empty(L.lib().crk_vq_tc_blob_floats(K, D), dtype=_f32, device=dev)
                L.call("crk_vq_pack_tc", L.ptr(Wc), L.ptr(blob), K, D)
                L.call("crk_vq_argmin_tc", L.ptr(x), ldx, L.ptr(Wc), L.ptr(blob), L.ptr(wn), L.ptr(idx),
                       L.ptr(e), D, L.ptr(qx), D, B * T, K, D)
            else:
                L.call("crk_vq_argmin", L.ptr(x), ldx, L.ptr(Wc), L.ptr(WT), L.ptr(wn), L.ptr(idx),
                       L.ptr(e), D, L.ptr(qx), D, B * T, K, D)
        ctx.save_for_backward(idx)
        ctx.KD = (K, D)
        ctx.mark_non_differentiable(idx)
        return e, qx, idx

    @staticmethod
    def backward(ctx, ge, gqx, _gidx):
        (idx,) = ctx.saved_tensors
        K, D = ctx.KD
        gx = gqx if ctx.needs_input_grad[0] else None
        gW = None
        if ctx.needs_input_grad[1] and ge is not None:
            ge, ldg = panel(ge)
            gW = torch.zeros(K, D, dtype=_f32, device=ge.device)
            L.call("crk_vq_scatter_grad", L.ptr(ge), ldg, L.ptr(idx), L.ptr(gW), idx.numel(), K, D)
        return gx, gW, None, None


def vq_ema_update(x, idx, ema_size, ema_w, W, decay, eps, reduce_fn=None, runner=None, opblob=None):
    """EMA codebook update (vqvae2.py:315-330).  `reduce_fn(flat_stats)` sums [counts | esum]
    across data-parallel ranks before the normalisation; `runner(fn, stats)` may run
    that reduction and the EMA kernel somewhere else (the communication stream) instead of inline.
    D = 64: statistics (2 launches) + ONE fused EMA launch (crk_vq_ema_fused) that also refreshes `opblob`, the
    operand blob the next crk_vq_argmin_fast call reads."""
    x, ldx = panel(x)
    B, T, D = x.shape
    K = W.shape[0]
    dev = x.device
    ws = _empty(L.lib().crk_vq_stats_ws_floats(B * T, K, D), dev)
    if D == 64 and K % 8 == 0:
        stats = torch.empty(L.lib().crk_vq_stats_floats(K, D), dtype=_f32, device=dev)
        L.call("crk_vq_stats_fused", L.ptr(x), ldx, L.ptr(idx), L.ptr(stats), L.ptr(ws), B * T, K, D)
        red = stats[:K + D * K]
        blob = opblob if (opblob is not None and vq_fast_ok(K, D)) else None

        def finish():
            if reduce_fn is not None:
                reduce_fn(red)
            L.call("crk_vq_ema_fused", L.ptr(stats), L.ptr(ema_size), L.ptr(ema_w), L.ptr(W), L.ptr(blob),
                   C.c_float(decay), C.c_float(eps), K, D)
    else:
        stats = torch.empty(K + D * K, dtype=_f32, device=dev)
        counts = stats[:K]
        esum = stats[K:]
        L.call("crk_vq_stats", L.ptr(x), ldx, L.ptr(idx), L.ptr(counts), L.ptr(esum), L.ptr(ws),
               B * T, K, D)

        def finish():
            if reduce_fn is not None:
                reduce_fn(stats)
            L.call("crk_vq_ema", L.ptr(counts), L.ptr(esum), L.ptr(ema_size), L.ptr(ema_w), L.ptr(W),
                   C.c_float(decay), C.c_float(eps), K, D)

    if runner is not None:
        runner(finish, stats)
    else:
        finish()


# ---------------------------------------------------------------------------------------------
def _mask_u8(mask, B, T):
    if mask is None:
        return None
    m = mask.reshape(B, T)
    if m.dtype == torch.bool:
        m = m.contiguous().view(torch.uint8)
    elif m.dtype != torch.uint8:
        m = (m != 0).view(torch.uint8)
    return m.contiguous()


class MaskedLossFn(torch.autograd.Function):
    """(mean |x-y|, mean (x-y)^2) over mask-selected frames with causal shift.  y may be a python
    float (LSGAN targets).  Returns two 0-dim tensors; no host sync."""

    @staticmethod
    def forward(ctx, x, y, mask, shift):
        x, ldx = panel(x)
        B, T, D = x.shape
        yconst = 0.0
        yt, ldy = None, 0
        if isinstance(y, torch.Tensor):
            yt, ldy = panel(y.detach())
        else:
            yconst = float(y)
        m = _mask_u8(mask, B, T)
        out = torch.empty(3, dtype=_f32, device=x.device)
        ws = _empty(L.lib().crk_masked_loss_ws_floats(B, T, D), x.device)
        L.call("crk_masked_loss_fwd", L.ptr(x), ldx, L.ptr(yt), ldy, C.c_float(yconst), L.ptr(m),
               B, T, D, int(shift), L.ptr(out), L.ptr(ws))
        ctx.meta = (B, T, D, ldx, ldy, yconst, int(shift))
        ctx.has_y = yt is not None
        ctx.has_m = m is not None
        saved = [x, out]
        if yt is not None:
            saved.append(yt)
        if m is not None:
            saved.append(m)
        ctx.save_for_backward(*saved)
        cnt = out[2]
        ctx.mark_non_differentiable(cnt)
        return out[0], out[1], cnt

    @staticmethod
    def backward(ctx, g1, g2, _gcnt):
        saved = list(ctx.saved_tensors)
        x, out = saved[:2]
        rest = saved[2:]
        yt = rest.pop(0) if ctx.has_y else None
        m = rest.pop(0) if ctx.has_m else None
        B, T, D, ldx, ldy, yconst, shift = ctx.meta
        if g1 is not None:
            g1 = g1.contiguous()
        if g2 is not None:
            g2 = g2.contiguous()
        dx = torch.empty(B, T, D, dtype=_f32, device=x.device)
        L.call("crk_masked_loss_bwd", L.ptr(x), ldx, L.ptr(yt), ldy, C.c_float(yconst), L.ptr(m),
               B, T, D, shift, L.ptr(out), L.ptr(g1), L.ptr(g2), L.ptr(dx), D)
        return dx, None, None, None


# Data-parallel exactness hook (set by crank_b200.net._dp.enable): a masked mean computed on this rank
# over `count` selected elements becomes this rank's share of the GLOBAL mean  sum_r(num_r) / sum_r(count_r)
# when it is multiplied by  world * count / sum_r(count_r)  (the gradient all-reduce then divides by world).
#   hook(key_tensor, shift, count) -> 0-dim weight tensor, or None when the weight is exactly 1
_mean_weight_hook = None


def set_mean_weight_hook(fn):
    global _mean_weight_hook
    _mean_weight_hook = fn


def masked_l1_mse(x, y, mask=None, shift=0):
    l1, mse, cnt = MaskedLossFn.apply(x, y, mask, shift)
    if _mean_weight_hook is not None and mask is not None:
        w = _mean_weight_hook(mask, int(shift), cnt)
        if w is not None:
            return l1 * w, mse * w
    return l1, mse


class StftLossFn(torch.autograd.Function):
    """STFT-magnitude trajectory L1 for ONE (n_fft, hop, win) resolution; returns (mag, logmag) means."""

    @staticmethod
    def forward(ctx, x, y, n_fft, hop, win):
        x, ldx = panel(x)
        y, ldy = panel(y.detach())
        B, T, D = x.shape
        out = torch.empty(2, dtype=_f32, device=x.device)
        ws = _empty(L.lib().crk_stft_loss_ws_floats(B, T, D, n_fft, hop), x.device)
        L.call("crk_stft_loss_fwd", L.ptr(x), ldx, L.ptr(y), ldy, B, T, D, n_fft, hop, win,
               L.ptr(out), L.ptr(ws))
        ctx.meta = (B, T, D, ldx, ldy, n_fft, hop, win)
        ctx.save_for_backward(x, y)
        ctx.set_materialize_grads(False)      # an unused output's gradient arrives as None, not zeros
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g, glog):
        x, y = ctx.saved_tensors
        B, T, D, ldx, ldy, n_fft, hop, win = ctx.meta
        dx = torch.empty(B, T, D, dtype=_f32, device=x.device)
        g = g.contiguous() if g is not None else None
        glog = glog.contiguous() if glog is not None else None
        if g is None and glog is None:
            return torch.zeros_like(dx), None, None, None, None
        L.call("crk_stft_loss_bwd", L.ptr(x), ldx, L.ptr(y), ldy, B, T, D, n_fft, hop, win,
               L.ptr(g), L.ptr(glog), C.c_float(1.0), L.ptr(dx), D, 0)
        return dx, None, None, None, None


class CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, ignore_index):
        L.require_cuda(logits, labels)
        assert logits.dim() == 2
        if logits.stride(1) != 1:
            logits = logits.contiguous()
        logits = logits.float()
        labels = labels.contiguous()
        Fn, S = logits.shape
        out = torch.empty(2, dtype=_f32, device=logits.device)
        ws = _empty(L.lib().crk_ce_ws_floats(Fn), logits.device)
        L.call("crk_ce_fwd", L.ptr(logits), logits.stride(0), L.ptr(labels), Fn, S,
               int(ignore_index), L.ptr(out), L.ptr(ws))
        ctx.meta = (Fn, S, int(ignore_index))
        ctx.save_for_backward(logits, labels, out)
        cnt = out[1]
        ctx.mark_non_differentiable(cnt)
        return out[0], cnt

    @staticmethod
    def backward(ctx, g, _gcnt):
        logits, labels, out = ctx.saved_tensors
        Fn, S, ign = ctx.meta
        dl = torch.empty(Fn, S, dtype=_f32, device=logits.device)
        L.call("crk_ce_bwd", L.ptr(logits), logits.stride(0), L.ptr(labels), Fn, S, ign,
               L.ptr(out), L.ptr(g.contiguous()), L.ptr(dl), S)
        return dl, None, None


def cross_entropy(logits, labels, ignore_index=-100):
    ce, cnt = CrossEntropyFn.apply(logits, labels, ignore_index)
    if _mean_weight_hook is not None:
        w = _mean_weight_hook(labels, 0, cnt)
        if w is not None:
            return ce * w
    return ce


# ---------------------------------------------------------------------------------------------
def adam_step(p, g, m, v, lr, beta1, beta2, eps, step_count):
    L.require_cuda(p, g, m, v)
    L.call("crk_adam_step", L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), C.c_float(lr),
           C.c_float(beta1), C.c_float(beta2), C.c_float(eps), int(step_count))


def radam_step(p, g, m, v, lr, beta1, beta2, eps, step_count):
    L.require_cuda(p, g, m, v)
    L.call("crk_radam_step", L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), C.c_float(lr), C.c_float(beta1),
           C.c_float(beta2), C.c_float(eps), int(step_count))


def lamb_step(p, g, m, v, upd, seg_off, seg_len, trust, lr, beta1, beta2, eps):
    L.require_cuda(p, g, m, v, upd, seg_off, seg_len, trust)
    if not p.is_contiguous():
        raise ValueError("lamb_step needs a contiguous parameter")
    L.call("crk_lamb_step", L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), L.ptr(upd), L.ptr(seg_off), L.ptr(seg_len),
           int(seg_off.numel()), L.ptr(trust), C.c_float(lr), C.c_float(beta1), C.c_float(beta2), C.c_float(eps))


def adam_step_dev(p, g, m, v, lr, beta1, beta2, eps, step_dev):
    """Adam with the step counter in device memory (int64 tensor of one element, incremented by the call)."""
    L.require_cuda(p, g, m, v, step_dev)
    L.call("crk_adam_step_dev", L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), C.c_float(lr),
           C.c_float(beta1), C.c_float(beta2), C.c_float(eps), L.ptr(step_dev))


_MEL_BANDS = {}
MEL_FUSED_MAXNNZ = 2048


def _runs(w):
    """per column of w: first/last non-zero row as (start, len, off, vals)"""
    start, length, off, vals = [], [], [], []
    for m in range(w.shape[1]):
        nz = w[:, m].nonzero()[0]
        lo, hi = (int(nz[0]), int(nz[-1]) + 1) if len(nz) else (0, 0)
        start.append(lo), length.append(hi - lo), off.append(len(vals))
        vals.extend(w[lo:hi, m].tolist())
    return start, length, off, vals


def _mel_bands(mel_basis):
    """Non-zero runs of a (bins, n_mels) mel basis, per mel channel (band_start, band_len, band_off, band_w, nnz) and
    transposed, per bin (bin_start, bin_len, bin_off, bin_w), on its device; None when the basis is not banded enough
    for the fused kernels.  Cached per basis version (one host sync)."""
    key = (mel_basis.data_ptr(), mel_basis._version, tuple(mel_basis.shape), str(mel_basis.device))
    hit = _MEL_BANDS.get(key)
    if hit is not None:
        return hit[0]
    w = mel_basis.detach().float().cpu().numpy()
    start, length, off, vals = _runs(w)
    tstart, tlength, toff, tvals = _runs(w.T.copy())
    bands = None
    if 0 < len(vals) <= MEL_FUSED_MAXNNZ:
        dev = mel_basis.device
        it = lambda a: torch.tensor(a, dtype=torch.int32, device=dev)      # noqa: E731
        ft = lambda a: torch.tensor(a, dtype=_f32, device=dev)             # noqa: E731
        bands = (it(start), it(length), it(off), ft(vals), len(vals), it(tstart), it(tlength), it(toff), ft(tvals))
    if len(_MEL_BANDS) > 16:
        _MEL_BANDS.clear()
    _MEL_BANDS[key] = (bands, mel_basis)  # keeps the basis alive so the data_ptr key cannot be recycled
    return bands


class LogMelFn(torch.autograd.Function):
    """Fused front end with a backward to the STFT window and / or the waveform (crk_logmel_fused_{fwd,bwd}): the
    learnable "param" / "conv" windows of crank/net/module/mlfb.py:72-90."""

    @staticmethod
    def forward(ctx, wav, window, mel_basis, n_fft, hop, eps, mean, std):
        L.require_cuda(wav, window, mel_basis)
        wav = wav.float().contiguous()
        window = window.float().contiguous()
        bands = _mel_bands(mel_basis) if (n_fft == 1024 and hop <= 512 and mel_basis.shape[1] <= 128 and
                                          mel_basis.shape[0] == 513) else None
        if bands is None:
            raise ValueError("learnable STFT windows need the fused log-mel kernel: n_fft = 1024, hop <= 512, a banded basis")
        B, n = wav.shape
        M = 1 + (n - n_fft) // hop
        out = torch.empty(B, M, mel_basis.shape[1], dtype=_f32, device=wav.device)
        st, ln, of, bw, nnz = bands[:5]
        L.call("crk_logmel_fused_fwd", L.ptr(wav), B, n, L.ptr(window), L.ptr(st), L.ptr(ln), L.ptr(of), L.ptr(bw),
               nnz, n_fft, hop, out.shape[2], C.c_float(eps), L.ptr(mean), L.ptr(std), L.ptr(out))
        ctx.save_for_backward(wav, window)
        ctx.misc = (bands, n_fft, hop, eps, mean, std)
        return out

    @staticmethod
    def backward(ctx, dout):
        wav, window = ctx.saved_tensors
        bands, n_fft, hop, eps, mean, std = ctx.misc
        st, ln, of, bw, nnz, tst, tln, tof, tbw = bands
        B, n = wav.shape
        dout = dout.float().contiguous()
        dwav = torch.zeros_like(wav) if ctx.needs_input_grad[0] else None
        dwin = torch.empty_like(window) if ctx.needs_input_grad[1] else None
        ws = _empty(L.lib().crk_logmel_bwd_ws_floats(B, n, hop), wav.device) if dwin is not None else None
        L.call("crk_logmel_fused_bwd", L.ptr(wav), B, n, L.ptr(window), L.ptr(st), L.ptr(ln), L.ptr(of), L.ptr(bw), nnz,
               L.ptr(tst), L.ptr(tln), L.ptr(tof), L.ptr(tbw), n_fft, hop, dout.shape[2], C.c_float(eps), L.ptr(mean),
               L.ptr(std), L.ptr(dout), L.ptr(dwin), L.ptr(dwav), L.ptr(ws))
        return dwav, dwin, None, None, None, None, None, None


def logmel_learnable(wav, window, mel_basis, n_fft, hop, eps=1e-10, mean=None, std=None):
    return LogMelFn.apply(wav, window, mel_basis, n_fft, hop, eps, mean, std)


def logmel(wav, window, mel_basis, n_fft, hop, eps=1e-10, mean=None, std=None, fused=None):
    """wav (B, n_samples) -> (B, n_frames, n_mels);  frames start at m*hop (no centring here).

    n_fft = 1024 with a banded basis (every recipe) runs the one-kernel front end `crk_logmel_fused_fwd`;
    anything else (or fused=False) the cuFFT path `crk_logmel_fwd`."""
    L.require_cuda(wav, window, mel_basis)
    wav = wav.float().contiguous()
    B, n = wav.shape
    n_mels = mel_basis.shape[1]
    M = 1 + (n - n_fft) // hop
    out = torch.empty(B, M, n_mels, dtype=_f32, device=wav.device)
    can_fuse = n_fft == 1024 and hop <= 512 and n_mels <= 128 and mel_basis.shape[0] == 513
    bands = _mel_bands(mel_basis) if (can_fuse and fused is not False) else None
    if fused and bands is None:
        raise ValueError("the fused log-mel kernel needs n_fft = 1024, hop <= 512, n_mels <= 128 and a banded basis")
    if bands is not None:
        st, ln, of, bw, nnz = bands[:5]
        L.call("crk_logmel_fused_fwd", L.ptr(wav), B, n, L.ptr(window), L.ptr(st), L.ptr(ln), L.ptr(of), L.ptr(bw),
               nnz, n_fft, hop, n_mels, C.c_float(eps), L.ptr(mean), L.ptr(std), L.ptr(out))
        return out
    ws = _empty(L.lib().crk_logmel_ws_floats(B, M, n_fft), wav.device)
    L.call("crk_logmel_fwd", L.ptr(wav), B, n, L.ptr(window), L.ptr(mel_basis), n_fft, hop,
           n_mels, C.c_float(eps), L.ptr(mean), L.ptr(std), L.ptr(out), L.ptr(ws))
    return out


SQRT_HALF = math.sqrt(0.5)
