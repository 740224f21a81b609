"""Drop-in `parallel_wavegan.models` for crank, backed by the sm_100a kernels.

Same constructor keywords, forward signatures ((B, C, T) in / out), attributes and state-dict
keys as the three classes crank imports from the third-party package:

  ParallelWaveGANGenerator             <- crank/net/module/vqvae2.py:17,236-273 (kwargs :237-253)
  ParallelWaveGANDiscriminator         <- crank/bin/train.py:78-89, crank/net/module/spkradv.py:49-60
  ResidualParallelWaveGANDiscriminator <- crank/bin/train.py:107-115

Differences under the hood (B200-first):
  * all Conv1d parameters of a network live in ONE flat parameter `theta` ([weight_g | weight_v |
    bias] per conv, PyTorch weight_norm layout), so an optimizer step / gradient all-reduce is a
    single kernel / collective.  `state_dict()` / `load_state_dict()` still speak the reference's
    per-conv keys (`conv_layers.3.conv.weight_g` ...), so old checkpoints load.
  * weight-norm is evaluated once per parameter version (not at every conv call) into a packed
    effective-weight buffer in the kernels' layout.
  * the whole stack is one autograd node; activations are channels-last (B, T, C) internally
    (`forward_cl`), the (B, C, T) API transposes at the edge only.
There is no CPU / eager fallback: forward raises unless tensors are CUDA and the library loads.
"""

import ctypes as C
import math
from collections import OrderedDict

import torch

from .. import lib as L
from ..net import _dp
from ..ops import ConvstackFn, WavenetFn

_ACT = {"none": 0, "ReLU": 1, "LeakyReLU": 2}


class _PackedConvNet(torch.nn.Module):
    """Flat-parameter container + weight-norm cache + reference-keyed state dict."""

    def _setup(self, descs, theta_floats, weff_floats, names, use_weight_norm=True):
        self._descs = descs
        self._names = names
        self._weff_floats = int(weff_floats)
        self._use_weight_norm = use_weight_norm
        self._weight_norm_removed = not use_weight_norm
        self.theta = torch.nn.Parameter(torch.zeros(int(theta_floats)))
        # the reference's parameter tensors inside the flat pack (weight_g / weight_v / bias of every conv): per-tensor
        # optimizers (LAMB's trust ratio) work on these segments
        segs = []
        for d in descs:
            segs.append((int(d.g_off), int(d.cout)))
            segs.append((int(d.v_off), int(d.cout * d.cin * d.k)))
            if d.b_off >= 0:
                segs.append((int(d.b_off), int(d.cout)))
        self.theta._crk_segments = segs
        self._weff = None
        self._weff_key = None
        self.reset_parameters()

    # -- initialisation: kaiming_normal_(relu) weights, zero bias (parallel_wavegan.layers.Conv1d),
    #    then weight_norm's decomposition g = ||v||, v = w
    def reset_parameters(self):
        with torch.no_grad():
            th = self.theta
            for d in self._descs:
                n = d.cout * d.cin * d.k
                std = math.sqrt(2.0 / (d.cin * d.k))
                w = torch.randn(d.cout, d.cin * d.k) * std
                th[d.v_off : d.v_off + n] = w.reshape(-1).to(th.device)
                th[d.g_off : d.g_off + d.cout] = w.norm(dim=1).to(th.device)
                if d.b_off >= 0:
                    th[d.b_off : d.b_off + d.cout] = 0.0

    def _launch_weights(self, theta, weff):
        raise NotImplementedError

    def effective_weights(self):
        th = self.theta
        _dp.wait_for((th,))              # data parallel: a pending side-stream all-reduce + Adam of this pack
        key = (th._version, th.data_ptr())
        if self._weff is None or self._weff_key != key:
            L.require_cuda(th)
            weff = torch.zeros(max(self._weff_floats, 4), dtype=torch.float32, device=th.device)
            self._launch_weights(th.detach(), weff)
            self._weff, self._weff_key = weff, key
        return self._weff

    # -- reference-keyed state dict ----------------------------------------------------------
    def _conv_views(self):
        th = self.theta.detach()
        for name, d in zip(self._names, self._descs):
            g = th[d.g_off : d.g_off + d.cout].view(d.cout, 1, 1)
            v = th[d.v_off : d.v_off + d.cout * d.cin * d.k].view(d.cout, d.cin, d.k)
            b = th[d.b_off : d.b_off + d.cout] if d.b_off >= 0 else None
            yield name, g, v, b

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        _dp.flush()
        for name, g, v, b in self._conv_views():
            if b is not None:
                destination[prefix + name + ".bias"] = b.clone()
            if self._weight_norm_removed:
                destination[prefix + name + ".weight"] = (v * (g / v.flatten(1).norm(dim=1).view(-1, 1, 1))).clone()
            else:
                destination[prefix + name + ".weight_g"] = g.clone()
                destination[prefix + name + ".weight_v"] = v.clone()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs):
        with torch.no_grad():
            for name, g, v, b in self._conv_views():
                kg, kv, kw, kb = (prefix + name + s for s in (".weight_g", ".weight_v", ".weight", ".bias"))
                try:
                    if kg in state_dict and kv in state_dict:
                        g.copy_(state_dict[kg].reshape(g.shape))
                        v.copy_(state_dict[kv].reshape(v.shape))
                    elif kw in state_dict:
                        w = state_dict[kw].reshape(v.shape).to(v.device, v.dtype)
                        v.copy_(w)
                        g.copy_(w.flatten(1).norm(dim=1).view(g.shape))
                    else:
                        missing_keys.extend([kg, kv])
                    if b is not None:
                        if kb in state_dict:
                            b.copy_(state_dict[kb].reshape(b.shape))
                        else:
                            missing_keys.append(kb)
                except RuntimeError as err:
                    error_msgs.append(f"size mismatch for {prefix + name}: {err}")
            self.theta.add_(0.0)  # bump the version: invalidate the weight-norm cache
        own = set()
        for name in self._names:
            own.update(prefix + name + s for s in (".weight_g", ".weight_v", ".weight", ".bias"))
        if strict:
            for k in state_dict:
                if k.startswith(prefix) and k not in own:
                    unexpected_keys.append(k)

    def named_conv_grads(self):
        """{reference parameter key: gradient view} sliced out of theta.grad (tests / inspection)."""
        out = {}
        if self.theta.grad is None:
            return out
        gr = self.theta.grad
        for name, d in zip(self._names, self._descs):
            out[name + ".weight_g"] = gr[d.g_off : d.g_off + d.cout].view(d.cout, 1, 1)
            out[name + ".weight_v"] = gr[d.v_off : d.v_off + d.cout * d.cin * d.k].view(d.cout, d.cin, d.k)
            if d.b_off >= 0:
                out[name + ".bias"] = gr[d.b_off : d.b_off + d.cout]
        return out

    def apply_weight_norm(self):
        self._weight_norm_removed = False

    def remove_weight_norm(self):
        """Fold g into v (w = g*v/||v|| becomes the stored weight; the function is unchanged)."""
        with torch.no_grad():
            for _, g, v, _b in self._conv_views():
                w = v * (g / v.flatten(1).norm(dim=1).view(-1, 1, 1))
                v.copy_(w)
                g.copy_(w.flatten(1).norm(dim=1).view(g.shape))
            self.theta.add_(0.0)
        self._weight_norm_removed = True


class _WavenetBase(_PackedConvNet):
    def _build(self, in_channels, out_channels, aux_channels, layers, stacks, kernel_size,
               use_causal_conv, first_act, head_act, slope, dropout, first_name, use_weight_norm):
        assert layers % stacks == 0
        if int(in_channels) > 256 or max(int(out_channels), int(aux_channels)) > 128:
            raise NotImplementedError(
                f"WaveNet stack with {in_channels} input / {out_channels} output / {aux_channels} aux channels: the "
                "sm_100a kernels take at most 256 input and 128 output / aux channels (n_vq_stacks = 3 needs 192 inputs)")
        self.cfg = L.WavenetCfg(
            in_ch=in_channels, out_ch=out_channels, aux_ch=max(int(aux_channels), 0), layers=layers,
            stacks=stacks, kernel_size=kernel_size, causal=int(bool(use_causal_conv)),
            first_act=first_act, head_act=head_act, slope=slope,
        )
        descs, th, we = L.describe_wavenet(self.cfg)
        names = [first_name]
        for l in range(layers):
            names.append(f"conv_layers.{l}.conv")
            if self.cfg.aux_ch > 0:
                names.append(f"conv_layers.{l}.conv1x1_aux")
            names.append(f"conv_layers.{l}.conv1x1_out")
            names.append(f"conv_layers.{l}.conv1x1_skip")
        names += ["last_conv_layers.1", "last_conv_layers.3"]
        assert len(names) == len(descs)
        self.dropout = float(dropout)
        self._setup(descs, th, we, names, use_weight_norm)

    def _launch_weights(self, theta, weff):
        L.call("crk_wavenet_weights", C.byref(self.cfg), L.ptr(theta), L.ptr(weff))

    def _dropmul(self, B, T, device):
        if self.dropout <= 0.0 or not self.training:
            return None
        keep = 1.0 - self.dropout
        m = torch.rand(self.cfg.layers, B * T, 64, device=device) < keep
        return m.float().mul_(1.0 / keep)

    def forward_cl(self, x, c=None, dropmul=None):
        """x (B,T,in) [, c (B,T,aux)] -> (B,T,out), channels-last."""
        L.require_cuda(x, self.theta)
        if dropmul is None:
            dropmul = self._dropmul(x.shape[0], x.shape[1], x.device)
        if self.cfg.aux_ch > 0 and c is None:
            raise AssertionError("this stack was built with aux_channels > 0: `c` is required")
        if self.cfg.aux_ch == 0:
            c = None
        # under torch.no_grad() (or when nothing upstream requires a gradient) the inference entry is used:
        # the (tanh, sigmoid) pairs backward would need are not written (40 % of the block's output traffic)
        needs = self.theta.requires_grad or x.requires_grad or (c is not None and c.requires_grad)
        return WavenetFn.apply(self, x, c, dropmul, self.theta, bool(torch.is_grad_enabled() and needs))

    @property
    def receptive_field_size(self):
        lps = self.cfg.layers // self.cfg.stacks
        dil = [2 ** (i % lps) for i in range(self.cfg.layers)]
        return (self.cfg.kernel_size - 1) * sum(dil) + 1


class ParallelWaveGANGenerator(_WavenetBase):
    def __init__(self, in_channels=1, out_channels=1, kernel_size=3, layers=30, stacks=3,
                 residual_channels=64, gate_channels=128, skip_channels=64, aux_channels=80,
                 aux_context_window=2, dropout=0.0, bias=True, use_weight_norm=True,
                 use_causal_conv=False, upsample_conditional_features=True,
                 upsample_net="ConvInUpsampleNetwork", upsample_params=None):
        super().__init__()
        if (residual_channels, gate_channels, skip_channels) != (64, 128, 64) or not bias:
            raise NotImplementedError("kernels are specialised to 64/128/64 channels with bias (all crank uses)")
        if upsample_conditional_features:
            raise NotImplementedError("crank always passes upsample_conditional_features=False")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.aux_channels, self.aux_context_window = aux_channels, aux_context_window
        self.layers, self.stacks, self.kernel_size = layers, stacks, kernel_size
        self._build(in_channels, out_channels, aux_channels, layers, stacks, kernel_size,
                    use_causal_conv, _ACT["none"], _ACT["ReLU"], 0.0, dropout, "first_conv",
                    use_weight_norm)

    def forward(self, x, c=None):
        """x (B, in, T), c (B, aux, T) or None -> (B, out, T)   [reference layout]"""
        y = self.forward_cl(x.transpose(1, 2), None if c is None else c.transpose(1, 2))
        return y.transpose(1, 2)


class ResidualParallelWaveGANDiscriminator(_WavenetBase):
    def __init__(self, in_channels=1, out_channels=1, kernel_size=3, layers=30, stacks=3,
                 residual_channels=64, gate_channels=128, skip_channels=64, dropout=0.0, bias=True,
                 use_weight_norm=True, use_causal_conv=False, nonlinear_activation="LeakyReLU",
                 nonlinear_activation_params={"negative_slope": 0.2}):
        super().__init__()
        if (residual_channels, gate_channels, skip_channels) != (64, 128, 64) or not bias:
            raise NotImplementedError("kernels are specialised to 64/128/64 channels with bias")
        if nonlinear_activation != "LeakyReLU":
            raise NotImplementedError("only LeakyReLU")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.layers, self.stacks, self.kernel_size = layers, stacks, kernel_size
        slope = float(nonlinear_activation_params.get("negative_slope", 0.01))
        self._build(in_channels, out_channels, 0, layers, stacks, kernel_size, use_causal_conv,
                    _ACT["LeakyReLU"], _ACT["LeakyReLU"], slope, dropout, "first_conv.0",
                    use_weight_norm)

    def forward(self, x):
        return self.forward_cl(x.transpose(1, 2)).transpose(1, 2)


class ParallelWaveGANDiscriminator(_PackedConvNet):
    def __init__(self, in_channels=1, out_channels=1, kernel_size=3, layers=10, conv_channels=64,
                 dilation_factor=1, nonlinear_activation="LeakyReLU",
                 nonlinear_activation_params={"negative_slope": 0.2}, bias=True,
                 use_weight_norm=True):
        super().__init__()
        assert (kernel_size - 1) % 2 == 0, "Not support even number kernel size."
        assert dilation_factor > 0, "Dilation factor must be > 0."
        if nonlinear_activation != "LeakyReLU" or not bias:
            raise NotImplementedError("only LeakyReLU with bias (all crank uses)")
        slope = float(nonlinear_activation_params.get("negative_slope", 0.01))
        self.cfg = L.ConvstackCfg(in_ch=in_channels, out_ch=out_channels, layers=layers,
                                  kernel_size=kernel_size, conv_ch=conv_channels,
                                  dilation_factor=dilation_factor, slope=slope)
        descs, th, we = L.describe_convstack(self.cfg)
        names = [f"conv_layers.{2 * i}" for i in range(layers)]
        self._setup(descs, th, we, names, use_weight_norm)

    def _launch_weights(self, theta, weff):
        L.call("crk_convstack_weights", C.byref(self.cfg), L.ptr(theta), L.ptr(weff))

    def forward_cl(self, x, grad_scale=1.0):
        """x (B,T,in) -> (B,T,out).  grad_scale multiplies the gradient flowing back into x."""
        L.require_cuda(x, self.theta)
        return ConvstackFn.apply(self, x, self.theta, grad_scale)

    def forward(self, x):
        return self.forward_cl(x.transpose(1, 2)).transpose(1, 2)


__all__ = [
    "ParallelWaveGANGenerator",
    "ParallelWaveGANDiscriminator",
    "ResidualParallelWaveGANDiscriminator",
]
_ = OrderedDict  # (kept for state-dict typing parity)
