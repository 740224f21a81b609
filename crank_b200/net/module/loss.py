"""Feature losses on the fused loss kernels (sum/count reductions: no masked_select, no sync).

API mirror of crank/net/module/loss.py:18-114 -- including its argument plumbing quirk:
`MultiSizeSTFTLoss.__init__` passes `(fft, hop, win)` positionally into
`STFTLoss(fft_size, win_size, hop_size)` (loss.py:99-102), so with the recipes' yaml
(fft [64,128], win [64,128], hop [16,32]) the effective torch.stft parameters are
n_fft=64/hop_length=64/win_length=16 and 128/128/32.  STFTLoss on its own is
not swapped (its two internal swaps cancel, loss.py:79 vs :50-57).
"""

import torch
import torch.nn as nn

from ... import ops


class CustomFeatureLoss(nn.Module):
    def __init__(self, loss_type="l1", causal=False, stft_params={}, device="cuda"):
        super().__init__()
        self.loss_type = loss_type
        self.causal = causal
        if loss_type == "stft":
            self.loss_func = MultiSizeSTFTLoss(**stft_params, device=device)
        elif loss_type not in ("l1", "mse"):
            raise ValueError(f"unknown loss_type {loss_type}")

    def forward(self, x, y, mask=None, causal_size=0):
        shift = causal_size if self.causal else 0
        if self.loss_type == "stft":
            if shift > 0:
                x, y = x[:, shift:], y[:, :-shift]
            elif shift < 0:
                x, y = x[:, :shift], y[:, -shift:]
            return self.loss_func(x, y)
        l1, mse = ops.masked_l1_mse(x, y, mask, shift)
        return l1 if self.loss_type == "l1" else mse


def stft(x, fft_size, hop_size, win_size, window=None):
    raise NotImplementedError(
        "the STFT magnitudes are never materialised here: use STFTLoss (fused STFT->|.|->L1 kernel)"
    )


class STFTLoss(nn.Module):
    def __init__(self, fft_size=32, win_size=20, hop_size=10, logratio=0.0, device="cuda"):
        super().__init__()
        self.fft_size = fft_size
        self.win_size = win_size
        self.hop_size = hop_size
        self.logratio = logratio

    def forward(self, x, y):
        """x, y (B, T, D): L1 between STFT magnitudes of every feature-dimension trajectory."""
        mag, lmag = ops.StftLossFn.apply(x, y, int(self.fft_size), int(self.hop_size), int(self.win_size))
        if self.logratio == 0:
            return mag                     # == (1 - 0) * mag + 0 * lmag exactly; skips the log-term backward
        return (1 - self.logratio) * mag + self.logratio * lmag


class MultiSizeSTFTLoss(nn.Module):
    def __init__(self, fft_sizes=[32, 128, 256], win_sizes=[20, 80, 160], hop_sizes=[10, 20, 30],
                 logratio=0.0, device="cuda"):
        super().__init__()
        self.loss_layers = nn.ModuleList()
        for fft_size, win_size, hop_size in zip(fft_sizes, win_sizes, hop_sizes):
            # positional order kept from the reference: (fft, hop, win) into (fft, win, hop)
            self.loss_layers.append(STFTLoss(fft_size, hop_size, win_size, logratio=logratio, device=device))

    def forward(self, x, y):
        losses = [layer(x, y) for layer in self.loss_layers]
        return sum(losses) / len(losses)


class MaskedMSELoss(nn.Module):
    """criterion["mse"] / ["l1"]: plain (x, y) like nn.MSELoss / nn.L1Loss, plus the fused
    `(x, y, mask=...)` form the trainers use instead of masked_select (y may be a float)."""

    def __init__(self, kind="mse"):
        super().__init__()
        self.kind = kind

    def forward(self, x, y, mask=None):
        if x.dim() != 3:
            x = x.reshape(1, -1, 1)
            if isinstance(y, torch.Tensor):
                y = y.reshape(1, -1, 1)
        l1, mse = ops.masked_l1_mse(x, y, mask, 0)
        return mse if self.kind == "mse" else l1


class CrossEntropyLoss(nn.Module):
    def __init__(self, ignore_index=-100):
        super().__init__()
        self.ignore_index = ignore_index

    def forward(self, logits, labels):
        return ops.cross_entropy(logits, labels, self.ignore_index)
