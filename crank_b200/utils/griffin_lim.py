"""On-device Griffin-Lim for the eval / reconstruction path: log-mel -> linear spectrogram -> waveform.

API mirror of crank/utils/utils.py:94-107 (`mlfb2wav`), :210-234 (`logmelspc_to_linearspc`) and :237-269
(`griffin_lim`, which calls librosa.core.griffinlim: the "fast" Griffin-Lim of Perraudin et al. with momentum 0.99,
random initial phases, centred reflect-padded STFT, window-envelope normalised ISTFT).  The reference runs it per
utterance on CPU workers (joblib, basetrainer.py:403); here a whole batch of equal-length utterances iterates on the
GPU: the STFT / ISTFT pairs are batched cuFFT transforms (`torch.stft` / `torch.istft` -- library FFTs, this is the eval
path, not the train step), the phase update is fused elementwise torch arithmetic on the device, nothing returns to
the host until the waveform is done.

Differences from the reference that cannot be avoided: librosa draws the initial phases from numpy's global RNG
(unseeded), so two runs of the reference differ as well; pass `init_angles` (or a torch Generator) for reproducible
output.  With the same initial phases the result matches the numpy restatement oracle/griffinlim.py
(tests/test_cpu_griffinlim.py, tests/test_gpu_eval.py).
"""
import math

import numpy as np
import torch

from ..net.module.mlfb import mel_basis

_PINV = {}


def _inv_mel_basis(fs, n_fft, n_mels, fmin, fmax, device):
    key = (fs, n_fft, n_mels, fmin, fmax, str(device))
    if key not in _PINV:
        basis = mel_basis(fs, n_fft, n_mels, fmin, fmax)              # float32 (n_mels, bins), as librosa returns it
        _PINV[key] = torch.from_numpy(np.linalg.pinv(basis)).to(device)      # (bins, n_mels)
    return _PINV[key]


def logmelspc_to_linearspc(lmspc, fs, n_mels, n_fft, fmin=None, fmax=None):
    """(..., T, n_mels) log10-mel -> (..., T, n_fft // 2 + 1) linear magnitude estimate: pinv(mel_basis) . 10**lmspc."""
    lmspc = torch.as_tensor(lmspc)
    assert lmspc.shape[-1] == n_mels
    fmin = 0 if fmin is None else fmin
    fmax = fs / 2 if fmax is None else fmax
    inv = _inv_mel_basis(fs, n_fft, n_mels, fmin, fmax, lmspc.device)
    return torch.matmul(torch.pow(10.0, lmspc.to(inv.dtype)), inv.t())


def _window(name, win_length, device, dtype):
    if not isinstance(name, str):
        return torch.as_tensor(name, device=device, dtype=dtype)
    if name not in ("hann", "hanning"):
        raise ValueError(f"window {name!r}: only hann is built (every recipe: default.yml:8)")
    return torch.hann_window(win_length, periodic=True, device=device, dtype=dtype)     # scipy get_window(fftbins=True)


def griffin_lim(spc, n_fft, n_shift, win_length, window="hann", n_iters=100, momentum=0.99, init_angles=None,
                generator=None):
    """spc (T, bins) or (B, T, bins) linear magnitude -> waveform (N,) / (B, N), clipped like the reference.

    Same iteration as librosa.core.griffinlim (librosa 0.8, `init="random"`):
        angles <- rebuilt - momentum / (1 + momentum) * previous rebuilt;  angles <- angles / (|angles| + 1e-16)
    `init_angles`: optional complex tensor (B, bins, T) of unit phasors replacing the random initialisation."""
    spc = torch.as_tensor(spc)
    single = spc.dim() == 2
    if single:
        spc = spc[None]
    assert spc.shape[-1] == n_fft // 2 + 1
    S = spc.abs().transpose(1, 2).to(torch.float32)                 # (B, bins, T)
    dev = S.device
    win = _window(window, win_length, dev, torch.float32)
    if init_angles is None:
        ph = torch.rand(S.shape, device=dev, generator=generator)
        angles = torch.polar(torch.ones_like(ph), 2.0 * math.pi * ph)
    else:
        angles = torch.as_tensor(init_angles, device=dev).to(torch.complex64).reshape(S.shape)
    kw = dict(n_fft=n_fft, hop_length=n_shift, win_length=win_length, window=win, center=True)
    rebuilt = None
    coef = momentum / (1.0 + momentum)
    for _ in range(n_iters):
        tprev = rebuilt
        inverse = torch.istft(S * angles, **kw)
        rebuilt = torch.stft(inverse, pad_mode="reflect", return_complex=True, **kw)
        angles = rebuilt if tprev is None else rebuilt - coef * tprev
        angles = angles / (angles.abs() + 1e-16)
    y = torch.istft(S * angles, **kw).clamp_(-1.0, 0.999969482421875)
    return y[0] if single else y


def mlfb2wav(mlfb, fs=22050, n_mels=80, fftl=1024, win_length=1024, hop_size=220, fmin=80, fmax=7600, window="hann",
             n_iters=100, **kw):
    """crank/utils/utils.py:94-107 on the device: (T, n_mels) or (B, T, n_mels) log-mel -> waveform."""
    spc = logmelspc_to_linearspc(mlfb, fs, n_mels, fftl, fmin=fmin, fmax=fmax)
    return griffin_lim(spc, fftl, hop_size, win_length, window=window, n_iters=n_iters, **kw)
