"""Offline log-mel extraction on the GPU (the `mlfb` branch of
`crank.bin.extract_feature`): the counterpart of `Feature._analyze_mlfb` (crank/feature/feature.py:126-145), i.e.
`parallel_wavegan.bin.preprocess.logmelfilterbank` with the window `scipy.signal.hann(win_length)` builds --
SYMMETRIC, unlike the periodic `torch.hann_window` of the on-the-fly layer (mlfb.py:100-101) -- centred frames with
reflect padding, Slaney mel basis, `log10(max(eps, |STFT| . mel))`.

Same kernels as the training front end (`crk_logmel_fwd`: frame + window -> cuFFT R2C -> |.| -> mel GEMM -> log10);
only the window samples and the padding differ.  fp32 on the device against the reference's float64 on the CPU:
the reference's own test of this boundary allows 1e-3 / 1e-5 (test/test_feature_pytorch.py:80-127).
Everything else of `crank/feature` (WORLD analysis, mcep, Griffin-Lim, HDF5 I/O) stays out of scope.
"""
import numpy as np
import torch

from . import ops
from .net.module.mlfb import mel_basis

EPS = 1e-10


def symmetric_hann(n):
    """scipy.signal.hann(n) == scipy.signal.windows.hann(n, sym=True)"""
    if n == 1:
        return np.ones(1)
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / (n - 1))


def extract_mlfb(wav, fs=22050, fft_size=1024, hop_size=128, win_length=1024, n_mels=80, fmin=80, fmax=7600,
                 window=None, eps=EPS, device="cuda"):
    """wav: (n_samples,) or (B, n_samples) float array / tensor in [-1, 1] -> (frames, n_mels) or (B, frames, n_mels)
    float32 tensor on `device`; frames = 1 + n_samples // hop_size (centred)."""
    x = torch.as_tensor(np.asarray(wav, dtype=np.float32) if not isinstance(wav, torch.Tensor) else wav).float()
    single = x.dim() == 1
    if single:
        x = x[None]
    x = x.to(device)
    win = symmetric_hann(win_length) if window is None else np.asarray(window, dtype=np.float64)
    if len(win) != win_length:
        raise ValueError(f"window has {len(win)} samples, win_length is {win_length}")
    if win_length < fft_size:                      # librosa.stft centres a short window inside the FFT frame
        lpad = (fft_size - win_length) // 2
        win = np.pad(win, (lpad, fft_size - win_length - lpad))
    wt = torch.from_numpy(win.astype(np.float32)).to(device)
    basis = torch.from_numpy(mel_basis(fs, fft_size, n_mels, fmin, fmax).T.copy()).float().to(device)    # (bins, mels)
    x = torch.nn.functional.pad(x.unsqueeze(1), (fft_size // 2, fft_size // 2), mode="reflect").squeeze(1)
    out = ops.logmel(x, wt, basis, fft_size, hop_size, eps=eps)
    return out[0] if single else out
