"""On-GPU log-mel front end: frame+window -> cuFFT R2C -> |.| -> mel GEMM -> clamp/log10 -> scaler.

API mirror of crank/net/module/mlfb.py:19-171 (LogMelFilterBankLayer and its three sub-layers).
The mel basis is the Slaney-scale / Slaney-norm filterbank `librosa.filters.mel` would build
(mlfb.py:27-33); librosa is not a dependency here, the basis is computed in `mel_basis()` below.
n_fft = 1024 (every recipe) runs ONE fused kernel (crk_logmel_fused_fwd: framing + window + FFT + banded mel + log10 +
scaler); the learnable "param" / "conv" windows of mlfb.py:72-90 get their gradients from crk_logmel_fused_bwd.
"""

import numpy as np
import torch
import torch.nn as nn

from ... import ops


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    lin = f / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) / logstep, lin)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    logstep = np.log(6.4) / 27.0
    return np.where(m >= 15.0, 1000.0 * np.exp(logstep * (m - 15.0)), f_sp * m)


def mel_basis(sr, n_fft, n_mels=80, fmin=0.0, fmax=None):
    """(n_mels, 1+n_fft//2) float32 triangular filters, Slaney mel scale, area-normalised."""
    fmax = float(sr) / 2 if fmax is None else fmax
    freqs = np.linspace(0, float(sr) / 2, 1 + n_fft // 2)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    width = np.diff(edges)
    ramps = edges[:, None] - freqs[None, :]
    lower = -ramps[:-2] / width[:-1, None]
    upper = ramps[2:] / width[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return w.astype(np.float32)


class MLFBLayer(nn.Module):
    def __init__(self, fs=22050, fft_size=1024, n_mels=80, fmin=None, fmax=None, eps=1.0e-10):
        super().__init__()
        fmin = 0 if fmin is None else fmin
        fmax = fs / 2 if fmax is None else fmax
        self.eps = eps
        basis = mel_basis(fs, fft_size, n_mels, fmin, fmax)
        self.register_buffer("mel_basis", torch.from_numpy(basis.T.copy()).float())  # (bins, n_mels)

    def forward(self, x):
        """(…, bins) magnitude spectrogram -> log10 mel (mlfb.py:36-43).  Stand-alone use only (plain torch ops on whatever
        device x lives on); inside LogMelFilterBankLayer the projection is part of the fused kernel."""
        return torch.clamp(torch.matmul(x, self.mel_basis), min=self.eps).log10()


class STFTLayer(nn.Module):
    """Holds the STFT geometry and window of crank/net/module/mlfb.py:45-110; the transform itself runs inside the
    fused log-mel kernel (LogMelFilterBankLayer.forward).  Window types (mlfb.py:64-90):
      "hann"   fixed periodic hann;
      "param"  the window is a learnable parameter `window` (initialised to scipy get_window("hann") = periodic hann);
               its gradient comes from crk_logmel_fused_bwd;
      "conv"   a learnable pre-filter `window_conv` = Conv1d(1, 24, 65, padding 32) + Sigmoid, averaged over its 24
               channels, replaces the waveform and the STFT uses a rectangular window; the 24-channel pre-filter is a
               stock torch conv (cuDNN, off the default path: `raw_window_type: hann` in every recipe), the gradient
               reaches it through the d wav output of crk_logmel_fused_bwd."""

    def __init__(self, fs=22050, hop_size=256, fft_size=1024, win_length=None, window="hann",
                 center=True, pad_mode="reflect", return_complex=False):
        super().__init__()
        if window not in ("hann", "param", "conv"):
            raise NotImplementedError(f"window {window!r}: hann / param / conv are built")
        self.hop_size = hop_size
        self.fft_size = fft_size
        self.win_length = fft_size if win_length is None else win_length
        self.center = center
        self.pad_mode = pad_mode
        self.window_type = window
        if window == "param":
            if self.win_length != fft_size:
                raise NotImplementedError("the learnable window spans the whole FFT frame (win_length == fft_size)")
            self.register_parameter("window", nn.Parameter(torch.hann_window(self.win_length), requires_grad=True))
        elif window == "conv":
            kernel_size = 65
            self.window_conv = nn.Sequential(
                nn.Conv1d(in_channels=1, out_channels=24, kernel_size=kernel_size, stride=1,
                          padding=(kernel_size - 1) // 2),
                nn.Sigmoid(),
            )
        win = torch.hann_window(self.win_length) if window != "conv" else torch.ones(self.win_length)
        if self.win_length < fft_size:
            lpad = (fft_size - self.win_length) // 2
            win = torch.nn.functional.pad(win, (lpad, fft_size - self.win_length - lpad))
        self.register_buffer("window_padded", win, persistent=False)

    @property
    def learnable(self):
        return self.window_type in ("param", "conv")

    def forward(self, x):
        """(B, n_samples) -> (B, frames, bins, 2) real / imaginary STFT like the reference layer (mlfb.py:92-110).  Stand-alone
        use only (`torch.stft`, a library FFT): LogMelFilterBankLayer.forward never calls it, its transform is in-kernel."""
        window = None
        if self.window_type == "param":
            window = self.window
        elif self.window_type == "conv":
            x = self.window_conv(x.unsqueeze(1)).mean(dim=1)
        else:
            window = torch.hann_window(self.win_length, dtype=x.dtype, device=x.device)
        st = torch.stft(x, n_fft=self.fft_size, win_length=self.win_length, hop_length=self.hop_size, window=window,
                        center=self.center, pad_mode=self.pad_mode, return_complex=True)
        return torch.view_as_real(st).transpose(1, 2).float()


class MLFBScalerLayer(nn.Module):
    def __init__(self, scaler):
        super().__init__()
        self.register_parameter(
            "mean", nn.Parameter(torch.from_numpy(scaler.mean_).float(), requires_grad=False))
        self.register_parameter(
            "std", nn.Parameter(torch.from_numpy(scaler.var_).float().sqrt(), requires_grad=False))

    def forward(self, x):
        return (x - self.mean) / self.std


class LogMelFilterBankLayer(nn.Module):
    def __init__(self, fs=22050, hop_size=256, fft_size=1024, win_length=None, window="hann",
                 center=True, pad_mode="reflect", n_mels=80, fmin=None, fmax=None, scaler=None):
        super().__init__()
        self.stft_layer = STFTLayer(fs, hop_size, fft_size, win_length, window, center=center,
                                    pad_mode=pad_mode)
        self.mlfb_layer = MLFBLayer(fs, fft_size, n_mels, fmin, fmax)
        self.scaler_layer = MLFBScalerLayer(scaler) if scaler is not None else None

    def forward(self, x):
        """x (B, n_samples) raw waveform -> (B, n_frames, n_mels) log10 mel (optionally standardised)."""
        st = self.stft_layer
        x = x.float()
        mean = std = None
        if self.scaler_layer is not None:
            mean, std = self.scaler_layer.mean.data, self.scaler_layer.std.data
        if st.learnable:
            if st.window_type == "conv":      # mlfb.py:93-96
                x = st.window_conv(x.unsqueeze(1)).mean(dim=1)
            if st.center:
                x = torch.nn.functional.pad(x.unsqueeze(1), (st.fft_size // 2, st.fft_size // 2),
                                            mode=st.pad_mode).squeeze(1)
            window = st.window if st.window_type == "param" else st.window_padded
            return ops.logmel_learnable(x, window, self.mlfb_layer.mel_basis, st.fft_size, st.hop_size,
                                        eps=self.mlfb_layer.eps, mean=mean, std=std)
        with torch.no_grad():
            if st.center:
                x = torch.nn.functional.pad(x.unsqueeze(1), (st.fft_size // 2, st.fft_size // 2),
                                            mode=st.pad_mode).squeeze(1)
            return ops.logmel(x, st.window_padded, self.mlfb_layer.mel_basis, st.fft_size, st.hop_size,
                              eps=self.mlfb_layer.eps, mean=mean, std=std)
