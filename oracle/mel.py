"""ORACLE (test infrastructure, NOT product code) -- mel front-end restatement in numpy.

Restates the two third-party functions the reference's front end calls and that are absent
from /root/reference and this image:

  * `librosa.filters.mel` (Slaney mel scale + Slaney area norm)  <- crank/net/module/mlfb.py:27-33
  * `parallel_wavegan.bin.preprocess.logmelfilterbank`           <- crank/feature/feature.py:134-145

Pinned against the reference's own fixture: the `mlfb` dataset stored in
/root/reference/test/data/SF1/SF1_10001.feats.h5 (float64, 1057x80, computed from
test/data/SF1_10001.wav by the reference's extract_feature) -- tests/test_oracle_mel.py
checks `logmelfilterbank` against the committed copy of that matrix
(tests/golden/ref_fixture_mlfb.npz, made by tests/golden/make_golden.py).
"""

import numpy as np


def hz_to_mel(f):
    """Slaney (htk=False) mel scale."""
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore", invalid="ignore"):
        log_t = f >= min_log_hz
        mels = np.where(
            log_t, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels
        )
    return mels


def mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    log_t = m >= min_log_mel
    return np.where(log_t, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_basis(sr, n_fft, n_mels=80, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm='slaney') -> (n_mels, 1+n_fft//2) float32."""
    if fmax is None:
        fmax = float(sr) / 2
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0, float(sr) / 2, n_bins, endpoint=True)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights.astype(np.float32)


def hann(n, periodic):
    """periodic=True == torch.hann_window(n); periodic=False == scipy.signal.hann(n) (symmetric)."""
    k = np.arange(n, dtype=np.float64)
    den = n if periodic else n - 1
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / den)


def stft_mag(x, n_fft, hop, window, center=True):
    """|STFT| -> (frames, 1+n_fft//2); `window` already has length n_fft."""
    x = np.asarray(x, dtype=np.float64)
    if center:
        x = np.pad(x, n_fft // 2, mode="reflect")
    n_frames = 1 + (len(x) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = x[idx] * window[None, :]
    return np.abs(np.fft.rfft(frames, n=n_fft, axis=1))


def logmelfilterbank(
    audio,
    sampling_rate,
    fft_size=1024,
    hop_size=256,
    win_length=None,
    window="hann",
    num_mels=80,
    fmin=None,
    fmax=None,
    eps=1e-10,
):
    """parallel_wavegan's offline log-mel: librosa.stft(center=True, reflect) -> |.| -> mel -> log10.

    `window` may be a name ("hann": scipy get_window(..., fftbins=True) == periodic, as librosa does
    for a string) or an explicit array (crank passes scipy's *symmetric* hann, feature.py:174).
    """
    if win_length is None:
        win_length = fft_size
    if isinstance(window, str):
        assert window == "hann"
        win = hann(win_length, periodic=True)
    else:
        win = np.asarray(window, dtype=np.float64)
    if win_length < fft_size:  # librosa pad_center
        lpad = (fft_size - win_length) // 2
        win = np.pad(win, (lpad, fft_size - win_length - lpad))
    spc = stft_mag(audio, fft_size, hop_size, win, center=True)
    fmin = 0 if fmin is None else fmin
    fmax = sampling_rate / 2 if fmax is None else fmax
    basis = mel_basis(sampling_rate, fft_size, num_mels, fmin, fmax)
    return np.log10(np.maximum(eps, np.dot(spc, basis.T.astype(np.float64))))
