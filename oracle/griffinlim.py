"""TEST INFRASTRUCTURE (oracle): numpy restatement of the Griffin-Lim path of crank/utils/utils.py:210-269.

`griffin_lim` there is `librosa.core.griffinlim(S, n_iter, hop_length, win_length, window)` (librosa is an absent,
un-pinned dependency: tools/requirements.txt; 0.8.x era) clipped to [-1, 1 - 2^-15].  Restated from librosa's published
algorithm: stft = reflect-centred frames x periodic hann -> rfft; istft = irfft x window, overlap-add, divided by the
window sum-square envelope where it exceeds `tiny`, trimmed by n_fft // 2 at both ends; fast Griffin-Lim update with
momentum 0.99.  The random initial phases are an explicit argument here (librosa draws them from numpy's global RNG).
Parity: unpinned by the reference (it has no test of this function); pinned structurally by the round trip
istft(stft(x)) == x below and used as the checker of crank_b200/utils/griffin_lim.py.
"""
import numpy as np

from . import mel


def stft(y, n_fft, hop, win):
    y = np.pad(np.asarray(y, dtype=np.float64), n_fft // 2, mode="reflect")
    n_frames = 1 + (len(y) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    return np.fft.rfft(y[idx] * win[None, :], axis=1).T              # (bins, frames)


def istft(D, n_fft, hop, win):
    n_frames = D.shape[1]
    ytmp = np.fft.irfft(D.T, n=n_fft, axis=1) * win[None, :]
    n = n_fft + hop * (n_frames - 1)
    y = np.zeros(n)
    env = np.zeros(n)
    for m in range(n_frames):
        y[m * hop:m * hop + n_fft] += ytmp[m]
        env[m * hop:m * hop + n_fft] += win ** 2
    nz = env > np.finfo(np.float32).tiny
    y[nz] /= env[nz]
    return y[n_fft // 2:n - n_fft // 2]


def logmelspc_to_linearspc(lmspc, fs, n_mels, n_fft, fmin=None, fmax=None):
    fmin = 0 if fmin is None else fmin
    fmax = fs / 2 if fmax is None else fmax
    basis = mel.mel_basis(fs, n_fft, n_mels, fmin, fmax)
    return np.matmul(np.linalg.pinv(basis), np.power(10.0, lmspc).T).T


def griffin_lim(spc, n_fft, n_shift, win_length, init_angles, n_iters=100, momentum=0.99):
    """spc (T, bins); init_angles complex (bins, T)."""
    assert win_length == n_fft
    S = np.abs(spc.T).astype(np.float64)
    win = mel.hann(win_length, periodic=True)
    angles = np.asarray(init_angles, dtype=np.complex128)
    rebuilt = 0.0
    for _ in range(n_iters):
        tprev = rebuilt
        inverse = istft(S * angles, n_fft, n_shift, win)
        rebuilt = stft(inverse, n_fft, n_shift, win)
        angles = rebuilt - (momentum / (1 + momentum)) * tprev
        angles = angles / (np.abs(angles) + 1e-16)
    return np.clip(istft(S * angles, n_fft, n_shift, win), -1, 0.999969482421875)
